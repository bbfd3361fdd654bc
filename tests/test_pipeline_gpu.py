"""Pipelined projection launches (ccp_project_batch_pipelined / ccp_project_flush): the samples a launch parks when
its seed list runs dry are finished by its successors, with results bit-identical to the complete-mode launch."""
import numpy as np
import pytest

from conftest import make_oracles

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _bits(t):
    return t.cpu().numpy().view(np.uint64)


@pytest.fixture(scope="module")
def setup():
    import closed_chain_motion_planner_b200 as pkg

    assert torch.cuda.is_available()
    cfg, A, B = make_oracles("dumbbell")
    c = pkg.KinematicChainConstraint.from_config("dumbbell", device=0)
    return pkg, c, A


@pytest.mark.parametrize("layout", ["aos", "soa"])
@pytest.mark.parametrize("count", [300_000, 5_000, 37])
def test_pipelined_equals_complete(setup, layout, count):
    pkg, c, A = setup
    lay = pkg.CCP_LAYOUT_AOS if layout == "aos" else pkg.CCP_LAYOUT_SOA
    batches = []
    for b in range(3):
        s = A.seeds_uniform(0, b * count, count)
        s = s if layout == "aos" else np.ascontiguousarray(s.T)
        batches.append(torch.from_numpy(s).cuda())
    ref = [c.projectBatch(x, layout=lay) for x in batches]
    torch.cuda.synchronize()
    n = 14
    compact = torch.zeros((3 * count, n), dtype=torch.float64, device="cuda")
    n_ok = torch.zeros(1, dtype=torch.int64, device="cuda")
    res = [c.projectBatch(x, layout=lay, compact=compact, n_ok=n_ok, pipelined=True) for x in batches]
    assert c.pipelineOpen()
    with pytest.raises(Exception):
        c.setTolerance(1e-3, 5e-3)  # setters are refused while samples are parked
    c.flush(compact=compact, n_ok=n_ok)
    torch.cuda.synchronize()
    assert not c.pipelineOpen()
    total_ok = 0
    ok_rows = []
    for r, p in zip(ref, res):
        assert np.array_equal(_bits(r.x), _bits(p.x))
        assert torch.equal(r.ok, p.ok) and torch.equal(r.converged, p.converged) and torch.equal(r.iters, p.iters)
        assert np.array_equal(_bits(r.resid), _bits(p.resid))
        total_ok += int(r.ok.sum())
        xa = r.x if layout == "aos" else r.x.T
        ok_rows.append(xa[r.ok.bool()].cpu().numpy())
    # the compacted stream holds exactly the ok states of the three batches (order unspecified)
    assert int(n_ok.item()) == total_ok
    got = compact[:total_ok].cpu().numpy()
    want = np.concatenate(ok_rows)
    key = lambda a: a[np.lexsort(a.T[::-1])]
    assert np.array_equal(key(got).view(np.uint64), key(want).view(np.uint64))


def test_complete_launch_closes_the_pipeline(setup):
    pkg, c, A = setup
    x0 = torch.from_numpy(A.seeds_uniform(0, 0, 100_000)).cuda()
    x1 = torch.from_numpy(A.seeds_uniform(0, 100_000, 100_000)).cuda()
    r0, r1 = c.projectBatch(x0), c.projectBatch(x1)
    p0 = c.projectBatch(x0, pipelined=True)
    assert c.pipelineOpen()
    p1 = c.projectBatch(x1)  # adopts what p0 parked, completes everything
    torch.cuda.synchronize()
    assert not c.pipelineOpen()
    for r, p in ((r0, p0), (r1, p1)):
        assert np.array_equal(_bits(r.x), _bits(p.x)) and torch.equal(r.iters, p.iters) and torch.equal(r.ok, p.ok)


def test_layout_change_flushes(setup):
    pkg, c, A = setup
    s = A.seeds_uniform(0, 0, 60_000)
    xa = torch.from_numpy(s).cuda()
    xs = torch.from_numpy(np.ascontiguousarray(s.T)).cuda()
    ra = c.projectBatch(xa)
    pa = c.projectBatch(xa, pipelined=True)
    ps = c.projectBatch(xs, layout=pkg.CCP_LAYOUT_SOA, pipelined=True)  # other instantiation: AOS samples are flushed first
    c.flush()
    torch.cuda.synchronize()
    assert np.array_equal(_bits(ra.x), _bits(pa.x))
    assert np.array_equal(_bits(ra.x), _bits(ps.x.T.contiguous()))


def test_many_small_pipelined_launches(setup):
    """A capped sample (250 iterations) is carried through many tiny launches; the slot ring stays consistent."""
    pkg, c, A = setup
    s = A.seeds_uniform(0, 0, 80 * 256)
    full = c.projectBatch(torch.from_numpy(s).cuda())
    parts = []
    for b in range(80):
        parts.append(c.projectBatch(torch.from_numpy(s[b * 256:(b + 1) * 256]).cuda(), pipelined=True))
    c.flush()
    torch.cuda.synchronize()
    x = torch.cat([p.x for p in parts])
    it = torch.cat([p.iters for p in parts])
    assert np.array_equal(_bits(full.x), _bits(x)) and torch.equal(full.iters, it)
    assert int(full.iters.max()) == 250


def _np(v):
    return v.cpu().numpy() if hasattr(v, "cpu") else np.asarray(v)


def _same(r, g):
    assert np.array_equal(_np(r.x).view(np.uint64), _np(g.x).view(np.uint64))
    for name in ("ok", "converged", "iters"):
        assert np.array_equal(_np(getattr(r, name)), _np(getattr(g, name))), name
    assert np.array_equal(_np(r.resid).view(np.uint64), _np(g.resid).view(np.uint64))


@pytest.mark.parametrize("pinned", [False, True])
@pytest.mark.parametrize("count", [1_000_000, 200_000])
def test_chunked_host_call_equals_device_launch(setup, count, pinned):
    """ccp_project_batch_host on a batch large enough to be chunked: a sample is carried through several of the
    call's launches, and every chunk's copy-out must wait for the launch that completes it.  A first call with other
    seeds leaves different results in the stage, so a chunk copied early would show."""
    pkg, c, A = setup
    c.projectBatch(A.seeds_uniform(7, 0, count), pinned=pinned)
    x = A.seeds_uniform(0, 12345, count)
    got = c.projectBatch(x, pinned=pinned)  # numpy in: the host entry point
    ref = c.projectBatch(torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    _same(ref, got)


@pytest.mark.parametrize("pinned", [False, True])
@pytest.mark.parametrize("count", [1_000_000, 70_000, 13])
def test_streaming_host_submit_wait(setup, count, pinned):
    """ccp_project_batch_host_submit / _wait: two host batches in flight; each result bit-identical to the complete
    device launch's, whatever launch finished the stragglers."""
    pkg, c, A = setup
    batches = [A.seeds_uniform(0, b * count, count) for b in range(4)]
    ref = [c.projectBatch(torch.from_numpy(x).cuda()) for x in batches]
    torch.cuda.synchronize()
    pending = []
    got = []
    for x in batches:
        pending.append(c.submitHostBatch(x, want_resid=True, pinned=pinned))
        if len(pending) == 2:
            t, r = pending.pop(0)
            c.waitHostBatch(t)
            got.append(r)
    while pending:
        t, r = pending.pop(0)
        c.waitHostBatch(t)
        got.append(r)
    assert not c.pipelineOpen()
    for r, g in zip(ref, got):
        _same(r, g)
    # a ticket never issued is an error and leaves the handle usable; a synchronous call may follow a submit
    with pytest.raises(Exception):
        c.waitHostBatch(10_000_000)
    t, r = c.submitHostBatch(batches[1], pinned=pinned)
    if pinned:
        torch.cuda.synchronize()  # a device-wide synchronisation between submit and wait is allowed
    _same(ref[0], c.projectBatch(batches[0], pinned=not pinned))
    c.waitHostBatch(t)
    assert np.array_equal(_np(ref[1].ok), r.ok)


def test_streaming_host_random_sizes_and_buffer_kinds(setup):
    """Batches of very different sizes, page-locked and pageable results mixed (the two copy schedules hand over to
    each other), a blocking call and a device-wide synchronisation thrown in: every result equals the device launch's."""
    pkg, c, A = setup
    c_ref = pkg.KinematicChainConstraint.from_config("dumbbell", device=0)
    rng = np.random.default_rng(11)
    sizes = [1, 100, 600, 5_000, 250_000, 500_000, 1_200_000]
    pending = []
    first = 0
    for b in range(16):
        count = int(rng.choice(sizes))
        x = A.seeds_uniform(2, first, count)
        first += count
        # reference results from a SECOND handle: device-pointer projections on `c` are refused while one of its host
        # tickets is pending (and every submit now really hands over to the previous batch: no launch in between)
        ref = c_ref.projectBatch(torch.from_numpy(x).cuda(), want_resid=False)
        pinned = bool(rng.integers(2))
        pending.append((c.submitHostBatch(x, pinned=pinned), ref))
        if b == 5:
            torch.cuda.synchronize()
        if b == 9:
            xs = A.seeds_uniform(5, 0, 3_000)
            rs = c.projectBatch(xs)  # blocking host call: completes the pending tickets first
            rd = c_ref.projectBatch(torch.from_numpy(xs).cuda())
            assert np.array_equal(_np(rd.x).view(np.uint64), rs.x.view(np.uint64))
        while len(pending) > (0 if b == 15 else 1):
            (t, r), ref = pending.pop(0)
            c.waitHostBatch(t)
            assert np.array_equal(_np(ref.x).view(np.uint64), r.x.view(np.uint64)), (b, count, pinned)
            assert np.array_equal(_np(ref.ok), r.ok) and np.array_equal(_np(ref.iters), r.iters)
    assert not c.pipelineOpen()


def test_registered_caller_buffers(setup):
    """ccp_host_register: memory the caller already owns is page-locked in place and the chunked host call then takes
    the completion-count schedule; ccp_host_alloc memory likewise.  Results equal the device launch's."""
    import ctypes as C

    pkg, c, A = setup
    lib, h = c._lib, c._h
    count, n = 300_000, 14
    x = A.seeds_uniform(9, 0, count)
    ref = c.projectBatch(torch.from_numpy(x).cuda(), want_resid=False)
    torch.cuda.synchronize()
    xo = np.zeros((count, n))
    ok = np.zeros(count, np.uint8)
    it = np.zeros(count, np.int32)
    for a in (x, xo, ok, it):
        assert lib.ccp_host_register(a.ctypes.data, a.nbytes) == 0
    try:
        assert lib.ccp_project_batch_host(h, x.ctypes.data, count, xo.ctypes.data, ok.ctypes.data, None, it.ctypes.data, None) == 0
    finally:
        for a in (x, xo, ok, it):
            assert lib.ccp_host_unregister(a.ctypes.data) == 0
    assert np.array_equal(_np(ref.x).view(np.uint64), xo.view(np.uint64))
    assert np.array_equal(_np(ref.ok), ok) and np.array_equal(_np(ref.iters), it)
    p = C.c_void_p()
    assert lib.ccp_host_alloc(C.byref(p), xo.nbytes) == 0 and p.value
    try:
        xa = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(count, n))
        assert lib.ccp_project_batch_host(h, x.ctypes.data, count, p, None, None, None, None) == 0
        assert np.array_equal(_np(ref.x).view(np.uint64), xa.view(np.uint64))
    finally:
        lib.ccp_host_free(p)


def test_device_calls_refused_while_host_ticket_pending(setup):
    """ccp.h: no other projection call on the handle while a ticket is pending.  The library enforces it
    (CCP_ERR_STATE) instead of letting a device launch on the caller's stream adopt the host batch's parked samples."""
    pkg, c, A = setup
    x = A.seeds_uniform(4, 0, 400_000)
    xd = torch.from_numpy(x[:1000]).cuda()
    ref = c.projectBatch(torch.from_numpy(x).cuda(), want_resid=False)
    torch.cuda.synchronize()
    t, r = c.submitHostBatch(x, pinned=True)
    with pytest.raises(pkg.CcpError, match="host batch is pending"):
        c.projectBatch(xd)
    with pytest.raises(pkg.CcpError, match="host batch is pending"):
        c.projectBatch(xd, pipelined=True)
    with pytest.raises(pkg.CcpError, match="host batch is pending"):
        c.flush()
    c.waitHostBatch(t)
    assert np.array_equal(_np(ref.x).view(np.uint64), r.x.view(np.uint64)) and np.array_equal(_np(ref.ok), r.ok)
    r2 = c.projectBatch(xd)  # usable again
    torch.cuda.synchronize()
    assert np.array_equal(_np(r2.x).view(np.uint64), r.x[:1000].view(np.uint64))


@pytest.mark.parametrize("pinned", [True, False])
def test_streaming_host_growing_batches_back_to_back(pinned):
    """Batches of growing size submitted back to back with nothing in between: every submit outgrows its slot's device
    stage while the previous batch still has chunks in flight that only THIS batch's launches finish.  (The stage used to
    be regrown with cudaFree/cudaMalloc — a device-wide synchronisation — after the previous batch's stream waits were
    already enqueued: a hang.)  A fresh handle, so that no stage exists beforehand."""
    import closed_chain_motion_planner_b200 as pkg
    from conftest import make_oracles as mk

    cfg, A, B = mk("dumbbell")
    c = pkg.KinematicChainConstraint.from_config("dumbbell", device=0)
    c_ref = pkg.KinematicChainConstraint.from_config("dumbbell", device=0)
    sizes = [240_000, 320_000, 420_000, 560_000, 740_000, 1_000_000]
    pending, first = [], 0
    for count in sizes:
        x = A.seeds_uniform(6, first, count)
        first += count
        ref = c_ref.projectBatch(torch.from_numpy(x).cuda(), want_resid=False)
        pending.append((c.submitHostBatch(x, pinned=pinned), ref))
        if len(pending) == 2:
            (t, r), rf = pending.pop(0)
            c.waitHostBatch(t)
            assert np.array_equal(_np(rf.x).view(np.uint64), r.x.view(np.uint64)), count
            assert np.array_equal(_np(rf.ok), r.ok) and np.array_equal(_np(rf.iters), r.iters)
    (t, r), rf = pending.pop(0)
    c.waitHostBatch(t)
    assert np.array_equal(_np(rf.x).view(np.uint64), r.x.view(np.uint64))
    assert not c.pipelineOpen()
    del c  # ccp_destroy synchronises the device: must not hang on a leftover stream wait


def test_handles_on_two_devices_in_one_process():
    """cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: a second handle on another GPU of the same process
    must raise the limit there too (it used to be done once per process)."""
    import closed_chain_motion_planner_b200 as pkg
    from conftest import make_oracles as mk

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in one process")
    cfg, A, B = mk("dumbbell")
    x = A.seeds_uniform(1, 0, 20_000)
    out = []
    for dev in (0, 1):
        c = pkg.KinematicChainConstraint.from_config("dumbbell", device=dev)
        r = c.projectBatch(torch.from_numpy(x).to(f"cuda:{dev}"))
        torch.cuda.synchronize(dev)
        out.append((r.x.cpu().numpy(), r.ok.cpu().numpy()))
    assert np.array_equal(out[0][0].view(np.uint64), out[1][0].view(np.uint64)) and np.array_equal(out[0][1], out[1][1])


def _compact_rows_equal(ref_x, ref_ok, res, n_ok):
    """packed rows == the reference launch's ok rows, matched through the packed seed indices"""
    okm = _np(ref_ok).astype(bool)
    assert n_ok == int(okm.sum())
    idx = res.index[:n_ok]
    assert np.array_equal(np.sort(idx), np.nonzero(okm)[0])  # every ok seed exactly once
    assert np.array_equal(res.states[:n_ok].view(np.uint64), _np(ref_x)[idx].view(np.uint64))


@pytest.mark.parametrize("pinned", [True, False])
@pytest.mark.parametrize("count", [1_000_000, 70_000, 13])
def test_compact_host_batches_streaming(setup, count, pinned):
    """ccp_host_batch_submit with compact outputs: only the ok states come back, packed per batch (a straggler finished
    by the NEXT batch's launch still lands in its own batch's rows), with the seed index of every row."""
    pkg, c, A = setup
    c_ref = pkg.KinematicChainConstraint.from_config("dumbbell", device=0)
    batches = [A.seeds_uniform(0, b * count, count) for b in range(4)]
    refs = [c_ref.projectBatch(torch.from_numpy(x).cuda(), want_resid=False) for x in batches]
    torch.cuda.synchronize()
    pending, got = [], []
    for x in batches:
        pending.append(c.submitCompactBatch(x, want_flags=True, pinned=pinned))  # pageable: the launch-lag copy schedule
        if len(pending) == 2:
            t, r = pending.pop(0)
            got.append((r, c.waitCompactBatch(t, r)))
    while pending:
        t, r = pending.pop(0)
        got.append((r, c.waitCompactBatch(t, r)))
    assert not c.pipelineOpen()
    for rf, (r, nk) in zip(refs, got):
        _compact_rows_equal(rf.x, rf.ok, r, nk)
        assert np.array_equal(_np(rf.ok), r.ok) and np.array_equal(_np(rf.iters), r.iters)
    # a capacity smaller than the number of ok states: the count is still exact, the rows that fit are valid ok states
    t, r = c.submitCompactBatch(batches[0], capacity=5)
    nk = c.waitCompactBatch(t, r)
    assert nk == int(_np(refs[0].ok).sum())
    k = min(nk, 5)
    okrows = _np(refs[0].x)[_np(refs[0].ok).astype(bool)]
    assert np.all(_np(refs[0].ok)[r.index[:k]] == 1)
    assert np.array_equal(r.states[:k].view(np.uint64), _np(refs[0].x)[r.index[:k]].view(np.uint64))
    assert okrows.shape[0] == nk


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_seeded_compact_host_batches_equal_sampler(setup, mode):
    """Sampler arguments instead of host states: the seeds are generated on the device (no H2D), the packed ok states
    equal those of ccp_sample_project_batch on the same counter range, wrap included."""
    import ctypes as C

    from closed_chain_motion_planner_b200 import _capi

    pkg, c, A = setup
    count = 300_000
    near = c.config.start.copy()
    nearp = near.ctypes.data_as(C.POINTER(C.c_double))
    outs = []
    for b in range(3):
        a = _capi.SamplerArgs(rng_seed=5, first_index=b * count, mode=mode, wrap_bounds=1, distance=0.4, near_host=nearp)
        x = torch.empty((count, 14), dtype=torch.float64, device="cuda")
        ok = torch.empty(count, dtype=torch.uint8, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        assert c._lib.ccp_sample_project_batch(c._h, C.byref(a), count, 0, x.data_ptr(), ok.data_ptr(), None, None, None, st) == 0
        torch.cuda.synchronize()
        outs.append((x.cpu().numpy(), ok.cpu().numpy()))
    pending, got = [], []
    for b in range(3):
        a = _capi.SamplerArgs(rng_seed=5, first_index=b * count, mode=mode, wrap_bounds=1, distance=0.4, near_host=nearp)
        pending.append(c.submitCompactBatch(sampler=a, count=count))
        if len(pending) == 2:
            t, r = pending.pop(0)
            got.append((r, c.waitCompactBatch(t, r)))
    while pending:
        t, r = pending.pop(0)
        got.append((r, c.waitCompactBatch(t, r)))
    for (x, ok), (r, nk) in zip(outs, got):
        _compact_rows_equal(x, ok, r, nk)
    assert not c.pipelineOpen()


def test_three_arm_compact_and_seeded_host_batches():
    """21-DoF states through the compact-output and device-seeded forms of the streaming host path."""
    import closed_chain_motion_planner_b200 as pkg
    from closed_chain_motion_planner_b200 import _capi

    c = pkg.KinematicChainConstraint.from_config("stefan_three_arm", device=0)
    c_ref = pkg.KinematicChainConstraint.from_config("stefan_three_arm", device=0)
    rng = np.random.default_rng(5)
    count = 260_000
    x = c.config.start[None, :] + 0.08 * rng.standard_normal((count, 21))
    ref = c_ref.projectBatch(torch.from_numpy(x).cuda(), want_resid=False)
    torch.cuda.synchronize()
    t0, r0 = c.submitCompactBatch(x[: count // 2], want_flags=True)
    t1, r1 = c.submitCompactBatch(x[count // 2:], want_flags=True)
    n0 = c.waitCompactBatch(t0, r0)
    n1 = c.waitCompactBatch(t1, r1)
    okm = _np(ref.ok).astype(bool)
    h = count // 2
    assert n0 == int(okm[:h].sum()) and n1 == int(okm[h:].sum()) and n0 > 0
    assert np.array_equal(r0.states[:n0].view(np.uint64), _np(ref.x)[:h][r0.index[:n0]].view(np.uint64))
    assert np.array_equal(r1.states[:n1].view(np.uint64), _np(ref.x)[h:][r1.index[:n1]].view(np.uint64))
    assert np.array_equal(np.sort(r1.index[:n1]), np.nonzero(okm[h:])[0])
    # device-generated seeds: the packed rows equal those of the device sampler call on the same counter range
    import ctypes as C

    a = _capi.SamplerArgs(rng_seed=2, first_index=10, mode=0, wrap_bounds=0, distance=0.0, near_host=None)
    xs = torch.empty((50_000, 21), dtype=torch.float64, device="cuda")
    oks = torch.empty(50_000, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    assert c_ref._lib.ccp_sample_project_batch(c_ref._h, C.byref(a), 50_000, 0, xs.data_ptr(), oks.data_ptr(), None, None, None, st) == 0
    torch.cuda.synchronize()
    t, r = c.submitCompactBatch(sampler=a, count=50_000)
    nk = c.waitCompactBatch(t, r)
    _compact_rows_equal(xs, oks, r, nk)
