"""Edge cases of the projection entry points: empty / tiny / ragged batches, odd SoA strides and misaligned seed
pointers (the TMA staging must fall back to plain loads), NaN seeds, in-place projection, the chunked sampler path."""
import ctypes as C

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
    cfg, A, B = make_oracles("Wine_Bottle")
    c = pkg.KinematicChainConstraint.from_config("Wine_Bottle", device=0)
    return pkg, c, A, B


@pytest.mark.parametrize("count", [0, 1, 31, 32, 33, 1023, 148 * 384 + 1])
@pytest.mark.parametrize("layout", ["aos", "soa"])
def test_ragged_counts_bit_exact(setup, count, layout):
    pkg, c, A, B = setup
    seeds = A.seeds_uniform(0, 0, max(count, 1))[:count]
    rb = B.project(seeds, nthreads=8) if count else None
    if layout == "aos":
        r = c.projectBatch(torch.from_numpy(seeds).cuda().reshape(count, 14))
        x = r.x.cpu().numpy()
    else:
        r = c.projectBatch(torch.from_numpy(np.ascontiguousarray(seeds.T)).cuda().reshape(14, count), layout=pkg.CCP_LAYOUT_SOA)
        x = r.x.cpu().numpy().T
    torch.cuda.synchronize()
    assert r.ok.shape[0] == count
    if count:
        assert np.array_equal(x.view(np.uint64), rb["x"].view(np.uint64))
        assert np.array_equal(r.iters.cpu().numpy(), rb["iters"]) and np.array_equal(r.ok.cpu().numpy(), rb["ok"])


def test_misaligned_seed_pointer_and_in_place(setup):
    """Seeds starting 8 bytes off a 16-byte boundary cannot be bulk-copied: the kernel must take the plain-load path,
    with identical results; out may alias the seeds."""
    pkg, c, A, B = setup
    n = 70_000
    seeds = A.seeds_uniform(0, 0, n)
    ref = c.projectBatch(torch.from_numpy(seeds).cuda())
    buf = torch.empty(n * 14 + 1, dtype=torch.float64, device="cuda")
    view = buf[1:].view(n, 14)
    assert view.data_ptr() % 16 == 8
    view.copy_(torch.from_numpy(seeds))
    r = c.projectBatch(view, out=view)  # misaligned AND in place
    torch.cuda.synchronize()
    assert np.array_equal(_bits(r.x), _bits(ref.x)) and torch.equal(r.iters, ref.iters)


def test_nan_and_far_seeds_terminate(setup):
    pkg, c, A, B = setup
    seeds = A.seeds_uniform(0, 0, 4096)
    seeds[5, 3] = np.nan
    seeds[77, :] = np.inf
    seeds[100, 0] = 1e9  # a joint angle far outside any sane range
    r = c.projectBatch(torch.from_numpy(seeds).cuda())
    torch.cuda.synchronize()
    ok = r.ok.cpu().numpy()
    assert ok[5] == 0 and ok[77] == 0
    rb = B.project(seeds, nthreads=8)
    good = np.ones(4096, bool)
    good[[5, 77]] = False
    assert np.array_equal(r.x.cpu().numpy()[good].view(np.uint64), rb["x"][good].view(np.uint64))
    assert np.array_equal(ok, rb["ok"])


@pytest.mark.parametrize("layout", ["aos", "soa"])
def test_chunked_sampler_equals_explicit_seeds(setup, layout):
    """ccp_sample_project_batch cuts batches above 2 M seeds into pipelined launches over a scratch seed buffer: the
    per-seed outputs must equal projecting the explicitly generated seeds, also in SoA (strided chunk outputs)."""
    pkg, c, A, B = setup
    from closed_chain_motion_planner_b200 import _capi

    n = 2 * 1024 * 1024 + 70_001
    lay = pkg.CCP_LAYOUT_AOS if layout == "aos" else pkg.CCP_LAYOUT_SOA
    shape = (n, 14) if layout == "aos" else (14, n)
    st = torch.cuda.current_stream().cuda_stream
    a = _capi.SamplerArgs(rng_seed=5, first_index=123, mode=0, wrap_bounds=0, distance=0.0, near_host=None)
    seeds = torch.empty(shape, dtype=torch.float64, device="cuda")
    assert c._lib.ccp_generate_seeds(c._h, C.byref(a), n, lay, seeds.data_ptr(), st) == 0
    ref = c.projectBatch(seeds, layout=lay, want_resid=False)
    x = torch.empty(shape, dtype=torch.float64, device="cuda")
    ok = torch.empty(n, dtype=torch.uint8, device="cuda")
    it = torch.empty(n, dtype=torch.int32, device="cuda")
    n_ok = torch.zeros(1, dtype=torch.int64, device="cuda")
    assert c._lib.ccp_sample_project_batch(c._h, C.byref(a), n, lay, x.data_ptr(), ok.data_ptr(), it.data_ptr(), None,
                                           n_ok.data_ptr(), st) == 0
    torch.cuda.synchronize()
    assert not c.pipelineOpen()
    assert torch.equal(ok, ref.ok) and torch.equal(it, ref.iters) and np.array_equal(_bits(x), _bits(ref.x))
    assert int(n_ok) == int(ref.ok.sum())


def test_opt_in_damping_and_clamping(setup):
    """The north star's damped-least-squares step and joint-limit clamping are opt-in modes the reference does not
    have (parity keeps them off): bit-exact against the host build with the same options, every clamped result
    inside [lb, ub], and switching them off restores the reference behaviour."""
    import closed_chain_motion_planner_b200 as pkg

    _, c0, A, B = setup
    c = pkg.KinematicChainConstraint.from_config("Wine_Bottle", device=0)
    seeds = A.seeds_uniform(9, 0, 20_000)
    base = c.projectBatch(torch.from_numpy(seeds).cuda())
    lb = np.tile(c.lb_, 2)
    ub = np.tile(c.ub_, 2)
    try:
        c.setOptions(damping=1e-4, clamp=True)
        B.set_options(damping=1e-4, clamp=True)
        r = c.projectBatch(torch.from_numpy(seeds).cuda())
        rb = B.project(seeds, nthreads=8)
        assert np.array_equal(_bits(r.x), rb["x"].view(np.uint64))
        assert np.array_equal(r.iters.cpu().numpy(), rb["iters"]) and np.array_equal(r.ok.cpu().numpy(), rb["ok"])
        x = r.x.cpu().numpy()
        assert np.all(x >= lb) and np.all(x <= ub)  # clamped every step
        # clamping keeps iterates inside the limits, so far more converged states are also jointValid
        assert int(r.ok.sum()) > 1.2 * int(base.ok.sum())  # measured +38 %: states parked ON a limit still fail the 1e-3 margin
        ok = r.ok.cpu().numpy().astype(bool)
        f = A.function(x[ok])
        assert np.all(f[:, 0] <= 1e-3 * (1 + 1e-9)) and np.all(f[:, 1] < 5e-3 * (1 + 1e-9))
        with pytest.raises(Exception):
            c.setOptions(damping=-1.0)
    finally:
        B.set_options()
    c.setOptions()
    again = c.projectBatch(torch.from_numpy(seeds).cuda())
    assert np.array_equal(_bits(again.x), _bits(base.x)) and torch.equal(again.ok, base.ok)


def test_round2_entry_points_reject_bad_arguments(setup):
    """Argument checking of the entry points added in round 2 (error codes, no crash, handle usable afterwards)."""
    import ctypes as C

    import closed_chain_motion_planner_b200 as pkg
    from closed_chain_motion_planner_b200 import _capi

    _, c, A, B = setup
    lib, h = c._lib, c._h
    INVALID, STATE = -1, -3
    x = np.ascontiguousarray(A.seeds_uniform(0, 0, 64))
    xo = np.zeros_like(x)
    t = C.c_int64(0)
    nk = C.c_int64(0)
    # ccp_host_batch_submit: exactly one source of seeds, a capacity with compact outputs, a positive count
    b = _capi.HostBatch()
    b.count = 64
    assert lib.ccp_host_batch_submit(h, C.byref(b), C.byref(t)) == INVALID  # neither seeds nor sampler
    sa = _capi.SamplerArgs(rng_seed=0, first_index=0, mode=0, wrap_bounds=0, distance=0.0, near_host=None)
    b.seeds_host = x.ctypes.data
    b.sampler = C.pointer(sa)
    assert lib.ccp_host_batch_submit(h, C.byref(b), C.byref(t)) == INVALID  # both
    b.sampler = None
    b.compact_host = xo.ctypes.data
    assert lib.ccp_host_batch_submit(h, C.byref(b), C.byref(t)) == INVALID  # compact output without a capacity
    b.compact_capacity = 64
    b.count = 0
    assert lib.ccp_host_batch_submit(h, C.byref(b), C.byref(t)) == INVALID
    assert lib.ccp_host_batch_submit(h, None, C.byref(t)) == INVALID and lib.ccp_host_batch_submit(h, C.byref(b), None) == INVALID
    bad_mode = _capi.SamplerArgs(rng_seed=0, first_index=0, mode=7, wrap_bounds=0, distance=0.0, near_host=None)
    b2 = _capi.HostBatch()
    b2.count = 64
    b2.sampler = C.pointer(bad_mode)
    b2.compact_host = xo.ctypes.data
    b2.compact_capacity = 64
    assert lib.ccp_host_batch_submit(h, C.byref(b2), C.byref(t)) == INVALID  # sampler mode out of range
    assert lib.ccp_host_batch_wait(h, 123456789, C.byref(nk)) == INVALID  # a ticket never issued
    # a good batch still goes through, and waiting twice for it is harmless
    b.count = 64
    assert lib.ccp_host_batch_submit(h, C.byref(b), C.byref(t)) == 0
    assert lib.ccp_host_batch_wait(h, t.value, C.byref(nk)) == 0 and 0 <= nk.value <= 64
    assert lib.ccp_host_batch_wait(h, t.value, C.byref(nk)) == 0
    rb = B.project(x)
    assert nk.value == int(rb["ok"].sum())
    # multicast needs peers first; peers need aligned pools
    assert lib.ccp_set_gather_multicast(h, 0x1000) == STATE
    pools = (C.c_uint64 * 1)(0x1008)
    assert lib.ccp_set_gather_peers(h, 1, 0, pools, 16) == INVALID  # not 16-byte aligned
    assert lib.ccp_set_gather_peers(h, 0, 0, None, 0) == 0
    # peer groups and the NCCL gather
    g = C.c_void_p()
    assert lib.ccp_peer_group_create(None, 1, 16, C.byref(g)) == INVALID
    hs = (C.c_void_p * 1)(h)
    assert lib.ccp_peer_group_create(hs, 0, 16, C.byref(g)) == INVALID and lib.ccp_peer_group_create(hs, 9, 16, C.byref(g)) == INVALID
    assert lib.ccp_peer_group_create(hs, 1, 0, C.byref(g)) == INVALID and b"capacity" in lib.ccp_peer_group_last_error(None)
    assert lib.ccp_peer_group_create(hs, 1, 4096, C.byref(g)) == 0 and lib.ccp_peer_group_world(g) == 1
    cnt = (C.c_int64 * 1)()
    assert lib.ccp_peer_group_sample_project(g, None, 100, cnt) == INVALID
    assert lib.ccp_peer_group_sample_project(g, C.byref(sa), 2000, cnt) == 0 and 0 < cnt[0] < 2000  # a group of one works
    rows = np.zeros((4096, 14))
    got = C.c_int64(0)
    assert lib.ccp_peer_group_gather_host(g, 0, rows.ctypes.data, 4096, cnt, C.byref(got)) == 0 and got.value == cnt[0]
    assert lib.ccp_peer_group_gather_host(g, 0, rows.ctypes.data, 3, cnt, C.byref(got)) == INVALID  # host buffer too small
    assert lib.ccp_peer_group_gather_host(g, 5, rows.ctypes.data, 4096, cnt, C.byref(got)) == INVALID
    f = A.function(rows[:got.value])
    assert np.all(f[:, 0] <= 1e-3 * (1 + 1e-9)) and np.all(f[:, 1] < 5e-3)
    lib.ccp_peer_group_destroy(g)
    lib.ccp_peer_group_destroy(None)
    assert lib.ccp_allgather_converged(h, None, 2, None, None, 16, None, None, None) == INVALID
    assert lib.ccp_set_coop_threshold(None, 5) == INVALID and lib.ccp_set_coop_threshold(h, -1) == 0
    assert lib.ccp_device_count() >= 1
    it, tl = C.c_double(), C.c_double()
    assert lib.ccp_algorithmic_flops_ik(C.byref(it), C.byref(tl)) == 0 and it.value == 1070.0 and tl.value == 540.0
    # the handle is still fine
    r = c.projectBatch(x)
    assert np.array_equal(r.x.view(np.uint64), rb["x"].view(np.uint64))


def test_host_entry_points_from_several_threads(setup):
    """ccp.h: host-pointer calls on one handle may come from several threads (serialised inside, as the reference's
    graphMutex_ serialises its callers); handles are independent of one another.  ctypes drops the GIL during the calls."""
    import threading

    import closed_chain_motion_planner_b200 as pkg

    _, c, A, B = setup
    c2 = pkg.KinematicChainConstraint.from_config("Wine_Bottle", device=0)
    sets = [A.seeds_uniform(20 + t, 0, n) for t, n in enumerate((1, 300, 5_000, 40_000, 260_000, 7))]
    want = [B.project(s, nthreads=4) for s in sets]
    errors = []

    def worker(tid, handle):
        try:
            for rep in range(3):
                for k in range(len(sets)):
                    s = sets[(k + tid) % len(sets)]
                    w = want[(k + tid) % len(sets)]
                    r = handle.projectBatch(s)
                    if not (np.array_equal(r.x.view(np.uint64), w["x"].view(np.uint64)) and np.array_equal(r.ok, w["ok"])):
                        errors.append((tid, rep, k, "projectBatch"))
                    f = handle.functionBatch(s[:50])
                    if not np.array_equal(f.view(np.uint64), B.function(s[:50]).view(np.uint64)):
                        errors.append((tid, rep, k, "functionBatch"))
        except Exception as e:  # noqa: BLE001
            errors.append((tid, repr(e)))

    threads = [threading.Thread(target=worker, args=(t, c if t < 3 else c2)) for t in range(5)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:5]
