"""Multi-GPU path on real devices: sharded sample->project->compact and the NCCL gather of converged states.
Needs >= 2 GPUs (skipped on a single-GPU box; the host logic is covered on CPU by tests/test_dist_cpu.py)."""
import socket
import subprocess
import sys
import textwrap

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

WORKER = textwrap.dedent(
    """
    import os, sys
    sys.path.insert(0, %(root)r)
    import numpy as np, torch, torch.distributed as dist
    import closed_chain_motion_planner_b200 as pkg
    from closed_chain_motion_planner_b200.dist import ShardedSampleProjector, shard_range
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    c = pkg.KinematicChainConstraint.from_config("stefan", device=local)
    total = 40001
    # fused path (the kernel stores into every rank's symmetric pool over NVLink) against the NCCL gather
    sp = ShardedSampleProjector(c)
    states, counts = sp.sample_project(rng_seed=3, first_index=1000, total=total)
    torch.cuda.synchronize()
    assert sp.fused, "fused peer-store gather unavailable: " + str(sp.fused_error)
    sn = ShardedSampleProjector(c, fused=False)
    states_n, counts_n = sn.sample_project(rng_seed=3, first_index=1000, total=total)
    torch.cuda.synchronize()
    assert torch.equal(counts, counts_n)
    off = 0
    for rr in range(world):  # same rows per rank region (order within a region is unspecified)
        k = int(counts[rr])
        a, b = states[off:off + k].cpu().numpy(), states_n[off:off + k].cpu().numpy()
        srt = lambda m: m[np.lexsort(m.T[::-1])]
        assert np.array_equal(srt(a), srt(b))
        off += k
    states2, counts2 = sp.sample_project(rng_seed=4, first_index=0, total=total // 2)  # pool reuse
    torch.cuda.synchronize()
    assert int(counts2.sum()) == states2.shape[0]
    assert counts.shape == (world,) and int(counts.sum()) == states.shape[0]
    # every rank holds the same gathered pool
    chk = states.sum(dim=0).clone()
    ref = chk.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(chk, ref)
    # the pool equals a single-GPU projection of the whole stream slice (as a set of rows)
    if rank == 0:
        from closed_chain_motion_planner_b200 import _capi
        import ctypes as C
        seeds = torch.empty((total, 14), dtype=torch.float64, device="cuda")
        a = _capi.SamplerArgs(rng_seed=3, first_index=1000, mode=0, wrap_bounds=0, distance=0.0, near_host=None)
        st = torch.cuda.current_stream().cuda_stream
        assert c._lib.ccp_generate_seeds(c._h, C.byref(a), total, 0, seeds.data_ptr(), st) == 0
        r = c.projectBatch(seeds)
        want = r.x[r.ok.bool()].cpu().numpy()
        got = states.cpu().numpy()
        key = lambda m: m[np.lexsort(m.T[::-1])]
        assert got.shape == want.shape and np.array_equal(key(got), key(want))
        # per-rank counts match the shard ranges
        okc = r.ok.cpu().numpy()
        for rr in range(world):
            f, n = shard_range(total, rr, world)
            assert int(counts[rr]) == int(okc[f:f + n].sum())
    dist.barrier()
    dist.destroy_process_group()
    print("rank" + str(rank) + "-ok", flush=True)
    """
)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_sample_project_and_gather(tmp_path):
    n = min(torch.cuda.device_count(), 4)
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    for k in range(n):
        assert f"rank{k}-ok" in r.stdout


def test_sharded_projector_world1():
    """With no process group the sharded projector degrades to the local kernel call."""
    import closed_chain_motion_planner_b200 as pkg
    from closed_chain_motion_planner_b200.dist import ShardedSampleProjector

    c = pkg.KinematicChainConstraint.from_config("stefan", device=0)
    states, counts = ShardedSampleProjector(c).sample_project(rng_seed=3, first_index=0, total=5000)
    assert states.shape[0] == int(counts[0]) and 0.15 * 5000 < states.shape[0] < 0.35 * 5000
    assert bool(c.isSatisfiedBatch(states.contiguous()).all()) and bool(c.jointValidBatch(states.contiguous()).all())
