"""Host logic of the multi-GPU path on CPU: shard ranges and the converged-pool gather, world_size 2, gloo."""
import os
import socket
import subprocess
import sys
import textwrap

import pytest

from conftest import ROOT

from closed_chain_motion_planner_b200.dist import gather_capacity, shard_range


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 10, 1_000_000, 10_000_003):
        for world in (1, 2, 3, 4, 8):
            nxt = 0
            for r in range(world):
                first, cnt = shard_range(total, r, world)
                assert first == nxt and cnt >= 0
                nxt = first + cnt
            assert nxt == total
            sizes = [shard_range(total, r, world)[1] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_gather_capacity_bounds():
    assert gather_capacity(1_000_000) == 400_064
    assert gather_capacity(10) == 10 and gather_capacity(0) == 1


WORKER = textwrap.dedent(
    """
    import os, sys
    sys.path.insert(0, %(root)r)
    import torch, torch.distributed as dist
    from closed_chain_motion_planner_b200.dist import gather_converged, unpack_pool, shard_range
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n, cap = 14, 16
    # rank r "converged" (r+1)*3 states whose entries encode (rank, row)
    k = (rank + 1) * 3
    compact = torch.full((32, n), -1.0, dtype=torch.float64)
    for i in range(k):
        compact[i] = 100.0 * rank + i
    n_ok = torch.tensor([k], dtype=torch.int64)
    pool, counts = gather_converged(compact, n_ok, cap)
    assert counts.tolist() == [(r + 1) * 3 for r in range(world)], counts
    states = unpack_pool(pool, counts)
    assert states.shape == (sum((r + 1) * 3 for r in range(world)), n)
    row = 0
    for r in range(world):
        for i in range((r + 1) * 3):
            assert float(states[row, 0]) == 100.0 * r + i and float(states[row, n - 1]) == 100.0 * r + i
            row += 1
    # overflow is detected, never silent
    try:
        unpack_pool(pool, counts * 10)
        raise SystemExit("overflow not detected")
    except OverflowError:
        pass
    # shard ranges of all ranks tile the stream
    spans = [shard_range(1001, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][0] + spans[-1][1] == 1001
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


def test_gather_converged_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "rank0-ok" in r.stdout and "rank1-ok" in r.stdout


def test_bench_reference_arm_runs_on_cpu():
    """--impl reference never touches CUDA: it times the CPU restatement and prints the contract's line."""
    import json

    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-sample", "200"], capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["unit"] == "converged projections/s"
