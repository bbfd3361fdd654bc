"""Multi-GPU entry points of the C ABI driven from compiled C++ programs in ONE process (the reference planner is one
C++ process): the peer group with the gather fused into the projection kernels (tests/cpp/test_multi_gpu.cpp, no CUDA
header), and the NCCL fallback on communicators the host owns (tests/cpp/test_multi_nccl.cu).  Both need >= 2 GPUs and
skip on a single-GPU box; the gathered pools are compared with a single-GPU projection of the same stream slices."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_cfg

pytestmark = pytest.mark.gpu

LIB_DIR = os.path.join(ROOT, "closed_chain_motion_planner_b200", "csrc")


def _rows_sorted(m):
    return m[np.lexsort(m.T[::-1])]


def _run(exe, tmp_path, world=None):
    cfg = load_cfg("dumbbell")
    (tmp_path / "start.bin").write_bytes(cfg.start.tobytes())
    args = [str(exe), str(tmp_path / "start.bin"), str(tmp_path / "out.bin")] + ([str(world)] if world else [])
    r = subprocess.run(args, capture_output=True, text=True)
    if r.returncode == 77:
        pytest.skip(r.stdout.strip())
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    return open(tmp_path / "out.bin", "rb").read()


def test_peer_group_from_cpp(tmp_path):
    exe = tmp_path / "test_multi_gpu"
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_multi_gpu.cpp"), "-o", str(exe),
                           "-L", LIB_DIR, "-lccp", f"-Wl,-rpath,{LIB_DIR}"])
    raw = _run(exe, tmp_path)
    hdr = np.frombuffer(raw, np.int64, 9)
    world, got1, got2, n1, n2, same_pool, rc_dup, rc_overflow, single_ok = (int(v) for v in hdr)
    off = 72
    counts1 = np.frombuffer(raw, np.int64, world, off)
    counts2 = np.frombuffer(raw, np.int64, world, off + 8 * world)
    off += 16 * world
    rows = np.frombuffer(raw, np.float64, (got1 + got2) * 14, off).reshape(-1, 14)
    off += rows.nbytes
    single = np.frombuffer(raw, np.float64, (n1 + n2) * 14, off).reshape(-1, 14)
    assert world >= 2 and same_pool == 1 and single_ok == 1
    assert rc_dup == -1 and rc_overflow == -1  # CCP_ERR_INVALID: duplicate device / capacity overflow reported
    assert counts1.sum() == got1 == n1 and counts2.sum() == got2 == n2
    # the gathered pools hold exactly the states one GPU finds in the same slices of the stream (order unspecified)
    assert np.array_equal(_rows_sorted(rows[:got1]).view(np.uint64), _rows_sorted(single[:n1]).view(np.uint64))
    assert np.array_equal(_rows_sorted(rows[got1:]).view(np.uint64), _rows_sorted(single[n1:]).view(np.uint64))
    # wrapped into [-pi, pi) (the sampler's enforceBounds)
    assert rows.min() >= -np.pi and rows.max() < np.pi


@pytest.mark.skipif(shutil.which("nvcc") is None or not os.path.exists("/usr/include/nccl.h"), reason="needs nvcc and nccl.h")
def test_nccl_allgather_from_cpp(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    exe = tmp_path / "test_multi_nccl"
    subprocess.check_call(["nvcc", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_multi_nccl.cu"), "-o", str(exe),
                           "-L", LIB_DIR, "-lccp", "-lnccl", "-Xlinker", f"-rpath={LIB_DIR}"])
    raw = _run(exe, tmp_path)
    world, same, cap = (int(v) for v in np.frombuffer(raw, np.int64, 3))
    counts = np.frombuffer(raw, np.int64, world, 24)
    rows = np.frombuffer(raw, np.float64, int(counts.sum()) * 14, 24 + 8 * world).reshape(-1, 14)
    assert world >= 2 and same == 1 and counts.max() <= cap
    # against this process's own single-GPU projection of the same counter ranges
    import ctypes as C

    import closed_chain_motion_planner_b200 as pkg
    from closed_chain_motion_planner_b200 import _capi

    c = pkg.KinematicChainConstraint.from_config("dumbbell", device=0)
    off = 0
    for d in range(world):
        a = _capi.SamplerArgs(rng_seed=33, first_index=d * 50000, mode=0, wrap_bounds=0, distance=0.0, near_host=None)
        comp = np.zeros((50000, 14))
        nk = C.c_int64(0)
        assert c._lib.ccp_sample_project_batch_host(c._h, C.byref(a), 50000, None, None, None, comp.ctypes.data, C.byref(nk)) == 0
        assert nk.value == counts[d]
        assert np.array_equal(_rows_sorted(rows[off:off + nk.value]).view(np.uint64), _rows_sorted(comp[:nk.value]).view(np.uint64))
        off += nk.value
