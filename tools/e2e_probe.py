"""Diagnostic (not part of the product): PCIe copy rates on the box and the host-path pipeline vs chunk count."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import closed_chain_motion_planner_b200 as pkg
from closed_chain_motion_planner_b200 import _capi

n = 1_000_000
c = pkg.KinematicChainConstraint.from_config("dumbbell")
hs = torch.empty((n, 14), dtype=torch.float64).pin_memory()
hx = torch.empty((n, 14), dtype=torch.float64).pin_memory()
d = torch.empty((n, 14), dtype=torch.float64, device="cuda")
a = _capi.SamplerArgs(rng_seed=0, first_index=0, mode=0, wrap_bounds=0, distance=0.0, near_host=None)
c._lib.ccp_generate_seeds(c._h, C.byref(a), n, 0, d.data_ptr(), torch.cuda.current_stream().cuda_stream)
hs.copy_(d)
torch.cuda.synchronize()
for name, fn in (("H2D", lambda: d.copy_(hs, non_blocking=True)), ("D2H", lambda: hx.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"{name} 112 MB pinned: {dt*1e3:.2f} ms  {0.112/dt:.1f} GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1):
        d.copy_(hs, non_blocking=True)
    with torch.cuda.stream(s2):
        hx.copy_(d, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print(f"H2D+D2H concurrent: {dt*1e3:.2f} ms")
hok = torch.empty(n, dtype=torch.uint8).pin_memory()
hit = torch.empty(n, dtype=torch.int32).pin_memory()

# --- do chunk kernels on different streams overlap their tails? (data resident, no copies) ---
x_out = torch.empty_like(d)
ok = torch.empty(n, dtype=torch.uint8, device="cuda")
it = torch.empty(n, dtype=torch.int32, device="cuda")
lib, h = c._lib, c._h


def run_chunks(parts, nstreams):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    chunk = (n + parts - 1) // parts
    for i in range(parts):
        off = i * chunk
        cnt = min(chunk, n - off)
        st = streams[i % nstreams]
        rc = lib.ccp_project_batch(h, d.data_ptr() + off * 112, cnt, 0, x_out.data_ptr() + off * 112, ok.data_ptr() + off, None,
                                   it.data_ptr() + off * 4, None, None, None, st.cuda_stream)
        assert rc == 0
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3


for parts, ns in ((1, 1), (4, 1), (4, 4), (8, 1), (8, 4), (16, 4)):
    run_chunks(parts, ns)
    print(f"resident data, {parts} chunk kernels on {ns} stream(s): {min(run_chunks(parts, ns) for _ in range(3)):.2f} ms")

for parts in (1, 2, 4, 6, 8):
    os.environ["CCP_HOST_CHUNKS"] = str(parts)
