#!/bin/bash
# Launch-configuration variants of the projection kernel (CCP_PROJ_VARIANT, compiled only with -DCCP_TUNE=1) against the
# shipped configuration.  Build the tuning library first (it is not part of the product):
#   cd closed_chain_motion_planner_b200/csrc && mkdir -p /tmp/tune && for kp in "2 1" "3 1"; do set -- $kp; \
#     nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo --fmad=false -Xcompiler -fPIC -I../../include -I. \
#       -diag-suppress 177 -DCCP_TUNE=1 -DCCP_TU_K=$1 -DCCP_TU_PANDA=$2 -c -o /tmp/tune/ccp_project_k$1p$2.o ccp_project.cu; done
#   nvcc -shared -gencode arch=compute_100a,code=sm_100a -o libccp_tune.so build/ccp_api.o build/ccp_multi.o build/ccp_coop.o \
#     build/ccp_geodesic.o build/ccp_ik.o build/ccp_project_k2p0.o /tmp/tune/ccp_project_k2p1.o build/ccp_project_k3p0.o \
#     /tmp/tune/ccp_project_k3p1.o -ldl
# Round-2 result on B200 (ms per pipelined 1 M-seed dumbbell step / fraction of the measured FP64 peak):
#   0 shipped 384 x 168 regs           3.291  0.820      K = 3, 1 M seeds:  0 shipped 256 x 255 regs   18.5 ms  0.765
#   10 384, lever arms in smem         3.406  0.792                         7 320 thr, lever arms smem 32.4     0.436
#   9  416 x 128 regs, lever arms smem 4.112  0.655                         8 384 thr, lever arms smem 25.6     0.553
#   8  448 x 128 regs, lever arms smem 4.318  0.623                         9 288 thr, lever arms smem 31.5     0.449
#   7  448, no seed staging            3.776  0.713                         5 288 thr                  26.6     0.532
#   6  512 x 128 regs, no seed staging 3.658  0.737                         6 320 thr                  25.7     0.549
# More warps per scheduler need <= 128 registers, which spills ~250 B per thread (K = 2) / ~600 B (K = 3): slower every time.
for v in 0 10 9 8 7 6; do
  echo "K2 variant $v: $(CCP_LIB=$PWD/closed_chain_motion_planner_b200/csrc/libccp_tune.so CCP_PROJ_VARIANT=$v python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-configs | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), round(d['roofline']['kernel_ms_per_launch'],4), round(d['roofline']['frac'],4))")"
done
for v in 0 7 8 9 5 6; do
  echo "K3 variant $v: $(CCP_LIB=$PWD/closed_chain_motion_planner_b200/csrc/libccp_tune.so CCP_PROJ_VARIANT=$v python tools/sweep.py stefan_three_arm 1000000 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms'],3), round(d['frac_of_measured_fp64_peak'],4))")"
done
