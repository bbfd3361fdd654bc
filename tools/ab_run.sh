#!/bin/bash
# A/B of the projection kernel on one box: libccp_old.so (a build of the previous commit) against the working tree,
# over the epilogue queue's batch size (CCP_FIN_BATCH).
J='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["ms_per_step"],4), round(d["roofline"]["kernel_ms_per_launch"],4), round(d["roofline"]["frac"],4))'
J2='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["ms"],3), round(d["frac_of_measured_fp64_peak"],4))'
OLD=$PWD/closed_chain_motion_planner_b200/csrc/libccp_old.so
echo "K2 old: $(CCP_LIB=$OLD python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-configs | python -c "$J")"
for v in ${FINS:-1 2 4 6 8}; do
  echo "K2 fin_batch $v: $(CCP_FIN_BATCH=$v python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-configs | python -c "$J")"
done
echo "K2 old: $(CCP_LIB=$OLD python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-configs | python -c "$J")"
echo "K3 old: $(CCP_LIB=$OLD python tools/sweep.py stefan_three_arm 1000000 | python -c "$J2")"
for v in ${FINS:-1 2 4 6 8}; do
echo "K3 fin_batch $v: $(CCP_FIN_BATCH=$v python tools/sweep.py stefan_three_arm 1000000 | python -c "$J2")"
done
echo "wine old: $(CCP_LIB=$OLD python tools/sweep.py Wine_Bottle 4000000 | python -c "$J2")"
for v in ${FINS:-1 2 4 6 8}; do
echo "wine fin_batch $v: $(CCP_FIN_BATCH=$v python tools/sweep.py Wine_Bottle 4000000 | python -c "$J2")"
done
