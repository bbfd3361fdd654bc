#!/bin/bash
# ncu evidence of round 2 (run on the GPU box from the repo root; writes gpurun_out/).  Every profiled command is first
# run plain (exit 0) and no number printed under ncu is used as a bench value.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-configs"
$B > gpurun_out/r02_plain.json 2> gpurun_out/r02_plain.err || exit 1
# 1. launch list of the bench command (share of the step per kernel)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_launches.log 2>&1
# 2. full capture of two consecutive pipelined projection launches
ncu --set full --clock-control none --import-source on -k regex:'^ccp_project_kernel' --launch-skip 6 -c 2 -o gpurun_out/r02_project -f $B --no-e2e > gpurun_out/r02_project.log 2>&1
# 3. the cooperative kernel on 1000 seeds (latency regime)
python tools/coop_probe.py > gpurun_out/r02_coop_probe.log 2>&1
ncu --set full --clock-control none -k regex:'^ccp_project_coop_kernel' --launch-skip 12 -c 1 -o gpurun_out/r02_coop -f python tools/coop_probe.py > gpurun_out/r02_coop.log 2>&1
# 4. the lane-refill IK kernels, the geodesic kernel, the coalesced seed kernel
python tools/profile_aux.py > gpurun_out/r02_aux_plain.json 2> gpurun_out/r02_aux_plain.err
for k in ccp_geodesic_kernel ccp_ik_kernel ccp_ik_sample_kernel; do
  ncu --set full --clock-control none -k regex:"^$k" --launch-skip 1 -c 1 -o gpurun_out/r02_aux_$k -f python tools/profile_aux.py > gpurun_out/r02_aux_$k.log 2>&1
done
ncu --set full --clock-control none -k regex:'^ccp_seed_kernel' -c 1 -o gpurun_out/r02_aux_ccp_seed_kernel -f python tools/profile_aux.py > gpurun_out/r02_aux_seed.log 2>&1
ls -la gpurun_out/r02_*
