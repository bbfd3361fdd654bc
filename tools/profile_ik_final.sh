#!/bin/bash
# ncu --set full of the final IK kernels and of the three-arm cooperative kernel (one launch each, after a plain run)
set -u
mkdir -p gpurun_out
python tools/profile_aux.py > gpurun_out/r02_aux_final_plain.json 2> gpurun_out/r02_aux_final_plain.err || exit 1
for k in ccp_ik_kernel ccp_ik_sample_kernel; do
  ncu --set full --clock-control none -k regex:"^$k" --launch-skip 1 -c 1 -o gpurun_out/r02_auxf_$k -f python tools/profile_aux.py > gpurun_out/r02_auxf_$k.log 2>&1
done
COOP_PROBE_COUNTS=1000 python tools/coop_probe.py stefan_three_arm > gpurun_out/r02_coop3_plain.log 2>&1
COOP_PROBE_COUNTS=1000 ncu --set full --clock-control none -k regex:'^ccp_project_coop3_kernel' --launch-skip 3 -c 1 -o gpurun_out/r02_coop3 -f python tools/coop_probe.py stefan_three_arm > gpurun_out/r02_coop3.log 2>&1
ls -la gpurun_out/r02_auxf_* gpurun_out/r02_coop3.ncu-rep
