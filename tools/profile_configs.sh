#!/bin/bash
# ncu --set full of the projection kernel on the other two BASELINE configurations (one launch each, after an unprofiled
# run of the same command); writes gpurun_out/r02_k3.ncu-rep and gpurun_out/r02_wine.ncu-rep
set -u
mkdir -p gpurun_out
python tools/sweep.py stefan_three_arm 1000000 > gpurun_out/r02_k3_plain.json || exit 1
python tools/sweep.py Wine_Bottle 4000000 > gpurun_out/r02_wine_plain.json || exit 1
ncu --set full --clock-control none -k regex:'^ccp_project_kernel' --launch-skip 1 -c 1 -o gpurun_out/r02_k3 -f python tools/sweep.py stefan_three_arm 1000000 > gpurun_out/r02_k3.log 2>&1
ncu --set full --clock-control none -k regex:'^ccp_project_kernel' --launch-skip 1 -c 1 -o gpurun_out/r02_wine -f python tools/sweep.py Wine_Bottle 4000000 > gpurun_out/r02_wine.log 2>&1
ls -la gpurun_out/r02_k3.ncu-rep gpurun_out/r02_wine.ncu-rep
