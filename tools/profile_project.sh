#!/bin/bash
# ncu evidence of the shipped projection kernel alone (see tools/profile_round2.sh for the rest); writes gpurun_out/
set -u
mkdir -p gpurun_out
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-configs"
$B > gpurun_out/r02_plain.json 2> gpurun_out/r02_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'^ccp_project_kernel' --launch-skip 6 -c 2 -o gpurun_out/r02_project -f $B --no-e2e > gpurun_out/r02_project.log 2>&1
ls -la gpurun_out/r02_project.ncu-rep gpurun_out/r02_launches.csv
