#!/bin/bash
# 2-GPU evidence of the fused gather's peer traffic: ONE process driving two devices (tests/cpp/test_multi_gpu.cpp: the C
# ABI's peer group, 16-byte stores per peer) under ncu with the L2 peer-aperture and NVLink counters.
# Run on a box with >= 2 GPUs from the repo root; writes gpurun_out/.
set -u
mkdir -p gpurun_out
L=closed_chain_motion_planner_b200/csrc
g++ -std=c++17 -O1 -I include tests/cpp/test_multi_gpu.cpp -o /tmp/test_multi_gpu -L $L -lccp -Wl,-rpath,$PWD/$L || exit 1
python - <<'PY'
import sys
sys.path.insert(0, '.')
from closed_chain_motion_planner_b200 import grasping_point
open('/tmp/start.bin', 'wb').write(grasping_point().loadConfig('dumbbell').start.tobytes())
PY
/tmp/test_multi_gpu /tmp/start.bin /tmp/out.bin 2 || exit 1
ncu --query-metrics 2>/dev/null | grep -i -E "nvl(rx|tx)__bytes|aperture_peer" | awk '{print $1}' | sort -u > gpurun_out/r02_peer_metric_names.txt
M=$(grep -E "^(nvltx__bytes|nvlrx__bytes|lts__t_sectors_aperture_peer|lts__t_sectors_srcunit_tex_aperture_peer|lts__t_sectors_aperture_peer_op_write|lts__t_bytes_aperture_peer)$" gpurun_out/r02_peer_metric_names.txt | sed 's/$/.sum/' | paste -sd, -)
echo "metrics: $M"
ncu --metrics gpu__time_duration.sum,$M --clock-control none -k regex:'^ccp_project_kernel' -c 8 --csv --log-file gpurun_out/r02_peer_traffic.csv /tmp/test_multi_gpu /tmp/start.bin /tmp/out.bin 2 > gpurun_out/r02_peer_traffic.log 2>&1
tail -5 gpurun_out/r02_peer_traffic.log
head -c 3000 gpurun_out/r02_peer_traffic.csv
