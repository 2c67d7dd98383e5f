#!/bin/bash
# ncu --set full + SASS-level counts of the shade kernels of the first wavefront iteration (scene 6 FHD, 4.15 M camera rays)
set -u
TAG=${1:-r2_x}; O=gpurun_out; mkdir -p $O
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_shade --launch-skip 0 --launch-count 12 -f \
  -o /tmp/prof_shade_$TAG python tools/perf_probe.py 6:1920:2 > $O/ncu_capture_shade.log 2>&1
ncu -i /tmp/prof_shade_$TAG.ncu-rep --page source --csv 2>/dev/null | gzip > $O/${TAG}_src_k_shade.csv.gz
ls -la $O/${TAG}_src_k_shade.csv.gz
