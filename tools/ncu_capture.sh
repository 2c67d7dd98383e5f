#!/bin/bash
# ncu --set full of the first two wavefront iterations of a scene-6 FHD render (GPU box).  Iteration 1 = 4.15 M camera
# rays (k_top<PRIMARY>), iteration 2 = the 2.54 M survivors (k_top on old paths): both launch sizes are known exactly, so
# per-ray figures can be derived.  Reports stay in /tmp (too large for gpurun_out); what comes back are the raw-page CSVs
# (all metrics per launch) and gzipped source-page CSVs of the hot kernels.
#   tools/ncu_capture.sh TAG
set -u
TAG=${1:-r2_x}; O=gpurun_out; mkdir -p $O
python tools/perf_probe.py 6:1920:2 > /dev/null 2>&1 || { echo "probe failed"; exit 1; }
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_top|k_mesh|k_shade" --launch-skip 0 --launch-count 20 -f \
  -o /tmp/prof_$TAG python tools/perf_probe.py 6:1920:2 > $O/ncu_capture.log 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > $O/${TAG}_raw.csv 2>/dev/null
python tools/ncu_summary.py /tmp/prof_$TAG.ncu-rep "$TAG: first two wavefront iterations of scene 6 FHD (4.15 M camera rays, then 2.54 M survivors)" > $O/${TAG}_kernels_ncu.md
for K in k_top k_mesh_walk k_mesh_enter k_mesh_multi; do
  ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv -k regex:$K 2>/dev/null | gzip > $O/${TAG}_src_$K.csv.gz
done
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv -k regex:"k_shade<.int.2|k_shade<.int.5" 2>/dev/null | gzip > $O/${TAG}_src_k_shade.csv.gz
ls -la $O/${TAG}_* /tmp/prof_$TAG.ncu-rep
