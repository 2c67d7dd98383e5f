#!/bin/bash
# One-GPU evidence run for profiles/ (run under gpurun; every ncu pass only after the same command ran clean without it):
#   tools/collect_evidence.sh TAG bench   -> gpurun_out/TAG_bench_1gpu.json, TAG_bench_reference_arm.json, TAG_matrix/, TAG_launches.csv
#   tools/collect_evidence.sh TAG quick   -> the same without the per-scene matrix
#   tools/collect_evidence.sh TAG ncu     -> tools/ncu_capture.sh: TAG_raw.csv, TAG_kernels_ncu.md, TAG_kernels_ncu.json (bench.py's roofline constants)
#   tools/collect_evidence.sh TAG checked -> tools/checked_probe.sh (memory-safety evidence: the range-checked build)
set -u
TAG=${1:-r2_x}; MODE=${2:-bench}
O=gpurun_out
mkdir -p $O
if [ "$MODE" = "bench" ] || [ "$MODE" = "quick" ]; then
  python __graft_entry__.py smoke 2>&1 | tail -1
  python -m pytest tests -m gpu -x -q 2>&1 | tail -3
  python bench.py --steps 5 --warmup 3 > $O/${TAG}_bench_1gpu.json 2> $O/${TAG}_bench_1gpu.err || echo "bench failed"
  python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference_arm.json 2> $O/${TAG}_ref.err || echo "reference arm failed"
  [ "$MODE" = "bench" ] && tools/scene_matrix.sh $O/${TAG}_matrix "3 1 5 7 70" "1" 400 5
  # launch list of a bench step (cold-cache, serialised per-launch times: only the kernels' SHARES are meaningful)
  python bench.py --steps 1 --warmup 3 --spp 32 --no-cpu-baseline --no-target-render > /dev/null 2>&1 && \
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/${TAG}_launches.csv \
      python bench.py --steps 1 --warmup 3 --spp 32 --no-cpu-baseline --no-target-render > $O/ncu_bench.log 2>&1
  cut -c1-400 $O/${TAG}_bench_1gpu.json
elif [ "$MODE" = "ncu" ]; then
  tools/ncu_capture.sh $TAG
  python tools/ncu_constants.py $O/${TAG}_raw.csv > $O/${TAG}_kernels_ncu.json
elif [ "$MODE" = "checked" ]; then
  tools/checked_probe.sh $TAG
fi
