#!/bin/bash
# One-GPU evidence run for profiles/ (run under gpurun; every ncu pass only after the same command ran clean without it):
#   tools/collect_evidence.sh TAG bench   -> gpurun_out/TAG_*.json, TAG_matrix/, TAG_launches.csv
#   tools/collect_evidence.sh TAG quick   -> the same without the per-scene matrix (about 3.5 GPU-minutes)
#   tools/collect_evidence.sh TAG ncu     -> gpurun_out/prof_{trace,shade}_TAG.ncu-rep + TAG_k_{trace,shade}_ncu.md
# gpurun copies back at most 64 MiB, so the two halves are separate calls and the captures are kept small.
set -u
TAG=${1:-r1_x}; MODE=${2:-bench}
O=gpurun_out
mkdir -p $O
if [ "$MODE" = "bench" ] || [ "$MODE" = "quick" ]; then
  python __graft_entry__.py smoke 2>&1 | tail -1
  python -m pytest tests -m gpu -x -q 2>&1 | tail -3
  python bench.py --steps 5 --warmup 3 > $O/${TAG}_bench_1gpu.json 2> $O/${TAG}_bench_1gpu.err || echo "bench failed"
  python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference_arm.json 2> $O/${TAG}_ref.err || echo "reference arm failed"
  [ "$MODE" = "bench" ] && tools/scene_matrix.sh $O/${TAG}_matrix "3 1 5 7 70" "1" 400 5
  # launch list of a bench step (cold-cache, serialised per-launch times: only the kernels' SHARES are meaningful)
  python bench.py --steps 1 --warmup 3 --spp 32 --no-cpu-baseline > /dev/null 2>&1 && \
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/${TAG}_launches.csv \
      python bench.py --steps 1 --warmup 3 --spp 32 --no-cpu-baseline > $O/ncu_bench.log 2>&1
  cut -c1-300 $O/${TAG}_bench_1gpu.json
else
  python tools/perf_probe.py 6:1920:8 > /dev/null 2>&1 || { echo "probe failed"; exit 1; }
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace --launch-skip 4 --launch-count 2 -f \
    -o $O/prof_trace_${TAG} python tools/perf_probe.py 6:1920:8 > $O/ncu_trace.log 2>&1
  timeout 900 ncu --set full --clock-control none -k regex:k_shade --launch-skip 7 --launch-count 6 -f \
    -o $O/prof_shade_${TAG} python tools/perf_probe.py 6:1920:8 > $O/ncu_shade.log 2>&1
  python tools/ncu_summary.py $O/prof_trace_${TAG}.ncu-rep "k_trace<DEFER> + k_trace_blas_refill round 0 (${TAG}) - scene 6 FHD, perf_probe.py 6:1920:8, first bounce iteration" > $O/${TAG}_k_trace_ncu.md
  python tools/ncu_summary.py $O/prof_shade_${TAG}.ncu-rep "k_shade<class> (${TAG}) - scene 6 FHD, perf_probe.py 6:1920:8, launches 8-13" > $O/${TAG}_k_shade_ncu.md
  rm -f $O/prof_shade_${TAG}.ncu-rep   # ~40 MB: the summary is what profiles/ keeps; the trace report stays for the source page
  ls -la $O/*${TAG}*; du -sh $O
fi
