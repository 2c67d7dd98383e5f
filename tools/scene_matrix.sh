#!/bin/bash
# Per-scene bench matrix (BASELINE.json configs): tools/scene_matrix.sh OUT_DIR "SCENES" "GPU_COUNTS" [SPP] [STEPS]
#   e.g. tools/scene_matrix.sh gpurun_out/matrix "3 1 5 7 70" "1 8" 400 5
# Every cell is one bench.py run (its JSON line lands in OUT_DIR/scene<S>_n<N>.json); N = 1 also times the CPU baseline.
set -u
OUT=${1:-gpurun_out/matrix}; SCENES=${2:-"3 1 5 7 70"}; GPUS=${3:-"1"}; SPP=${4:-400}; STEPS=${5:-5}
mkdir -p "$OUT"
PORT=29611
for s in $SCENES; do
  for n in $GPUS; do
    f="$OUT/scene${s}_n${n}.json"
    if [ "$n" = "1" ]; then
      python bench.py --gpus 1 --steps "$STEPS" --warmup 3 --scene "$s" --spp "$SPP" > "$f" 2> "$OUT/scene${s}_n${n}.err"
    else
      PORT=$((PORT + 1))
      python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port "$PORT" \
        bench.py --gpus "$n" --steps "$STEPS" --warmup 3 --scene "$s" --spp "$SPP" > "$f" 2> "$OUT/scene${s}_n${n}.err"
    fi
    python - "$f" <<'EOF'
import json, sys
try:
    l = json.loads([x for x in open(sys.argv[1]) if x.startswith("{")][-1])
    cb = l.get("cpu_baseline") or {}
    r = l["roofline"]; oc = r["survey_model"]["oracle_counters_per_segment"]
    print(f'{l["config"]["workload"]:44s} N={l["n_gpus"]} {l["value"]:9.1f} Mrays/s {l["samples_per_s"]:.3e} samples/s e2e {l["e2e"]["value"]:9.1f} '
          f'seg/path {l["segments_per_path"]:.2f} stages {r["stage_ms_per_step"]} oracle counters/seg boxes {oc["boxes"]:.1f} sph {oc["spheres"]:.2f} quad {oc["quads"]:.2f} tri {oc["triangles"]:.2f} '
          f'cpu {cb.get("value", float("nan")):.2f} Mrays/s x{cb.get("cores", 0)} cores')
except Exception as e:
    print(sys.argv[1], "FAILED", e)
EOF
  done
done
