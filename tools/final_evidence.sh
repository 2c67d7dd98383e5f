set -u
O=gpurun_out; T=r1_l
python __graft_entry__.py smoke 2>&1 | tail -1
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > $O/${T}_bench_1gpu.json 2> $O/${T}_bench_1gpu.err || echo "bench failed"
python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_reference_arm.json 2> $O/${T}_ref.err || echo "reference arm failed"
python bench.py --steps 1 --warmup 3 --spp 32 --no-cpu-baseline > /dev/null 2>&1 && \
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/${T}_launches.csv \
    python bench.py --steps 1 --warmup 3 --spp 32 --no-cpu-baseline > $O/ncu_bench.log 2>&1
cut -c1-200 $O/${T}_bench_1gpu.json
