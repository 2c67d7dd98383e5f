#!/bin/bash
# compute-sanitizer over every kernel flavour (GPU box): tools/sanitize.sh TAG -> gpurun_out/TAG_sanitizer/{memcheck,racecheck,synccheck}.log
set -u
TAG=${1:-r2}; O=gpurun_out/${TAG}_sanitizer; mkdir -p $O
python tools/sanitize_probe.py > $O/plain.log 2>&1 || { echo "probe failed without the sanitizer"; tail -5 $O/plain.log; exit 1; }
for TOOL in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $TOOL --print-limit 20 python tools/sanitize_probe.py > $O/$TOOL.full.log 2>&1
  echo "exit code $?" >> $O/$TOOL.full.log
  grep -E "^=========|exit code" $O/$TOOL.full.log | grep -vE "^========= *$" | head -60 > $O/$TOOL.log
  tail -2 $O/$TOOL.log
done
rm -f $O/*.full.log.tmp; ls -la $O
