#!/bin/bash
# Memory-safety evidence without compute-sanitizer (closed on this GPU pool): the checked build (-DPT_CHECKED=1: every table
# index, queue slot and stack push range-checked on the device, violations printed) runs every kernel flavour
# (tools/sanitize_probe.py) and a part of the parity tests.  tools/checked_probe.sh TAG -> gpurun_out/TAG_sanitizer/checked_build.log
set -u
TAG=${1:-r2}; O=gpurun_out/${TAG}_sanitizer; mkdir -p $O
L=$PWD/thu-acg-f2024-path-tracer_b200/lib_checked
[ -f $L/libptb200.so ] || { echo "build it first: make -C thu-acg-f2024-path-tracer_b200 LIB=lib_checked BIN=bin_checked EXTRA=-DPT_CHECKED=1"; exit 1; }
{
  echo "# checked build (-DPT_CHECKED=1), $(date -u +%FT%TZ), $(nvidia-smi --query-gpu=name --format=csv,noheader | head -1)"
  echo "## tools/sanitize_probe.py"
  PT_B200_LIBDIR=$L python tools/sanitize_probe.py 2>&1
  echo "## pytest -m gpu -k 'traversal_stage or start_of_path or two_pass or exact_tie or ragged or volumes or every_light'"
  PT_B200_LIBDIR=$L python -m pytest tests -m gpu -x -q -k "traversal_stage or start_of_path or two_pass or exact_tie or ragged or volumes or every_light" 2>&1 | tail -4
} > $O/checked_build.full.log 2>&1
N=$(grep -c "PT_CHECK failed" $O/checked_build.full.log)
grep -v "PT_CHECK failed" $O/checked_build.full.log > $O/checked_build.log
grep "PT_CHECK failed" $O/checked_build.full.log | sort | uniq -c | head -20 >> $O/checked_build.log
echo "## PT_CHECK violations: $N" >> $O/checked_build.log
rm -f $O/checked_build.full.log
tail -8 $O/checked_build.log
