#!/bin/bash
# A/B probe (GPU box): tools/ab_probe.sh "lib lib_x lib_y" "6:1920:32 70:600:100" -> one perf_probe block per library variant
# (variants are built with `make LIB=lib_x BIN=bin_x EXTRA=-DPT_SOMETHING=1` and loaded through PT_B200_LIBDIR)
for L in $1; do
  echo "=== $L"
  PT_B200_LIBDIR=$PWD/thu-acg-f2024-path-tracer_b200/$L python tools/perf_probe.py $2 2>&1 | grep -v "prof=2" | sed 's/^.*prof=/prof=/' | cut -d'|' -f1,3,5-
done
