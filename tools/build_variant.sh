#!/bin/bash
# A/B variant of the CUDA library that differs from lib/ in ONE translation unit's flags:
#   tools/build_variant.sh lib_x "trace" "-DPT_WALK_TOS=0"      (objects of the other units are copied from lib/, which must be up to date)
set -eu
V=$1; UNITS=$2; EXTRA=$3
cd "$(dirname "$0")/../thu-acg-f2024-path-tracer_b200"
mkdir -p $V/obj
cp -p lib/obj/*.o lib/obj/*.ptxas.log $V/obj/
for u in $UNITS; do rm -f $V/obj/$u.o; done
make -j8 LIB=$V BIN=${V/lib/bin} EXTRA="$EXTRA" 2>&1 | grep -iE "error" || true
ls -la $V/libptb200.so | cut -c20-
