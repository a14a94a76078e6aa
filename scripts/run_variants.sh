#!/bin/bash
# Developer tool (GPU box): time each built fast-kernel variant.  Usage: run_variants.sh scale name...
scale=$1; shift
for v in "$@"; do
  ISCE3_B200_LIB=isce3_b200/csrc/build/variants/lib_$v.so python scripts/perf_fast.py $scale $v 2>&1 | tail -1
done
