#!/bin/bash
# A/B timing of generic-kernel build variants on ONE box (libraries built into variants/)
for lib in mettagrid_b200/libmettagrid_b200.so variants/lib_*.so; do
  echo "== $lib"
  METTAGRID_B200_LIB=$PWD/$lib python tools/quick_bench_generic.py c3 16384 2>&1 | tail -1
  METTAGRID_B200_LIB=$PWD/$lib python tools/quick_bench_generic.py c4 8192 2>&1 | tail -1
  METTAGRID_B200_NO_FAST=1 METTAGRID_B200_LIB=$PWD/$lib python tools/quick_bench.py 4096 16 flush 2>&1 | tail -1
done
