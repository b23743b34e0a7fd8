#!/bin/bash
# A/B timing of generic-kernel build variants on ONE box (libraries built into variants/ by
# python -m mettagrid_b200.build --variant NAME -DFLAG...)
shopt -s nullglob
for lib in mettagrid_b200/libmettagrid_b200.so variants/lib_*.so; do
  case $lib in *prof*) continue;; esac
  echo "== $lib"
  METTAGRID_B200_LIB=$PWD/$lib python tools/quick_bench_generic.py c3 16384 flush 2>&1 | tail -1
  METTAGRID_B200_LIB=$PWD/$lib python tools/quick_bench_generic.py c4 8192 flush 2>&1 | tail -1
  METTAGRID_B200_LIB=$PWD/$lib python tools/quick_bench_generic.py toy 4096 flush 2>&1 | tail -1
done
