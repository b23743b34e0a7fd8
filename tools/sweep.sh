#!/bin/bash
# A/B timing of kernel build variants on ONE box (scratch; variant libraries are built into variants/)
shopt -s nullglob
for lib in mettagrid_b200/libmettagrid_b200.so variants/lib_*.so; do
  echo "== $lib"
  for rep in 1 2; do
    METTAGRID_B200_LIB=$PWD/$lib python tools/quick_bench.py 4096 16 flush 2>&1 | tail -1
  done
  METTAGRID_B200_LIB=$PWD/$lib python tools/quick_bench.py 32768 16 flush 2>&1 | tail -1
done
