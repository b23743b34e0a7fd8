#!/bin/bash
# timing sweep over kernel build variants (scratch; libs are built into variants/)
for lib in mettagrid_b200/libmettagrid_b200.so variants/lib_*.so; do
  echo "== $lib"
  METTAGRID_B200_LIB=$PWD/$lib python tests/quick_bench.py 4096 16 2>&1 | tail -1
  METTAGRID_B200_LIB=$PWD/$lib python tests/quick_bench.py 32768 16 2>&1 | tail -1
done
