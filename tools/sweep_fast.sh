#!/bin/bash
# A/B timing of k_step_fast build variants on ONE box (variants/ built by python -m mettagrid_b200.build --variant NAME -DFLAG...)
shopt -s nullglob
for lib in mettagrid_b200/libmettagrid_b200.so variants/lib_*.so; do
  echo "== $lib"
  for n in ${SWEEP_ENVS:-4096}; do
    METTAGRID_B200_LIB=$PWD/$lib python tools/quick_bench.py $n 16 flush 2>&1 | tail -1
  done
  METTAGRID_B200_LIB=$PWD/$lib python tools/quick_bench_generic.py toy 4096 flush 2>&1 | tail -1
done
