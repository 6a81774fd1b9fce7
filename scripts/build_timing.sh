#!/bin/bash
# Instrumented build (per-stage clock64 counters, -DYALPS_TIMING) -> yalps_b200/libyalps_timing.so, for scripts/phase_timing.py
set -e
cd "$(dirname "$0")/../yalps_b200/csrc"
mkdir -p _build_timing
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -DYALPS_TIMING"
for u in yalps_b200 ktab_base ktab_split_a ktab_split_b ktab_split_c ktab_cluster ktab_tmem ktab_bnb; do
  nvcc $FLAGS -c -o _build_timing/$u.o $u.cu > _build_timing/$u.log 2>&1 &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libyalps_timing.so _build_timing/*.o
