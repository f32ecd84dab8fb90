#!/bin/bash
# Debug build for `compute-sanitizer --tool racecheck`: identical sources, but the step hand-offs of k_march_tiled2 and k_march_rows use named barriers
# (bar.sync / bar.arrive, which racecheck models) instead of the mbarrier split arrive / wait (which it reports as a hazard).
#   MMH_LIBRARY=mrmustard_b200/csrc/libmmhermite_barsync.so compute-sanitizer --tool racecheck python scripts/sanitizer_cases.py
set -e
cd "$(dirname "$0")/../mrmustard_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=false -Xcompiler -fPIC -shared -DMMH_T2_HANDOFF_BAR=1 -DMMH_RW_HANDOFF_BAR=1 \
     -o libmmhermite_barsync.so mmh_api.cu mmh_forward.cu mmh_march.cu mmh_lanes.cu mmh_box.cu mmh_tiled.cu mmh_rows.cu mmh_stable_boxes.cu mmh_vjp.cu mmh_diagonal.cu \
     mmh_diagonal_rolling.cu mmh_gates.cu mmh_autoshape.cu mmh_einsum.cu -lcudart
echo built libmmhermite_barsync.so
