#!/bin/bash
# cfg2 with different tile grids for stage 0 (debug aid)
for g in "5,5,5" "11,13,1" "13,11,1" "7,7,3" "10,5,3" "10,3,5" "12,4,3" "9,8,2" "8,9,2"; do
  echo -n "stage0 grid $g: "; MMH_TILE_STAGE=0 MMH_TILE_G=$g python scripts/quick_cfg2.py
done
