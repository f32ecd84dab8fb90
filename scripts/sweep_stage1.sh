#!/bin/bash
# cfg2 with different tile grids for stage 1 (debug aid)
for g in "3,3" "2,2" "2,3" "3,2" "3,4" "4,3" "4,4" "2,4" "4,2" "3,5" "1,3" "3,1"; do
  echo -n "stage1 grid $g: "; MMH_TILE_STAGE=1 MMH_TILE_G=$g python scripts/quick_cfg2.py
done
