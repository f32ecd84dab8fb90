#!/bin/bash
# cfg2 with different tile grids for stage 1 (debug aid)
for g in "3,3" "4,4" "5,5" "6,6" "7,7" "2,5" "5,2" "3,5" "5,3" "4,6" "2,2"; do
  echo -n "stage1 grid $g: "; MMH_TILE_STAGE=1 MMH_TILE_G=$g python scripts/quick_cfg2.py
done
