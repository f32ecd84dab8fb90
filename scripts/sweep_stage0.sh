#!/bin/bash
# cfg2 with different tile grids for stage 0 (debug aid)
for g in "5,5,5" "9,4,4" "9,5,3" "9,3,5" "6,6,4" "6,4,6" "4,6,6"; do
  echo -n "stage0 grid $g: "; MMH_TILE_STAGE=0 MMH_TILE_G=$g python scripts/quick_cfg2.py
done
