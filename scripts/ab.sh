#!/bin/bash
# A/B two builds of libmmhermite.so on the same box: default build vs mrmustard_b200/csrc/alt/libmmhermite.so
for rep in 1 2; do
  echo "== default"; python scripts/quick_cfg2.py; python scripts/one_tile.py 450
  cp mrmustard_b200/csrc/libmmhermite.so /tmp/default.so; cp mrmustard_b200/csrc/alt/libmmhermite.so mrmustard_b200/csrc/libmmhermite.so
  echo "== alt"; python scripts/quick_cfg2.py; python scripts/one_tile.py 450
  cp /tmp/default.so mrmustard_b200/csrc/libmmhermite.so
done
