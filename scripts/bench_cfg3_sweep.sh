#!/bin/bash
# sweep K2 launch shapes on cfg3 (run under gpurun)
for L in ${LS:-4 8 12 16 24}; do for R in ${RS:-1 2 4}; do
  echo -n "L=$L R=$R : "
  MMH_K2_L=$L MMH_K2_R=$R python bench.py --workload cfg3 --steps 10 --no-extras --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['roofline']['frac'],4))"
done; done
