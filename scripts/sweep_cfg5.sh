#!/bin/bash
# cfg5 (40,)^4 forward with different stage-0 tile grids (debug aid)
cat > /tmp/q5.py <<'PY'
import os, sys, hashlib
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from mrmustard_b200 import _lib
gold = np.load("tests/golden/vanilla_golden.npz")
dev = torch.device("cuda:0")
dA, db, dc = (torch.from_numpy(np.ascontiguousarray(gold[k])).to(dev) for k in ("cfg5_A", "cfg5_b", "cfg5_c")); dc = dc.reshape(1)
shape = (40,) * 4; sh = _lib.shape_array(shape)
dG = torch.empty(shape, dtype=torch.complex128, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run(): _lib.check(_lib.lib.mmh_forward(4, sh, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), dG.data_ptr(), 0, None))
for _ in range(3): run()
torch.cuda.synchronize(); ms = []
for _ in range(15):
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(); b.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
ok = hashlib.sha256((dG.cpu().numpy() + 0.0).tobytes()).hexdigest() == str(gold["cfg5_G40_sha"])
print(f"median {np.median(ms)*1e3:.1f} us parity {'OK' if ok else 'MISMATCH'}")
PY
echo -n "default: "; python /tmp/q5.py
for g in "4,4,4" "5,5,5" "5,5,4" "5,4,4" "4,5,5" "6,5,4" "7,5,4" "5,5,3" "4,4,5"; do
  echo -n "stage0 grid $g: "; MMH_TILE_STAGE=0 MMH_TILE_G=$g python /tmp/q5.py
done
for g in "2,2" "3,3" "4,4" "3,2"; do
  echo -n "stage1 grid $g: "; MMH_TILE_STAGE=1 MMH_TILE_G=$g python /tmp/q5.py
done
