"""cfg4 as written (8-mode diagonal strategy, cutoff 12) timed device-resident; MMH_LIBRARY selects the build."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mrmustard_b200 as mm
from mrmustard_b200 import _lib
g = np.load(os.path.join(ROOT, "tests", "golden", "vanilla_golden.npz"))
A, b, c = g["cfg4_A"], g["cfg4_b"], complex(g["cfg4_c"])
Adm = np.zeros((16, 16), complex); Adm[:8, :8] = np.conj(A); Adm[8:, 8:] = A
bdm = np.concatenate([np.conj(b), b])
A2, b2 = (np.ascontiguousarray(x) for x in mm.backend.reorder_AB_bargmann(Adm, bdm))
dev = torch.device("cuda:0")
to = lambda x: torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.complex128))).to(dev)
dA, dB, dG0 = to(A2), to(b2), to(np.array([abs(c) ** 2]))
cut = tuple(int(x) for x in (sys.argv[1:] or ["12"] * 8))
if len(cut) == 1: cut = cut * 8
out = torch.empty(int(np.prod(cut)), dtype=torch.complex128, device=dev)
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
f = lambda: _lib.check(_lib.lib.mmh_diagonal(8, _lib.shape_array(cut), dA.data_ptr(), dB.data_ptr(), 0, dG0.data_ptr(), out.data_ptr(), sp))
f(); torch.cuda.synchronize()
ms = []
for _ in range(3):
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); f(); e.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(e))
print(os.environ.get("MMH_LIBRARY", "default"), cut, "ms:", [round(x, 1) for x in ms], "checksum", float(out.real.sum()))
