"""cfg1 (the latency config: one (200,) chain) timed per call with CUDA events, for `ncu --metrics gpu__time_duration.sum`."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrmustard_b200 import _lib
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "vanilla_golden.npz"))
dev = torch.device("cuda:0")
dA, db, dc = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (g["cfg1_A"], g["cfg1_b"], g["cfg1_c"].reshape(1)))
dG = torch.empty(200, dtype=torch.complex128, device=dev)
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
f = lambda: _lib.check(_lib.lib.mmh_forward(1, _lib.shape_array((200,)), dA.data_ptr(), db.data_ptr(), dc.data_ptr(), dG.data_ptr(), 0, sp))
for _ in range(3): f()
torch.cuda.synchronize()
ms = []
for _ in range(10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); f(); b.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(b) * 1e3)
print("cfg1 event-to-event us:", [round(x, 1) for x in ms])
assert np.array_equal(dG.cpu().numpy(), g["cfg1_G"])
