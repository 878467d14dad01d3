"""Developer tool: phase time stamps inside v_update_objective_kernel (needs tools/libprmf_dbg.so built with
-DPRMF_TAIL_TIMING).  python tools/tail_timing.py"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prmf_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libprmf_dbg.so")
import prmf_b200.build as b
b.is_stale = lambda: False
from prmf_b200 import CudaEngine, pack_pathways, synth
m, n, k, P = 37032, 6750, 10, 300
rng = np.random.Generator(np.random.PCG64(0))
import torch
X = torch.rand((m, n), dtype=torch.float64, device="cuda")
Gs = synth.random_pathway_graphs(rng, n, P)
eng = CudaEngine(m, m, n, k)
eng.set_X(X); eng.set_pathways(pack_pathways(Gs, list(range(n))))
eng.set_UV(3 * (1 - rng.random((m, k))), 3 * (1 - rng.random((n, k)))); eng.set_active(list(range(k)))
lib = _lib.load()
acc = np.zeros(5); acc2 = np.zeros(5); cnt = 0
for it in range(30):
    eng.step(1, 900.0, 1e-3)
    st = (ctypes.c_ulonglong * 16)()
    lib.prmf_debug_tail_stamps(st)
    t = np.array([st[i] for i in range(6)], dtype=np.float64)
    if it >= 5:
        acc += np.diff(t) / 1e3; cnt += 1
        u = np.array([st[4], st[6], st[7], st[8], st[9], st[5]], dtype=np.float64)
        acc2 += np.diff(u) / 1e3
print("us: Gu-sum %.1f | elements %.1f | gram+store+blocksum %.1f | wait-for-last %.1f | objective %.1f" % tuple(acc / cnt))
print("objective us: prefetch %.1f | Gv-sum %.1f | entries %.1f | block-sum %.1f | final %.1f" % tuple(acc2 / cnt))
