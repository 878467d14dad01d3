"""Developer tool: where the warps of fused_xvu_kernel spend their cycles (tools/libprmf_dbg.so built with
-DPRMF_FUSED_TIMING).  PRMF_FUSED=1 python tools/fused_timing.py"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PRMF_FUSED"] = "1"
from prmf_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libprmf_dbg.so")
import prmf_b200.build as b
b.is_stale = lambda: False
from prmf_b200 import CudaEngine, pack_pathways, synth
import torch
m, n, k, P = 37032, 6750, 10, 300
rng = np.random.Generator(np.random.PCG64(0))
X = torch.rand((m, n), dtype=torch.float64, device="cuda")
Gs = synth.random_pathway_graphs(rng, n, P)
eng = CudaEngine(m, m, n, k)
eng.set_X(X); eng.set_pathways(pack_pathways(Gs, list(range(n))))
eng.set_UV(3 * (1 - rng.random((m, k))), 3 * (1 - rng.random((n, k)))); eng.set_active(list(range(k)))
lib = _lib.load()
for it in range(5):
    eng.step(1, 900.0, 1e-3)
st = (ctypes.c_ulonglong * 32)()
lib.prmf_debug_fused(st)
v = np.array(list(st), dtype=np.float64) / 1e3
print("exchange warp 0 (kcycles): wait pa_full %.0f | sum+publish %.0f | poll+sum %.0f | U update+arrive %.0f | loop %.0f" % (v[0], v[1], v[2], v[3], v[7]))
for base, name in ((8, "consumer (panel 0, group 0, warp 0)"), (16, "consumer (panel 5, group 3, warp 3)")):
    print("%s (kcycles): wait TMA %.0f | wait pa_empty %.0f | phase A %.0f | wait U_new %.0f | phase B %.0f | loop %.0f" % (
        name, v[base], v[base + 1], v[base + 2], v[base + 3], v[base + 4], v[base + 7]))
