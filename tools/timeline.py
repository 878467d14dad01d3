"""Developer tool: device timeline of the fused-tail inner steps (tools/libprmf_dbg.so built with -DPRMF_EPI_TIMING
-DPRMF_TAIL_TIMING): for every X-stream launch the time the first CTA started, the last CTA left its main loop and
the last CTA ended, so the gaps BETWEEN launches and the cost of the fused tails can be read off directly."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prmf_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libprmf_dbg.so")
import prmf_b200.build as b
b.is_stale = lambda: False
import torch
from prmf_b200 import CudaEngine, pack_pathways, synth
m, n, k, P = 37032, 6750, 10, 300
rng = np.random.Generator(np.random.PCG64(0))
X = torch.rand((m, n), dtype=torch.float64, device="cuda")
Gs = synth.random_pathway_graphs(rng, n, P)
eng = CudaEngine(m, m, n, k)
eng.set_X(X); eng.set_pathways(pack_pathways(Gs, list(range(n))))
eng.set_UV(3 * (1 - rng.random((m, k))), 3 * (1 - rng.random((n, k)))); eng.set_active(list(range(k)))
lib = _lib.load()
eng.step(10, 900.0, 1e-3)
buf = (ctypes.c_ulonglong * 256)()
lib.prmf_debug_timeline(buf, 1)
eng.step(20, 900.0, 1e-3)                      # 40 X-stream launches -> slots wrap at 64: all distinct
lib.prmf_debug_timeline(buf, 0)
t = np.array(list(buf), dtype=np.float64).reshape(64, 4)[:, :3]
t = t[t[:, 2] > 0]
t = t[np.argsort(t[:, 0])]
dur = (t[:, 2] - t[:, 0]) / 1e3
main = (t[:, 1] - t[:, 0]) / 1e3
tail = (t[:, 2] - t[:, 1]) / 1e3
gap = (t[1:, 0] - t[:-1, 2]) / 1e3
print("launches %d | kernel %.1f us (main loop %.1f + fused tail %.1f) | gap to the next X-stream launch: mean %.1f us"
      " (even->odd %.1f, odd->even incl. objective %.1f)" % (len(t), dur.mean(), main.mean(), tail.mean(), gap.mean(),
                                                              gap[0::2].mean(), gap[1::2].mean()))
print("step = %.1f us" % ((t[-1, 2] - t[0, 0]) / 1e3 / (len(t) / 2)))
