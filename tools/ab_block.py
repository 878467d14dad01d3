"""Developer tool: A/B of the persistent step kernel (PRMF_BLOCK=1) against two launches per step (PRMF_BLOCK=0) in one
process, alternating, at BASELINE config 2 (or --m/--n/--k).  Prints ms per inner step for both."""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=37032); ap.add_argument("--n", type=int, default=6750)
ap.add_argument("--k", type=int, default=10); ap.add_argument("--pathways", type=int, default=300)
ap.add_argument("--rounds", type=int, default=5); ap.add_argument("--blocks", type=int, default=20)
ap.add_argument("--stages", default="")
a = ap.parse_args()
import torch
from prmf_b200 import CudaEngine, pack_pathways, synth
rng = np.random.Generator(np.random.PCG64(0))
gen = torch.Generator(device="cuda"); gen.manual_seed(7)
X = torch.rand((a.m, a.n), dtype=torch.float64, device="cuda", generator=gen)
Gs = synth.random_pathway_graphs(rng, a.n, a.pathways)
packed = pack_pathways(Gs, list(range(a.n)))
U0 = 3 * (1 - rng.random((a.m, a.k))); V0 = 3 * (1 - rng.random((a.n, a.k)))
engs = {}
import threading
import pynvml
pynvml.nvmlInit()
_h = pynvml.nvmlDeviceGetHandleByIndex(0)
class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True); self.stop = False; self.clk = []; self.pw = []
    def run(self):
        while not self.stop:
            self.clk.append(pynvml.nvmlDeviceGetClockInfo(_h, pynvml.NVML_CLOCK_SM)); self.pw.append(pynvml.nvmlDeviceGetPowerUsage(_h) / 1e3)
            time.sleep(0.003)
MODES = ("1", "1f", "0")
for mode in MODES:
    os.environ["PRMF_BLOCK"] = mode[0]
    os.environ["PRMF_BLOCK_FLAGS"] = "1" if mode.endswith("f") else "0"
    e = CudaEngine(a.m, a.m, a.n, a.k)
    e.set_X(X); e.set_pathways(packed); e.set_UV(U0, V0); e.set_active([f % a.pathways for f in range(a.k)])
    engs[mode] = e
del X
normX = float(np.sqrt(engs["1"].normX_sq)); g, d = normX / a.k, 10 / normX
res = {m_: [] for m_ in MODES}
clk = {m_: [] for m_ in MODES}
for r in range(a.rounds):
    for mode in MODES:
        e = engs[mode]
        os.environ["PRMF_BLOCK_FLAGS"] = "1" if mode.endswith("f") else "0"
        e.step(10, g, d)                       # warm
        torch.cuda.synchronize()
        sm = Sampler(); sm.start()
        t0 = time.perf_counter()
        for _ in range(a.blocks):
            e.step_async(10, g, d)
        parts, _, _ = e.step_collect(10)
        torch.cuda.synchronize()
        res[mode].append((time.perf_counter() - t0) * 1e3 / (10 * a.blocks))
        sm.stop = True; sm.join()
        clk[mode].append((np.median(sm.clk), np.max(sm.pw)))
for mode in MODES:
    v = np.array(res[mode])
    print("PRMF_BLOCK=%s: ms per inner step: median %.4f  min %.4f  max %.4f   (%s)  sm MHz / W: %s" % (
        mode, np.median(v), v.min(), v.max(), " ".join("%.4f" % x for x in v), " ".join("%d/%d" % c for c in clk[mode])))
