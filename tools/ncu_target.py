"""Short profiling target: a few inner steps of the shipping path at a BASELINE configuration, on one GPU.

    python tools/ncu_target.py [--config 2|4] [--x-dtype f64|tf32] [--m M] [--n N] [--k K] [--pathways P] [--steps S] [--scores]

Used under `ncu` (launch list, and `--replay-mode application` captures of the cooperative X-stream kernels, whose
CTAs wait for each other and therefore cannot be replayed one kernel at a time)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--m", type=int, default=None)
ap.add_argument("--n", type=int, default=None)
ap.add_argument("--k", type=int, default=None)
ap.add_argument("--pathways", type=int, default=None)
ap.add_argument("--x-dtype", default=None)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--scores", action="store_true")
a = ap.parse_args()
CFG = {2: (37032, 6750, 10, 300, "f64"), 4: (37032, 6750, 64, 2000, "f64"), 5: (65536, 20000, 128, 2000, "tf32")}[a.config]
m, n, k, P, xd = (a.m or CFG[0], a.n or CFG[1], a.k or CFG[2], a.pathways or CFG[3], a.x_dtype or CFG[4])

import torch
from prmf_b200 import CudaEngine, pack_pathways, synth

rng = np.random.Generator(np.random.PCG64(0))
gen = torch.Generator(device="cuda"); gen.manual_seed(7)
X = torch.rand((m, n), dtype=torch.float32 if xd == "tf32" else torch.float64, device="cuda", generator=gen)
Gs = synth.random_pathway_graphs(rng, n, P)
eng = CudaEngine(m, m, n, k, x_dtype=xd)
eng.set_X(X)
del X
eng.set_pathways(pack_pathways(Gs, list(range(n))))
eng.set_UV(3 * (1 - rng.random((m, k))), 3 * (1 - rng.random((n, k))))
eng.set_active([f % P for f in range(k)])
normX = float(np.sqrt(eng.normX_sq))
parts, _, _ = eng.step(a.steps, normX / k, 10 / normX)
if a.scores:
    eng.scores()
print("ncu_target ok: m=%d n=%d k=%d P=%d %s obj=%r launches=%d" % (m, n, k, P, xd, float(parts[-1, 4]), eng.launch_count))
eng.close()
