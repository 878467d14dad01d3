"""Developer tool: warm device time of the k x P score tables (scores_kernel) at BASELINE configs 2 and 4."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prmf_b200 import CudaEngine, pack_pathways, synth
for n, k, P in [(6750, 10, 300), (6750, 64, 2000), (20000, 128, 2000)]:
    rng = np.random.Generator(np.random.PCG64(0))
    Gs = synth.random_pathway_graphs(rng, n, P)
    with CudaEngine(0, 0, n, k) as eng:
        eng.set_pathways(pack_pathways(Gs, list(range(n))))
        eng.set_UV(None, 3 * (1 - rng.random((n, k))))
        for _ in range(3):
            eng.scores()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 50
        for _ in range(reps):
            eng.scores()                      # launch + D2H of the three tables + sync
        dt = (time.perf_counter() - t0) / reps
        print("n=%d k=%d P=%d: prmf_scores call (launch + %0.1f KB D2H + sync): %.1f us" % (n, k, P, 3 * k * P * 8 / 1e3, dt * 1e6))
