"""Developer tool: where the wall time of one whole solve goes (nmf_pathway on the bench's planted instance, host X)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = [sys.argv[0]]
import importlib.util
spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
import torch
from prmf_b200 import CudaEngine, nmf_pathway
from prmf_b200.solver import init_latent_to_pathway_data

a = bench.parse_args()
Gs, nodelist, packed = bench.make_pathways(a)
Xh = torch.empty((a.m, a.n), dtype=torch.float64, pin_memory=True)
bench.host_rows(0, a.m, a.n, out=Xh.numpy())
bench.plant_signal(Xh.numpy(), Gs, 0, a.m)
X = Xh.numpy()
torch.cuda.synchronize()

def tick(label, t0):
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print("  %-28s %8.2f ms" % (label, (t1 - t0) * 1e3))
    return t1

for rep in range(3):
    print("rep", rep)
    t0 = time.perf_counter()
    eng = CudaEngine(a.m, a.m, a.n, a.k)
    t0 = tick("engine create", t0)
    eng.set_X(X)
    t0 = tick("set_X (pinned host)", t0)
    _ = eng.normX_sq
    t0 = tick("normX readback", t0)
    eng.set_pathways(packed)
    t0 = tick("set_pathways", t0)
    np.random.seed(1)
    U0 = 3 * (1 - np.random.rand(a.m, a.k)); V0 = 3 * (1 - np.random.rand(a.n, a.k))
    t0 = tick("draw U0 V0", t0)
    eng.set_UV(U0, V0)
    t0 = tick("set_UV", t0)
    cands = init_latent_to_pathway_data(a.k, packed.P)
    t0 = tick("init candidates", t0)
    eng.set_active(list(range(a.k)))
    eng.step_async(10, 900.0, 1e-3)
    eng.block_end(10, want_scores=True, prefetch=True)
    t0 = tick("first block (10 steps)", t0)
    eng.step_async(10, 900.0, 1e-3)
    eng.block_end(10, want_scores=True, prefetch=True)
    t0 = tick("second block", t0)
    eng.get_UV()
    t0 = tick("get_UV", t0)
    eng.close()
    t0 = tick("close", t0)
    np.random.seed(1)
    t0 = time.perf_counter()
    trace = {"keep_blocks": 0}
    nmf_pathway(X, packed, k_latent=a.k, nodelist=nodelist, quiet=True, trace=trace)
    tick("nmf_pathway whole (%d steps)" % len(trace["obj_parts"]), t0)
    Xd = torch.from_numpy(X).cuda()
    np.random.seed(1)
    t0 = time.perf_counter()
    nmf_pathway(Xd, packed, k_latent=a.k, nodelist=nodelist, quiet=True)
    tick("nmf_pathway, X on device", t0)
    del Xd
