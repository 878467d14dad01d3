"""Developer tool: per-CTA time stamps of the in-kernel exchange (EPI 4) of the pass-2 X-stream kernel; needs
tools/libprmf_dbg.so built with -DPRMF_EPI_TIMING.  Launch with torchrun on >= 2 GPUs:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/epi_timing.py"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prmf_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libprmf_dbg.so")
import prmf_b200.build as b
b.is_stale = lambda: False
import torch
from prmf_b200 import CudaEngine, pack_pathways, synth
from prmf_b200.dist import DistContext, row_block
from prmf_b200.engine import attach_collectives

ctx = DistContext.from_env()
torch.cuda.set_device(ctx.local_rank)
m, n, k, P = 37032, 6750, 10, 300
lo, hi = row_block(m, ctx.world, ctx.rank)
rng = np.random.Generator(np.random.PCG64(0))
Gs = synth.random_pathway_graphs(rng, n, P)
X = torch.rand((hi - lo, n), dtype=torch.float64, device="cuda")
eng = CudaEngine(hi - lo, m, n, k, device=ctx.local_rank)
attach_collectives(eng, ctx)
eng.set_X(X); eng.set_pathways(pack_pathways(Gs, list(range(n))))
U0 = 3 * (1 - rng.random((m, k))); V0 = 3 * (1 - rng.random((n, k)))
eng.set_UV(U0[lo:hi], V0); eng.set_active(list(range(k)))
lib = _lib.load()
ncta = 147
acc = []
for it in range(20):
    eng.step(1, 900.0, 1e-3)
    st = (ctypes.c_ulonglong * (ncta * 8))()
    lib.prmf_debug_epi_stamps(st, ncta * 8)
    a = np.array(list(st), dtype=np.float64).reshape(ncta, 8)[:, :6]
    if it >= 5:
        acc.append(a)
a = np.stack(acc)                                # [it, cta, stamp]
t0 = a[:, :, 0].min(axis=1, keepdims=True)       # first CTA out of the main loop
names = ["main loop done", "panel barrier", "local sums+publish+fence", "peer flags", "peer sums", "V update+gram"]
if ctx.rank == 0:
    print("exchange mode:", eng.exchange_mode)
    rel = (a - t0[:, :, None]) / 1e3
    print("us after the first CTA left its main loop (mean over launches of: mean CTA | last CTA)")
    for i, nm in enumerate(names):
        print("  %-28s %7.1f | %7.1f" % (nm, rel[:, :, i].mean(), rel[:, :, i].max(axis=1).mean()))
    d = np.diff(a, axis=2) / 1e3
    print("per-phase us (mean CTA | max CTA):")
    for i, nm in enumerate(names[1:]):
        print("  %-28s %7.1f | %7.1f" % (nm, d[:, :, i].mean(), d[:, :, i].max(axis=1).mean()))
ctx.barrier()
eng.close()
