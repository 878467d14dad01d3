"""Developer tool: where the persistent step kernel spends its time (tools/libprmf_dbg.so built with
`python -m prmf_b200.build --debug PRMF_BLOCK_TIMING`).  One GPU: python tools/block_timing.py [--m M];
several: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/block_timing.py"""
import argparse, ctypes, os, sys
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
ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=37032); ap.add_argument("--n", type=int, default=6750); ap.add_argument("--k", type=int, default=10)
a = ap.parse_args()
ctx = DistContext.from_env()
torch.cuda.set_device(ctx.local_rank)
lo, hi = row_block(a.m, ctx.world, ctx.rank)
rng = np.random.Generator(np.random.PCG64(0))
Gs = synth.random_pathway_graphs(rng, a.n, 300)
X = torch.rand((hi - lo, a.n), dtype=torch.float64, device="cuda")
eng = CudaEngine(hi - lo, a.m, a.n, a.k, device=ctx.local_rank)
attach_collectives(eng, ctx)
eng.set_X(X); eng.set_pathways(pack_pathways(Gs, list(range(a.n))))
U0 = 3 * (1 - rng.random((a.m, a.k))); V0 = 3 * (1 - rng.random((a.n, a.k)))
eng.set_UV(U0[lo:hi], V0); eng.set_active(list(range(a.k)))
lib = _lib.load()
lib.prmf_debug_block_stamps.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
H = 24
for _ in range(3):
    eng.step(10, 900.0, 1e-3)
lib.prmf_debug_block_stamps(None, 0, 1)
ctx.barrier()
eng.step(10, 900.0, 1e-3)
buf = (ctypes.c_ulonglong * (160 * H * 8))()
assert lib.prmf_debug_block_stamps(buf, 160 * H * 8, 0) == 0
t = np.array(list(buf), dtype=np.float64).reshape(160, H, 8)
nh = 20
def us(x): return x / 1e3
lines = []
for i in range(2, nh):                       # skip the first step (cold)
    st = t[:, i, :]
    live = st[:, 0] > 0
    inn = live & (st[:, 2] > 0) & (st[:, 3] > 0)
    if not inn.any():
        continue
    s = st[inn]
    t_begin = s[:, 0].min()
    row = {"half": i, "pass": 1 + (i & 1),
           "first data after entry (max)": us((s[:, 1] - s[:, 0]).max()),
           "main loop (median)": us(np.median(s[:, 2] - s[:, 1])),
           "main loop end spread": us(s[:, 2].max() - s[:, 2].min()),
           "panel barrier wait (median)": us(np.median(s[:, 3] - s[:, 2])),
           "exchange (median)": us(np.median((s[:, 4] - s[:, 3])[s[:, 4] > 0])) if (s[:, 4] > 0).any() else 0.0,
           "update (median)": us(np.median(s[:, 5] - np.maximum(s[:, 3], s[:, 4]))),
           "gram+done+fold (median)": us(np.median(s[:, 6] - s[:, 5])),
           "half total first-entry -> last-exit": us(s[:, 6].max() - t_begin),
           "tail total: last main-loop end -> last exit": us(s[:, 6].max() - s[:, 2].max()),
           "producer dep satisfied after last exit of prev half": 0.0}
    if i > 0:
        prev = t[:, i - 1, :]
        pl = prev[:, 6] > 0
        if pl.any() and (st[:, 7] > 0).any():
            row["producer dep satisfied after last exit of prev half"] = us(np.median(st[st[:, 7] > 0, 7]) - prev[pl, 6].max())
            row["entry after last exit of prev half (median)"] = us(np.median(s[:, 0]) - prev[pl, 6].max())
    lines.append(row)
if not lines:
    print("rank %d: no complete stamps" % ctx.rank); raise SystemExit(0)
keys = [k for k in lines[0] if k not in ("half", "pass")]
for ps in (1, 2):
    sel = [r for r in lines if r["pass"] == ps]
    print("rank %d pass %d (%d halves)" % (ctx.rank, ps, len(sel)))
    for kx in keys + ["entry after last exit of prev half (median)"]:
        v = [r[kx] for r in sel if kx in r]
        if v:
            print("   %-56s mean %8.2f us   min %8.2f  max %8.2f" % (kx, np.mean(v), np.min(v), np.max(v)))
span = t[:, 2:nh, :]
ok = span[:, :, 6] > 0
print("rank %d: %d halves in %.1f us -> %.1f us per step" % (ctx.rank, nh - 2, us(span[:, :, 6][ok].max() - span[:, :, 0][span[:, :, 0] > 0].min()),
                                                         us(span[:, :, 6][ok].max() - span[:, :, 0][span[:, :, 0] > 0].min()) / ((nh - 2) / 2)))
ctx.barrier()
eng.close()
