"""Worker for the multi-process tests: run nmf_pathway on a golden fixture under torch.distributed and
write this rank's result.  Launched by torch.distributed.run (RANK / WORLD_SIZE / LOCAL_RANK in env).

    mp_worker.py <case> <outdir> <cpu|gpu>
"""
import contextlib
import io
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def synthetic_case(k, m=301, n=523, P=9, seed=11):
    """A planted instance that is not a fixture of the reference: used to compare sharded against single-GPU runs
    (e.g. large k, where the tails take the tiled kernels)."""
    from prmf_b200 import synth
    X, nodelist, Gs = synth.small_instance(m=m, n=n, k_true=3, n_pathways=P, pathway_size=14, seed=seed)
    meta = {"seed": 5, "gamma_in": 1.0, "delta_in": 1.0, "tradeoff": None, "k_latent": k, "max_iter": 30}
    return {"X": X, "Gs": Gs, "nodelist": nodelist, "meta": meta}


def main():
    case, outdir, mode = sys.argv[1], sys.argv[2], sys.argv[3]
    from conftest import load_golden
    from prmf_b200 import nmf_pathway
    from prmf_b200.dist import DistContext
    ctx = DistContext.from_env(backend="gloo" if mode == "cpu" else "nccl")
    if case.startswith("synth:"):
        g = synthetic_case(int(case.split(":")[1]))
    else:
        g = load_golden(case)
    meta = g["meta"]
    np.random.seed(meta["seed"]); random.seed(meta["seed"])
    kw = {}
    if mode == "cpu":
        import cpu_engine
        kw["engine_factory"] = cpu_engine.factory
    trace = {"keep_blocks": 0}
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        U, V, od = nmf_pathway(g["X"].copy(), [G.copy() for G in g["Gs"]], gamma=meta["gamma_in"],
                               delta=meta["delta_in"], tradeoff=meta["tradeoff"], k_latent=meta["k_latent"],
                               nodelist=list(g["nodelist"]), max_iter=meta["max_iter"], trace=trace, ctx=ctx, **kw)
    fmap = {str(k): [int(p) for p, _ in v] for k, v in od["latent_to_pathway_data"].items()}
    np.savez(os.path.join(outdir, "rank%d.npz" % ctx.rank), U=U, V=V, obj_parts=np.array(trace["obj_parts"]),
             sampled=np.array(trace["sampled"]), fmap=np.array(json.dumps(fmap)), obj=od["obj"])
    ctx.barrier()
    if ctx.world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
