"""Developer tool: accuracy and speed of the TF32 tensor-core X streams (x_dtype="tf32") on one GPU.
    python tools/tf32_check.py [--speed]"""
import contextlib, io, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prmf_b200 import CudaEngine, nmf_manifold_vec_update, pack_pathways, synth
from oracle import prmf_oracle as O


def tf32_round(a):
    """cvt.rna.tf32.f32 on the host: fp32, round to nearest (ties away) to 10 mantissa bits."""
    b = np.asarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    b = ((b + 0x1000) & 0xFFFFE000).astype(np.uint32)
    return b.view(np.float32).astype(np.float64)


def accuracy(m, n, k, P, steps=1, seed=0):
    X, nodelist, Gs = synth.small_instance(m=m, n=n, k_true=min(3, P), n_pathways=P, pathway_size=12, seed=seed)
    rng = np.random.Generator(np.random.PCG64(seed + 100))
    U = 3 * (1 - rng.random((m, k))); V = 3 * (1 - rng.random((n, k)))
    active = [int(rng.integers(0, P)) for _ in range(k)]
    Xr = tf32_round(X)
    tables = O.PathwayTables(Gs, nodelist)
    Uo, Vo, odo, _, _ = O.update_block(Xr, U.copy(), V.copy(), tables, active, steps, 2.5, 0.3)
    eng = CudaEngine(m, m, n, k, x_dtype="tf32")
    eng.set_X(X); eng.set_pathways(pack_pathways(Gs, nodelist))
    with contextlib.redirect_stdout(io.StringIO()):
        Ug, Vg, od = nmf_manifold_vec_update(X, U, V, Gs, active, n_steps=steps, gamma=2.5, delta=0.3,
                                             nodelist=nodelist, engine=eng)
    nx2 = eng.normX_sq
    eng.close()
    eu = np.max(np.abs(Ug - Uo) / (np.abs(Uo) + 1e-300)); ev = np.max(np.abs(Vg - Vo) / (np.abs(Vo) + 1e-300))
    print("m=%d n=%d k=%d steps=%d: U rel %.2e  V rel %.2e  obj rel %.2e recon rel %.2e  normX rel %.1e" % (
        m, n, k, steps, eu, ev, abs(od["obj"] - odo["obj"]) / abs(odo["obj"]),
        abs(od["recon"] - odo["recon"]) / abs(odo["recon"]), abs(nx2 - (Xr ** 2).sum()) / (Xr ** 2).sum()))
    return eu, ev


def speed(m, n, k, P, steps=10, dtype="tf32"):
    rng = np.random.Generator(np.random.PCG64(0))
    Gs = synth.random_pathway_graphs(rng, n, P)
    X = torch.rand((m, n), dtype=torch.float32 if dtype == "tf32" else torch.float64, device="cuda")
    eng = CudaEngine(m, m, n, k, x_dtype=dtype)
    eng.set_X(X); del X
    eng.set_pathways(pack_pathways(Gs, list(range(n))))
    eng.set_UV(3 * (1 - rng.random((m, k))), 3 * (1 - rng.random((n, k)))); eng.set_active(list(range(k)))
    eng.step(3, 900.0, 1e-3)
    eng.set_profiling(True); eng.kernel_times(reset=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    eng.step(steps, 900.0, 1e-3)
    dt = (time.perf_counter() - t0) / steps
    kt = eng.kernel_times(reset=True)
    eng.close()
    sz = 4 if dtype == "tf32" else 8
    ph = {a: 1e3 * t / max(1, c) for a, (t, c) in kt.items()}
    print("%s m=%d n=%d k=%d: %.3f ms/step | xv %.0f us (%.0f GB/s) xtu %.0f us (%.0f GB/s) u %.0f v %.0f" % (
        dtype, m, n, k, dt * 1e3, ph["xv"], m * n * sz / ph["xv"] / 1e3, ph["xtu"], m * n * sz / ph["xtu"] / 1e3,
        ph["u_update"], ph["v_update"]))


if __name__ == "__main__":
    if "--only128" in sys.argv:
        speed(65536, 20000, 128, 300, steps=2)
        sys.exit(0)
    for shape in ((128, 64, 16, 4), (300, 700, 10, 8), (129, 257, 3, 5), (64, 2100, 64, 6), (500, 1500, 128, 6), (37, 131, 17, 5)):
        accuracy(*shape)
    accuracy(300, 700, 10, 8, steps=5)
    if "--speed" in sys.argv:
        speed(37032, 6750, 10, 300)
        speed(37032, 6750, 64, 300)
        speed(65536, 20000, 128, 300, steps=5)
