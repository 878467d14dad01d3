"""Developer tool: speed of the TF32 tensor-core X streams (x_dtype="tf32") on one GPU, per phase of the inner step
(accuracy against the oracle lives in tests/test_tf32.py).
    python tools/tf32_check.py [--only128]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prmf_b200 import CudaEngine, pack_pathways, synth


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
    speed(37032, 6750, 10, 300)
    speed(37032, 6750, 64, 300)
    speed(65536, 20000, 128, 300, steps=5)
