"""GPU tests of the opt-in TF32 storage mode (x_dtype="tf32"; BASELINE config 5's tensor-core path).

This mode is NOT the parity mode: X, and the U / V operands of the two X products, are rounded to tf32
(10-bit mantissa, relative 4.9e-4 per element) and accumulated in fp32 on the tensor cores; the updates, the
objective and the reductions stay fp64.  Stated tolerance against the fp64 oracle run on the SAME rounded X:
U, V max relative error <= 1e-3 per element over a 10-step block (observed <= 3e-4), objective parts relative
<= 1e-3 (x 5 for instances with fewer than 64 rows or columns, where sums do not average the rounding out); factor -> pathway maps identical on the planted fixture.
"""
import contextlib
import io

import numpy as np
import pytest

from conftest import load_golden
from helpers import oracle_block, run_product, tf32_round

pytestmark = pytest.mark.gpu

RTOL_UV = 1e-3
RTOL_OBJ = 1e-3


def _instance(m, n, k, P, seed):
    from prmf_b200 import synth
    X, nodelist, Gs = synth.small_instance(m=m, n=n, k_true=min(3, P), n_pathways=P, pathway_size=12, seed=seed)
    rng = np.random.Generator(np.random.PCG64(seed + 100))
    U = 3 * (1 - rng.random((m, k)))
    V = 3 * (1 - rng.random((n, k)))
    active = [int(rng.integers(0, P)) for _ in range(k)]
    return X, nodelist, Gs, U, V, active


def _run(X_for_engine, X_shape_like, U, V, Gs, nodelist, active, steps, gamma, delta):
    from prmf_b200 import CudaEngine, nmf_manifold_vec_update, pack_pathways
    m, n = X_shape_like.shape
    with CudaEngine(m, m, n, V.shape[1], x_dtype="tf32") as eng:
        eng.set_X(X_for_engine)
        eng.set_pathways(pack_pathways(Gs, nodelist))
        with contextlib.redirect_stdout(io.StringIO()):
            Ug, Vg, od = nmf_manifold_vec_update(X_shape_like, U, V, Gs, active, n_steps=steps, gamma=gamma,
                                                 delta=delta, nodelist=nodelist, engine=eng)
        nx2 = eng.normX_sq
        r2 = eng.residual_sq()
    return Ug, Vg, od, nx2, r2


@pytest.mark.parametrize("m,n,k,P,steps", [
    (128, 64, 16, 4, 1),       # one tile, two K blocks
    (300, 700, 10, 8, 10),     # the benchmark's k (N = 16 with 6 zero columns), ragged tiles, a whole block
    (129, 257, 3, 5, 3),       # one row past a tile, one column past a K block
    (37, 131, 17, 5, 3),       # k just above a multiple of 16
    (64, 2100, 64, 6, 3),      # BASELINE config 4's k, several column chunks
    (500, 1500, 128, 6, 3),    # BASELINE config 5's k
    (5, 33, 1, 3, 2),          # tiny: sums of a few terms do not average the rounding out (tolerance x 5)
])
def test_tf32_steps_close_to_oracle_on_rounded_X(m, n, k, P, steps):
    scale = 5.0 if min(m, n) < 64 else 1.0
    X, nodelist, Gs, U, V, active = _instance(m, n, k, P, seed=m + n + k)
    Xr = tf32_round(X)
    Uo, Vo, parts_o, _, _, _ = oracle_block(Xr, U, V, Gs, nodelist, active, steps, 2.5, 0.3)
    Ug, Vg, od, nx2, r2 = _run(X, X, U, V, Gs, nodelist, active, steps, 2.5, 0.3)
    np.testing.assert_allclose(nx2, (Xr ** 2).sum(), rtol=1e-12)          # ||X||^2 of the stored values
    np.testing.assert_allclose(Ug, Uo, rtol=scale * RTOL_UV, atol=1e-9)
    np.testing.assert_allclose(Vg, Vo, rtol=scale * RTOL_UV, atol=1e-9)
    for key, col in (("recon", 0), ("manifold", 1), ("ignore", 2), ("fro", 3), ("obj", 4)):
        np.testing.assert_allclose(od[key], parts_o[-1, col], rtol=scale * RTOL_OBJ)
    # the explicit residual pass over the stored fp32 X agrees with U, V that came back (fp64 arithmetic)
    np.testing.assert_allclose(r2, np.linalg.norm(Xr - Ug @ Vg.T) ** 2, rtol=1e-10)


def test_tf32_input_flavours_agree_bitwise():
    """fp64 host, fp32 host and fp32 device inputs store the same rounded matrix."""
    import torch
    X, nodelist, Gs, U, V, active = _instance(200, 300, 6, 5, seed=3)
    X32 = X.astype(np.float32)
    ref = _run(X, X, U, V, Gs, nodelist, active, 2, 2.0, 0.5)
    for flavour in (X32, torch.from_numpy(X32).cuda(), torch.from_numpy(X).cuda(), np.asfortranarray(X)):
        got = _run(flavour, X, U, V, Gs, nodelist, active, 2, 2.0, 0.5)
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])


def test_tf32_config5_slice_one_outer_iteration():
    """BASELINE config 5 scaled down (k = 128, 40 pathways over 2048 genes, 4096 samples): one outer iteration
    (10 inner steps + score tables) against the oracle."""
    from prmf_b200 import CudaEngine, pack_pathways, synth
    from oracle import prmf_oracle as O
    m, n, k, P = 4096, 2048, 128, 40
    rng = np.random.Generator(np.random.PCG64(5))
    X = rng.random((m, n), dtype=np.float32)
    Gs = synth.random_pathway_graphs(rng, n, P)
    nodelist = list(range(n))
    U = 3 * (1 - rng.random((m, k)))
    V = 3 * (1 - rng.random((n, k)))
    active = [int(rng.integers(0, P)) for _ in range(k)]
    Xr = tf32_round(X)
    normX = np.linalg.norm(Xr)
    gamma, delta = normX / k, 10 / normX
    tables = O.PathwayTables(Gs, nodelist)
    Uo, Vo, odo, _, _ = O.update_block(Xr, U.copy(), V.copy(), tables, active, 10, gamma, delta)
    with CudaEngine(m, m, n, k, x_dtype="tf32") as eng:
        eng.set_X(X)
        eng.set_pathways(pack_pathways(Gs, nodelist))
        eng.set_UV(U, V)
        eng.set_active(active)
        parts, _, _ = eng.step(10, gamma, delta)
        mass, qn, qr = eng.scores()
        Ug, Vg = eng.get_UV()
    np.testing.assert_allclose(Ug, Uo, rtol=RTOL_UV, atol=1e-9)
    np.testing.assert_allclose(Vg, Vo, rtol=RTOL_UV, atol=1e-9)
    np.testing.assert_allclose(parts[-1, 4], odo["obj"], rtol=RTOL_OBJ)
    score = np.sqrt(mass[3, 7]) + 1 - qn[3, 7]
    np.testing.assert_allclose(score, O.score_match(tables, Vo[:, 3], 7), rtol=RTOL_OBJ)


def test_tf32_whole_loop_keeps_the_reference_assignments():
    """The planted fixture recorded from the unmodified reference: the pathways sampled every outer iteration and
    the final factor -> pathway map are the reference's in TF32 mode too; the rounding may move the convergence
    test (relative objective change < 1e-3) by one outer iteration.  Objective within 1e-3 of the reference's."""
    g = load_golden("small_planted")
    U, V, od, trace, _ = run_product(g, x_dtype="tf32")
    meta = g["meta"]
    common = min(len(trace["sampled"]), len(meta["sampled"]))
    assert abs(len(trace["sampled"]) - len(meta["sampled"])) <= 1
    assert trace["sampled"][:common] == meta["sampled"][:common]
    fm = {int(k): [p for p, _ in v] for k, v in meta["final_map"].items()}
    assert {k: [p for p, _ in v] for k, v in od["latent_to_pathway_data"].items()} == fm
    np.testing.assert_allclose(od["obj"], meta["final"]["obj"], rtol=1e-3)


def test_f32_input_needs_tf32_engine():
    from prmf_b200 import CudaEngine
    from prmf_b200._lib import PrmfLibraryError
    import ctypes
    with CudaEngine(8, 8, 16, 2) as eng:
        assert eng.lib.prmf_x_dtype(eng.h) == 0
        buf = np.zeros((8, 16), dtype=np.float32)
        rc = eng.lib.prmf_set_X_f32(eng.h, buf.ctypes.data_as(ctypes.c_void_p), 16, 0)
        assert rc != 0
        with pytest.raises(PrmfLibraryError):
            eng._ck(rc)
