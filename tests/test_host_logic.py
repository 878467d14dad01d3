"""CPU tests of the host-side logic: pathway packing, candidate bookkeeping, percentile, C-ABI symbols."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden


def test_library_exports_every_header_symbol():
    """libprmf_b200.so loads without a GPU and exports every function include/prmf_b200.h declares."""
    from prmf_b200 import _lib
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "prmf_b200.h")).read()
    declared = set(re.findall(r"\b(prmf_[a-z_A-Z0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), "library does not export %s" % name
    assert declared == set(_lib.SYMBOLS), "ctypes table and header disagree: %s" % (declared ^ set(_lib.SYMBOLS))
    assert lib.prmf_abi_version() == 1


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from prmf_b200 import CudaEngine
    from prmf_b200._lib import PrmfLibraryError
    with pytest.raises(PrmfLibraryError, match="no CPU fallback|no CUDA"):
        CudaEngine(10, 10, 20, 3)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "prmf_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f


@pytest.mark.parametrize("n", [2, 3, 5, 6, 7, 11, 37, 100, 240, 299, 300, 301, 2000])
def test_percentile_matches_numpy_bitwise(n):
    from prmf_b200.solver import percentile_19_9
    rng = np.random.Generator(np.random.PCG64(n))
    for trial in range(50):
        x = rng.random(n) * (10.0 ** rng.integers(-3, 4))
        if trial % 5 == 0:
            x = np.round(x, 1)                      # many ties
        if trial % 7 == 0:
            x[:] = x[0]                             # all equal
        assert percentile_19_9(x) == np.percentile(x, 19.9)


def test_pack_pathways_matches_oracle_tables():
    from oracle import prmf_oracle as O
    from prmf_b200 import pack_pathways, synth
    X, nodelist, Gs = synth.small_instance(m=10, n=150, k_true=3, n_pathways=12, pathway_size=15, seed=4,
                                           weighted=True)
    packed = pack_pathways(Gs, nodelist)
    tables = O.PathwayTables(Gs, nodelist)
    assert packed.P == len(tables)
    for p in range(packed.P):
        supp = packed.supports[p]
        assert list(supp) == tables.supports[p]
        W = np.zeros((len(nodelist), len(nodelist)))
        beg = packed.path_ptr[p]
        for r in range(beg, packed.path_ptr[p + 1]):
            cols = packed.col_local[packed.row_ptr[r]:packed.row_ptr[r + 1]]
            gcols = supp[cols]
            assert np.all(np.diff(gcols) > 0), "row entries must be sorted by gene index"
            W[packed.support_idx[r], gcols] = packed.w[packed.row_ptr[r]:packed.row_ptr[r + 1]]
        np.testing.assert_array_equal(W, tables.Ws[p].toarray())


def test_pack_drops_nodes_outside_nodelist_and_keeps_isolated():
    import networkx as nx
    from prmf_b200 import pack_pathways
    G = nx.Graph()
    G.add_edge("a", "b"); G.add_edge("b", "zzz"); G.add_node("c"); G.add_edge("a", "a", weight=0.5)
    packed = pack_pathways([G], ["c", "b", "a", "d"])
    assert list(packed.supports[0]) == [2, 1, 0]          # graph node order a, b, c -> gene indices
    rows = {int(packed.support_idx[r]): (packed.col_local[packed.row_ptr[r]:packed.row_ptr[r + 1]].tolist(),
                                         packed.w[packed.row_ptr[r]:packed.row_ptr[r + 1]].tolist())
            for r in range(packed.S)}
    assert rows[0] == ([], [])                            # isolated node stays in the support
    assert rows[1] == ([0], [1.0])                        # b - a only; 'zzz' dropped
    assert rows[2] == ([1, 0], [1.0, 0.5])                # a: neighbour b (gene 1) then self loop (gene 2)


def test_sample_active_matches_scipy_multinomial_stream():
    from scipy.stats import multinomial
    from prmf_b200.solver import sample_active
    rng = np.random.Generator(np.random.PCG64(3))
    cands = {k: [(int(p), float(s)) for p, s in zip(rng.permutation(40)[:n], rng.random(n) + 0.1)]
             for k, n in enumerate([40, 17, 1, 5])}
    np.random.seed(7)
    mine = sample_active(cands, 4)
    tail_mine = np.random.rand()
    np.random.seed(7)
    ref = []
    for k in range(4):
        ids = [p for p, _ in cands[k]]
        scores = np.array([s for _, s in cands[k]])
        draw = multinomial.rvs(1, scores / np.sum(scores))
        ref.append(ids[np.where(draw != 0)[0][0]])
    assert mine == ref
    assert tail_mine == np.random.rand()                   # same amount of RNG state consumed


def test_restrict_and_force_from_tables_match_oracle():
    from oracle import prmf_oracle as O
    from prmf_b200.solver import force_distinct_from_tables, restrict_from_tables
    from prmf_b200 import pack_pathways
    g = load_golden("kernel_vectors")
    tables = O.PathwayTables(g["Gs"], g["nodelist"])
    V = g["V"]
    K, P = g["score"].shape
    mass = np.array([[np.sum((V[tables.supports[p], k] / np.linalg.norm(V[:, k])) ** 2) for p in range(P)]
                     for k in range(K)])
    cands = {k: [(p, 1) for p in range(P)] for k in range(K)}
    got = restrict_from_tables(mass, g["quad_norm"], cands)
    for k, lst in g["meta"]["restricted"].items():
        assert [p for p, _ in got[int(k)]] == [p for p, _ in lst]
        np.testing.assert_allclose([s for _, s in got[int(k)]], [s for _, s in lst], rtol=1e-12)
    packed = pack_pathways(g["Gs"], g["nodelist"])
    active = [0, 1, 2, 3]
    small = {k: [(p, 1.0) for p in range(3)] for k in range(K)}
    mine = force_distinct_from_tables(g["quad_raw"], V, packed.supports, active, {k: list(v) for k, v in small.items()}, 2.0, 0.5)
    ref = O.force_distinct(V, tables, {k: list(v) for k, v in small.items()}, active, 2.0, 0.5)
    assert {k: [p for p, _ in v] for k, v in mine.items()} == {k: [p for p, _ in v] for k, v in ref.items()}


def test_native_restrict_is_bit_identical_to_numpy():
    """prmf_host_restrict_batch (plain C++ in libprmf_b200.so, no GPU involved) against the numpy expressions of
    restrict (:123-125, :159, :171): same survivors, same scores, bit for bit -- including ties at the threshold,
    tiny candidate lists and repeated pruning."""
    from prmf_b200.solver import _CandArrays, init_latent_to_pathway_data, restrict_from_tables
    rng = np.random.Generator(np.random.PCG64(0))
    for k, P in ((10, 300), (3, 7), (64, 2000), (1, 2), (5, 41)):
        mass = rng.random((k, P))
        qn = rng.random((k, P))
        dup = np.arange(0, P - 1, 5)                                    # duplicates -> ties around the percentile
        mass[:, dup] = mass[:, dup + 1]
        qn[:, dup] = qn[:, dup + 1]
        a = b = init_latent_to_pathway_data(k, P)
        for _ in range(6):
            try:
                a = restrict_from_tables(mass, qn, a, native=False)
            except ValueError:                       # the duplicates left a factor with equal scores only
                with pytest.raises(ValueError):
                    restrict_from_tables(mass, qn, b, native=True)
                break
            b = restrict_from_tables(mass, qn, b, native=True)
            assert list(a) == list(b)
            for f in a:
                assert np.array_equal(np.asarray(a[f].ids), np.asarray(b[f].ids))
                assert np.array_equal(np.asarray(a[f].scores), np.asarray(b[f].scores))
            if all(len(v) <= 2 for v in a.values()):
                break
    # a factor with a single candidate is passed through; equal scores raise as the reference does
    c = {0: _CandArrays(np.array([3]), np.array([1.0])), 1: _CandArrays(np.arange(4), np.ones(4))}
    flat = np.full((2, 6), 0.25)
    for native in (False, True):
        with pytest.raises(ValueError):
            restrict_from_tables(flat, flat, c, native=native)
    # non-contiguous / non-float64 tables take the numpy route and agree
    m32 = rng.random((4, 30)).astype(np.float32)
    c4 = init_latent_to_pathway_data(4, 30)
    x = restrict_from_tables(m32, m32, c4)
    y = restrict_from_tables(m32.astype(np.float64), m32.astype(np.float64), c4)
    for f in x:
        assert np.array_equal(np.asarray(x[f].ids), np.asarray(y[f].ids))


def test_sample_active_raises_like_seterr_divide_raise():
    """np.seterr(divide='raise') (:22): all-zero scores of a factor make the reference's `scores / sum` raise --
    also for a factor with a single candidate, whose draw is otherwise skipped."""
    from prmf_b200.solver import sample_active
    np.random.seed(0)
    for cands in ({0: [(3, 0.0)]}, {0: [(1, 0.0), (2, 0.0)]}):
        with pytest.raises(FloatingPointError):
            sample_active(cands, 1)
    assert sample_active({0: [(5, 2.0)], 1: [(7, 1e-300)]}, 2) == [5, 7]
