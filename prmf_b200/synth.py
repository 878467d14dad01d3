"""Synthetic PRMF instances: the reference's own test instance and recount2-shape data.

All generators are seeded and independent of the networkx / scipy version except `test1_instance`,
which follows the reference test byte for byte (legacy `np.random.seed` + `scipy.stats.gamma.rvs`).
"""
import numpy as np
import networkx as nx


def test1_instance(unmeasured=False):
    """The instance of reference `test/script/PRMF/test_inferred_nodelist_1/test_inferred_nodelist_1.py:9-45`
    (100 samples x 1000 genes, k=6, six path graphs over each true factor's top-5% genes).
    `unmeasured=True` adds one graph-only gene per pathway as in `test_inferred_nodelist_2.py:28-54`
    (X is zero-padded for those genes, the job of `prmf.embed_arr`).
    Returns X (m x n), nodelist (list[str]), Gs (list[nx.Graph] keyed by gene name)."""
    from scipy.stats import gamma
    state = np.random.get_state()
    try:
        np.random.seed(seed=1)
        m, n, k = 100, 1000, 6
        U = gamma.rvs(5, size=m * k).reshape(m, k)
        V = gamma.rvs(5, size=n * k).reshape(n, k)
    finally:
        np.random.set_state(state)
    X = U.dot(V.transpose())
    nodelist = ["ENSP%d" % i for i in range(n)]
    Gs = []
    for kk in range(k):
        inds = np.where(V[:, kk] > np.percentile(V[:, kk], 95))[0]
        G = nx.Graph()
        names = ["ENSP%d" % j for j in inds]
        G.add_nodes_from(names)
        for a, b in zip(names[:-1], names[1:]):
            G.add_edge(a, b)
        if unmeasured:
            G.add_node("ENSP%d" % (n + kk))
        Gs.append(G)
    if unmeasured:
        nodelist = nodelist + ["ENSP%d" % (n + kk) for kk in range(k)]
        X = np.concatenate([X, np.zeros((m, k))], axis=1)
    return X, nodelist, Gs


def random_pathway_edges(rng, n_genes, size, extra_edges=None, weighted=False):
    """One random pathway: `size` distinct genes, a random recursive tree plus `extra_edges`
    uniformly random distinct pairs (mean degree ~4 when extra_edges == size).  Returns
    (genes int64[size], edges int64[e,2] over gene ids, weights float64[e])."""
    genes = rng.choice(n_genes, size=size, replace=False)
    pairs = set()
    for i in range(1, size):
        j = int(rng.integers(0, i))
        pairs.add((j, i))
    if extra_edges is None:
        extra_edges = size
    tries = 0
    while extra_edges > 0 and tries < 20 * size and size > 2:
        a, b = (int(x) for x in rng.integers(0, size, size=2))
        tries += 1
        if a == b:
            continue
        key = (min(a, b), max(a, b))
        if key in pairs:
            continue
        pairs.add(key)
        extra_edges -= 1
    pairs = np.array(sorted(pairs), dtype=np.int64).reshape(-1, 2)
    w = rng.uniform(0.2, 1.0, size=len(pairs)) if weighted else np.ones(len(pairs))
    return genes, genes[pairs], w


def random_pathway_graphs(rng, n_genes, n_pathways, median_size=80, sigma=0.6, lo=10, hi=400,
                          weighted=False, nodelist=None):
    """KEGG-size random pathway graphs: sizes ~ clip(lognormal(ln median, sigma))."""
    Gs = []
    hi = min(hi, n_genes)
    lo = min(lo, hi)
    for _ in range(n_pathways):
        size = int(np.clip(round(rng.lognormal(np.log(median_size), sigma)), lo, hi))
        genes, edges, w = random_pathway_edges(rng, n_genes, size, weighted=weighted)
        G = nx.Graph()
        name = (lambda g: nodelist[g]) if nodelist is not None else (lambda g: int(g))
        G.add_nodes_from(name(g) for g in genes)
        for (a, b), ww in zip(edges, w):
            if weighted:
                G.add_edge(name(a), name(b), weight=float(ww))
            else:
                G.add_edge(name(a), name(b))
        Gs.append(G)
    return Gs


def small_instance(m=60, n=300, k_true=4, n_pathways=24, pathway_size=30, seed=0, weighted=False,
                   noise=0.05, dangling=True):
    """A small planted instance: the first `k_true` pathways carry a rank-1 bump on their genes.
    With `dangling`, two pathways get an isolated node and one gets a self loop, and one pathway
    references a gene that is absent from the nodelist (exercises prmf_runner.py:670-671)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    nodelist = ["g%d" % i for i in range(n)]
    Gs = random_pathway_graphs(rng, n, n_pathways, median_size=pathway_size, sigma=0.3,
                               lo=5, hi=max(6, n // 3), weighted=weighted, nodelist=nodelist)
    U = rng.gamma(2.0, size=(m, k_true))
    V = rng.gamma(1.0, size=(n, k_true)) * 0.3
    for kk in range(k_true):
        for g in Gs[kk].nodes():
            V[nodelist.index(g), kk] += 3.0 + rng.gamma(2.0)
    X = U.dot(V.T) + noise * rng.uniform(size=(m, n))
    if dangling and n_pathways >= 3:
        Gs[-1].add_node(nodelist[int(rng.integers(0, n))])
        Gs[-2].add_node(nodelist[int(rng.integers(0, n))])
        g0 = list(Gs[-3].nodes())[0]
        Gs[-3].add_edge(g0, g0)
        Gs[-2].add_edge(list(Gs[-2].nodes())[0], "not_in_nodelist")
    return X, nodelist, Gs


def recount2_shape(m=37032, n=6750, n_pathways=300, seed=0, plant=10, dtype=np.float64):
    """C2/C4 synthetic instance: X iid U(0,1) (the distribution `quantile_transform` outputs), integer
    node ids, optional rank-1 bumps on the first `plant` pathways so assignments are not degenerate."""
    rng = np.random.Generator(np.random.PCG64(seed))
    X = rng.random((m, n), dtype=np.float64)
    nodelist = list(range(n))
    Gs = random_pathway_graphs(rng, n, n_pathways)
    for p in range(min(plant, n_pathways)):
        genes = np.fromiter(Gs[p].nodes(), dtype=np.int64)
        u = rng.random(m) * 0.5
        X[:, genes] += u[:, None]
    if dtype != np.float64:
        X = X.astype(dtype)
    return X, nodelist, Gs
