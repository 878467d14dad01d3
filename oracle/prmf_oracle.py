"""CPU oracle for the PRMF hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy/scipy restatement of the reference's alternating factorisation loop
(`/root/reference/script/prmf_runner.py`).  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it, and only as the checker or the
timed CPU baseline; nothing under `prmf_b200/` does.

Parity status: PINNED.  `tests/test_oracle_golden.py` checks this file against fixtures recorded from
the unmodified reference (`tests/golden/*.npz`, written by `tests/golden/make_golden.py`) and, when
`/root/reference` is present, against the reference itself run side by side in the same process.

Every function cites the reference lines it restates.  The arithmetic deliberately keeps the
reference's evaluation order and cost structure (dense residual for `recon`, n x n scipy CSR matvecs,
normalised Laplacian rebuilt for every objective evaluation) so that timing it is a fair stand-in for
timing the reference on the same host; `cache_normalized=True` is the only shortcut and is off by
default.
"""
import math

import numpy as np
import scipy.sparse as sp
import networkx as nx

EPSILON = np.finfo(np.float32).eps          # prmf_runner.py:23
PERCENTILE = 19.9                           # prmf_runner.py:159


# ------------------------------------------------------------------------------------------------
# pathway tables  (prmf_runner.py:670-696)
# ------------------------------------------------------------------------------------------------
def adjacency(G, nodelist, index=None):
    """Symmetric n x n CSR adjacency of `G` over `nodelist` (`nx.adjacency_matrix` as the pinned
    networkx 1.11 behaves at prmf_runner.py:679): missing nodes are ignored, 'weight' defaults to 1,
    a self loop is counted once."""
    if index is None:
        index = {g: i for i, g in enumerate(nodelist)}
    n = len(nodelist)
    r, c, w = [], [], []
    for a, b, d in G.edges(data=True):
        if a in index and b in index:
            ww = d.get("weight", 1)
            r.append(index[a]); c.append(index[b]); w.append(ww)
            if a != b:
                r.append(index[b]); c.append(index[a]); w.append(ww)
    return sp.csr_matrix(sp.coo_matrix((np.asarray(w, dtype=np.float64), (r, c)), shape=(n, n)))


class PathwayTables:
    """Ws, Ds, Ls (n x n sparse), supports, and the normalised Laplacians (prmf_runner.py:673-696)."""

    def __init__(self, Gs, nodelist):
        n = len(nodelist)
        index = {g: i for i, g in enumerate(nodelist)}
        self.n = n
        self.Ws, self.Ds, self.Ls, self.supports = [], [], [], []
        for G in Gs:
            # :670-671  drop nodes (and incident edges) that are not in the nodelist
            nodes = [g for g in G.nodes() if g in index]
            W = adjacency(G, nodelist, index)
            deg = np.asarray(W.sum(axis=0)).ravel()                       # :680
            D = sp.dia_matrix((deg[None, :], np.array([0])), shape=(n, n)).tocsr()   # :681-682
            L = sp.csr_matrix(D - W)                                      # :683
            self.Ws.append(W); self.Ds.append(D); self.Ls.append(L)
            self.supports.append([index[g] for g in nodes])               # :689-690
        self.Lns = [normalize_laplacian(L, s) for L, s in zip(self.Ls, self.supports)]   # :693-696

    def __len__(self):
        return len(self.Ls)


def normalize_laplacian(L, support):
    """D^-1/2 L D^-1/2 with D = diag(L) on the support; zero diagonal -> 0 (prmf_runner.py:56-65)."""
    diag = L.diagonal()
    vals = np.zeros(L.shape[0])
    for ind in support:
        v = diag[ind]
        if v != 0:
            vals[ind] = v ** (-1 / 2)
    Dmh = sp.dia_matrix((vals[None, :], np.array([0])), shape=L.shape)
    return Dmh.dot(L.dot(Dmh))


# ------------------------------------------------------------------------------------------------
# objective and inner update  (prmf_runner.py:336-451, :497-554)
# ------------------------------------------------------------------------------------------------
def objective(X, U, V, tables, active, gamma, delta, cache_normalized=False):
    """recon (un-squared Frobenius norm), manifold, ignore, fro (squared) and their weighted sum
    (prmf_runner.py:336-372, `normal = True` branch)."""
    recon = np.linalg.norm(X - U.dot(V.transpose()))                      # :337
    manifold = 0.0
    ignore = 0.0
    for k, p in enumerate(active):
        v_unit = V[:, k] / np.linalg.norm(V[:, k])                        # :345
        support = tables.supports[p]
        Ln = tables.Lns[p] if cache_normalized else normalize_laplacian(tables.Ls[p], support)  # :349
        manifold += Ln.dot(v_unit).dot(v_unit)                            # :350
        ignore += np.sum(np.power(v_unit[support] + 1, -1))               # :352
    fro = np.sum(np.multiply(U, U))                                       # :359
    obj = recon + gamma * manifold + delta * ignore + fro                 # :362
    return {"recon": recon, "manifold": manifold, "ignore": ignore, "fro": fro,
            "gamma": gamma, "delta": delta, "obj": obj}


def update_step(X, U, V, tables, active, gamma, delta):
    """One multiplicative update of U then V (prmf_runner.py:420-444).  Returns new (U, V)."""
    n, k_latent = V.shape
    num = X.dot(V)                                                        # :420
    den = U.dot(V.transpose().dot(V)) + U                                 # :421
    U = np.multiply(U, np.divide(num, den, out=np.ones_like(num), where=den != 0))   # :422  0/0 := 1
    B = X.transpose().dot(U)                                              # :424
    C = V.dot(U.transpose().dot(U))                                       # :425
    n_man = np.zeros((n, k_latent)); d_man = np.zeros((n, k_latent)); n_ign = np.zeros((n, k_latent))
    for k, p in enumerate(active):                                        # :431-438
        n_man[:, k] = gamma * tables.Ws[p].dot(V[:, k])
        d_man[:, k] = gamma * tables.Ds[p].dot(V[:, k])
        s = tables.supports[p]
        n_ign[s, k] = delta * np.power(V[s, k] + 1, -2)
    v_num = B + (n_man + n_ign)                                           # :440
    v_den = C + d_man                                                     # :441
    v_den[v_den < EPSILON] = EPSILON                                      # :442
    V = np.multiply(V, np.divide(v_num, v_den, out=np.ones_like(v_num), where=v_den != 0))   # :443
    V[V < EPSILON] = EPSILON                                              # :444
    return U, V


def update_block(X, U, V, tables, active, n_steps=10, gamma=1.0, delta=1.0, tradeoff=None,
                 log=None, cache_normalized=False):
    """`n_steps` inner steps with fixed active pathways (prmf_runner.py:374-451); with `tradeoff`,
    gamma = delta = (1-t)*recon/(t*manifold) is fed back after every step (:497-554).
    Returns U, V, obj_data, gamma, delta; appends every step's obj_data to `log` if given."""
    obj_data = None
    for _ in range(n_steps):
        U, V = update_step(X, U, V, tables, active, gamma, delta)
        obj_data = objective(X, U, V, tables, active, gamma, delta, cache_normalized)   # :446
        if log is not None:
            log.append(obj_data)
        if tradeoff is not None:                                          # :542-548
            den = tradeoff * obj_data["manifold"]
            gamma = 1 if den == 0 else ((1 - tradeoff) * obj_data["recon"]) / den
            delta = gamma
    return U, V, obj_data, gamma, delta


# ------------------------------------------------------------------------------------------------
# candidate scoring, pruning, matching  (prmf_runner.py:115-258)
# ------------------------------------------------------------------------------------------------
def score_match(tables, v, p):
    """sqrt(mass of unit v on the support) + (1 - unit-v quadratic form of the normalised Laplacian)
    (prmf_runner.py:115-127)."""
    v_unit = v / np.linalg.norm(v)
    support = tables.supports[p]
    mass = np.sqrt(np.sum(np.power(v_unit[support], 2)))
    manifold = 1 - tables.Lns[p].dot(v_unit).dot(v_unit)
    return mass + manifold


def restrict(V, tables, cands):
    """Keep, per factor, the candidates scoring above the 19.9th percentile (prmf_runner.py:129-194)."""
    out = {}
    for k in range(V.shape[1]):
        data = cands[k]
        if len(data) > 1:
            ids = [p for p, _ in data]
            scores = np.array([score_match(tables, V[:, k], p) for p in ids])
            keep = np.where(scores > np.percentile(scores, PERCENTILE))[0]       # :171
            if len(keep) == 0:
                # :173-183 draws ceil(n*(1-19.9)/100) < 0 samples with np.random.choice -> ValueError
                raise ValueError("restrict: no candidate above the percentile (reference raises here too)")
            out[k] = [(ids[i], scores[i]) for i in keep]
        else:
            out[k] = data
    return out


def force_distinct(V, tables, cands, active, gamma, delta):
    """Max-weight bipartite matching of factors to remaining candidates (prmf_runner.py:209-258).
    The ignore penalty is, as in the reference (:234-235 overwrites instead of accumulating), that of
    the LAST factor only and therefore the same constant on every edge."""
    k_latent = V.shape[1]
    ign = 0
    for k2 in range(k_latent):
        s = tables.supports[active[k2]]
        ign = np.sum(np.power(V[s, k2] + 1, -1))
    G = nx.Graph()
    for k in cands:
        for p, _ in cands[k]:
            man = tables.Ls[p].dot(V[:, k]).dot(V[:, k])                  # :232
            den = gamma * man + delta * ign                               # :237
            G.add_edge("k%d" % k, "l%d" % p, weight=0 if den == 0 else 1 / den)
    for a, b in nx.max_weight_matching(G):                                # :246
        kn, ln = (a, b) if a[0] == "k" else (b, a)
        cands[int(kn[1:])] = [(int(ln[1:]), 2)]                           # :257
    return cands


def find_mins(V, Ls):
    """argmin_p v_k^T L_p v_k, first minimum wins (prmf_runner.py:37-54; dead code in the reference)."""
    out = -1 * np.ones(V.shape[1])
    for k in range(V.shape[1]):
        v = V[:, k]
        best = np.inf
        for i, L in enumerate(Ls):
            pen = L.dot(v).dot(v)
            if pen < best:
                best, out[k] = pen, i
    return out


# ------------------------------------------------------------------------------------------------
# driver  (prmf_runner.py:556-792)
# ------------------------------------------------------------------------------------------------
def nmf_pathway(X, Gs, gamma=1.0, delta=1.0, tradeoff=None, k_latent=6, tol=1e-3, max_iter=1000,
                nodelist=None, modulus=10, U_init=None, V_init=None, trace=None,
                cache_normalized=False):
    """Whole alternating loop; consumes the global legacy NumPy RNG exactly like the reference."""
    norm_X = np.linalg.norm(X)                                            # :640
    gamma = gamma * norm_X / k_latent                                     # :644
    delta = delta * 10 / norm_X                                           # :645
    m, n = X.shape
    U = 3 * (1 - np.random.rand(m, k_latent)) if U_init is None else U_init   # :650-653
    V = 3 * (1 - np.random.rand(n, k_latent)) if V_init is None else V_init
    if U.shape != (m, k_latent) or V.shape != (n, k_latent):
        raise ValueError("invalid U_init / V_init shape")                 # :656-659
    tables = PathwayTables(Gs, nodelist)
    P = len(tables)
    cands = {k: [(p, 1) for p in range(P)] for k in range(k_latent)}     # :700
    converged, candidates_remain = False, True
    obj, i = math.inf, 0
    obj_data = {}
    best = {"obj": np.inf}
    while i < max_iter and (candidates_remain or not converged):          # :715
        active = []
        for k in range(k_latent):                                         # :717-729
            ids = [p for p, _ in cands[k]]
            scores = np.array([s for _, s in cands[k]])
            with np.errstate(divide="raise"):
                prob = scores / np.sum(scores)
            draw = np.random.multinomial(1, prob)      # == scipy.stats.multinomial.rvs(1, prob)
            active.append(ids[int(np.where(draw != 0)[0][0])])
        log = [] if trace is not None else None
        U, V, obj_data, g2, d2 = update_block(X, U, V, tables, active, modulus, gamma, delta,
                                              tradeoff, log, cache_normalized)
        if tradeoff is not None:
            gamma, delta = g2, d2
        i += modulus
        if obj_data["obj"] < best["obj"]:                                 # :747-750
            best = {"obj": obj_data["obj"], "U": U, "V": V, "obj_data": obj_data}
        kind = None
        if candidates_remain:                                             # :754-768
            if min(len(v) for v in cands.values()) <= k_latent:
                cands = force_distinct(V, tables, cands, active, gamma, delta)
                candidates_remain = False
                kind = "force"
            else:
                cands = restrict(V, tables, cands)
                candidates_remain = any(len(v) > 1 for v in cands.values())
                kind = "restrict"
        if trace is not None:
            trace.setdefault("sampled", []).append(list(active))
            trace.setdefault("obj_parts", []).extend(
                [[d["recon"], d["manifold"], d["ignore"], d["fro"], d["obj"]] for d in log])
            trace.setdefault("blocks", []).append((U, V))
            if kind:
                trace.setdefault("cands", []).append(
                    {"kind": kind, "data": {k: list(v) for k, v in cands.items()}})
        prev_obj, obj = obj, obj_data["obj"]
        converged = abs(obj - prev_obj) / obj < tol                       # :772-774
    if best["obj"] < obj_data["obj"]:                                     # :778-782
        U, V, obj_data = best["U"], best["V"], best["obj_data"]
    obj_data = dict(obj_data)
    obj_data["latent_to_pathway_data"] = cands
    return U, V, obj_data
