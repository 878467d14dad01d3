"""Pathway graphs -> one block-diagonal packed CSR over local support indices.

Replaces the per-pathway n x n scipy matrices W, D, L and the support lists that the reference builds
at `script/prmf_runner.py:670-696`.  The layout is the one `prmf_set_pathways` takes
(include/prmf_b200.h): for P pathways with supports of s_p nodes and e_p undirected edges it stores
S = sum(s_p) support entries and E = sum(2 e_p - selfloops_p) weighted entries instead of 4P sparse
n x n matrices.
"""
import numpy as np


class PackedPathways:
    __slots__ = ("P", "n", "path_ptr", "support_idx", "row_ptr", "col_local", "w", "supports")

    def __init__(self, P, n, path_ptr, support_idx, row_ptr, col_local, w):
        self.P, self.n = P, n
        self.path_ptr = np.ascontiguousarray(path_ptr, dtype=np.int64)
        self.support_idx = np.ascontiguousarray(support_idx, dtype=np.int32)
        self.row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
        self.col_local = np.ascontiguousarray(col_local, dtype=np.int32)
        self.w = np.ascontiguousarray(w, dtype=np.float64)
        self.supports = [self.support_idx[self.path_ptr[p]:self.path_ptr[p + 1]] for p in range(P)]

    @property
    def S(self):
        return int(self.path_ptr[-1])

    @property
    def E(self):
        return int(self.row_ptr[-1])


def pack_pathways(Gs, nodelist):
    """Pack graphs (anything with `.nodes()` and `.edges(data=True)`, e.g. networkx.Graph).

    Semantics follow the reference with its pinned networkx 1.11:
      * nodes that are not in `nodelist` are dropped with their edges (`G.subgraph(nodelist)`, :670-671);
      * the support of a pathway is every remaining node, isolated ones included, in the graph's node
        order (:689-690);
      * an undirected edge contributes w to both (a,b) and (b,a); a self loop is counted once; the edge
        attribute 'weight' defaults to 1 (`nx.adjacency_matrix(G, nodelist)`, :679).
    Row entries are sorted by the neighbour's gene index, the order scipy's CSR product sums them in.
    """
    index = {g: i for i, g in enumerate(nodelist)}
    n = len(nodelist)
    # one pass over the graphs collects flat edge lists (Python only touches every node and edge once); everything
    # else -- mirroring, sorting rows by neighbour gene, row pointers -- is done for all pathways at once in numpy
    P = len(Gs)
    sizes = np.zeros(P, dtype=np.int64)
    supp_all, e_path, e_a, e_b, e_w = [], [], [], [], []
    get = index.get
    for p, G in enumerate(Gs):
        local = {}
        for g in G.nodes():
            i = get(g)
            if i is not None:
                local[g] = len(local)
                supp_all.append(i)
        sizes[p] = len(local)
        lget = local.get
        for a, b, d in G.edges(data=True):
            ia, ib = lget(a), lget(b)
            if ia is not None and ib is not None:
                e_path.append(p); e_a.append(ia); e_b.append(ib); e_w.append(d.get("weight", 1))
    path_ptr = np.zeros(P + 1, dtype=np.int64)
    np.cumsum(sizes, out=path_ptr[1:])
    support_idx = np.asarray(supp_all, dtype=np.int64)
    S = int(path_ptr[-1])
    if e_path:
        ep = np.asarray(e_path, dtype=np.int64)
        ea = np.asarray(e_a, dtype=np.int64); eb = np.asarray(e_b, dtype=np.int64)
        ew = np.asarray(e_w, dtype=np.float64)
        off = ea != eb                                           # an undirected edge counts both ways, a self loop once
        rows = np.concatenate([ea, eb[off]]) + np.concatenate([path_ptr[ep], path_ptr[ep[off]]])   # packed row ids
        cols = np.concatenate([eb, ea[off]])
        vals = np.concatenate([ew, ew[off]])
        base = np.concatenate([path_ptr[ep], path_ptr[ep[off]]])
        order = np.lexsort((support_idx[base + cols], rows))     # by packed row, then by the neighbour's gene index
        rows, cols, vals = rows[order], cols[order], vals[order]
        counts = np.bincount(rows, minlength=S)
    else:
        cols = np.zeros(0, dtype=np.int64); vals = np.zeros(0); counts = np.zeros(S, dtype=np.int64)
    row_ptr = np.zeros(S + 1, dtype=np.int64)
    np.cumsum(counts, out=row_ptr[1:])
    return PackedPathways(P, n, path_ptr, support_idx.astype(np.int32), row_ptr, cols.astype(np.int32), vals)
