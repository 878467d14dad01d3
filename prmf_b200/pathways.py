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
    path_ptr = [0]
    row_ptr = [0]
    supp_all, col_all, w_all = [], [], []
    for G in Gs:
        nodes = [g for g in G.nodes() if g in index]
        local = {g: i for i, g in enumerate(nodes)}
        gidx = np.fromiter((index[g] for g in nodes), dtype=np.int64, count=len(nodes))
        rows, cols, vals = [], [], []
        for a, b, d in G.edges(data=True):
            if a in local and b in local:
                ww = d.get("weight", 1)
                ia, ib = local[a], local[b]
                rows.append(ia); cols.append(ib); vals.append(ww)
                if ia != ib:
                    rows.append(ib); cols.append(ia); vals.append(ww)
        s = len(nodes)
        if rows:
            rows = np.asarray(rows, dtype=np.int64)
            cols = np.asarray(cols, dtype=np.int64)
            vals = np.asarray(vals, dtype=np.float64)
            order = np.lexsort((gidx[cols], rows))          # by row, then by neighbour gene index
            rows, cols, vals = rows[order], cols[order], vals[order]
            counts = np.bincount(rows, minlength=s)
        else:
            cols = np.zeros(0, dtype=np.int64); vals = np.zeros(0); counts = np.zeros(s, dtype=np.int64)
        base = row_ptr[-1]
        row_ptr.extend((base + np.cumsum(counts)).tolist())
        supp_all.append(gidx); col_all.append(cols); w_all.append(vals)
        path_ptr.append(path_ptr[-1] + s)
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dtype=dt)
    return PackedPathways(len(Gs), n, path_ptr, cat(supp_all, np.int32), row_ptr,
                          cat(col_all, np.int32), cat(w_all, np.float64))
