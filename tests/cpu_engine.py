"""A numpy stand-in for `prmf_b200.engine.CudaEngine`, used ONLY by the CPU tests of the host logic
(row sharding, RNG consistency across ranks, candidate bookkeeping, gathers) where no GPU exists.
It follows the same sharded algorithm as the device engine -- local X.V, local U update, partial
X^T U / U^T U summed over ranks, replicated V update, pass-free objective -- with the oracle's formulas.
Test infrastructure: never imported by the product."""
import numpy as np
import scipy.sparse as sp

from oracle import prmf_oracle as O


class _Tables:
    """The oracle's PathwayTables rebuilt from the packed arrays the product hands to the engine."""

    def __init__(self, packed):
        n = packed.n
        self.Ws, self.Ds, self.Ls, self.supports = [], [], [], []
        for p in range(packed.P):
            beg, end = packed.path_ptr[p], packed.path_ptr[p + 1]
            supp = packed.support_idx[beg:end].astype(np.int64)
            r, c, w = [], [], []
            for row in range(beg, end):
                e0, e1 = packed.row_ptr[row], packed.row_ptr[row + 1]
                r += [int(packed.support_idx[row])] * int(e1 - e0)
                c += supp[packed.col_local[e0:e1]].tolist()
                w += packed.w[e0:e1].tolist()
            W = sp.csr_matrix(sp.coo_matrix((np.asarray(w, dtype=np.float64), (r, c)), shape=(n, n)))
            deg = np.asarray(W.sum(axis=0)).ravel()
            D = sp.dia_matrix((deg[None, :], np.array([0])), shape=(n, n)).tocsr()
            self.Ws.append(W); self.Ds.append(D); self.Ls.append(sp.csr_matrix(D - W))
            self.supports.append(supp.tolist())
        self.Lns = [O.normalize_laplacian(L, s) for L, s in zip(self.Ls, self.supports)]

    def __len__(self):
        return len(self.Ls)


class NumpyShardEngine:
    def __init__(self, m_local, m_global, n, k, ctx):
        self.m, self.m_global, self.n, self.k, self.ctx = m_local, m_global, n, k, ctx
        self.P = 0
        self.launch_count = 0

    def close(self):
        pass

    def set_X(self, X):
        self.X = np.array(X, dtype=np.float64)
        self._normX_sq = float(self.ctx.all_reduce_sum(np.array([np.sum(self.X * self.X)]))[0])

    @property
    def normX_sq(self):
        return self._normX_sq

    def set_pathways(self, packed):
        self.tables = _Tables(packed)
        self.P = packed.P

    def set_UV(self, U=None, V=None):
        if U is not None:
            self.U = np.array(U, dtype=np.float64)
        if V is not None:
            self.V = np.array(V, dtype=np.float64)

    def get_UV(self, want_U=True, want_V=True):
        return (self.U.copy() if want_U else None), (self.V.copy() if want_V else None)

    def set_active(self, active):
        self.active = [int(a) for a in active]

    def step(self, n_steps, gamma, delta, tradeoff=None):
        parts = np.zeros((n_steps, 8))
        n, k = self.n, self.k
        T = self.tables
        for s in range(n_steps):
            U, V = self.U, self.V
            num = self.X.dot(V)
            den = U.dot(V.T.dot(V)) + U
            U = U * np.divide(num, den, out=np.ones_like(num), where=den != 0)
            packed = np.concatenate([self.X.T.dot(U).ravel(), U.T.dot(U).ravel()])
            packed = self.ctx.all_reduce_sum(packed)
            B, Gu = packed[:n * k].reshape(n, k), packed[n * k:].reshape(k, k)
            C = V.dot(Gu)
            n_man = np.zeros((n, k)); d_man = np.zeros((n, k)); n_ign = np.zeros((n, k))
            for c, p in enumerate(self.active):
                n_man[:, c] = gamma * T.Ws[p].dot(V[:, c])
                d_man[:, c] = gamma * T.Ds[p].dot(V[:, c])
                sidx = T.supports[p]
                n_ign[sidx, c] = delta * np.power(V[sidx, c] + 1, -2)
            v_num = B + (n_man + n_ign)
            v_den = C + d_man
            v_den[v_den < O.EPSILON] = O.EPSILON
            V = V * (v_num / v_den)
            V[V < O.EPSILON] = O.EPSILON
            self.U, self.V = U, V
            r2 = self._normX_sq - 2 * np.sum(V * B) + np.sum(Gu * V.T.dot(V))
            recon = np.sqrt(max(r2, 0.0))
            man = ign = 0.0
            for c, p in enumerate(self.active):
                vu = V[:, c] / np.linalg.norm(V[:, c])
                man += T.Lns[p].dot(vu).dot(vu)
                ign += np.sum(np.power(vu[T.supports[p]] + 1, -1))
            fro = np.trace(Gu)
            parts[s] = [recon, man, ign, fro, recon + gamma * man + delta * ign + fro, gamma, delta, r2]
            if tradeoff is not None:
                d = tradeoff * man
                gamma = 1 if d == 0 else ((1 - tradeoff) * recon) / d
                delta = gamma
        return parts, gamma, delta

    def step_async(self, n_steps, gamma, delta, tradeoff=None):
        self._pending = self.step(n_steps, gamma, delta, tradeoff)

    def step_collect(self, n_steps):
        return self._pending

    def scores(self):
        K, P = self.k, self.P
        mass = np.zeros((K, P)); qn = np.zeros((K, P)); qr = np.zeros((K, P))
        for c in range(K):
            v = self.V[:, c]
            vu = v / np.linalg.norm(v)
            for p in range(P):
                mass[c, p] = np.sum(np.power(vu[self.tables.supports[p]], 2))
                qn[c, p] = self.tables.Lns[p].dot(vu).dot(vu)
                qr[c, p] = self.tables.Ls[p].dot(v).dot(v)
        return mass, qn, qr

    def block_end(self, n_steps, want_scores=True, prefetch=True):
        parts, g2, d2 = self.step_collect(n_steps)
        return parts, g2, d2, (self.scores() if want_scores else None)

    def snapshot_best(self):
        self._best = (self.U.copy(), self.V.copy())

    def restore_best(self):
        self.U, self.V = self._best[0].copy(), self._best[1].copy()


def factory(m_local, m_global, n, k, ctx):
    return NumpyShardEngine(m_local, m_global, n, k, ctx)
