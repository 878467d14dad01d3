"""Transfer of a fitted gene x factor matrix to new samples -- the step after the hot path in the reference's recount2
workflow (`script/transfer_learning.R:98-110`):

    B = (Z^T Z + l2 I)^-1 Z^T Y          Z: genes x k (the V of a PRMF run), Y: genes x samples, B: k x samples

with the reference's default l2 = 100 * (genes * samples) / (k * samples) (:93-96).  In the orientation of this
package the new data come as X = Y^T (samples x genes), so Z^T Y = (X Z)^T is the same skinny product as the U
update's X.V: it runs on the GPU through the pass-1 X-stream kernel (`prmf_project`); the k x k solve is done on
the host in fp64.
"""
import numpy as np

from .engine import CudaEngine


def default_l2(n_genes, n_samples, k):
    """`ratio * 100` of transfer_learning.R:93-96."""
    return (n_genes * n_samples) / (k * n_samples) * 100


def ridge_transfer(X, Z, l2=None, device=None):
    """X: new samples x genes, Z: genes x k.  Returns B^T (samples x k), the rows of the reference's
    `sample_by_latent_transfer.csv` (it writes t(b_matrix), :110)."""
    import torch
    X = np.asarray(X, dtype=np.float64)
    Z = np.ascontiguousarray(Z, dtype=np.float64)
    if X.ndim != 2 or Z.ndim != 2 or X.shape[1] != Z.shape[0]:
        raise ValueError("Incompatible dimensions: X %s (samples x genes), Z %s (genes x k)" % (X.shape, Z.shape))
    m, n = X.shape
    k = Z.shape[1]
    if l2 is None:
        l2 = default_l2(n, m, k)
    dev = torch.cuda.current_device() if device is None else int(device)
    with CudaEngine(m, m, n, k, device=dev) as eng:
        eng.set_X(X)
        eng.set_UV(None, Z)
        A = eng.project()                                    # X Z = (Z^T Y)^T, on the device
    G = Z.T @ Z + l2 * np.eye(k)                             # k x k, host
    return np.linalg.solve(G, A.T).T                         # (G^-1 Z^T Y)^T


def main(argv=None):
    """CLI counterpart of transfer_learning.R for inputs that already share gene identifiers:
    --data Y.csv (samples x genes, header = genes, first column = sample names), --V V.csv (as written by
    prmf_runner.py), --outdir; writes sample_by_latent_transfer.csv."""
    import argparse
    import os
    import pandas as pd
    ap = argparse.ArgumentParser(description=main.__doc__)
    ap.add_argument("--data", required=True)
    ap.add_argument("--V", required=True)
    ap.add_argument("--outdir", required=True)
    ap.add_argument("--l2", type=float, default=None)
    a = ap.parse_args(argv)
    Y = pd.read_csv(a.data, index_col=0)
    V = pd.read_csv(a.V, index_col=0)
    common = [g for g in V.index if g in set(Y.columns)]
    if not common:
        raise SystemExit("No common genes, quitting")                         # transfer_learning.R:79-82
    B = ridge_transfer(Y[common].fillna(0).to_numpy(), V.loc[common].to_numpy(), l2=a.l2)
    os.makedirs(a.outdir, exist_ok=True)
    pd.DataFrame(B, index=Y.index, columns=V.columns).to_csv(os.path.join(a.outdir, "sample_by_latent_transfer.csv"))


if __name__ == "__main__":
    main()
