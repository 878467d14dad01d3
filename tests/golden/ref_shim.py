"""Load the UNMODIFIED reference solver (`/root/reference/script/prmf_runner.py`) in this container.

Test infrastructure only.  The reference pins networkx<2.0 / numpy 1.18 (reference
`requirements.txt:5`, `env/minimal-environment.yml:6-10`); this image has networkx 3.x / numpy 2.x, so
six call sites fail.  The shims below patch *library* behaviour around the reference
module; not one line of reference source is edited or copied.  `/root/reference` exists only in the build
container, so nothing under `-m gpu`, `smoke()` or `bench.py` may import this file: it is used by
`make_golden.py` (to write the committed fixtures) and by CPU tests that skip when the tree is absent.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import scipy.sparse as sp
import networkx as nx

REF_ROOT = os.environ.get("PRMF_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REF_ROOT, "script", "prmf_runner.py"))


def _adjacency_matrix_nx1(G, nodelist=None, weight="weight"):
    """networkx-1.11 semantics of `nx.adjacency_matrix(G, nodelist)` as a scipy csr_matrix:
    nodelist may be a superset of G's nodes, edges with an endpoint outside nodelist are dropped,
    undirected edges are mirrored, self loops are counted once (used at prmf_runner.py:679)."""
    if nodelist is None:
        nodelist = list(G.nodes())
    index = {node: i for i, node in enumerate(nodelist)}
    n = len(nodelist)
    rows, cols, vals = [], [], []
    for u, v, d in G.edges(data=True):
        if u in index and v in index:
            w = d.get(weight, 1)
            rows.append(index[u]); cols.append(index[v]); vals.append(w)
            if u != v:
                rows.append(index[v]); cols.append(index[u]); vals.append(w)
    M = sp.coo_matrix((np.asarray(vals, dtype=np.float64), (rows, cols)), shape=(n, n))
    return sp.csr_matrix(M)


def _max_weight_matching_dict(G, maxcardinality=False, weight="weight"):
    """networkx>=2 returns a set of pairs; the reference iterates `mate.items()` (prmf_runner.py:246-247)."""
    mate = {}
    for a, b in nx.max_weight_matching(G, maxcardinality=maxcardinality, weight=weight):
        mate[a] = b
        mate[b] = a
    return mate


_PATCHED = False


def _patch_libraries():
    global _PATCHED
    if _PATCHED:
        return
    if not hasattr(np, "Inf"):
        np.Inf = np.inf                                      # prmf_runner.py:713
    orig_add_edge = nx.Graph.add_edge

    def add_edge(self, u, v, attr_dict=None, **attr):        # prmf_runner.py:245 passes a positional dict
        if attr_dict is not None:
            attr = dict(attr_dict, **attr)
        return orig_add_edge(self, u, v, **attr)

    nx.Graph.add_edge = add_edge
    orig_add_node = nx.Graph.add_node

    def add_node(self, n, attr_dict=None, **attr):
        if attr_dict is not None:
            attr = dict(attr_dict, **attr)
        return orig_add_node(self, n, **attr)

    nx.Graph.add_node = add_node
    nx.Graph.nodes_iter = lambda self: iter(self.nodes())    # prmf_runner.py:984
    nx.Graph.edges_iter = lambda self, *a, **k: iter(self.edges(*a, **k))
    nx.Graph.node = property(lambda self: self.nodes)        # prmf/__init__.py:72
    _PATCHED = True


_MODULE = None


def load_reference():
    """Return the reference module object, with `module.nx` replaced by a tolerant namespace."""
    global _MODULE
    if _MODULE is not None:
        return _MODULE
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    _patch_libraries()
    sys.dont_write_bytecode = True                           # the tree is read-only
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    spec = importlib.util.spec_from_file_location(
        "prmf_runner_reference", os.path.join(REF_ROOT, "script", "prmf_runner.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ns = types.SimpleNamespace(**{k: getattr(nx, k) for k in dir(nx) if not k.startswith("__")})
    ns.adjacency_matrix = _adjacency_matrix_nx1
    ns.max_weight_matching = _max_weight_matching_dict
    mod.nx = ns
    _MODULE = mod
    return mod


def reset_globals(mod):
    """`NORMALIZED_LAPLACIANS` is only ever appended to (prmf_runner.py:29,:696)."""
    mod.NORMALIZED_LAPLACIANS = []
    mod.LAPLACIANS = []
    mod.PATHWAY_TO_SUPPORT = None
