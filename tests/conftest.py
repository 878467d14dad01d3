import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU tests never run silently on a box without a device: they are skipped with a loud reason
    unless selected on a machine that has one."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (runs under gpurun)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    """Fixture recorded from the unmodified reference (tests/golden/make_golden.py)."""
    import networkx as nx
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    Gs = []
    for g in meta["graphs"]:
        G = nx.Graph()
        G.add_nodes_from(g["nodes"])
        for u, v, w in g["edges"]:
            if g["weighted"]:
                G.add_edge(u, v, weight=w)
            else:
                G.add_edge(u, v)
        Gs.append(G)
    out = {k: z[k] for k in z.files if k != "meta"}
    out["meta"] = meta
    out["Gs"] = Gs
    out["nodelist"] = meta["nodelist"]
    return out


GOLDEN_CASES = ["test1_raw", "test1_norm", "test2_raw", "small_planted", "small_tradeoff", "small_bigk"]
