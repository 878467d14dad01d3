"""prmf_runner.py drop-in: flags, exit codes, file formats (CPU) and end-to-end parity with the outputs of
the reference's own main() recorded in tests/golden/cli_test1_* (GPU)."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, load_golden

sys.path.insert(0, GOLDEN)


def _write_inputs(tmp_path):
    from make_golden_io import write_cli_inputs
    g = load_golden("test1_raw")
    fps = write_cli_inputs(str(tmp_path), g["X"], g["nodelist"], g["Gs"])
    return [os.path.basename(f) for f in fps]


def _main(argv, cwd):
    from prmf_b200.prmf_runner import main
    old = os.getcwd()
    os.chdir(cwd)
    try:
        with contextlib.redirect_stdout(io.StringIO()) as out, contextlib.redirect_stderr(io.StringIO()) as err:
            try:
                main(argv)
                code = 0
            except SystemExit as e:
                code = e.code
    finally:
        os.chdir(old)
    return code, out.getvalue(), err.getvalue()


def test_flag_schema_matches_reference():
    import argparse
    from prmf_b200 import prmf_args
    p = argparse.ArgumentParser()
    prmf_args.add_prmf_arguments(p)
    flags = {s for a in p._actions for s in a.option_strings}
    expected = {"--data", "--manifolds", "--manifolds-file", "--manifolds-init", "--node-attribute", "--outdir",
                "--nodelist", "--k-latent", "-k", "--tolerence", "--seed", "--gamma", "--delta", "--tradeoff",
                "--high-dimensional", "--no-normalize", "--delimiter", "--m-samples", "--cross-validation", "-c",
                "--verbose", "-v", "--normalize", "-h", "--help"}
    assert flags == expected
    a = p.parse_args(["--data", "x", "--outdir", "o"])
    assert (a.k_latent, a.gamma, a.delta, a.tradeoff, a.high_dimensional, a.delimiter, a.tolerence) == (
        6, 1.0, 1.0, -1, None, ",", 1e-3)              # no transpose unless asked (the reference defaults to True)
    assert p.parse_args(["--data", "x", "--outdir", "o", "--high-dimensional", "False"]).high_dimensional is False
    assert p.parse_args(["--data", "x", "--outdir", "o", "--high-dimensional", "True"]).high_dimensional is True
    assert p.parse_args(["--data", "x", "--outdir", "o", "--normalize"]).no_normalize is False


def test_exit_codes(tmp_path):
    rel = _write_inputs(tmp_path)
    base = ["--data", "data.tsv", "--outdir", ".", "--delimiter", "\t"]
    assert _main(base, tmp_path)[0] == 22                                       # neither manifold flag
    with open(tmp_path / "mf.txt", "w") as fh:
        fh.write("\n".join(rel))
    assert _main(base + ["--manifolds"] + rel + ["--manifolds-file", "mf.txt"], tmp_path)[0] == 23
    assert _main(base + ["--manifolds"] + rel, tmp_path)[0] == 25              # no nodelist, no header
    with open(tmp_path / "bad_nodelist.txt", "w") as fh:
        fh.write("\n".join("X%d" % i for i in range(1000)))
    code, _, err = _main(base + ["--manifolds"] + rel + ["--nodelist", "bad_nodelist.txt"], tmp_path)
    assert code == 24 and "Invalid manifolds" in err
    assert _main(["--outdir", "."], tmp_path)[0] == 2                           # argparse: --data required


def test_sniffers_and_embed(tmp_path):
    from prmf_b200.prmf_runner import check_header, check_row_names, embed_arr, parse_nodelist
    _write_inputs(tmp_path)
    assert check_header(str(tmp_path / "data.tsv"), "\t") is False
    assert check_header(str(tmp_path / "data_header.tsv"), "\t") is True
    assert check_row_names(str(tmp_path / "data_header.tsv"), "\t", True) is False
    with open(tmp_path / "rn.csv", "w") as fh:
        fh.write("id,a,b\ns1,1.0,2.0\ns2,3.0,4.0\n")
    assert check_header(str(tmp_path / "rn.csv"), ",") and check_row_names(str(tmp_path / "rn.csv"), ",", True)
    arr = np.arange(6.0).reshape(2, 3)
    out = embed_arr(["q", "b", "a", "z", "c"], ["a", "b", "c"], arr)
    ref = np.zeros((2, 5))
    for i in range(2):
        for j, name in enumerate(["a", "b", "c"]):
            ref[i, ["q", "b", "a", "z", "c"].index(name)] = arr[i, j]
    np.testing.assert_array_equal(out, ref)
    with open(tmp_path / "nl.txt", "w") as fh:
        fh.write("a b\nc\n  d\te\n")
    with open(tmp_path / "nl.txt") as fh:
        assert parse_nodelist(fh) == ["a", "b", "c", "d", "e"]


def _compare_outputs(tmp_path, tag):
    import pandas as pd
    exp = os.path.join(GOLDEN, "cli_test1_" + tag)
    got_obj = open(tmp_path / "obj.txt").read().splitlines()
    exp_obj = open(os.path.join(exp, "obj.txt")).read().splitlines()
    nkv = sum(1 for line in exp_obj if " = " in line)
    assert len(got_obj) == len(exp_obj)
    assert got_obj[nkv:] == exp_obj[nkv:], "factor -> graphml lines differ"
    for a, b in zip(got_obj[:nkv], exp_obj[:nkv]):
        ka, va = a.split(" = "); kb, vb = b.split(" = ")
        assert ka == kb
        np.testing.assert_allclose(float(va), float(vb), rtol=1e-6, atol=2e-5)    # values are printed to 5 decimals
    for f in ("U.csv", "V.csv"):
        got = pd.read_csv(tmp_path / f)
        ref = pd.read_csv(os.path.join(exp, f))
        assert list(got.columns) == list(ref.columns)
        if f == "V.csv":
            assert list(got.iloc[:, 0]) == list(ref.iloc[:, 0])
            got, ref = got.iloc[:, 1:], ref.iloc[:, 1:]
        np.testing.assert_allclose(got.to_numpy(), ref.to_numpy(), rtol=1e-6, atol=1e-10)
    raw_u = open(tmp_path / "U.csv").readline().strip()
    assert raw_u == '"LV0","LV1","LV2","LV3","LV4","LV5"'                         # QUOTE_NONNUMERIC header


@pytest.mark.gpu
@pytest.mark.parametrize("tag,extra", [("nonorm", ["--no-normalize"]), ("norm", [])])
def test_cli_matches_reference_main(tmp_path, tag, extra):
    rel = _write_inputs(tmp_path)
    argv = ["--data", "data.tsv", "--manifolds"] + rel + ["--node-attribute", "name", "--nodelist", "nodelist.txt",
                                                           "--outdir", ".", "--delimiter", "\t", "--seed", "1"] + extra
    code, out, err = _main(argv, tmp_path)
    assert code == 0, err[-2000:]
    lines = out.splitlines()
    exp_head = open(os.path.join(GOLDEN, "cli_test1_" + tag, "stdout_head.txt")).read().splitlines()
    assert lines[:7] == exp_head[:7]                                             # manifold coverage report
    assert lines[7].startswith("norm(X) = ")
    np.testing.assert_allclose(float(lines[7].split("=")[1]), float(exp_head[7].split("=")[1]), rtol=1e-13)
    _compare_outputs(tmp_path, tag)
    assert "Before restrict" not in err                                          # 6 pathways, k = 6: matched at once


@pytest.mark.gpu
def test_cli_inferred_nodelist_gives_same_result(tmp_path):
    """Second half of test_inferred_nodelist_1.py (:52-57): header-inferred nodelist, same obj.txt."""
    rel = _write_inputs(tmp_path)
    argv = ["--data", "data_header.tsv", "--manifolds"] + rel + ["--node-attribute", "name", "--outdir", ".",
                                                                  "--delimiter", "\t", "--seed", "1", "--no-normalize",
                                                                  "--normalize"]
    code, out, err = _main(argv, tmp_path)
    assert code == 0, err[-2000:]
    _compare_outputs(tmp_path, "nonorm")


@pytest.mark.gpu
def test_cli_manifolds_init_matches_reference(tmp_path):
    """--manifolds-init with exactly k files: pathway_to_vec + two NNLS solves per factor (:272-334, :1023-1064)."""
    rel = _write_inputs(tmp_path)
    order = [rel[i] for i in (3, 1, 5, 0, 2, 4)]
    argv = ["--data", "data.tsv", "--manifolds"] + rel + ["--node-attribute", "name", "--nodelist", "nodelist.txt",
                                                           "--outdir", ".", "--delimiter", "\t", "--seed", "1",
                                                           "--no-normalize", "--manifolds-init"] + order
    code, out, err = _main(argv, tmp_path)
    assert code == 0, err[-2000:]
    assert "Using the following manifolds for initialization:" in out
    assert open(tmp_path / "init_pathways.txt").read() == open(
        os.path.join(GOLDEN, "cli_test1_init", "init_pathways.txt")).read()
    _compare_outputs(tmp_path, "init")


@pytest.mark.gpu
def test_cli_cross_validation_matches_reference(tmp_path):
    """--cross-validation 0.2: first KFold split held out, per-sample NNLS test error (:1004-1013, :1074-1079)."""
    rel = _write_inputs(tmp_path)
    argv = ["--data", "data.tsv", "--manifolds"] + rel + ["--node-attribute", "name", "--nodelist", "nodelist.txt",
                                                           "--outdir", ".", "--delimiter", "\t", "--seed", "1",
                                                           "--no-normalize", "--cross-validation", "0.2"]
    code, out, err = _main(argv, tmp_path)
    assert code == 0, err[-2000:]
    _compare_outputs(tmp_path, "cv")
    got = np.loadtxt(tmp_path / "test_error.csv", delimiter=",")
    ref = np.loadtxt(os.path.join(GOLDEN, "cli_test1_cv", "test_error.csv"), delimiter=",")
    np.testing.assert_allclose(got, ref, rtol=1e-6)


def test_random_restart_commands():
    """nmf_pathway_rr.py:24-37: every restart gets the shared arguments, its own --outdir run<i> and a bare
    --manifolds-init; --condor / --n-runs / the outer --outdir are not passed through."""
    import argparse
    from prmf_b200 import prmf_args, restarts
    p = argparse.ArgumentParser()
    p.add_argument("--n-runs", type=int, default=2)
    p.add_argument("--condor", action="store_true")
    p.add_argument("--gpus", type=int, default=None)
    prmf_args.add_prmf_arguments(p)
    a = p.parse_args(["--n-runs", "3", "--data", "d.tsv", "--outdir", "out", "--manifolds", "a.graphml", "b.graphml",
                      "-k", "4", "--delimiter", "\t", "--no-normalize", "--seed", "7"])
    cmds = restarts.build_commands(a)
    assert [c[0] for c in cmds] == [os.path.join("out", "run%d" % i) for i in range(3)]
    argv = cmds[1][1]
    assert argv[-3:] == ["--outdir", os.path.join("out", "run1"), "--manifolds-init"]
    assert "--condor" not in argv and "--n-runs" not in argv and "--gpus" not in argv
    q = argparse.ArgumentParser()
    prmf_args.add_prmf_arguments(q)
    b = q.parse_args(argv)                                          # the child accepts what the parent emits
    assert (b.data, b.k_latent, b.delimiter, b.no_normalize, b.seed, b.manifolds, b.manifolds_init) == (
        "d.tsv", 4, "\t", True, "7", ["a.graphml", "b.graphml"], [])
    assert b.outdir == os.path.join("out", "run1") and b.high_dimensional is None


@pytest.mark.gpu
def test_random_restarts_end_to_end(tmp_path):
    """Two restarts on the reference's test files: both runs finish, write U.csv / V.csv / obj.txt and are listed."""
    import subprocess
    rel = _write_inputs(tmp_path)
    out = tmp_path / "rr"
    out.mkdir()
    cmd = [sys.executable, os.path.join(ROOT, "script", "nmf_pathway_rr.py"), "--n-runs", "2", "--data", "data.tsv",
           "--delimiter", "\t", "--manifolds"] + rel + ["--node-attribute", "name", "--nodelist", "nodelist.txt",
                                                        "--outdir", str(out), "-k", "3", "--no-normalize"]
    res = subprocess.run(cmd, cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    for i in range(2):
        for f in ("U.csv", "V.csv", "obj.txt", "init_pathways.txt", "nmf_pathway.out", "nmf_pathway.err"):
            assert (out / ("run%d" % i) / f).exists(), f
    lines = (out / "runs.tsv").read_text().strip().splitlines()
    assert len(lines) == 3 and all(l.split("\t")[1] == "0" for l in lines[1:])
