"""bench.py's driver-facing contract, checked without a GPU: the synthetic matrix is the same for every partition of the
rows (strong scaling compares ONE instance), both arms print the same `config` dict, and the reference arm produces a
well-formed line from the CPU implementation of the path (the unmodified reference functions when `baseline/_ref` is
installed, else the oracle port)."""
import importlib.util
import json
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT


def _bench():
    argv = sys.argv
    sys.argv = ["bench.py"]
    try:
        spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


def test_synthetic_matrix_does_not_depend_on_the_partition():
    b = _bench()
    full = b.host_rows(0, 9000, 37)
    for world in (2, 3, 8):
        per = -(-9000 // world)
        parts = [b.host_rows(min(9000, r * per), min(9000, (r + 1) * per), 37) for r in range(world)]
        assert np.array_equal(full, np.concatenate(parts))
    assert np.array_equal(b.host_rows(100, 5000, 37, np.float32), full[100:5000].astype(np.float32))
    out = np.empty((4096, 37))
    assert b.host_rows(4096, 8192, 37, out=out) is out and np.array_equal(out, full[4096:8192])


def test_both_arms_describe_the_same_config():
    b = _bench()
    sys_argv = sys.argv
    try:
        sys.argv = ["bench.py", "--gpus", "4"]
        ours = b.config_dict(b.parse_args())
        sys.argv = ["bench.py", "--gpus", "4", "--impl", "reference"]
        ref = b.config_dict(b.parse_args())
    finally:
        sys.argv = sys_argv
    assert ours == ref and "37032x6750" in ours["workload"] and "k=10" in ours["workload"]


def test_reference_arm_line(tmp_path):
    env = dict(os.environ, RANK="0", WORLD_SIZE="1", OMP_NUM_THREADS="1")       # as under torchrun: the arm must undo this
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--m", "400", "--n", "500",
                          "--pathways", "12", "--k", "4", "--steps", "2", "--warmup", "1", "--ref-budget-s", "20"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=str(tmp_path))
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "prmf_outer_iterations_per_sec" and line["unit"] == "outer_it/s"
    assert line["higher_is_better"] is True and line["steps"] == 2 and line["warmup"] == 1 and line["value"] > 0
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["value"] == line["value"] and cb["cores"] >= 1 and cb["sample"]
    assert 1 <= cb["inner_steps_per_sample"] <= 10
    assert line["e2e"] == {"value": line["value"], "unit": "outer_it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # ms_per_step is the wall time of a sample, value the rate of a whole outer iteration (10 inner steps + restrict)
    assert line["ms_per_step"] <= line["ms_per_outer_iteration"] * 1.05
    assert os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "bin", "prmf_runner.py")) == (cb["kind"] == "reference")
    # ranks other than 0 exit quietly
    res2 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--m", "100"],
                          capture_output=True, text=True, timeout=120, env=dict(env, RANK="1", WORLD_SIZE="2"))
    assert res2.returncode == 0 and res2.stdout.strip() == ""
