"""Random restarts -- the caller pattern above the hot path (script/nmf_pathway_rr.py:24-37).

The reference builds a job graph of `--n-runs` independent `nmf_pathway.py` processes, each with the same
arguments, its own `--outdir <outdir>/run<i>` and `--manifolds-init` (so every run draws its own random set of
initialising pathways), and runs them one after the other (or submits them to Condor).  The runs are independent, so
here they are spread over the visible B200s: one `prmf_runner.py` process per GPU at a time (CUDA_VISIBLE_DEVICES),
stdout / stderr of run i in `<outdir>/run<i>/nmf_pathway.out|.err` as the reference's job runner leaves them.
`--condor` is accepted and refused (there is no scheduler here).  Extra: `<outdir>/runs.tsv` lists every run with
its exit code and final objective, best first.
"""
import argparse
import os
import subprocess
import sys

from . import prmf_args

RUNNER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "script", "prmf_runner.py")


def args_to_list(args_dict):
    """`prmf.script_utils.args_to_list` (prmf/script_utils.py:9-22): sorted keys, `--flag value...`, None skipped.
    Booleans are emitted as switches only when set (the reference would emit `--no-normalize False`, which its own
    argparse rejects)."""
    rv = []
    for k in sorted(args_dict.keys()):
        v = args_dict[k]
        flag = "--" + k.replace("_", "-")
        if v is None:
            continue
        if isinstance(v, bool):
            if k == "high_dimensional":
                rv += [flag, str(v)]
            elif v:
                rv.append(flag)
        elif isinstance(v, list):
            rv.append(flag)
            rv += [str(x) for x in v]
        else:
            rv += [flag, str(v)]
    return rv


def build_commands(args):
    """One (run_outdir, argv) per restart, as nmf_pathway_rr.py:24-37 assembles them."""
    d = dict(vars(args))
    outdir = d.pop("outdir")
    d.pop("condor", None)
    d.pop("manifolds_init", None)
    n_runs = d.pop("n_runs")
    d.pop("gpus", None)
    cmds = []
    for i in range(n_runs):
        run_outdir = os.path.join(outdir, "run{}".format(i))
        argv = args_to_list(d) + ["--outdir", run_outdir, "--manifolds-init"]
        cmds.append((run_outdir, argv))
    return cmds


def visible_gpus(limit=None):
    env = os.environ.get("CUDA_VISIBLE_DEVICES")
    if env is not None and env.strip() != "":
        ids = [x.strip() for x in env.split(",") if x.strip() != ""]
    else:
        try:
            import torch
            ids = [str(i) for i in range(torch.cuda.device_count())]
        except Exception:
            ids = []
    if not ids:
        raise SystemExit("nmf_pathway_rr: no CUDA device visible; there is no CPU execution path")
    return ids[:limit] if limit else ids


def _final_obj(run_outdir):
    try:
        with open(os.path.join(run_outdir, "obj.txt")) as fh:
            for line in fh:
                if line.startswith("obj ="):
                    return float(line.split("=")[1])
    except OSError:
        pass
    return float("nan")


def main(argv=None):
    parser = argparse.ArgumentParser(description="Run prmf_runner.py with different random restarts, one per GPU at a "
                                                 "time.  Arguments are passed through except --condor, "
                                                 "--manifolds-init, --n-runs, --gpus and --outdir.")
    parser.add_argument("--n-runs", type=int, default=2, help="Number of random restarts")
    parser.add_argument("--condor", action="store_true", help="(reference flag) not available here")
    parser.add_argument("--gpus", type=int, default=None, help="Use at most this many of the visible GPUs")
    prmf_args.add_prmf_arguments(parser)
    args = parser.parse_args(argv)
    if args.condor:
        sys.stderr.write("nmf_pathway_rr: --condor is not supported; the runs are scheduled over the local GPUs\n")
        sys.exit(26)
    cmds = build_commands(args)
    gpus = visible_gpus(args.gpus)
    from . import _lib
    _lib.load()                                                   # build once, before the workers fan out
    for run_outdir, _ in cmds:
        os.mkdir(run_outdir)                                      # as the reference: fails if it exists
    running, results, todo = {}, {}, list(enumerate(cmds))
    free = list(gpus)
    while todo or running:
        while todo and free:
            i, (run_outdir, run_argv) = todo.pop(0)
            gpu = free.pop(0)
            env = dict(os.environ, CUDA_VISIBLE_DEVICES=gpu)
            out = open(os.path.join(run_outdir, "nmf_pathway.out"), "w")
            err = open(os.path.join(run_outdir, "nmf_pathway.err"), "w")
            proc = subprocess.Popen([sys.executable, RUNNER] + run_argv, stdout=out, stderr=err, env=env)
            running[i] = (proc, gpu, out, err, run_outdir)
        for i in list(running):
            proc, gpu, out, err, run_outdir = running[i]
            try:
                code = proc.wait(timeout=0.2)
            except subprocess.TimeoutExpired:
                continue
            out.close(); err.close()
            results[i] = (run_outdir, code, _final_obj(run_outdir))
            free.append(gpu)
            del running[i]
    order = sorted(results, key=lambda i: (results[i][1] != 0, results[i][2] != results[i][2], results[i][2]))
    with open(os.path.join(args.outdir, "runs.tsv"), "w") as fh:
        fh.write("run\texit_code\tobj\toutdir\n")
        for i in order:
            run_outdir, code, obj = results[i]
            fh.write("{}\t{}\t{}\t{}\n".format(i, code, obj, run_outdir))
    if any(code != 0 for _, code, _ in results.values()):
        sys.exit(1)


if __name__ == "__main__":
    main()
