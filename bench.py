#!/usr/bin/env python
"""PRMF hot-path benchmark (contract: see the task statement / DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One *step* = one outer iteration of `nmf_pathway` (reference script/prmf_runner.py:715-774) on the
recount2-shape synthetic instance (BASELINE.json configs[1]: 37 032 samples x 6 750 genes, k = 10,
300 KEGG-size random pathways): k multinomial draws of the active pathways, `modulus` = 10 inner
multiplicative updates with their objectives, and one `restrict` over the full candidate table.
`value` is outer iterations per second with X resident in HBM; `e2e` is the same step driven with
host buffers (X, U, V uploaded from pinned memory and U, V, objective read back every step).
With N > 1 the rows of X and U are sharded over the ranks (strong scaling; one NCCL all-reduce of
[X^T U | U^T U] per inner step).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M, N_GENES, K, P = 37032, 6750, 10, 300
MODULUS = 10
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the X-stream kernel at config 2, from ncu --set full
NCU_TRAFFIC_CONFIG2 = {False: 2.0073e9,      # fp64 skinny_tma_kernel     (profiles/r1_v5_skinny_tma_ncu_summary.txt)
                       True: 1.0092e9}       # tf32 tc_rowdot_kernel      (profiles/r1_v5_tc_rowdot_ncu_summary.txt)
METRIC = "prmf_outer_iterations_per_sec"
UNIT = "outer_it/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--m", type=int, default=M)
    ap.add_argument("--n", type=int, default=N_GENES)
    ap.add_argument("--k", type=int, default=K)
    ap.add_argument("--pathways", type=int, default=P)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-inner-steps", type=int, default=2)
    ap.add_argument("--x-dtype", default="f64", choices=["f64", "tf32"],
                    help="f64: parity mode (default, the headline); tf32: opt-in tensor-core X streams (X stored fp32)")
    return ap.parse_args()


def workload_name(a):
    return "recount2-shape synthetic %dx%d %s, k=%d, %d random KEGG-size pathways" % (
        a.m, a.n, "fp64" if a.x_dtype == "f64" else "fp32 (tf32 tensor-core X streams)", a.k, a.pathways)


def make_pathways(a):
    from prmf_b200 import pack_pathways, synth
    rng = np.random.Generator(np.random.PCG64(0))
    Gs = synth.random_pathway_graphs(rng, a.n, a.pathways)
    nodelist = list(range(a.n))
    return Gs, nodelist, pack_pathways(Gs, nodelist)


# ------------------------------------------------------------------------------------------------
# clocks (nvidia-smi sampled during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None
        self.begin = 0

    def _lines(self):
        try:
            with open(self.path) as fh:
                return fh.readlines()
        except Exception:
            return []

    def mark_begin(self):
        """Samples before this point (set-up, warm-up) are not part of the timed region."""
        self.begin = len(self._lines())

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        try:
            lines = self._lines()
            lines = lines[self.begin:] if len(lines) > self.begin else lines[-3:]
            for line in lines:
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)), reasons=sorted(reasons),
                       samples=len(sm), power_w_max=float(max(power)))
        return out


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle (a port of the reference's numpy/scipy path) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_outer_iteration_rate(X, Gs, nodelist, k, n_inner, seed=1):
    """Time `n_inner` inner steps (update + objective, prmf_runner.py:419-449) and one `restrict`
    (:129-194) of the CPU oracle at the full shape and extrapolate to one outer iteration
    (10 inner steps + restrict).  Returns (outer_it_per_s, detail dict)."""
    from oracle import prmf_oracle as O
    rng = np.random.Generator(np.random.PCG64(seed))
    m, n = X.shape
    U = 3 * (1 - rng.random((m, k))); V = 3 * (1 - rng.random((n, k)))
    t0 = time.perf_counter()
    tables = O.PathwayTables(Gs, nodelist)
    t_tables = time.perf_counter() - t0
    normX = np.linalg.norm(X)
    gamma, delta = normX / k, 10 / normX
    active = list(range(k))
    t0 = time.perf_counter()
    for _ in range(n_inner):
        U, V = O.update_step(X, U, V, tables, active, gamma, delta)
        O.objective(X, U, V, tables, active, gamma, delta)
    t_inner = (time.perf_counter() - t0) / n_inner
    cands = {kk: [(p, 1) for p in range(len(tables))] for kk in range(k)}
    t0 = time.perf_counter()
    O.restrict(V, tables, cands)
    t_restrict = time.perf_counter() - t0
    t_outer = MODULUS * t_inner + t_restrict
    return 1.0 / t_outer, {"s_per_inner_step": t_inner, "s_restrict": t_restrict, "s_tables": t_tables}


def host_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        if n:
            return int(max(n))
    except Exception:
        pass
    return os.cpu_count() or 1


def run_reference_arm(a):
    """`--impl reference`: the reference's CPU implementation of the path (the oracle port; the Python
    reference itself needs networkx<2 and cannot run on the GPU box) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from prmf_b200 import synth  # noqa: F401  (instance generators are shared)
    rng = np.random.Generator(np.random.PCG64(1234))
    X = rng.random((a.m, a.n))
    Gs, nodelist, _ = make_pathways(a)
    rates = []
    total = a.warmup + a.steps
    # each step = a bounded sample: `cpu_inner_steps` inner steps + 1 restrict, extrapolated to 10 + 1
    for s in range(total):
        r, detail = cpu_outer_iteration_rate(X, Gs, nodelist, a.k, a.cpu_inner_steps, seed=s)
        if s >= a.warmup:
            rates.append(r)
        if time.perf_counter() - T_START > 240 and rates:
            break
    value = float(np.mean(rates))
    cores = host_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
        "steps": len(rates), "warmup": a.warmup, "ms_per_step": 1000.0 / value, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "step": "1 outer iteration = 10 inner steps + restrict"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d inner steps + 1 restrict at full shape per step, extrapolated to 10 + 1" % a.cpu_inner_steps},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    from prmf_b200 import CudaEngine
    from prmf_b200.dist import DistContext, row_block
    from prmf_b200.engine import attach_collectives
    from prmf_b200.solver import init_latent_to_pathway_data, restrict_from_tables, sample_active

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    ctx = DistContext.from_env()
    if ctx.world != a.gpus:
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run)" % (a.gpus, ctx.world))
    dev = ctx.local_rank if ctx.world > 1 else 0
    torch.cuda.set_device(dev)
    sampler = ClockSampler(dev)
    sampler.start()                      # nvidia-smi needs ~1 s to come up: start it before the data setup
    stream = torch.cuda.Stream(device=dev)
    lo, hi = row_block(a.m, ctx.world, ctx.rank)
    m_local = hi - lo
    Gs, nodelist, packed = make_pathways(a)

    gen = torch.Generator(device="cuda")
    gen.manual_seed(1234 + ctx.rank)
    tf32 = a.x_dtype == "tf32"
    xdt = torch.float32 if tf32 else torch.float64
    xsz = 4 if tf32 else 8
    Xd = torch.rand((m_local, a.n), dtype=xdt, device="cuda", generator=gen)
    eng = CudaEngine(m_local, a.m, a.n, a.k, device=dev, stream=stream.cuda_stream, x_dtype=a.x_dtype)
    attach_collectives(eng, ctx)
    eng.set_X(Xd)
    eng.set_pathways(packed)
    normX = float(np.sqrt(eng.normX_sq))
    gamma, delta = normX / a.k, 10 / normX

    np.random.seed(1)
    U0 = 3 * (1 - np.random.rand(a.m, a.k))[lo:hi]
    V0 = 3 * (1 - np.random.rand(a.n, a.k))
    eng.set_UV(U0, V0)
    full_cands = init_latent_to_pathway_data(a.k, packed.P)

    host_t = [0.0]

    def outer_iteration():
        t0 = time.perf_counter()
        active = sample_active(full_cands, a.k)
        eng.set_active(active)
        t1 = time.perf_counter()
        eng.step_async(MODULUS, gamma, delta)
        # score tables enqueued behind the 10 steps, one host wait; the next block's X.V pass is already
        # streaming while the host runs restrict and the multinomial draws (as prmf_b200.nmf_pathway does)
        parts, _, _, (mass, qn, _) = eng.block_end(MODULUS, want_scores=True, prefetch=True)
        t2 = time.perf_counter()
        restrict_from_tables(mass, qn, full_cands)
        host_t[0] += (t1 - t0) + (time.perf_counter() - t2)
        return parts

    def sync_all():
        if ctx.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):            # NOTE: the same number of steps on every rank (each holds collectives)
        outer_iteration()
    launches0 = eng.launch_count
    sync_all()
    sampler.mark_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # ---- timed region: exactly K steps, device events on the launching stream, max over ranks ----
    host_t[0] = 0.0
    e0.record(stream)
    for _ in range(a.steps):
        parts = outer_iteration()
    e1.record(stream)
    sync_all()
    host_ms = host_t[0] * 1e3 / a.steps      # sampling + set_active + scores + restrict per outer iteration
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count - launches0
    # ---- the same K steps again with an event pair around every phase of every inner step (the per-kernel
    #      durations of the roofline; the event records cost ~7 %, so they are kept out of `value`) ----
    eng.kernel_times(reset=True)
    eng.set_profiling(True)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(stream)
    for _ in range(a.steps):
        outer_iteration()
    p1.record(stream)
    sync_all()
    eng.set_profiling(False)
    clocks = sampler.stop()
    ms_profiled = p0.elapsed_time(p1) / a.steps
    kt = eng.kernel_times(reset=True)
    if ctx.world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / a.steps
    value = 1000.0 / ms_per_step

    # roofline of the X-stream kernels (algorithmic bytes per launch / mean launch duration)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    wsz = 4 if tf32 else 8                      # bytes per element of the W operand (fp32 transposed copy in tf32 mode)
    bytes_xv = m_local * a.n * xsz + a.n * a.k * wsz + m_local * a.k * 8
    bytes_xtu = m_local * a.n * xsz + m_local * a.k * wsz + a.n * a.k * 8
    phase_ms = {name: (tot / max(1, cnt)) for name, (tot, cnt) in kt.items()}
    xv_ms, xtu_ms = phase_ms["xv"], phase_ms["xtu"]
    ach_xv = bytes_xv / (xv_ms * 1e-3) / 1e9 if xv_ms > 0 else 0.0
    ach_xtu = bytes_xtu / (xtu_ms * 1e-3) / 1e9 if xtu_ms > 0 else 0.0
    step_bytes = bytes_xv + bytes_xtu
    inner_ms = ms_per_step / MODULUS
    kname = "tc_rowdot_kernel" if tf32 else "skinny_tma_kernel" if a.k <= 10 else "skinny_tma_gen_kernel"
    dominant = kname + (" pass 2 (X^T.U)" if xtu_ms >= xv_ms else " pass 1 (X.V)")
    ach = ach_xtu if xtu_ms >= xv_ms else ach_xv
    roofline = {
        "bound": "hbm", "kernel": dominant, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
        "traffic": NCU_TRAFFIC_CONFIG2[tf32] if (a.m, a.n, a.k, ctx.world) == (M, N_GENES, K, 1) else None,
        "traffic_source": "profiles/r1_v5_%s_ncu_summary.txt (dram__bytes_read.sum + dram__bytes_write.sum, mean of the two passes)"
                          % ("tc_rowdot" if tf32 else "skinny_tma"),
        "peak_source": peak_src,
        "note": ("launch durations come from a second, event-instrumented pass (an event pair around every launch makes "
                 "the step a few % slower than the timed pass); at k <= 10 the X-stream launches also contain the fused "
                 "U / V-update tails (12.7 us each, tools/timeline.py), so `frac` is a lower bound for the stream itself; "
                 "inner_step.frac is the whole timed step against the two-stream roofline"),
        "xv": {"ms": xv_ms, "GBps": ach_xv, "frac": ach_xv / peak, "bytes": bytes_xv},
        "xtu": {"ms": xtu_ms, "GBps": ach_xtu, "frac": ach_xtu / peak, "bytes": bytes_xtu},
        "inner_step": {"ms": inner_ms, "bytes": step_bytes, "GBps": step_bytes / (inner_ms * 1e-3) / 1e9,
                       "frac": step_bytes / (inner_ms * 1e-3) / 1e9 / peak,
                       "x_stream_share_of_step": (xv_ms + xtu_ms) / inner_ms if inner_ms > 0 else None,
                       "phase_ms": phase_ms, "host_ms_per_outer": host_ms,
                       "ms_per_outer_with_phase_events": ms_profiled},
    }

    # e2e: the same outer iteration with HOST buffers (pinned X, U, V in; U, V, objective out)
    e2e = None
    if not a.no_e2e:
        Xh = torch.empty((m_local, a.n), dtype=xdt, pin_memory=True)
        Xh.copy_(Xd)
        del Xd
        Uh = torch.empty((m_local, a.k), dtype=torch.float64, pin_memory=True)
        Vh = torch.empty((a.n, a.k), dtype=torch.float64, pin_memory=True)
        Uh.copy_(torch.from_numpy(np.ascontiguousarray(U0))); Vh.copy_(torch.from_numpy(V0))
        Xn, Un, Vn = Xh.numpy(), Uh.numpy(), Vh.numpy()

        def e2e_iteration():
            eng.set_X(Xn)                       # H2D of this step's input
            eng.set_UV(Un, Vn)
            active = sample_active(full_cands, a.k)
            eng.set_active(active)
            eng.step_async(MODULUS, gamma, delta)
            p, _, _, (mass, qn, _) = eng.block_end(MODULUS, want_scores=True, prefetch=False)
            restrict_from_tables(mass, qn, full_cands)
            Uo, Vo = eng.get_UV()               # D2H of the result
            return float(p[-1, 4]), Uo, Vo

        n_e2e = max(2, min(a.steps, 5))
        e2e_iteration()
        sync_all()
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(n_e2e):
            e2e_iteration()
        e1.record(stream)
        sync_all()
        wall_ms = (time.perf_counter() - t0) * 1e3 / n_e2e
        ems = max(e0.elapsed_time(e1) / n_e2e, wall_ms)
        if ctx.world > 1:
            import torch.distributed as dist
            t = torch.tensor([ems], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
        e2e = {"value": 1000.0 / ems, "unit": UNIT, "ms_per_step": ems,
               "h2d_bytes_per_step": int(m_local * a.n * xsz + (m_local + a.n) * a.k * 8),
               "d2h_bytes_per_step": int((m_local + a.n) * a.k * 8 + MODULUS * 64 + 2 * a.k * packed.P * 8),
               "steps": n_e2e}
        Xcpu = Xn
        # context: the whole solve through the public entry point, host arrays in and out, X uploaded ONCE
        # (engine creation, pathway packing, upload, transposed copy, 16 outer iterations, download)
        if ctx.world == 1:
            import contextlib
            import io
            from prmf_b200 import nmf_pathway
            times = []
            for _ in range(2):                      # the first call also pays one-off costs (lazy kernel loading)
                np.random.seed(1)
                with contextlib.redirect_stderr(io.StringIO()):
                    t0 = time.perf_counter()
                    nmf_pathway(Xn, list(Gs), k_latent=a.k, nodelist=nodelist, max_iter=16 * MODULUS, quiet=True,
                                x_dtype=a.x_dtype)
                    times.append(time.perf_counter() - t0)
            e2e["whole_solve"] = {"outer_iterations": 16, "seconds": times[1], "seconds_first_call": times[0],
                                  "value": 16 / times[1], "unit": UNIT,
                                  "note": "nmf_pathway(X_host, graphs) -> (U, V) on the host, X uploaded once "
                                          "(engine creation, pathway packing, upload, transposed copy, 16 outer iterations, download)"}
    else:
        Xcpu = None

    cpu = None
    if ctx.rank == 0 and ctx.world == 1 and not a.no_cpu_baseline and Xcpu is not None:
        rate, detail = cpu_outer_iteration_rate(np.asarray(Xcpu, dtype=np.float64), Gs, nodelist, a.k, a.cpu_inner_steps)
        cpu = {"value": rate, "unit": UNIT, "cores": host_threads(), "kind": "port",
               "sample": "%d inner steps + 1 restrict of the numpy/scipy oracle at full shape, extrapolated to 10 + 1"
                         % a.cpu_inner_steps, **detail}

    exch_mode = eng.exchange_mode
    ctx.barrier()
    eng.close()
    if ctx.rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64" if not tf32 else "tf32 X products (fp32 accumulate), f64 updates",
            "data": "synthetic",
            "config": {"workload": workload_name(a), "step": "1 outer iteration = 10 inner steps + scores/restrict",
                       "parallelism": "rows of X,U sharded over %d GPU(s)" % a.gpus + (
                           "" if a.gpus == 1 else (", per-step sum over ranks of [X^T U | U^T U]: " + {
                               "nccl": "ncclAllReduce", "p2p-v-update": "NVLink peer loads fused into the V-update kernel",
                               "p2p-pass2": "NVLink peer loads fused into the pass-2 X-stream kernel (exchange + V update)",
                               "none": "none"}[exch_mode])),
                       "l2": "inputs larger than L2 (X block %.2f GB per GPU per pass)" % (m_local * a.n * xsz / 1e9),
                       "inner_steps_per_s": 1000.0 / inner_ms},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "final_obj": float(parts[-1, 4]),
        }
        print(json.dumps(line))
    if ctx.world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


T_START = time.perf_counter()

if __name__ == "__main__":
    import faulthandler
    faulthandler.dump_traceback_later(420, exit=True)     # never hang a GPU box: die loudly instead
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)
