#!/usr/bin/env python
"""bench.py — VB iterations/sec on Bayesian PCA (N=10M, D=64, K=16), BASELINE.json's metric.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps K --warmup W      # CPU arm: the unmodified reference (oracle/_ref)

One "step" = one full ``Q.update()`` sweep over [X, C, alpha, tau] INCLUDING the lower-bound
evaluation VB.update always performs (vmp.py:154-172, :713).  Model and synthetic data follow
doc/source/examples/pca.rst:40-66 / SURVEY.md §8d.  The sample axis N is block-sharded over
the ranks (strong scaling: the total N stays 10M); the only exchange per sweep is one NCCL
all-reduce of the plate-summed statistics (1 296 doubles).

Timing: W untimed sweeps, then exactly K sweeps bracketed by barrier+sync, CUDA events on the
library's compute stream, max over ranks.  Inputs (5.12 GB of Y at N=10M) are larger than the
126 MB L2, so no explicit flush is needed between iterations.
"""
import argparse
import json
import os
import sys as _sys
if "reference" in _sys.argv:
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm may use every host thread BLAS wants (fixed across N)
    os.environ.pop("OMP_NUM_THREADS", None)
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

M_DIM, K_DIM = 64, 16
N_TOTAL = 10_000_000
METRIC = "VB iterations/sec on PCA N=10M D=64 K=16"
BYTES_PER_COL = M_DIM * 8 + K_DIM * 8                       # 640 B
FLOPS_PER_COL = 2 * M_DIM * K_DIM * 2 + 2 * K_DIM * K_DIM + K_DIM * (K_DIM + 1)   # ~4.9 kflop (algorithmic)
EXEC_FLOPS_PER_COL = 8 * 512                                # the sweep kernel executes 8 DMMA.8x8x4 per column


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


BLOCK = 65536      # columns per independently seeded block of synthetic data


def synth_shard(M, n0, n1, seed):
    """Columns [n0, n1) of y = w x^T + 0.1 eps (pca.rst:29-32 pattern).  The stream is seeded per fixed block of
    65 536 columns, so column n holds the same numbers whatever the world size or the shard bounds: every SCALE
    line works on the same 10M columns, and the reference arm's sample is a prefix of the GPU arm's data."""
    w = np.random.default_rng(seed).standard_normal((M, 4))
    out = np.empty((M, n1 - n0))
    for blk in range(n0 // BLOCK, (max(n1, n0 + 1) - 1) // BLOCK + 1):
        c0, c1 = blk * BLOCK, (blk + 1) * BLOCK
        rng = np.random.default_rng([seed, blk])
        x = rng.standard_normal((4, BLOCK))
        y = rng.standard_normal((M, BLOCK))
        y *= 0.1
        y += w @ x
        lo, hi = max(c0, n0), min(c1, n1)
        if hi > lo:
            out[:, lo - n0:hi - n0] = y[:, lo - c0:hi - c0]
    return out


def init_C(M, K):
    """Replicated initial <C>: identical on every rank and in the reference arm."""
    return np.random.RandomState(1).randn(M, 1, K)


# =============================================================================================
# reference arm / cpu_baseline: the UNMODIFIED reference package (oracle/_ref, staged by
# oracle/make_ref.py) running the same model script on a prefix of the same data; falls back to
# the oracle port (oracle/pca_ref.py) only if the staged package is missing.
# =============================================================================================
def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        return os.cpu_count() or 1


def reference_sweep_seconds(n_sample, steps, warmup):
    """(seconds per sweep, kind) of the reference's Q.update() on the first n_sample columns."""
    y = synth_shard(M_DIM, 0, n_sample, 1)
    try:
        from oracle import make_ref, ref_models
        if not make_ref.available():
            raise RuntimeError("oracle/_ref not staged")
        Q, _ = ref_models.pca(y, K_DIM, init_C(M_DIM, K_DIM))
        return ref_models.time_sweeps(Q, steps, warmup), "reference"
    except Exception as e:                                   # pragma: no cover - only without oracle/_ref
        sys.stderr.write("bench: reference package unavailable (%s); timing the oracle port instead\n" % e)
        from oracle.pca_ref import PcaOracle
        o = PcaOracle(y, K_DIM, init_C(M_DIM, K_DIM))
        for _ in range(warmup):
            o.sweep()
        t = time.perf_counter()
        for _ in range(steps):
            o.sweep()
        return (time.perf_counter() - t) / max(steps, 1), "port"


def reference_sample_size(steps, warmup, budget_s):
    """Columns per step such that warmup+steps sweeps of the reference take about budget_s seconds
    (calibrated on one 10k-column sweep; the reference is O(N) per sweep, ~45 us/col on this class of host)."""
    dt, _ = reference_sweep_seconds(10_000, 1, 1)
    per_col = dt / 10_000
    n = int(budget_s / max(steps + warmup, 1) / per_col)
    n = max(20_000, min(200_000, n))
    return (n // 1000) * 1000, per_col


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_total = args.n
    n_sample, _ = reference_sample_size(args.steps, args.warmup, args.ref_budget_s)
    dt, kind = reference_sweep_seconds(n_sample, args.steps, args.warmup)
    scale = n_total / n_sample
    value = 1.0 / (dt * scale)
    cores = blas_threads()
    what = ("the unmodified reference package (oracle/_ref, staged from /root/reference by oracle/make_ref.py)"
            if kind == "reference" else "oracle/pca_ref.py (NumPy/SciPy port of the reference sweep)")
    sample = ("%s running pca.rst:40-66 on the first %d of %d columns of the GPU arm's data, %d timed Q.update() sweeps "
              "after %d warm-up; it/s extrapolated linearly in N (x%.1f: the reference is O(N) per sweep and needs "
              "~9 KB/col of host RAM, N=1e7 does not fit a practical host run); np.einsum and the per-plate loops are "
              "single-threaded, BLAS threads=%d of %d host cores"
              % (what, n_sample, n_total, args.steps, args.warmup, scale, cores, os.cpu_count() or 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "it/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "Bayesian PCA N=%d M=64 K=16 fully observed (pca.rst:40-66), one VB sweep over "
                               "[X,C,alpha,tau] incl. lower bound" % n_total,
                   "n_total": n_total, "cpu_sample_columns": n_sample, "value_is_extrapolated": True,
                   "extrapolation_factor": scale, "ms_per_step_note": "measured, one step = one sweep over the sample",
                   "ms_per_step_full_n_extrapolated": 1e3 * dt * scale},
        "cpu_baseline": {"value": value, "unit": "it/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# =============================================================================================
# GPU arm
# =============================================================================================
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, interval_ms=200):
        self.gpu = gpu_index
        self.interval_ms = int(interval_ms)
        self.proc = None
        self.path = None

    def start(self):
        if self.interval_ms <= 0:
            return
        try:
            fd, self.path = tempfile.mkstemp(prefix="bpk_clocks_", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", str(self.interval_ms)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            try:
                self.proc.kill()
            except Exception:
                pass
        try:
            sm, mx, reasons = [], [], set()
            for ln in open(self.path):
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                   f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            if sm:
                out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                       "samples": len(sm)}
        except Exception:
            pass
        try:
            os.remove(self.path)
        except Exception:
            pass
        return out


def build_model(y_host, fused=True):
    from bayespy_b200.nodes import GaussianARD, Gamma, SumMultiply
    from bayespy_b200.inference import VB
    M, N = y_host.shape
    K = K_DIM
    X = GaussianARD(0, 1, plates=(1, N), shape=(K,), name="X")
    alpha = Gamma(1e-5, 1e-5, plates=(K,), name="alpha")
    C = GaussianARD(0, alpha, plates=(M, 1), shape=(K,), name="C")
    F = SumMultiply("d,d->", X, C, name="F")
    tau = Gamma(1e-5, 1e-5, name="tau")
    Y = GaussianARD(F, tau, name="Y")
    Y.observe(y_host)
    C.initialize_from_value(init_C(M, K))              # replicated factor: identical on every rank
    Q = VB(Y, X, C, alpha, tau, fused=fused)
    Q.ignore_bound_checks = True                       # time exactly K sweeps (no early convergence return)
    return Q, dict(X=X, C=C, alpha=alpha, tau=tau, Y=Y)


def expected_bound(n_total, sweeps):
    """Lower bound after `sweeps` sweeps as measured on ONE GPU (profiles/bench_expected.json), so that every
    multi-GPU line can show that N ranks computed what one rank computes."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "bench_expected.json")))
        return float(d["pca"]["%d:%d" % (n_total, sweeps)])
    except Exception:
        return None


def run_gpu(args):
    from bayespy_b200 import _bpk, parallel
    world, rank = parallel.init_from_env()
    be = _bpk.get()
    cpus = parallel.bind_to_gpu_numa()          # host threads and pinned buffers next to this rank's GPU
    if world != args.gpus and rank == 0:
        sys.stderr.write("warning: --gpus %d but WORLD_SIZE=%d\n" % (args.gpus, world))
    n_total = args.n
    n0, n1 = parallel.shard_bounds(n_total, world, rank)
    y = synth_shard(M_DIM, n0, n1, seed=1)
    Q, nodes = build_model(y)
    plan = Q.plans[0]
    steps, warmup = args.steps, max(args.warmup, 3)

    Q.update(repeat=warmup, verbose=False)
    parallel.barrier()

    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")), args.clock_interval_ms)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    plan.kernel_timers = [be.timer_create() for _ in range(steps)]
    plan._timer_pos = 0
    plan.kernel_timer_log = []
    t_all = be.timer_create()
    parallel.barrier_aligned()               # barrier + common start instant (one node: shared monotonic clock)
    l0 = be.launch_count()
    wall0 = time.perf_counter()
    be.timer_record(t_all, 0)
    Q.update(repeat=steps, verbose=False)
    be.timer_record(t_all, 1)
    be.sync()
    wall = time.perf_counter() - wall0
    launches = be.launch_count() - l0
    ms_dev = be.timer_elapsed_ms(t_all)
    host_timing = dict(getattr(plan, "host_timing", None) or {})      # host time around the launches of the timed region
    parallel.barrier()
    clocks = sampler.stop() if rank == 0 else None
    # the host drives the sweep synchronously (one bound read-back per sweep), so the slower of the
    # device-event span and the host wall clock is the honest step time; max over ranks
    ms_rank = max(ms_dev, 1e3 * wall)
    ms_max = float(np.max(parallel.allgather_scalar(ms_rank)))
    # sweep-kernel time per sweep: a launch may run a whole chunk of sweeps (timer brackets the launch)
    if getattr(plan, "kernel_timer_log", None):
        tot_ms = sum(be.timer_elapsed_ms(t) for ids, _ in plan.kernel_timer_log for t in ids)
        tot_sw = sum(n for _, n in plan.kernel_timer_log if n)
        kern_avg = tot_ms / max(tot_sw, 1)
        sweeps_per_launch = tot_sw / max(sum(len(ids) for ids, _ in plan.kernel_timer_log), 1)
    else:
        kern_ms = [be.timer_elapsed_ms(t) for t in plan.kernel_timers[:plan._timer_pos]]
        kern_avg = float(np.mean(kern_ms)) if kern_ms else float("nan")
        sweeps_per_launch = 1.0
    plan.kernel_timers = None
    kern_avg = float(np.max(parallel.allgather_scalar(kern_avg)))
    L_last = float(Q.L[Q.iter - 1])

    # ---- e2e: host buffers in, host scalars out, every step (public API: observe + update) ----
    e2e_steps = max(10, args.e2e_steps)
    nbytes = y.nbytes
    hptr = be.host_alloc(nbytes)
    import ctypes
    pinned = np.ctypeslib.as_array((ctypes.c_double * y.size).from_address(hptr)).reshape(y.shape)
    pinned[...] = y
    Y = nodes["Y"]
    for _ in range(2):                         # warm-up: pinned pages touched, copy engine and pool warm
        Y.observe(pinned)
        Q.update(repeat=1, verbose=False)
    e2e_t = []
    for _ in range(e2e_steps):
        parallel.barrier()
        t0 = time.perf_counter()
        Y.observe(pinned)                      # H2D of this step's inputs from pinned host memory
        Q.update(repeat=1, verbose=False)      # sweep + D2H of the per-node bound terms
        _ = float(Q.L[Q.iter - 1])
        be.sync()
        e2e_t.append(time.perf_counter() - t0)
    # per step the slowest rank counts; the reported figure is the median step
    e2e_mat = np.array([parallel.allgather_scalar(t) for t in e2e_t])          # (steps, world)
    e2e_s = float(np.median(e2e_mat.max(axis=1)))
    e2e_best = float(e2e_mat.max(axis=1).min())
    be.host_free(hptr)

    if rank != 0:
        return
    peak, peak_src = load_peaks()
    n_local_max = parallel.shard_bounds(n_total, world, 0)[1]
    achieved = BYTES_PER_COL * n_local_max / (kern_avg * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "pca_xsweep_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    value = steps / (ms_max * 1e-3)
    line = {
        "metric": METRIC if n_total == N_TOTAL else "VB iterations/sec on PCA N=%d D=64 K=16" % n_total, "value": value, "unit": "it/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_max / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "Bayesian PCA N=%d M=64 K=16 fully observed (pca.rst:40-66), one VB sweep over "
                               "[X,C,alpha,tau] incl. lower bound" % n_total,
                   "n_total": n_total, "n_per_gpu": n_local_max, "parallelism": "plate-shard x%d" % world,
                   "l2": "inputs (%.2f GB of Y per GPU) larger than the 126 MB L2; no flush" % (y.nbytes / 1e9),
                   "launch_note": "a chunk of up to 50 VB sweeps is ONE persistent launch of the fused sweep kernel "
                                  "(data pass + grid reduction + node updates + bound per sweep, grid barriers in between); "
                                  "gpu_launches counts launches, roofline.sweeps_per_launch the sweeps inside each",
                   "lower_bound_last": L_last, "device_ms_per_step": ms_dev / steps,
                   "host_wall_ms_per_step": 1e3 * wall / steps,
                   "host_us_around_launches": {k: (round(v, 1) if isinstance(v, float) else v)
                                               for k, v in host_timing.items()}},
        "clocks": clocks,
        "e2e": {"value": 1.0 / e2e_s, "unit": "it/s", "h2d_bytes_per_step": int(nbytes),
                "d2h_bytes_per_step": 8 * len(Q.model), "steps": e2e_steps, "statistic": "median over steps of the max over ranks",
                "best_step_its": 1.0 / e2e_best, "h2d_gbs_per_gpu": nbytes / e2e_s / 1e9,
                "numa_bound_cpus": (len(cpus) if cpus else None),
                "note": "per step: Y.observe(pinned host array) + Q.update() + read of L; the step is the H2D copy of "
                        "Y (PCIe-bound: %.2f GB per GPU per step) followed by a %.1f ms sweep" % (nbytes / 1e9, ms_max / steps)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "pca_xsweep_ws_kernel<...,FUSED> (persistent launch of sweeps_per_launch VB sweeps; per sweep: data pass + grid reduction + node updates + bound)", "kernel_ms": kern_avg,
                     "kernel_ms_note": "launch duration / sweeps_per_launch (CUDA events on the library stream)",
                     "traffic_note": "DRAM bytes per sweep, ncu --set full capture of a one-sweep launch (profiles/pca_xsweep_traffic.json)",
                     "kernel_share_of_step": kern_avg / (ms_max / steps),
                     "algorithmic_bytes_per_col": BYTES_PER_COL, "peak_source": peak_src,
                     "sweeps_per_launch": sweeps_per_launch,
                     "fp64_tflops_algorithmic": FLOPS_PER_COL * n_local_max / (kern_avg * 1e-3) / 1e12,
                     "fp64_tflops_executed": EXEC_FLOPS_PER_COL * n_local_max / (kern_avg * 1e-3) / 1e12,
                     "fp64_note": "algorithmic = %d flop/col (SURVEY 8d); executed = 8 DMMA.8x8x4 x 512 flop per column "
                                  "(S_xx is derived in the tail from S_yx and s_x); measured DMMA peak 36.9 TFLOP/s "
                                  "(profiles/r01_ubench_*)" % FLOPS_PER_COL},
    }
    exp = expected_bound(n_total, warmup + steps)
    if exp is not None:
        line["config"]["lower_bound_expected_1gpu"] = exp
        line["config"]["lower_bound_rel_err_vs_1gpu"] = abs(L_last - exp) / abs(exp)
    if world == 1 and not args.no_cpu_baseline:
        n_sample = 100_000
        dt, kind = reference_sweep_seconds(n_sample, 2, 1)
        scale = n_total / n_sample
        cores = blas_threads()
        line["cpu_baseline"] = {
            "value": 1.0 / (dt * scale), "unit": "it/s", "cores": cores, "kind": kind,
            "sample": "%s on the first %d of %d columns of the same data, 2 timed Q.update() sweeps after 1 warm-up "
                      "(%.1f s per sweep measured), extrapolated linearly in N (x%.0f); BLAS threads=%d of %d host cores, "
                      "the einsum / per-plate loops of the reference are single-threaded"
                      % ("unmodified reference package (oracle/_ref)" if kind == "reference" else "oracle port",
                         n_sample, n_total, dt, scale, cores, os.cpu_count() or 1)}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="pca", choices=["pca", "pca_masked", "gmm", "lssm"],
                    help="pca = the headline metric (BASELINE.json configs[1] scaled to N=1e7); gmm / lssm = configs[2] / [3] "
                         "(tools/bench_workloads.py)")
    ap.add_argument("--n", "--columns", dest="n", type=int, default=N_TOTAL,
                    help="total number of columns / rows / time steps of the workload (default: the metric's 1e7; under torchrun "
                         "spell it --columns: torchrun's own parser claims the prefix --n)")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--ref-budget-s", type=float, default=150.0,
                    help="--impl reference: wall-clock budget of the whole run; sets the column sample per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--clock-interval-ms", type=int, default=200,
                    help="nvidia-smi sampling period during the timed region (B200_PROFILING.md recipe: 200); 0 = off")
    args = ap.parse_args()
    if args.workload != "pca":
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_workloads as bw
        if args.impl == "reference":
            {"gmm": bw.run_gmm_reference}.get(args.workload, lambda a: print(json.dumps(
                {"impl": "reference", "unavailable": "no reference arm for workload %s" % a.workload})))(args)
        else:
            {"gmm": bw.run_gmm, "lssm": bw.run_lssm, "pca_masked": bw.run_pca_masked}[args.workload](args)
        return
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
