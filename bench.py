#!/usr/bin/env python
"""bench.py — VB iterations/sec on Bayesian PCA (N=10M, D=64, K=16), BASELINE.json's metric.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps K --warmup W      # CPU arm (oracle port of the reference)

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
FLOPS_PER_COL = 2 * M_DIM * K_DIM * 2 + 2 * K_DIM * K_DIM + K_DIM * (K_DIM + 1)   # ~4.9 kflop


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def synth_shard(M, n0, n1, seed):
    """Columns [n0, n1) of y = w x^T + 0.1 eps (pca.rst:29-32 pattern); w is shared by all
    ranks (seed), x / eps are per-shard streams."""
    rng_w = np.random.default_rng(seed)
    w = rng_w.standard_normal((M, 4))
    rng = np.random.default_rng([seed, n0])
    n = n1 - n0
    x = rng.standard_normal((4, n))
    y = rng.standard_normal((M, n))
    y *= 0.1
    y += w @ x
    return y


# =============================================================================================
# reference arm / cpu_baseline: the oracle port of the reference's NumPy/SciPy sweep
# =============================================================================================
def cpu_sweeps(n_sample, steps, warmup, seed=1):
    from oracle.pca_ref import PcaOracle
    y = synth_shard(M_DIM, 0, n_sample, seed)
    rs = np.random.RandomState(seed)
    C0 = rs.randn(M_DIM, 1, K_DIM)
    o = PcaOracle(y, K_DIM, C0)
    for _ in range(warmup):
        o.sweep()
    t = time.perf_counter()
    for _ in range(steps):
        o.sweep()
    dt = (time.perf_counter() - t) / max(steps, 1)
    return dt


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sample = 30_000
    dt = cpu_sweeps(n_sample, args.steps, args.warmup)
    scale = N_TOTAL / n_sample
    value = 1.0 / (dt * scale)
    cores = blas_threads()
    sample = ("oracle/pca_ref.py (NumPy/SciPy port of the reference sweep) on the first %d of %d columns, "
              "%d timed sweeps; it/s extrapolated linearly in N (the reference is O(N) per sweep and needs "
              "~9 KB/col, so N=10M does not fit a practical host run); np.einsum is single-threaded, BLAS "
              "threads=%d of %d host cores" % (n_sample, N_TOTAL, args.steps, cores, os.cpu_count() or 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "it/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt * scale,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "Bayesian PCA N=1e7 M=64 K=16 fully observed, one VB sweep incl. lower bound",
                   "cpu_sample_columns": n_sample},
        "cpu_baseline": {"value": value, "unit": "it/s", "cores": cores, "kind": "port", "sample": sample},
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
    rs = np.random.RandomState(1)                      # replicated factor: identical on every rank
    C.initialize_from_value(rs.randn(M, 1, K))
    Q = VB(Y, X, C, alpha, tau, fused=fused)
    Q.ignore_bound_checks = True                       # time exactly K sweeps (no early convergence return)
    return Q, dict(X=X, C=C, alpha=alpha, tau=tau, Y=Y)


def run_gpu(args):
    from bayespy_b200 import _bpk, parallel
    world, rank = parallel.init_from_env()
    be = _bpk.get()
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
    parallel.barrier()
    l0 = be.launch_count()
    wall0 = time.perf_counter()
    be.timer_record(t_all, 0)
    Q.update(repeat=steps, verbose=False)
    be.timer_record(t_all, 1)
    be.sync()
    wall = time.perf_counter() - wall0
    launches = be.launch_count() - l0
    ms_dev = be.timer_elapsed_ms(t_all)
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
    e2e_steps = max(2, min(steps, args.e2e_steps))
    nbytes = y.nbytes
    hptr = be.host_alloc(nbytes)
    import ctypes
    pinned = np.ctypeslib.as_array((ctypes.c_double * y.size).from_address(hptr)).reshape(y.shape)
    pinned[...] = y
    Y = nodes["Y"]
    Y.observe(pinned)
    Q.update(repeat=1, verbose=False)
    parallel.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        Y.observe(pinned)                      # H2D of this step's inputs from pinned host memory
        Q.update(repeat=1, verbose=False)      # sweep + D2H of the per-node bound terms
        _ = float(Q.L[Q.iter - 1])
    be.sync()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    e2e_s = float(np.max(parallel.allgather_scalar(e2e_s)))
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
        "metric": METRIC, "value": value, "unit": "it/s", "n_gpus": world, "steps": steps, "warmup": warmup,
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
                   "host_wall_ms_per_step": 1e3 * wall / steps},
        "clocks": clocks,
        "e2e": {"value": 1.0 / e2e_s, "unit": "it/s", "h2d_bytes_per_step": int(nbytes),
                "d2h_bytes_per_step": 8 * len(Q.model), "steps": e2e_steps,
                "note": "per step: Y.observe(pinned host array) + Q.update() + read of L"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "pca_xsweep_ws_kernel<...,FUSED> (persistent launch of sweeps_per_launch VB sweeps; per sweep: data pass + grid reduction + node updates + bound)", "kernel_ms": kern_avg,
                     "kernel_ms_note": "launch duration / sweeps_per_launch (CUDA events on the library stream)",
                     "traffic_note": "DRAM bytes per sweep, ncu --set full capture of a one-sweep launch (profiles/pca_xsweep_traffic.json)",
                     "kernel_share_of_step": kern_avg / (ms_max / steps),
                     "algorithmic_bytes_per_col": BYTES_PER_COL, "peak_source": peak_src,
                     "sweeps_per_launch": sweeps_per_launch,
                     "fp64_tflops": FLOPS_PER_COL * n_local_max / (kern_avg * 1e-3) / 1e12},
    }
    if world == 1 and not args.no_cpu_baseline:
        n_sample = 50_000
        dt = cpu_sweeps(n_sample, 2, 1)
        scale = n_total / n_sample
        cores = blas_threads()
        line["cpu_baseline"] = {
            "value": 1.0 / (dt * scale), "unit": "it/s", "cores": cores, "kind": "port",
            "sample": "oracle/pca_ref.py (NumPy/SciPy port of the reference sweep) on %d of %d columns, 2 timed "
                      "sweeps after 1 warm-up, extrapolated linearly in N; BLAS threads=%d of %d host cores"
                      % (n_sample, n_total, cores, os.cpu_count() or 1)}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=N_TOTAL, help="total number of columns (default: the metric's 1e7)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--clock-interval-ms", type=int, default=200,
                    help="nvidia-smi sampling period during the timed region (B200_PROFILING.md recipe: 200); 0 = off")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
