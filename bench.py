#!/usr/bin/env python
"""Benchmark of the GP fit+predict hot path (BASELINE.json metric: "GP fit+predict sec").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c3|c1|c5|...]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (one rank per GPU)

A step = MultiOutputGP fit(thetas) + predict(Xs, unc=True) of the whole job.  Default workload = BASELINE.json
configs[2] (C3): 32 outputs x n=4096 x d=10 SqExp, nugget 1e-6, 10000 test points -- the configuration the
headline target is quoted on.  With N ranks the 32 outputs are block-partitioned (strong scaling) and predict
ends with the single NCCL all-gather.  One JSON line on rank 0.

  value : seconds per step with the design matrix and targets already resident in HBM (an existing emulator
          object; per step only thetas / test points go in and posteriors come out)
  e2e   : seconds per step through the public API from host arrays: construct the emulator (H2D of X, Y),
          fit, predict, posteriors back on the host
  --impl reference : the reference's CPU algorithm (oracle port of mogp_emulator's numpy/scipy path) timed on
          this box's host cores on a bounded sample (one output per step, scaled to the job)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (outputs, n, d, m, kernel, nugget, seed)      SURVEY.md section 8d
    "c1": (1, 256, 4, 1000, "SquaredExponential", 1.0e-6, 0),
    "c3": (32, 4096, 10, 10000, "SquaredExponential", 1.0e-6, 2),
    "c3a": (32, 4096, 10, 10000, "SquaredExponential", "adaptive", 2),  # C3 with the API-default nugget (adaptive -> 0.0 here)
    "c4": (1, 16384, 20, 1000, "Matern52", "adaptive", 3),
    "c5": (256, 8192, 15, 10000, "SquaredExponential", 1.0e-6, 4),
    "c2": (1, 4096, 10, 10000, "SquaredExponential", 1.0e-6, 1),       # fit_GP_MAP (run_c2)
    "c2s": (1, 4096, 10, 10000, "SquaredExponential", 1.0e-6, 1),      # one output of the C2 shape: fit + predict
    "c3x4": (4, 4096, 10, 10000, "SquaredExponential", 1.0e-6, 2),     # one rank's share of C3 at 8 GPUs
    "tiny": (4, 512, 5, 700, "SquaredExponential", 1.0e-6, 9),
}


def make_workload(n, d, n_out, m, seed):
    """X~U[0,1)^d, Y[k] = sin(2*sum(x)+k) + 0.01*N(0,1), Xs~U[0,1)^d."""
    rng = np.random.default_rng(seed)
    X = rng.random((n, d))
    Y = np.stack([np.sin(2.0 * X.sum(axis=1) + k) + 0.01 * rng.standard_normal(n) for k in range(n_out)])
    Xs = rng.random((m, d))
    return X, Y, Xs


def make_thetas(n_out, d):
    """theta_corr = 1 + 0.01*k (a distinct kernel matrix per output, as after a MAP fit), theta_cov = 0."""
    thetas = np.zeros((n_out, d + 1))
    thetas[:, :d] = 1.0 + 0.01 * np.arange(n_out)[:, None]
    return thetas


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        busy = [v for v in sm if v > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def oracle_step(X, y, Xs, theta, kernel, nugget):
    """One output of the job on the CPU with the reference's algorithm (oracle port)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gp_oracle as orc
    t0 = time.perf_counter()
    gp = orc.OracleGP(X, y, kernel=kernel, nugget=nugget, priors="weak", chunked=X.shape[0] > 8192).fit(theta)
    mean, var = gp.predict(Xs)
    return time.perf_counter() - t0, mean, var


def _pool_predict(gp, Xs):
    return gp.predict(Xs)


def run_reference(args, wl):
    """--impl reference: the reference's CPU algorithm (oracle port) on the host cores; rank 0 only.

    BASELINE.md section 4.4: MultiOutputGP.fit (a serial loop over emulators, MultiOutputGP.py:331-360) + predict
    (a multiprocessing Pool over emulators, MultiOutputGP.py:306-309) on min(E, 8) outputs, scaled linearly to E.  A step is
    a bounded sample of that: the fit of ONE output (the loop is serial, so its cost is per output) plus the POOLED predict of
    k = min(E, 8) fitted outputs at ceil(m / k) test points each -- the flops of one output's predict (linear in m), run the
    way the reference runs them, k emulators at a time in k worker processes.  value = step time x E."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing
    cores = os.cpu_count()
    try:        # torchrun exports OMP_NUM_THREADS=1: give the BLAS its cores back so every N times the same baseline
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=cores)
    except ImportError:
        pass
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gp_oracle as orc
    E, n, d, m, kernel, nugget, seed = wl
    k = min(E, 8)
    X, Y, Xs = make_workload(n, d, k, m, seed)
    thetas = make_thetas(E, d)
    chunked = n > 8192
    t_setup = time.perf_counter()
    gps = [orc.OracleGP(X, Y[j], kernel=kernel, nugget=nugget, priors="weak", chunked=chunked).fit(thetas[j]) for j in range(k)]
    t_setup = time.perf_counter() - t_setup
    mk = -(-m // k)
    chunks = [Xs[j * mk:(j + 1) * mk] for j in range(k)]
    times, fit_t, pred_t = [], [], []
    for s_ in range(args.warmup + args.steps):
        j = s_ % k
        t0 = time.perf_counter()
        gps[j] = orc.OracleGP(X, Y[j], kernel=kernel, nugget=nugget, priors="weak", chunked=chunked).fit(thetas[j])
        t1 = time.perf_counter()
        if k > 1:
            with multiprocessing.Pool(None) as pool:          # processes=None, as MultiOutputGP.predict is called
                pool.starmap(_pool_predict, [(gps[q], chunks[q]) for q in range(k) if len(chunks[q])])
        else:
            gps[0].predict(Xs)
        t2 = time.perf_counter()
        if s_ >= args.warmup:
            times.append(t2 - t0)
            fit_t.append(t1 - t0)
            pred_t.append(t2 - t1)
    per_output = float(np.mean(times))
    value = per_output * E
    line = {
        "impl": "reference", "metric": "gp_fit_predict_seconds", "value": value, "unit": "s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": value * 1e3, "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, wl, args.gpus),
        "cpu_baseline": {"value": value, "unit": "s", "cores": cores, "kind": "port",
                         "sample": "each step = fit of 1 of the %d outputs (serial loop of MultiOutputGP.fit; mean %.2f s) + pooled predict "
                                   "of %d fitted outputs at %d test points each in a multiprocessing.Pool(None) (= one output's predict "
                                   "flops, run %d emulators at a time as MultiOutputGP.predict does; mean %.2f s), oracle port of the "
                                   "reference's numpy/scipy path, BLAS on %d threads; value = mean step time x %d outputs; the %d "
                                   "emulators of the pool were fitted once outside the timed steps (%.1f s)"
                                   % (E, float(np.mean(fit_t)), k, mk, k, float(np.mean(pred_t)), cores, E, k, t_setup)},
        "seconds_per_output": {"fit_serial": float(np.mean(fit_t)), "predict_pooled": float(np.mean(pred_t))},
        "e2e": {"value": value, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def i8_planes():
    """Planes per operand of the int8 predict path as mogp_create reads MOGP_TRSM_I8: '6' -> 6, anything else that enables it -> 7."""
    return 6 if os.environ.get("MOGP_TRSM_I8", "7")[:1] == "6" else 7


def ncu_traffic(workload, world):
    """DRAM bytes (read + write) of one launch of the dominant kernel from the committed `ncu --set full` capture
    of this same command (profiles/r02_traffic.json); None when that workload / GPU count was not captured."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            return json.load(f).get("%s@%d" % (workload, world))
    except (OSError, ValueError):
        return None


_I8_PEAKS = {}


def i8_peaks(libmogp, device):
    """int8 tcgen05 issue peak of this device, measured in this run the way MEASURED_PEAKS.json measures its bf16 figures (which
    has no int8 entry): BURST = best of 10 short runs of the M128 N256 K32 probe on an idle chip, SUSTAINED = the same probe back
    to back for ~3 s, mean rate of the last 2 s (the chip is then on its power cap, like the kernels inside a benchmark step)."""
    if device in _I8_PEAKS:
        return _I8_PEAKS[device]
    burst = max(libmogp.peak_i8_tops(device)[0] for _ in range(10))
    rates, t0 = [], time.perf_counter()
    while time.perf_counter() - t0 < 3.0:
        rates.append((time.perf_counter() - t0, libmogp.peak_i8_tops(device, iters=100000)[0]))
    late = [r for t, r in rates if t >= 1.0] or [r for t, r in rates]
    _I8_PEAKS[device] = {"burst": burst, "sustained": float(np.mean(late)), "sustained_samples": len(late)}
    return _I8_PEAKS[device]


def measured_bf16(key="bf16_tflops"):
    """Driver-measured dense bf16 burst (or sustained) TFLOP/s of this pool's B200s (MEASURED_PEAKS.json), for context only."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)[key])
    except (OSError, ValueError, KeyError):
        return 1590.0 if key == "bf16_tflops" else 1330.0


def run_c2(args, wl):
    """C2: single-output n=4096 d=10 SqExp, fit_GP_MAP (L-BFGS-B, maxiter 20, theta0 = 0, default priors) on one GPU.
    A step = one fit_GP_MAP call on an existing emulator; e2e = construct + fit_GP_MAP + predict(m) from host arrays."""
    from mogp_emulator_b200 import GaussianProcessGPU, fit_GP_MAP, libmogp
    from mogp_emulator_b200.rendezvous import env_rank_world
    rank, world, local_rank = env_rank_world()
    if rank != 0:
        return      # replicas only: a single-output optimisation does not shard
    if not libmogp.gpu_usable():
        raise RuntimeError("bench.py: libmogp_b200 not loaded or no sm_100 device: " + libmogp.last_error())
    E, n, d, m, kernel, nugget, seed = wl
    X, Y, Xs = make_workload(n, d, 1, m, seed)
    y = Y[0]
    theta0 = np.zeros(d + 1)
    opts = dict(n_tries=1, theta0=theta0, maxiter=20)
    gp = GaussianProcessGPU(X, y, kernel=kernel, nugget=nugget, device=local_rank)
    gp.priors     # default priors built once (host root finds), outside the timed region like the reference's __init__
    for _ in range(args.warmup):
        gp.theta = None
        fit_GP_MAP(gp, **opts)
    gp._handle.timings(reset=True)
    gp.n_fit_calls = gp.n_grad_calls = 0
    sampler = ClockSampler(local_rank)
    sampler.start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        gp.theta = None
        fit_GP_MAP(gp, **opts)
    per_step = (time.perf_counter() - t0) / args.steps
    clocks = sampler.stop()
    tm = gp._handle.timings(reset=True)
    nfit, ngrad = gp.n_fit_calls / args.steps, gp.n_grad_calls / args.steps
    theta_map = gp.theta.get_data().copy()
    gp.close()
    e2e_steps = 0 if args.no_e2e else max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        gp2 = GaussianProcessGPU(X, y, kernel=kernel, nugget=nugget, device=local_rank)
        fit_GP_MAP(gp2, **opts)
        res = gp2.predict(Xs, unc=True, deriv=False)
        gp2.close()
    e2e = (time.perf_counter() - t0) / e2e_steps if e2e_steps else None
    peak = libmogp.peak_dmma_tflops(local_rank)
    flops_eval = float(n) ** 3          # n^3/3 factor + 2n^3/3 inverse
    dev_ms = tm["fit_ms"] + tm["grad_ms"]
    achieved = flops_eval * (ngrad * args.steps) / (dev_ms * 1e-3) * 1e-12 if dev_ms else None
    line = {
        "metric": "gp_fit_map_seconds", "value": per_step, "unit": "s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": False, "scaling": "replicas only",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "c2: GaussianProcess n=%d d=%d %s nugget=%s, fit_GP_MAP(n_tries=1, theta0=0, L-BFGS-B maxiter=20, "
                               "default priors)" % (n, d, kernel, nugget), "n": n, "d": d, "m": m, "seed": seed,
                   "l2_policy": "each evaluation rewrites and factorises a 134 MB matrix (> 126 MB L2)"},
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": "s", "h2d_bytes_per_step": int(8 * (X.size + y.size + Xs.size + nfit * (d + 1))),
                "d2h_bytes_per_step": int(8 * (2 * m + nfit * 4 + ngrad * (d + 1)))},
        "gpu_launches": int(tm["n_launches"]),
        "evaluations_per_step": {"fit": nfit, "gradient": ngrad},
        "ms_per_evaluation": {"fit_device": tm["fit_ms"] / max(nfit * args.steps, 1), "cholesky": tm["chol_ms"] / max(nfit * args.steps, 1),
                              "gradient_device": tm["grad_ms"] / max(ngrad * args.steps, 1),
                              "wall": per_step * 1e3 / max(ngrad, 1)},
        "roofline": {"bound": "tensor", "kernel": "factor + L^-1 (TRSM on identity) + K^-1 tile reduction per evaluation",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": (achieved / peak) if achieved else None,
                     "traffic": None, "flops_per_evaluation": flops_eval,
                     "note": "single n=4096 factorisation is dependency-chain bound (32 block columns), see DESIGN.md"},
        "theta_map": theta_map.tolist(),
    }
    if not args.no_cpu:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import gp_oracle as orc
        t0 = time.perf_counter()
        ref = orc.OracleGP(X, y, kernel=kernel, nugget=nugget)
        lp = ref.logposterior(theta_map)
        g = ref.logpost_deriv(theta_map)
        dt = time.perf_counter() - t0
        gp3 = GaussianProcessGPU(X, y, kernel=kernel, nugget=nugget, device=local_rank)
        glp, gg = gp3.logposterior(theta_map), gp3.logpost_deriv(theta_map)
        gp3.close()
        line["cpu_baseline"] = {"value": dt * ngrad, "unit": "s", "cores": os.cpu_count(), "kind": "port",
                                "sample": "one logposterior + logpost_deriv evaluation at the MAP theta with the oracle port "
                                          "(%.2f s; its gradient uses one explicit inverse, O(n^3) -- the reference as written "
                                          "runs logdet_deriv through O(n) LAPACK calls per parameter under scipy >= 1.15 and "
                                          "is far slower), scaled x%.1f evaluations" % (dt, ngrad)}
        line["parity_vs_cpu_sample"] = {"logpost_rel": float(abs(glp - lp) / abs(lp)),
                                        "grad_max_rel": float(np.max(np.abs(gg - g)) / np.max(np.abs(g)))}
    print(json.dumps(line))


def other_configs(device, peak_dmma):
    """The other BASELINE configs that fit one GPU, each in well under a second of GPU time, so that the driver's record
    carries them next to the headline line: C1 (latency-bound small case), C2 (per-evaluation cost of the
    MAP objective and a capped fit_GP_MAP), C4 (n = 16384 adaptive-nugget factorisation: Cholesky TFLOP/s against the DMMA
    peak).  Device-resident timings of an existing emulator, parity of each against the oracle where the oracle is cheap."""
    from mogp_emulator_b200 import GaussianProcessGPU, fit_GP_MAP
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gp_oracle as orc
    out = {}
    # ---- C1 ----
    E, n, d, m, kernel, nugget, seed = WORKLOADS["c1"]
    X, Y, Xs = make_workload(n, d, 1, m, seed)
    theta = make_thetas(1, d)[0]
    gp = GaussianProcessGPU(X, Y[0], kernel=kernel, nugget=nugget, device=device)
    for _ in range(5):
        gp.fit(theta)
        r = gp.predict(Xs, unc=True, deriv=False)
    t0 = time.perf_counter()
    for _ in range(50):
        gp.fit(theta)
        r = gp.predict(Xs, unc=True, deriv=False)
    c1 = (time.perf_counter() - t0) / 50
    ref = orc.OracleGP(X, Y[0], kernel=kernel, nugget=nugget, priors="weak").fit(theta)
    t0 = time.perf_counter()
    ref = orc.OracleGP(X, Y[0], kernel=kernel, nugget=nugget, priors="weak").fit(theta)
    rm, rv = ref.predict(Xs)
    c1_cpu = time.perf_counter() - t0
    gp.close()
    out["c1"] = {"workload": "GaussianProcess n=%d d=%d %s nugget=%s, fit+predict(%d)" % (n, d, kernel, nugget, m),
                 "seconds_per_step": c1, "cpu_port_seconds": c1_cpu, "bound": "launch latency (7 kernels)",
                 "mean_max_rel_vs_cpu": float(np.max(np.abs(r.mean - rm)) / np.max(np.abs(rm))),
                 "var_max_abs_vs_cpu": float(np.max(np.abs(r.unc - rv)))}
    # ---- C2 ----
    E, n, d, m, kernel, nugget, seed = WORKLOADS["c2"]
    X, Y, Xs = make_workload(n, d, 1, m, seed)
    gp = GaussianProcessGPU(X, Y[0], kernel=kernel, nugget=nugget, device=device)
    gp.priors
    th = np.zeros(d + 1)
    for k in range(2):
        gp.logposterior(th + 0.01 * k)
        gp.logpost_deriv(th + 0.01 * k)
    gp._handle.timings(reset=True)
    t0 = time.perf_counter()
    for k in range(5):
        gp.logposterior(th + 0.1 + 0.01 * k)
        gp.logpost_deriv(th + 0.1 + 0.01 * k)
    ev = (time.perf_counter() - t0) / 5
    tm = gp._handle.timings(reset=True)
    t0 = time.perf_counter()
    gp.theta = None
    fit_GP_MAP(gp, n_tries=1, theta0=th, maxiter=20)
    c2_map = time.perf_counter() - t0
    gp.close()
    out["c2"] = {"workload": "GaussianProcess n=%d d=%d %s nugget=%s: logposterior + logpost_deriv; fit_GP_MAP(n_tries=1, theta0=0, "
                             "maxiter=20)" % (n, d, kernel, nugget),
                 "ms_per_evaluation": ev * 1e3, "device_ms_per_evaluation": {"fit": tm["fit_ms"] / 5, "cholesky": tm["chol_ms"] / 5,
                                                                             "gradient": tm["grad_ms"] / 5},
                 "fit_GP_MAP_seconds": c2_map, "flops_per_evaluation": float(n) ** 3,
                 "tflops": float(n) ** 3 / ((tm["fit_ms"] + tm["grad_ms"]) / 5 * 1e-3) * 1e-12,
                 "bound": "dependency chain of a single n=4096 factorisation (32 block columns)"}
    # ---- C4 ----
    E, n, d, m, kernel, nugget, seed = WORKLOADS["c4"]
    X, Y, Xs = make_workload(n, d, 1, m, seed)
    theta = make_thetas(1, d)[0]
    gp = GaussianProcessGPU(X, Y[0], kernel=kernel, nugget=nugget, device=device)
    gp.fit(theta)
    gp.predict(Xs, unc=True, deriv=False)
    gp._handle.timings(reset=True)
    t0 = time.perf_counter()
    for _ in range(2):
        gp.fit(theta)
        r = gp.predict(Xs, unc=True, deriv=False)
    c4 = (time.perf_counter() - t0) / 2
    tm = gp._handle.timings(reset=True)
    chol_tf = (n ** 3 / 3.0) / (tm["chol_ms"] / 2 * 1e-3) * 1e-12
    out["c4"] = {"workload": "GaussianProcess n=%d d=%d %s nugget=%s, fit+predict(%d)" % (n, d, kernel, nugget, m),
                 "seconds_per_step": c4, "nugget_chosen": float(gp.nugget),
                 "device_ms": {"kmat": tm["kmat_ms"] / 2, "cholesky": tm["chol_ms"] / 2, "fit_solves": tm["solve_ms"] / 2,
                               "kstar_and_mean": tm["kstar_ms"] / 2, "predict_trsm": tm["trsm_ms"] / 2},
                 "cholesky_tflops": chol_tf, "cholesky_frac_of_dmma_peak": chol_tf / peak_dmma,
                 "cholesky_path": ("int8 tcgen05 history products (8 planes per operand), FP64 diagonal tiles and epilogue: "
                                   "FP64-EQUIVALENT TFLOP/s") if tm.get("chol_i8_outputs", 0) > 0 else "FP64 DMMA",
                 "kmat_GBps": (8.0 * n * n + 8.0 * n * d) / (tm["kmat_ms"] / 2 * 1e-3) * 1e-9,
                 "all_finite": bool(np.all(np.isfinite(r.mean)) and np.all(np.isfinite(r.unc)))}
    gp.close()
    return out


def workload_config(name, wl, gpus):
    E, n, d, m, kernel, nugget, seed = wl
    return {"workload": "%s: MultiOutputGP %d outputs x n=%d x d=%d %s, nugget=%s, fit(thetas)+predict(%d points, unc=True)"
                        % (name, E, n, d, kernel, nugget, m),
            "outputs": E, "n": n, "d": d, "m": m, "kernel": kernel, "nugget": nugget, "seed": seed,
            "parallelism": "outputs block-partitioned over %d rank(s), one all-gather of posteriors" % gpus,
            "l2_policy": "working set (%.1f GB of factors + workspace per step) exceeds the 126 MB L2; no flush needed"
                         % (E * n * n * 8 / 1e9)}


def cholesky_report(libmogp, device, tm, n_outputs, n, steps, peak_dmma):
    """The factorisation phase of the timed steps: which kernel ran, FP64(-equivalent) TFLOP/s against the DMMA issue peak, and
    for the tcgen05 path its int8 tensor ops against the int8 issue peak (DESIGN.md section 3 states the count: per ROW / DIAG
    tile of block column j, 4 j K-steps x 36 plane pairs x 2*128*64*32)."""
    if not tm.get("chol_ms"):
        return {"path": "none"}
    ms = tm["chol_ms"] / steps
    T = (n + 127) // 128
    tf = n_outputs * (n ** 3) / 3.0 / (ms * 1e-3) * 1e-12
    out = {"ms": ms, "outputs": n_outputs, "fp64_equivalent_tflops": tf, "vs_dmma_peak": tf / peak_dmma, "dmma_peak_tflops": peak_dmma}
    if tm.get("chol_i8_outputs", 0) > 0:
        ops = n_outputs * sum(2 * (T - j) * 4 * j for j in range(T)) * 36 * 2.0 * 128 * 64 * 32
        pk = i8_peaks(libmogp, device)
        tops = ops / (ms * 1e-3) * 1e-12
        out.update(path="int8 tcgen05 (chol_i8_kernel: 8 signed 7-bit planes per operand, pairs t + u <= 9; diagonal tiles, "
                        "triangular solves against inv(L_jj) and recombination in FP64)",
                   int8_tops=tops, int8_peak_tops=pk["sustained"], int8_frac=tops / pk["sustained"],
                   int8_peak_burst_tops=pk["burst"], int8_frac_of_burst=tops / pk["burst"],
                   dram_bytes_per_launch_ncu=ncu_traffic(("c4" if n_outputs == 1 and n == 16384 else "c3" if n == 4096 and n_outputs == 32 else "-") + "-chol-i8", 1),
                   failures_rechecked_in_fp64=tm.get("chol_i8_failures_rechecked", 0.0),
                   failures_overturned_by_fp64=tm.get("chol_i8_failures_overturned", 0.0))
    else:
        out["path"] = "FP64 DMMA (chol_dataflow_kernel)"
    return out


def run_b200(args, wl):
    from mogp_emulator_b200 import MultiOutputGP_GPU, libmogp
    from mogp_emulator_b200.rendezvous import init_comm, env_rank_world
    from mogp_emulator_b200.sharding import shard_bounds

    rank, world, local_rank = env_rank_world()
    if not libmogp.gpu_usable():
        raise RuntimeError("bench.py: libmogp_b200 not loaded or no sm_100 device: " + libmogp.last_error())
    E, n, d, m, kernel, nugget, seed = wl
    X, Y, Xs = make_workload(n, d, E, m, seed)
    thetas = make_thetas(E, d)
    comm = init_comm(local_rank)
    device = local_rank

    def sync_max(t):
        return comm.allreduce_max(t) if comm is not None else t

    def barrier():
        if comm is not None:
            comm.allreduce_max(0.0)

    wall = {"fit": 0.0, "predict": 0.0}

    def step(gp):
        t_a = time.perf_counter()
        gp.fit(thetas)
        t_b = time.perf_counter()
        res = gp.predict(Xs, unc=True, deriv=False)      # the BASELINE metric is posterior mean + variance
        wall["fit"] += t_b - t_a
        wall["predict"] += time.perf_counter() - t_b
        return res

    gp = MultiOutputGP_GPU(X, Y, kernel=kernel, nugget=nugget, device=device, comm=comm)
    lo, hi, _ = shard_bounds(E, rank, world)
    for _ in range(args.warmup):
        res = step(gp)
    gp.timings(reset=True)
    wall["fit"] = wall["predict"] = 0.0
    sampler = ClockSampler(device)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = step(gp)
    elapsed = time.perf_counter() - t0     # every call synchronises its device before returning
    barrier()
    elapsed = sync_max(elapsed)
    tm = gp.timings(reset=True)
    clocks = sampler.stop() if rank == 0 else None
    per_step = elapsed / args.steps
    wall_ms = {k: v / args.steps * 1e3 for k, v in wall.items()}
    # slowest rank per device phase (ranks differ through per-GPU clocks under the power cap)
    phase_max = {k: sync_max(tm[k]) / args.steps for k in ("fit_ms", "kstar_ms", "trsm_ms")}

    # end to end from host arrays: construct (H2D of X and Y) + fit + predict (+ gather) + posteriors on the host
    gp.close()
    del gp
    e2e_steps = 0 if args.no_e2e else max(1, min(args.steps, 3))
    barrier()
    t0 = time.perf_counter()
    e2e_iters = []
    for _ in range(e2e_steps):
        t_i = time.perf_counter()
        gp2 = MultiOutputGP_GPU(X, Y, kernel=kernel, nugget=nugget, device=device, comm=comm)
        t_c = time.perf_counter()
        res2 = step(gp2)
        t_s = time.perf_counter()
        gp2.close()
        del gp2
        e2e_iters.append([round((t_c - t_i) * 1e3, 2), round((t_s - t_c) * 1e3, 2), round((time.perf_counter() - t_s) * 1e3, 2)])
    e2e = sync_max((time.perf_counter() - t0) / max(e2e_steps, 1)) if e2e_steps else None

    if rank != 0:
        return
    e_loc = hi - lo
    trsm_flops = float(e_loc) * n * n * m            # minimal count: one triangular solve per test point
    n_trsm = max(tm["n_trsm"], 1.0)
    trsm_ms = tm["trsm_ms"] / n_trsm
    units_per_launch = tm["n_trsm"] and (e_loc * args.steps / tm["n_trsm"])
    peak = libmogp.peak_dmma_tflops(device)
    achieved = trsm_flops * args.steps / tm["n_trsm"] / (trsm_ms * 1e-3) * 1e-12 if tm["n_trsm"] else None
    roofline = {"bound": "tensor", "kernel": "predict_trsm_kernel (V = L^-1 K*, DMMA)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": (achieved / peak) if achieved else None, "traffic": ncu_traffic(args.workload + "-dmma", world),
                "peak_source": "DMMA issue peak measured in this run (mogp_peak_dmma); MEASURED_PEAKS.json has "
                               "no FP64 entry; cuBLAS DGEMM 8192^3 on this pool: 36.1 TFLOP/s",
                "flops_per_launch": trsm_flops * args.steps / tm["n_trsm"] if tm["n_trsm"] else None,
                "ms_per_launch": trsm_ms, "outputs_per_launch": units_per_launch}
    if tm.get("i8_block_rows", 0) > 0:
        # the predict TRSM ran on the int8 tcgen05 path (csrc/trsm_i8.cu): the dominant kernel is i8_trsm_kernel, ONE persistent
        # launch per predict call.  Algorithmic work = its kind::i8 MMAs: panels x (4 K-steps per 128-block of history) x plane
        # pairs x 2*128*64*32 integer ops; DESIGN.md section 3 states the count.
        planes = i8_planes()
        pairs = planes * (planes + 1) // 2
        T = (n + 127) // 128
        panels = (m + 63) // 64
        ops_step = float(e_loc) * panels * (T * (T - 1) // 2) * 4 * pairs * 2.0 * 128 * 64 * 32
        rows_ms = tm["i8_rows_ms"] / args.steps
        pk = i8_peaks(libmogp, device)
        peak256 = pk["sustained"]
        ach = ops_step / (rows_ms * 1e-3) * 1e-12
        roofline = {"bound": "tensor", "kernel": "i8_trsm_kernel<%d> (V_i = inv(L_ii)(K*_i - sum_j L_ij V_j): tcgen05.mma kind::i8 on %d "
                                                 "plane pairs per K step, FP64 DMMA epilogue; one persistent launch)" % (planes, pairs),
                    "achieved": ach, "peak": peak256, "unit": "TOP/s", "frac": ach / peak256,
                    "traffic": ncu_traffic(args.workload + "-i8", world),
                    "peak_source": "SUSTAINED int8 tcgen05 issue peak (M128 N256 K32 MMAs from resident operands, mogp_peak_i8, back to "
                                   "back for 3 s: the chip on its power cap) measured in this run -- the kernel is timed inside a "
                                   "long step under the same cap (see clocks); peak_burst / frac_of_burst: best of 10 short runs on "
                                   "the idle chip.  MEASURED_PEAKS.json has no int8 entry; twice its bf16 figures would be %.0f "
                                   "(burst) / %.0f (sustained) TOP/s" % (2.0 * measured_bf16(), 2.0 * measured_bf16("bf16_tflops_sustained")),
                    "peak_burst": pk["burst"], "frac_of_burst": ach / pk["burst"],
                    "ops_per_launch": ops_step, "ms_per_launch": rows_ms,
                    "launches_per_step": 1.0, "outputs_per_launch": float(e_loc),
                    "fp64_equivalent_tflops": trsm_flops * (1.0 - 1.0 / T) / (rows_ms * 1e-3) * 1e-12,
                    "fp64_dmma_peak_tflops": peak,
                    "accuracy_check_ms": tm.get("i8_check_ms", 0.0) / args.steps, "fp64_fallbacks": tm.get("i8_fallbacks", 0.0),
                    "whole_trsm_phase": {"ms": tm["trsm_ms"] / args.steps, "fp64_equivalent_tflops": achieved,
                                         "vs_dmma_peak": achieved / peak if achieved else None}}
    config = workload_config(args.workload, wl, world)
    if tm.get("i8_block_rows", 0) > 0:
        config["trsm_path"] = ("int8 tcgen05, %d planes per operand: the O(n^2 m) products of the FP64 forward substitution are "
                               "evaluated exactly on signed 7-bit digits (error-free splitting), everything else -- kernel "
                               "matrices, Cholesky, solves, means, the diagonal-block products, recombination and norms -- is "
                               "IEEE FP64; every call is checked a posteriori against FP64 solves of sampled test points "
                               "(accuracy_check_ms) and falls back to the all-FP64 kernel when they disagree; see parity_vs_cpu_sample"
                               % i8_planes())
        dtype = "f64 (predict TRSM: %d x int8 planes on tcgen05, exact s32 accumulation, FP64 epilogue)" % i8_planes()
    else:
        config["trsm_path"] = "FP64 DMMA"
        dtype = "f64"
    cholesky = cholesky_report(libmogp, device, tm, e_loc, n, args.steps, peak)
    if cholesky["path"].startswith("int8"):
        config["cholesky_path"] = cholesky["path"]
        dtype = dtype.replace("f64 (", "f64 (Cholesky history products: 8 x int8 planes on tcgen05, exact s32 accumulation, FP64 diagonal tiles and epilogue; ") \
            if "(" in dtype else "f64 (Cholesky history products: 8 x int8 planes on tcgen05, exact s32 accumulation, FP64 diagonal tiles and epilogue)"
    else:
        config["cholesky_path"] = cholesky["path"]
    line = {
        "metric": "gp_fit_predict_seconds", "value": per_step, "unit": "s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": config,
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": "s",
                "h2d_bytes_per_step": int(8 * (X.size + Y[lo:hi].size + thetas[lo:hi].size + Xs.size)),
                "d2h_bytes_per_step": int(8 * 2 * m * (E if world > 1 else e_loc) + 8 * 4 * e_loc)},
        "e2e_iterations_ms": {"construct, fit+predict, destroy": e2e_iters},
        "gpu_launches": int(tm["n_launches"]),
        "roofline": roofline,
        "phases_ms_per_step": {"fit_all_outputs": tm["fit_ms"] / args.steps, "kmat": tm["kmat_ms"] / args.steps,
                               "cholesky": tm["chol_ms"] / args.steps, "fit_solves": tm["solve_ms"] / args.steps,
                               "kstar_and_mean": tm["kstar_ms"] / args.steps,
                               "predict_trsm": tm["trsm_ms"] / args.steps,
                               "trsm_i8_planes_of_L": tm.get("i8_prep_ms", 0.0) / args.steps,
                               "trsm_i8_accuracy_check": tm.get("i8_check_ms", 0.0) / args.steps,
                               "trsm_i8_kernel": tm.get("i8_rows_ms", 0.0) / args.steps},
        "phases_ms_per_step_max_over_ranks": phase_max,
        "host_wall_ms_per_step": dict(wall_ms, predict_device_part=tm["predict_device_wall_ms"] / args.steps,
                                      predict_copy_out=tm["predict_d2h_wall_ms"] / args.steps),
        "cholesky_tflops": (e_loc * (n ** 3) / 3.0) / (tm["chol_ms"] / args.steps * 1e-3) * 1e-12 if tm["chol_ms"] else None,
        "cholesky": cholesky,
        "fit_tflops": (e_loc * (n ** 3) / 3.0) / (tm["fit_ms"] / args.steps * 1e-3) * 1e-12 if tm["fit_ms"] else None,
    }
    if not args.no_cpu:
        # bounded CPU sample on the same box: ONE output of the job with the oracle port of the reference.  At N > 1 it is the
        # LAST output -- owned by the last rank, so the row rank 0 checks arrived through the all-gather.
        o_chk = E - 1 if world > 1 else 0
        dt, cmean, cvar = oracle_step(X, Y[o_chk], Xs, thetas[o_chk], kernel, nugget)
        if world == 1:
            line["cpu_baseline"] = {"value": dt * E, "unit": "s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": "fit+predict of output 0 only (%.2f s on %d host cores), scaled x%d outputs"
                                              % (dt, os.cpu_count(), E)}
        gmean, gvar = res.mean[o_chk], res.unc[o_chk]
        nug = float(nugget) if not isinstance(nugget, str) else 0.0
        line["parity_vs_cpu_sample"] = {"output": o_chk, "owner_rank": world - 1 if world > 1 else 0,
                                        "mean_max_rel": float(np.max(np.abs(gmean - cmean)) / np.max(np.abs(cmean))),
                                        "var_max_abs": float(np.max(np.abs(gvar - cvar))),
                                        "var_worst_vs_tolerance": float(np.max(np.abs(gvar - cvar) / (1e-4 * np.abs(cvar) + 1e-4 * nug + 1e-300))),
                                        "all_outputs_finite": bool(np.all(np.isfinite(res.mean)) and np.all(np.isfinite(res.unc)))}
    if world == 1 and tm.get("i8_block_rows", 0) > 0 and not args.no_other:
        # the same step on the all-FP64 path (MOGP_TRSM_I8=0 is read when an emulator is constructed), in the same run on the
        # same box: the number the int8 figures stand beside.  One warm-up step, two timed.
        saved = os.environ.get("MOGP_TRSM_I8")
        saved_c = os.environ.get("MOGP_CHOL_I8")
        os.environ["MOGP_TRSM_I8"] = "0"
        os.environ["MOGP_CHOL_I8"] = "0"
        try:
            gpf = MultiOutputGP_GPU(X, Y, kernel=kernel, nugget=nugget, device=device)
            step(gpf)
            gpf.timings(reset=True)
            t0 = time.perf_counter()
            for _ in range(2):
                rf = step(gpf)
            dtf = (time.perf_counter() - t0) / 2
            tf = gpf.timings(reset=True)
            gpf.close()
            line["all_fp64_path"] = {"value": dtf, "unit": "s", "predict_trsm_ms": tf["trsm_ms"] / 2, "cholesky_ms": tf["chol_ms"] / 2,
                                     "trsm_tflops": trsm_flops / (tf["trsm_ms"] / 2 * 1e-3) * 1e-12,
                                     "var_max_abs_int8_vs_fp64": float(np.max(np.abs(rf.unc - res.unc))),
                                     "mean_max_rel_int8_vs_fp64": float(np.max(np.abs(rf.mean - res.mean)) / np.max(np.abs(rf.mean))),
                                     "note": "MOGP_TRSM_I8=0 MOGP_CHOL_I8=0: predict_trsm_kernel and chol_dataflow_kernel (DMMA) instead "
                                             "of i8_trsm_kernel and chol_i8_kernel; everything else identical"}
        finally:
            if saved_c is None:
                os.environ.pop("MOGP_CHOL_I8", None)
            else:
                os.environ["MOGP_CHOL_I8"] = saved_c
            if saved is None:
                os.environ.pop("MOGP_TRSM_I8", None)
            else:
                os.environ["MOGP_TRSM_I8"] = saved
    if world == 1 and args.workload == "c3" and not args.no_other:
        line["other_configs"] = other_configs(device, peak)
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the bounded CPU-baseline sample")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs)")
    ap.add_argument("--no-other", action="store_true", help="skip the short C1 / C2 / C4 runs appended to the C3 line at N = 1")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    elif args.workload == "c2":
        run_c2(args, wl)
    else:
        run_b200(args, wl)


if __name__ == "__main__":
    main()
