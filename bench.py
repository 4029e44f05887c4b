#!/usr/bin/env python
"""bench.py -- MOHGP VB iterations/sec (BASELINE.json metric), config 3: N=1M series, D=64 time points,
K=64 clusters, FP64, method='HS', synthetic data (SURVEY.md section 8d recipe).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one pass of the collapsed-VB loop body (collapsed_vb.py:147-303): G rows (natural gradient + CG dots), S rows (search
direction, trial point, softmax, statistics), per-cluster posterior, bound and the accept / reject / termination decision -- all
phases of ONE persistent cooperative kernel (k_mohgp_loop); the trace is read back once after the K steps.  Rows are
sharded over the ranks (strong scaling: N_total is fixed at 1M); statistics and dots cross NVLink inside that kernel (no NCCL call
per pass; NCCL carries the set-up, the barriers and the gather of the parity record).

The JSON line carries, besides the contract keys:
  roofline      the kernel of the timed region, HBM-bound: algorithmic bytes of everything the launch did / its CUDA-event
                duration, against MEASURED_PEAKS.json hbm_gbs; roofline_by_kernel = the two row phases from in-kernel timers
  parity        the oracle evaluated at N=1M on the bench inputs, outside every timed region, on this rank count
  fp64_tensor   4*N*K*D flop per step / step time against the measured DMMA peak (profiles/fp64_peaks_r01.json)
  cpu_baseline  the oracle port of the reference's native path (weave C kernels + numpy/LAPACK) on the
                host cores, on a bounded row sample, scaled linearly in N
  e2e           the same metric through the public API with HOST buffers: construct MOHGP from numpy Y and
                numpy phi_ (H2D), optimize() for the same number of steps, read phi and the bound back (D2H)
"""
import argparse
import json
import os
import subprocess
import sys
import time

if "reference" in sys.argv[1:]:
    # --impl reference is a CPU run on ALL host cores at every --gpus N.  torchrun exports OMP_NUM_THREADS=1 to its workers, which
    # would time the OpenMP kernels (and numpy's BLAS) on one core: set the thread counts before numpy / libgomp are loaded.
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_TOTAL, D, K = 1000000, 64, 64
METRIC = "MOHGP VB iterations/sec (N=1M,D=64,K=64)"
UNIT = "iterations/s"


def config_of(method):
    """`config` of the JSON line -- identical in both arms (the driver compares them key by key)."""
    return {"workload": "MOHGP N=1M D=64 K=64 FP64 %s-CG (config 3)" % method, "N": N_TOTAL, "D": D, "K": K, "method": method,
            "l2": "inputs larger than L2: every step streams 11 N x K / N x D arrays (5.6 GB over all GPUs, >= 700 MB per GPU at 8 GPUs) "
                  "against 126 MB of L2; no flush between steps"}
FP64_DMMA_PEAK_TFLOPS = 37.1  # measured on this pool's B200 (tools/microbench/fp64_peaks.cu, profiles/fp64_peaks_r01.json)
def source_sha16():
    """sha16 over the CUDA sources and the C-ABI header, in name order: identifies what libgpclust_b200.so is built from."""
    import hashlib
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "gpclust_b200", "csrc")
    for f in sorted(os.listdir(csrc)) + [os.path.join("..", "..", "include", "gpclust_b200.h")]:
        with open(os.path.join(csrc, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def ncu_traffic(kernel, n_loc, n_evals, n_g):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed `ncu --set full` capture of the shipped
    binary (profiles/ncu_traffic_r02.json, written by tools/ncu_traffic.py next to the capture: bytes per G phase and per S phase at
    1M rows, and the sha16 of the CUDA sources the captured library was built from), scaled to this launch.  None when there is no capture for this binary."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic_r02.json")) as f:
            rec = json.load(f)
        sha = source_sha16()
        per = rec["kernels"].get(kernel)
        if per is None:
            return None, "no capture of %s in profiles/ncu_traffic_r02.json" % kernel
        scale = n_loc / float(rec["rows"])
        val = (per["dram_bytes_per_g_phase"] * n_g + per["dram_bytes_per_s_phase"] * n_evals) * scale
        src = "%s (ncu --set full at %d rows, scaled by rows per GPU and evaluations per launch; captured library built from sources %s, this run %s)" % (
            rec["source"], rec["rows"], rec["src_sha16"], "the same sources" if rec["src_sha16"] == sha else "DIFFERENT sources " + sha)
        return val, src
    except (OSError, KeyError, ValueError) as e:
        return None, "unavailable (%s)" % type(e).__name__


def _rbf(X, variance, lengthscale):
    """RBF covariance of the synthetic-data recipe (GPy's formulation: r^2 = -2 X X^T + |x|^2 + |x'|^2, zero diagonal, clipped at 0).
    bench.py's own copy: the GPU arm does not touch oracle/ to make its inputs."""
    Xsq = np.sum(np.square(X), 1)
    r2 = -2.0 * np.dot(X, X.T) + (Xsq[:, None] + Xsq[None, :])
    r2[np.diag_indices(X.shape[0])] = 0.0
    r = np.sqrt(np.clip(r2, 0.0, np.inf)) / lengthscale
    return variance * np.exp(-0.5 * r ** 2)


def synth(n_total, d, k, seed=0):
    """SURVEY.md section 8(d) config 3 generator (host numpy, deterministic)."""
    rng = np.random.RandomState(seed)
    X = np.linspace(0, 10, d)[:, None]
    Sf = _rbf(X, 1.0, 2.0)
    Sy = _rbf(X, 0.1, 1.0) + np.eye(d) * 0.05
    z = rng.randint(0, k, n_total)
    means = (np.linalg.cholesky(Sf + 1e-8 * np.eye(d)).dot(rng.randn(d, k))).T
    Y = means[z] + (np.linalg.cholesky(Sy).dot(rng.randn(d, n_total))).T
    np.random.seed(0)
    phi0 = np.random.randn(n_total, k)  # collapsed_mixture.py:41
    return X, np.ascontiguousarray(Y), phi0


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except OSError:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler(object):
    """SM clock and throttle reasons of one GPU DURING the timed region: an NVML polling thread (one sample per ~2 ms, so that
    even the 20 ms regions of the 8-GPU runs are covered); `nvidia-smi -lms` as the fallback when NVML bindings are missing."""
    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.thread, self.samples, self.stop_flag = index, None, None, [], False
        self.nvml, self.handle = None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            handle = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml, self.handle = pynvml, handle
        except Exception:
            self.nvml = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                reasons = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append((sm, int(reasons)))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nvml is not None:
            import threading
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(1.0)
            n = self.nvml
            bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            sm = [s[0] for s in self.samples]
            seen = 0
            for s in self.samples:
                seen |= s[1]
            try:
                mx = float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM))
            except Exception:
                mx = None
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                    "reasons": sorted(k for k, b in bits.items() if seen & b), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().split("\n"):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "samples": len(sm),
                "reasons": sorted(reasons), "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------------------ CPU legs
def cpu_reference_rate(X, Y, phi0, k, budget_s, passes=2):
    """Oracle port of the reference's native path on the host cores: weave C kernels (oracle/weave_kernels.c,
    OpenMP over rows as utilities.py:54,143) + numpy/LAPACK for everything else, host-vector CG loop.
    Runs `passes` loop-body passes on a row sample sized for ~budget_s seconds; returns iterations/s scaled
    linearly to N_TOTAL rows (every N-dependent term of the iteration is linear in N)."""
    from oracle import native
    from oracle.build import build_oracle, has_openmp
    from oracle.vb_oracle import MOHGPOracle
    from oracle import gpy_shim as G
    build_oracle()
    cores = os.cpu_count() if has_openmp() else 1

    def run(n_rows):
        kF = G.RBF(1, 1.0, 2.0)
        kY = G.RBF(1, 0.1, 1.0) + G.White(1, 0.05)
        np.random.seed(0)
        m = MOHGPOracle(X, kF, kY, Y[:n_rows], K=k, mahalanobis=native.multiple_mahalanobis)
        m.softmax = native.softmax_weave
        m.set_vb_param(phi0[:n_rows].flatten())
        t0 = time.perf_counter()
        m.iterate(passes, method="HS")
        return (time.perf_counter() - t0) / passes

    probe_rows = min(4000, Y.shape[0])
    t_probe = run(probe_rows)
    per_row = t_probe / probe_rows
    rows = int(min(Y.shape[0], max(probe_rows, budget_s / passes / per_row)))
    t_iter = run(rows) if rows > probe_rows else t_probe
    rate = 1.0 / (t_iter * (N_TOTAL / float(rows)))
    sample = "%d of %d rows, %d loop passes, %.2f s/pass on the sample, scaled linearly in N" % (rows, N_TOTAL, passes, t_iter)
    return rate, cores, sample, t_iter


def cpu_other_baselines(X, Y, phi0, k):
    """The two other CPU lines of BASELINE.md section 3, each on a bounded row sample and scaled linearly in N:
      vectorised_blas          the oracle with the GEMM form of multiple_mahalanobis (numpy / BLAS on all cores): the "fair CPU" line
      python3_fallback_loops   what the reference runs AS IS on Python 3: np_utilities.multiple_mahalanobis_numpy_loops, a Python
                               double loop with one np.linalg.solve per (n, k) pair (np_utilities.py:43-59)"""
    from oracle.vb_oracle import MOHGPOracle, mahalanobis_whitened_gemm, mahalanobis_pairs_solve
    from oracle import gpy_shim as G
    out = {}
    for name, fn, rows, passes in (("vectorised_blas", mahalanobis_whitened_gemm, 100000, 2), ("python3_fallback_loops", mahalanobis_pairs_solve, 200, 1)):
        kF, kY = G.RBF(1, 1.0, 2.0), G.RBF(1, 0.1, 1.0) + G.White(1, 0.05)
        np.random.seed(0)
        m = MOHGPOracle(X, kF, kY, Y[:rows], K=k, mahalanobis=fn)
        m.set_vb_param(phi0[:rows].flatten())
        t0 = time.perf_counter()
        m.iterate(passes, method="HS")
        t_pass = (time.perf_counter() - t0) / passes
        out[name] = {"value": 1.0 / (t_pass * N_TOTAL / float(rows)), "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                     "sample": "%d of %d rows, %d loop passes, %.2f s/pass on the sample, scaled linearly in N" % (rows, N_TOTAL, passes, t_pass)}
    return out


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port; the reference itself cannot be imported --
    GPy / scipy.weave are absent).  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    X, Y, phi0 = synth(min(N_TOTAL, 200000), D, K)  # the sample never needs more rows than this
    steps, warm = max(1, args.steps), max(0, args.warmup)
    budget = 150.0
    per_step_budget = budget / (steps + warm)
    t0 = time.perf_counter()
    from oracle import native
    from oracle.build import build_oracle, has_openmp
    from oracle.vb_oracle import MOHGPOracle
    from oracle import gpy_shim as G
    build_oracle()
    cores = os.cpu_count() if has_openmp() else 1
    # calibrate the sample size so that (steps + warmup) passes fit the budget
    probe = 2000
    kF, kY = G.RBF(1, 1.0, 2.0), G.RBF(1, 0.1, 1.0) + G.White(1, 0.05)

    def make(n_rows):
        m = MOHGPOracle(X, kF, kY, Y[:n_rows], K=K, mahalanobis=native.multiple_mahalanobis)
        m.softmax = native.softmax_weave
        m.set_vb_param(phi0[:n_rows].flatten())
        return m
    m = make(probe)
    t1 = time.perf_counter()
    m.iterate(1)
    per_row = (time.perf_counter() - t1) / probe
    rows = int(min(Y.shape[0], max(probe, per_step_budget / per_row)))
    m = make(rows)
    m.iterate(warm) if warm else None
    t1 = time.perf_counter()
    m.iterate(steps)
    dt = time.perf_counter() - t1
    t_step_full = dt / steps * (N_TOTAL / float(rows))
    value = 1.0 / t_step_full
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": t_step_full * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": config_of(args.method),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d of %d rows per step, %.3f s/step on the sample, scaled linearly in N" % (rows, N_TOTAL, dt / steps)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from gpclust_b200 import MOHGP, kern

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    X, Y, phi0 = synth(N_TOTAL, D, K)
    lo = rank * N_TOTAL // world
    hi = (rank + 1) * N_TOTAL // world
    Y_loc, phi_loc = Y[lo:hi], phi0[lo:hi]
    n_loc = hi - lo
    kF = lambda: kern.RBF(1, 1.0, 2.0).fix()  # the metric holds the kernel hyper-parameters fixed
    kY = lambda: (kern.RBF(1, 0.1, 1.0) + kern.White(1, 0.05)).fix()

    steps, warm = max(1, args.steps), max(3, args.warmup)
    method = args.method

    # ---------------- device-resident timing: K loop-body passes, state already in HBM
    m = MOHGP(X, kF(), kY(), Y_loc, K=K, phi_init=phi_loc)
    m.bound()

    # K loop-body passes = one enqueue of K fixed launch sequences driven by the device-resident loop state
    # (accept / reject decided on the device), one read-back of the trace at the end
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    m.run_passes(warm, method=method)
    barrier()
    launches0 = m.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    it_before = m._pass_state["iteration"]
    barrier()
    # the events are recorded by the library immediately around the launch(es) of these K passes: the timed region is the device
    # work of exactly K steps plus the launch, not the host-side set-up of the loop state or the read-back of the trace
    m._launch_events = (ev0, ev1)
    rows = m.run_passes(steps, method=method)
    m._launch_events = None
    barrier()
    assert len(rows) == steps, (len(rows), steps)
    last_bound = rows[-1][1]
    rejected = int(rows[-1][0] - it_before) - steps  # a rejected conjugate step costs a second bound evaluation
    elapsed_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    launches = m.kernel_launches - launches0

    fused = bool(m._fused_ok())
    if fused:
        # the loop is ONE persistent kernel (k_mohgp_loop): its phases are timed inside the kernel -- CTA 0 reads %globaltimer at the
        # phase boundaries and accumulates ns per phase -- over the same K passes again (the stamps are kept out of the region
        # `value` is taken from).  "posterior" = B2 + per-cluster posterior + decision + B3; "dots" = B1 + the dots (+ NVLink exchange).
        m._loop_timing = torch.zeros(32, dtype=torch.int64, device="cuda")
        it_before2 = m._pass_state["iteration"]
        rows2 = m.run_passes(steps, method=method)
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        tt = m._loop_timing.cpu().numpy().astype(np.float64)
        m._loop_timing = None
        n_eval, n_g = max(1.0, tt[4]), max(1.0, tt[5])
        kernel_ms = {"natgrad": tt[0] / n_g * 1e-6, "dots": tt[1] / n_g * 1e-6, "step_stats": tt[2] / n_eval * 1e-6, "posterior": tt[3] / n_eval * 1e-6}
        kernel_n = {"natgrad": int(n_g), "dots": int(n_g), "step_stats": int(n_eval), "posterior": int(n_eval)}
        for q, nm in enumerate(("post_barrier_wait", "post_row_reduce", "post_exchange", "post_algebra", "post_tail_barrier")):
            kernel_ms[nm] = tt[6 + q] / n_eval * 1e-6
    else:
        # per-kernel durations: the same K passes continue with a CUDA-event pair around every launch (the pairs cost a
        # few microseconds per launch, which is why they are kept out of the region `value` is taken from)
        m._kernel_events = {}
        it_before2 = m._pass_state["iteration"]
        rows2 = m.run_passes(steps, method=method)
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        rejected2 = int(rows2[-1][0] - it_before2) - steps
        kernel_ms = {name: float(np.mean([a.elapsed_time(b) for a, b in pairs])) for name, pairs in m._kernel_events.items()}
        kernel_n = {name: len(pairs) for name, pairs in m._kernel_events.items()}
        # one bound evaluation per enqueued sequence: the sequence after a rejected conjugate step skips its G pass (returns at
        # once) and runs S / posterior as the steepest retry -- report the G pass over the launches that did work
        # (and the sequences enqueued beyond the last accepted step return at once altogether)
        n_eval = steps + rejected2
        for name, keep in (("natgrad", steps), ("step_stats", n_eval), ("posterior", n_eval)):
            if name in m._kernel_events:
                d = sorted((a.elapsed_time(b) for a, b in m._kernel_events[name]), reverse=True)
                kernel_ms[name] = float(np.mean(d[:keep]))
                kernel_n[name] = min(keep, len(d))
                if len(d) > keep:
                    kernel_ms[name + "_skipped"] = float(np.mean(d[keep:]))
        m._kernel_events = None
    ms_per_step = elapsed_ms / steps
    value = 1e3 / ms_per_step

    # ---------------- end to end through the public API with host buffers
    # the caller's arrays are PLAIN (pageable) numpy: whatever staging the library does to move them happens inside the timed region
    Y_host, phi_host0 = np.ascontiguousarray(Y_loc), np.ascontiguousarray(phi_loc)
    del m
    torch.cuda.empty_cache()
    e2e_steps = steps

    # the call a user makes: construct from host arrays, optimize(), read the responsibilities and the bound back.
    # maxiter counts bound evaluations (collapsed_vb.py:177,191), so a rejected step uses two of them.
    def e2e_run():
        mm = MOHGP(X, kF(), kY(), Y_host, K=K, phi_init=phi_host0)
        mm.optimize(method=method, maxiter=e2e_steps, ftol=-1.0, gtol=-1.0, verbose=False)
        phi = mm.phi
        return mm.bound(), phi, len(mm.optimize_trace)

    e2e_run()  # warm-up (CUDA allocator, the library's own pinned staging buffers)
    barrier()
    t0 = time.perf_counter()
    b_e2e, phi_host, e2e_passes = e2e_run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = e2e_passes / e2e_s
    # for information: the same call sequence when the caller's arrays are page-locked already (no staging copy in the ingest)
    e2e_pinned = None
    if not args.no_pinned_e2e:
        Yp, pp_ = torch.from_numpy(Y_host).pin_memory().numpy(), torch.from_numpy(phi_host0).pin_memory().numpy()
        Y_host, phi_host0 = Yp, pp_
        e2e_run()
        barrier()
        t0 = time.perf_counter()
        _, _, passes_p = e2e_run()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        tp = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        e2e_pinned = passes_p / float(tp.item())
        del Yp, pp_
    h2d = (n_loc * D * 8 + n_loc * K * 8) * world / float(e2e_passes)
    d2h = (n_loc * K * 8) * world / float(e2e_passes) + 128.0 * world

    # ---------------- oracle parity at the metric's own shape, on this rank count (outside every timed region)
    parity = None
    if not args.no_parity:
        from helpers import headline_parity

        def gather_rows(a):
            """Row-concatenation of the ranks' (n_loc x C) arrays on rank 0, over NCCL point-to-point (device tensors)."""
            if world == 1:
                return a
            t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
            if rank != 0:
                dist.send(t, dst=0)
                return None
            parts = [a]
            for r in range(1, world):
                rows_r = (r + 1) * N_TOTAL // world - r * N_TOTAL // world
                buf = torch.empty((rows_r, a.shape[1]), dtype=t.dtype, device="cuda")
                dist.recv(buf, src=r)
                parts.append(buf.cpu().numpy())
            return np.concatenate(parts, 0)

        mp_ = MOHGP(X, kF(), kY(), Y_loc, K=K, phi_init=phi_loc)
        parity = headline_parity(X, Y, phi0, K, mp_, passes=args.parity_passes, method=method, gather=gather_rows, is_root=(rank == 0))
        del mp_
        torch.cuda.empty_cache()
        if parity is not None:
            parity["ranks"] = world
            parity["tolerance"] = 1e-9
            parity["ok"] = bool(parity["bound_rel"] < 1e-9 and parity["natgrad_rel"] < 1e-9 and parity["trace_iterations_equal"] and
                                parity["bound_rel_after"] < 1e-9 and parity["phi_rel_after"] < 1e-9 and parity["argmax_equal"])
        barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_kind = measured_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    # algorithmic bytes per launch: G pass reads F, x, x_old, ng_old and writes ng;
    # S pass reads F, x, ng, sd_old and writes sd, x_new  (steepest / FR drop the *_old reads)
    with_old = method in ("HS", "PR")
    g_bytes = n_loc * 8.0 * (D + K * (4 if with_old else 2))
    s_bytes = n_loc * 8.0 * (D + K * (5 if method != "steepest" else 4))
    alg = {"natgrad": g_bytes, "step_stats": s_bytes}
    # the two N-sized row phases share the step almost evenly (the S rows also run for every rejected conjugate step)
    total_ms = sum(kernel_ms.get(n, 0.0) * kernel_n.get(n, 0) for n in ("natgrad", "dots", "step_stats", "posterior"))
    what = "phase of k_mohgp_loop (in-kernel %globaltimer, CTA 0)" if fused else "kernel (CUDA events around each launch)"
    by_kernel = {("g_rows" if fused else "k_natgrad") if n == "natgrad" else ("s_rows" if fused else "k_step_stats"):
                 {"what": what, "ms": kernel_ms[n], "count": kernel_n[n], "share_of_step": kernel_ms[n] * kernel_n[n] / total_ms,
                  "alg_bytes": alg[n], "achieved_GBps": alg[n] / (kernel_ms[n] * 1e-3) / 1e9, "frac": alg[n] / (kernel_ms[n] * 1e-3) / 1e9 / hbm_peak}
                 for n in ("natgrad", "step_stats")}
    if fused:
        # ONE launch in the timed region: algorithmic bytes of everything it did / its CUDA-event duration
        n_evals_timed = steps + rejected
        launch_bytes = steps * g_bytes + n_evals_timed * s_bytes
        dom, dom_name, dom_ms = None, "k_mohgp_loop", elapsed_ms
        achieved = launch_bytes / (elapsed_ms * 1e-3) / 1e9
        dom_bytes = launch_bytes
    else:
        dom = max(("natgrad", "step_stats"), key=lambda n: kernel_ms.get(n, 0.0))
        dom_name, dom_ms, dom_bytes = "k_" + dom, kernel_ms[dom], alg[dom]
        achieved = alg[dom] / (kernel_ms[dom] * 1e-3) / 1e9
    iter_bytes = g_bytes + s_bytes
    traffic, traffic_source = ncu_traffic(dom_name, n_loc, steps + rejected if fused else 1, steps if fused else 1)
    tensor_tf = 4.0 * N_TOTAL * K * D / (ms_per_step * 1e-3) / 1e12  # aggregate over the `world` GPUs

    cpu, cpu_other = None, None
    if world == 1 and not args.no_cpu:
        rate, cores, sample, _ = cpu_reference_rate(X, Y, phi0, K, budget_s=args.cpu_budget)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        cpu_other = cpu_other_baselines(X, Y, phi0, K)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(method),
        "run": {"rows_per_gpu": n_loc,
                "parallelism": ("rows sharded x%d, statistics + CG dots exchanged over NVLink peer memory inside the kernels" % world) if world > 1 else "1 GPU",
                "rejected_steps": rejected, "final_bound": last_bound},
        "parity": parity,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_passes,
                "what": "MOHGP(X,kernF,kernY,Y_host,K,phi_init=host) + optimize(maxiter=%d evaluations = %d passes) + m.phi + m.bound(), "
                        "host numpy in/out" % (e2e_steps, e2e_passes),
                "bound": b_e2e,
                "with_pinned_inputs": e2e_pinned,
                "inputs": "plain (pageable) numpy arrays: the library stages them through its own pinned ring inside the timed region; "
                          "`with_pinned_inputs` = the same calls when the caller's arrays are page-locked already (informational)"},
        "gpu_launches": launches,
        "gpu_launches_what": "k_mohgp_loop: the whole timed region is one persistent cooperative launch" if fused else "launches of this package's kernels in the timed region",
        "roofline": {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "peak_kind": peak_kind, "traffic": traffic, "traffic_source": traffic_source,
                     "alg_bytes_per_launch": dom_bytes, "ms_per_launch": dom_ms,
                     "launch": ("one cooperative launch = %d passes (%d bound evaluations); bytes = per-GPU algorithmic bytes of all of them" % (steps, steps + rejected)) if fused else "per launch",
                     "iteration_hbm_frac": iter_bytes / (ms_per_step * 1e-3) / 1e9 / hbm_peak},
        "roofline_by_kernel": by_kernel,
        "kernels_ms": kernel_ms,
        "fp64_tensor": {"achieved": tensor_tf, "achieved_per_gpu": tensor_tf / world, "peak": FP64_DMMA_PEAK_TFLOPS, "peak_is": "per GPU, measured DMMA",
                        "unit": "TFLOP/s", "frac": tensor_tf / world / FP64_DMMA_PEAK_TFLOPS, "flops_per_step": 4.0 * N_TOTAL * K * D},
        "cpu_baseline": cpu,
        "cpu_baselines_other": cpu_other,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--method", default="HS", choices=["HS", "PR", "FR", "steepest"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-pinned-e2e", action="store_true", help="skip the informational e2e run with page-locked inputs")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle parity record (about a minute of CPU work at N=1M)")
    ap.add_argument("--parity-passes", type=int, default=5)
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
