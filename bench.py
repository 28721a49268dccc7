#!/usr/bin/env python
"""bench.py — hyperlikelihood logL+grad evaluations per second on the O5-shaped mock (BASELINE.json metric).

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the CPU path (oracle port) on the box's host cores, full workload

One "step" = one evaluation of the hot path (theta -> logL parts, both 14-parameter gradients, Neff's) over the
whole catalog, which is resident in HBM (uploaded once, exactly as the reference's jitted model closes over its
data; NUTS only ever changes theta).  `value` times K back-to-back evaluations with CUDA events on the stream
they are launched on; `e2e` times the public host call (host theta in, host result out) by wall clock.
Prints ONE JSON line on rank 0.  Besides the headline it carries, all measured in the same run:

  roofline      the streaming kernel against the FP64 pipe (the binding resource) with HBM alongside
  cpu_baseline  the fused C++/OpenMP port (oracle/bump_cpu.cpp) on the FULL workload, all host cores
  configs       the other BASELINE.json shapes (GWTC-3, O4, O5 with w0-wa), sharded like the headline
  nuts          (1 GPU) NUTS 4 x (1000 + 1000), dense mass, the reference's seed, on the GWTC-3 shape: ESS/s
  result_check  (N > 1) every output of the sharded evaluation against an unsharded evaluation on rank 0's GPU;
                the process exits non-zero if they differ by more than 1e-12
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "hyperlikelihood logL+grad evals/sec (O5 mock)"
UNIT = "evals/s"
ALGO_BYTES_PER_SAMPLE = 32           # m1_det, q, d_L, pdraw as fp64
# FP64-pipe warp instructions per 32 samples that the hot path needs, FROZEN at what the round-1 streaming kernel
# executed (ncu, profiles/r01b_o5_stream_kernel_opmix.txt): 8 exp, 4 reciprocals, the table lerps and the 19 shifted
# sums.  roofline.achieved = this x samples / kernel time, so it rises when the kernel gets faster, whatever the
# kernel then executes (SURVEY.md 8(d)'s a-priori estimate was 700).
ALGO_FP64_INST_PER_SAMPLE = 217.5
FP64_PEAK_FALLBACK = 1.84e13         # DFMA warp-lane instructions/s measured by bump_peak on this pool's B200s
MULTIRANK_TOL = 1e-12


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


class OneLineStdout:
    """The contract is ONE JSON line on stdout.  Libraries write there too at the C level (NCCL prints its version
    banner on communicator creation): park fd 1 on stderr for the run and emit the line on the real stdout."""

    def __init__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, obj):
        sys.stdout.flush()
        os.write(self.real, (json.dumps(obj) + "\n").encode())


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(index), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for name, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def fp64_peak(device):
    exe = os.path.join(ROOT, "bumpcosmology_b200", "bump_peak")
    try:
        out = subprocess.run([exe, str(device)], capture_output=True, text=True, timeout=60).stdout.strip()
        return json.loads(out.splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)}


def workload_catalog(name):
    from bumpcosmology_b200.catalogs import make_catalog
    t0 = time.time()
    cat = make_catalog(name)
    return cat, time.time() - t0


def config_of(args, cat):
    """Identical in both arms (the driver compares them): what is computed, not how."""
    return {"workload": f"{args.workload}: {cat.nobs} events x {cat.nsamp} samples + {cat.nsel} injections"
                        + (", w0-wa dark energy (15 parameters)" if args.wa else ""),
            "elements": cat.n_elements, "outputs": "loglike, log_mu_sel, log_mu2, neff_sel, neff[nobs], "
                                                   "d loglike/d theta[14], d log_mu_sel/d theta[14]",
            "l2": "inputs (1.92 GB at O5) exceed the 126 MB L2; every step re-reads them from HBM"
                  if 32 * cat.n_elements > 2.5e8 else "inputs fit in L2 (NUTS re-reads the same data every step)"}


# ------------------------------------------------------------------------------------------------ CPU arm
def time_cpu_port(cat, steps, warmup):
    """oracle/bump_cpu.cpp (fused single-pass C++/OpenMP port, -march=native build made on this host) on the FULL
    workload with every core this process may use.  Returns dict or None (no compiler / library)."""
    try:
        from oracle import bump_cpu
        port = bump_cpu.CpuPort(*cat.as_args(), native=True)
    except Exception as e:  # noqa: BLE001
        print(f"[bench] C++ CPU port unavailable: {e}", file=sys.stderr)
        return None
    from bumpcosmology_b200.catalogs import THETA_DEFAULT, draw_prior_thetas
    thetas = np.vstack([THETA_DEFAULT, draw_prior_thetas(15, seed=5)])
    threads = host_threads()
    for i in range(warmup):
        port.evaluate(thetas[i % len(thetas)], nthreads=threads)
    ts = []
    t_all = time.perf_counter()
    for i in range(steps):
        t0 = time.perf_counter()
        port.evaluate(thetas[i % len(thetas)], nthreads=threads)
        ts.append(time.perf_counter() - t0)
    t_all = time.perf_counter() - t_all
    lib = os.path.basename(bump_cpu.loaded_path or "")
    port.close()
    return {"value": steps / t_all, "s_per_eval_median": float(np.median(ts)), "s_total": t_all, "cores": threads,
            "host_cpus": os.cpu_count(), "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS"),
            "engine": f"cpp_openmp (oracle/bump_cpu.cpp: fused single pass, log space, libm; {lib})",
            "sample": f"the full workload ({cat.nobs} events x {cat.nsamp} samples + {cat.nsel} injections), "
                      f"{steps} evaluations after {warmup} warm-ups, {threads} OpenMP threads set explicitly"}


def time_cpu_torch(cat, evals, warm):
    """oracle/bump_oracle.py (eager torch fp64 + autograd), bounded sample: only if the C++ port cannot be built."""
    import torch

    from bumpcosmology_b200.catalogs import THETA_DEFAULT
    from oracle import bump_oracle as bo
    torch.set_num_threads(host_threads())
    frac = min(1.0, 900_000 / cat.n_elements)
    ne, ns = max(1, int(round(cat.nobs * frac))), max(1, int(round(cat.nsel * frac)))
    data = (cat.m1s_det[:ne], cat.qs[:ne], cat.dls[:ne], cat.pdraw[:ne], cat.m1s_det_sel[:ns], cat.qs_sel[:ns],
            cat.dls_sel[:ns], cat.pdraw_sel[:ns], cat.Ndraw)
    scale = (ne * cat.nsamp + ns) / cat.n_elements
    chunk = max(1, 2_000_000 // max(cat.nsamp, 1))
    for _ in range(warm):
        bo.evaluate(THETA_DEFAULT, data, grad=True, event_chunk=chunk)
    t0 = time.perf_counter()
    for _ in range(evals):
        bo.evaluate(THETA_DEFAULT, data, grad=True, event_chunk=chunk)
    t = (time.perf_counter() - t0) / evals
    return {"value": scale / t, "s_per_eval_median": t / scale, "cores": torch.get_num_threads(),
            "host_cpus": os.cpu_count(), "engine": "torch_eager (oracle/bump_oracle.py)",
            "sample": f"{ne} events + {ns} injections ({100 * scale:.1f} % of the elements), extrapolated linearly"}


def cpu_arm(cat, steps, warmup):
    r = time_cpu_port(cat, steps, warmup)
    return r if r is not None else time_cpu_torch(cat, max(2, min(steps, 3)), 1)


def run_reference(args, rank, out):
    """The reference arm: the reference's algorithm on the host cores.  The reference itself (JAX/numpyro) is not
    installable in this image, so this times the fused C++/OpenMP port of oracle/ on the same config, FULL workload
    per step, all host cores (set explicitly: torchrun exports OMP_NUM_THREADS=1)."""
    if rank != 0:
        return
    cat, _ = workload_catalog(args.workload)
    r = cpu_arm(cat, args.steps, args.warmup)
    v = r["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_of(args, cat),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"],
                         "s_per_eval_median": r["s_per_eval_median"], "host_cpus": r["host_cpus"],
                         "engine": r["engine"]},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference JAX stack is not installable here (no wheel, no network); CPU arm = " + r["engine"],
    }
    out.emit(line)


# ------------------------------------------------------------------------------------------------ GPU helpers
def _flat(res):
    return np.concatenate([[res.loglike, res.log_mu_sel, res.log_mu2, res.neff_sel], res.dloglike, res.dlog_mu_sel])


def time_config(name, wa, world, local_rank, dist, steps=100, warm=5):
    """One of the other BASELINE.json shapes, sharded like the headline: device-timed back-to-back evaluations and
    the end-to-end host call."""
    import torch

    from bumpcosmology_b200.catalogs import THETA_DEFAULT, make_catalog
    from bumpcosmology_b200.likelihood import Hyperlikelihood, ShardedHyperlikelihood
    cat = name if not isinstance(name, str) else make_catalog(name)
    theta = np.concatenate([THETA_DEFAULT, [0.3]]) if wa else THETA_DEFAULT
    if world > 1:
        like = ShardedHyperlikelihood(cat.as_args(), device=local_rank, wa=wa)
        local = like.local
    else:
        like = local = Hyperlikelihood(*cat.as_args(), device=local_rank, wa=wa)
    like(theta)
    local.time_evals(theta, warm)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ms, _ = local.time_evals(theta, steps)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    t0 = time.perf_counter()
    for _ in range(steps):
        res = like(theta)
    e2e = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = float(t.item())
    _, ker = local.time_evals(theta, 10, kernel=True)
    rec = {"workload": f"{cat.nobs} events x {cat.nsamp} samples + {cat.nsel} injections" + (", w0-wa" if wa else ""),
           "evals_per_s": steps / (ms * 1e-3), "us_per_eval": 1e3 * ms / steps, "e2e_evals_per_s": steps / e2e,
           "stream_kernel_us": 1e3 * ker / 10, "steps": steps, "logl": res.logl}
    like.close() if hasattr(like, "close") else local.close()
    return rec


def nuts_record(workload, warmup, samples, chains, device, cpu_s_per_eval=None):
    """NUTS(dense_mass=True), `chains` x (warmup + samples), seed 1652819403 (run_cosmo_fit.py:17-19,45-49) with the
    library's C++ driver, one context and host thread per chain."""
    from bumpcosmology_b200 import intensity_models as im, nuts
    from bumpcosmology_b200.catalogs import make_catalog
    cat = make_catalog(workload)
    first = im.pop_cosmo_model(*cat.as_args(), device=device)
    models = [first] + [first.clone() for _ in range(chains - 1)]   # one resident catalog, one context per chain
    t0 = time.perf_counter()
    r = nuts.run_mcmc(models, warmup, samples, chains, seed=1652819403, native=True)
    wall = time.perf_counter() - t0
    ess = r["ess_bulk"][:14]
    rec = {"workload": f"{workload}: {cat.nobs} events x {cat.nsamp} samples + {cat.nsel} injections",
           "chains": chains, "warmup": warmup, "samples": samples, "dense_mass": True, "seed": 1652819403,
           "driver": "c++ (bump_nuts_chain), chains in parallel host threads on one GPU",
           "wall_s": wall, "sampling_s": r["sampling_s"], "min_bulk_ess": float(ess.min()),
           "ess_per_s_total": float(ess.min() / wall), "ess_per_s_sampling": float(ess.min() / r["sampling_s"]),
           "rhat_max": float(r["rhat"][:14].max()),
           "divergences": int(sum(c["stats"]["diverging"].sum() for c in r["chains"])),
           "mean_tree_depth": float(np.mean([c["stats"]["depth"].mean() for c in r["chains"]])),
           "step_size": [float(c["step_size"]) for c in r["chains"]],
           "model_evals": int(r["n_leapfrog_total"]), "evals_per_s": r["n_leapfrog_total"] / wall}
    if cpu_s_per_eval:
        # the CPU arm of the sampler: the same driver, the same number of model evaluations, each costing one
        # evaluation of the CPU port (measured in this run at this shape); chains in sequence on all cores
        cpu_wall = r["n_leapfrog_total"] * cpu_s_per_eval
        rec["cpu_arm"] = {"kind": "derived", "s_per_eval": cpu_s_per_eval, "wall_s": cpu_wall,
                          "ess_per_s_total": float(ess.min() / cpu_wall),
                          "note": "same sampler and evaluation count driving the C++/OpenMP port: model evaluations "
                                  "x measured CPU seconds per evaluation at this shape (all host cores)"}
    for m in models:
        m.close()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--workload", default=os.environ.get("BUMP_BENCH_WORKLOAD", "o5"))
    ap.add_argument("--exchange", default=os.environ.get("BUMP_EXCHANGE", "p2p"), choices=("torch", "nccl", "p2p"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the GWTC-3 / O4 / w0-wa sub-records")
    ap.add_argument("--no-nuts", action="store_true", help="skip the NUTS sub-record (1 GPU only)")
    ap.add_argument("--wa", action="store_true",
                    help="w0-wa (CPL) dark energy variant (BASELINE.json config 5): 15 parameters, the d_L(z) grid is "
                         "rebuilt on the device every step by the cumulative-trapezoid tables kernel")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    out = OneLineStdout()
    if args.impl == "reference":
        return run_reference(args, rank, out)

    import torch

    from bumpcosmology_b200 import _lib
    from bumpcosmology_b200.catalogs import THETA_DEFAULT, draw_prior_thetas
    from bumpcosmology_b200.likelihood import Hyperlikelihood, ShardedHyperlikelihood

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference)")
    _lib.load()
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    cat, gen_s = workload_catalog(args.workload)
    thetas = np.vstack([THETA_DEFAULT, draw_prior_thetas(15, seed=5)])
    if args.wa:   # append wa: 0.3 at the fiducial point, seeded uniform(-1, 1) elsewhere
        wa_col = np.concatenate([[0.3], np.random.default_rng(11).uniform(-1, 1, len(thetas) - 1)])
        thetas = np.hstack([thetas, wa_col[:, None]])
    THETA_DEFAULT = thetas[0]
    t0 = time.time()
    if world > 1:
        like = ShardedHyperlikelihood(cat.as_args(), device=local_rank, exchange=args.exchange, wa=args.wa)
        local = like.local
    else:
        like = local = Hyperlikelihood(*cat.as_args(), device=local_rank, wa=args.wa)
    upload_s = time.time() - t0
    n_local = local.nobs * local.nsamp + local.nsel
    K, W = args.steps, args.warmup

    def all_max(x):
        if dist is None:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local_rank) if rank == 0 else None
    # ---- device-timed region: K back-to-back evaluations, inputs resident in HBM
    if world == 1:
        local.time_evals(THETA_DEFAULT, W)
        torch.cuda.synchronize()
        ms_total, _ = local.time_evals(THETA_DEFAULT, K)
        torch.cuda.synchronize()
    elif like.exchange in ("nccl", "p2p"):
        # the exchange is part of the library's CUDA graph: time K graph replays with events on the context stream
        like(THETA_DEFAULT)
        local.time_evals(THETA_DEFAULT, W)
        dist.barrier()
        torch.cuda.synchronize()
        ms_local, _ = local.time_evals(THETA_DEFAULT, K)
        torch.cuda.synchronize()
        dist.barrier()
        ms_total = all_max(ms_local)
    else:
        like(THETA_DEFAULT)
        for _ in range(W):
            like.launch()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            like.launch()
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        ms_total = all_max(e0.elapsed_time(e1))
    # ---- end to end through the public host API: host theta in, host result out, every step
    for i in range(W):
        like(thetas[i % len(thetas)])
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K):
        res = like(thetas[i % len(thetas)])
    torch.cuda.synchronize()
    e2e_s = all_max(time.perf_counter() - t0)
    # ---- keep the SAME load running (every rank the same number of evaluations: the exchange is a collective) until
    # nvidia-smi, which samples every 20 ms, has seen it for about half a second
    n_extra = int(min(5000, max(20, math.ceil(500.0 / max(ms_total / K, 1e-3)))))
    if world == 1 or like.exchange in ("nccl", "p2p"):
        local.time_evals(THETA_DEFAULT, n_extra)
    else:
        for _ in range(n_extra):
            like.launch()
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    if clocks is not None:
        clocks["note"] = (f"sampled over the timed region and {n_extra} further evaluations of the same load on every "
                          "rank (the timed region alone is shorter than a few sampling periods)")
    # ---- dominant kernel alone (events around each launch on its stream)
    nk = max(3, min(K, 10))
    _, ker_ms = local.time_evals(THETA_DEFAULT, nk, kernel=True)
    ker_ms /= nk
    ker_ms_max = all_max(ker_ms)
    # per-kernel timeline of one directly launched evaluation (GPU global timer; every rank launches it: the
    # peer-memory exchange inside the epilogue is a collective), median of 5
    tls = [local.timeline(THETA_DEFAULT) for _ in range(5)]
    timeline = {k: [round(float(np.median([t[k][i] for t in tls])), 2) for i in (0, 1)] for k in tls[0]}

    # ---- N > 1: the sharded result against an unsharded evaluation of the whole catalog on rank 0's GPU
    result_check = None
    if world > 1:
        # at a theta that differs from the one of the evaluation before it: a stale partial from the previous exchange
        # (same theta in the timed loops) would otherwise go unnoticed
        like(thetas[2])
        res_sh = like(thetas[1])
        neff_parts = [None] * world
        dist.all_gather_object(neff_parts, res_sh.neff)
        flat_sh = _flat(res_sh)
        ok = torch.ones(1, device="cuda")
        if rank == 0:
            full = Hyperlikelihood(*cat.as_args(), device=local_rank, wa=args.wa)
            r1 = full(thetas[1])
            full.close()
            f1 = _flat(r1)
            gs = max(1.0, float(np.max(np.abs(r1.dloglike))))
            floor = np.ones_like(f1)
            floor[4:4 + len(r1.dloglike)] = gs
            e_hdr = float(np.max(np.abs(flat_sh - f1) / np.maximum(np.abs(f1), floor)))
            neff_sh = np.concatenate(neff_parts)
            e_neff = float(np.max(np.abs(neff_sh - r1.neff) / np.maximum(np.abs(r1.neff), 1.0)))
            result_check = {"max_rel_vs_1gpu": max(e_hdr, e_neff), "header_and_gradients": e_hdr, "neff": e_neff,
                            "outputs_compared": int(len(f1) + len(neff_sh)), "tolerance": MULTIRANK_TOL,
                            "logl": res_sh.logl, "logl_1gpu": r1.logl,
                            "ok": bool(max(e_hdr, e_neff) <= MULTIRANK_TOL)}
            ok[0] = 1.0 if result_check["ok"] else 0.0
        dist.broadcast(ok, src=0)
        multirank_ok = bool(ok.item() > 0.5)
    else:
        multirank_ok = True
        result_check = {"logl": res.logl, "neff_sel": res.neff_sel}
    plan = local.plan()
    launches = local.launches_per_eval
    exchange = getattr(like, "exchange", "none") if world > 1 else "none"
    ntheta, nobs_local = local.ntheta, local.nobs
    like.close() if hasattr(like, "close") else local.close()

    # ---- the other BASELINE.json shapes (every rank takes part: they are sharded like the headline)
    configs = None
    if not args.no_configs and not args.wa and args.workload == "o5":
        configs = {}
        for key, name, wa in (("gwtc3", "gwtc3", False), ("o4", "o4", False), ("o5_wa", cat, True)):
            try:
                configs[key] = time_config(name, wa, world, local_rank, dist)
            except Exception as e:  # noqa: BLE001
                configs[key] = {"error": str(e)}
                if dist is not None:
                    raise
    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        if not multirank_ok:
            sys.exit(3)
        return

    hbm_peak, peak_src = measured_peaks()
    pk = fp64_peak(local_rank)
    fp64_rate = pk.get("dfma_per_s") or FP64_PEAK_FALLBACK
    inst = ALGO_FP64_INST_PER_SAMPLE * n_local / (ker_ms * 1e-3)
    hbm_achieved = ALGO_BYTES_PER_SAMPLE * n_local / (ker_ms * 1e-3) / 1e9
    padded = plan["padded_samples"]
    value = K / (ms_total * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config_of(args, cat),
        "run": {"sharding": f"whole events and a contiguous injection range per rank, {world} rank(s)",
                "exchange": exchange, "plan": plan, "catalog_gen_s": round(gen_s, 1), "upload_s": round(upload_s, 2),
                "stream_kernel_ms_max_over_ranks": ker_ms_max,
                "timeline_us": timeline,
                "timeline_note": "[first block start, last block end] of each kernel of one directly launched "
                                 "evaluation on rank 0, microseconds from the first kernel's start"},
        "e2e": {"value": K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 8 * ntheta,
                "d2h_bytes_per_step": 8 * (_lib.OUT_HEADER + nobs_local),
                "note": "public host call Hyperlikelihood.__call__(theta): theta from host memory, result to "
                        "host memory, wall clock; the catalog is uploaded once (upload_s), as the reference's "
                        "jitted model closes over its data"},
        "gpu_launches": K * launches,
        "clocks": clocks,
        "roofline": {"bound": "fp64 pipe", "achieved": inst, "peak": fp64_rate, "unit": "FP64 warp-lane inst/s",
                     "frac": inst / fp64_rate,
                     "traffic": 8 * 7 * padded,
                     "kernel": "stream_kernel", "kernel_ms": ker_ms,
                     "peak_source": "DFMA issue rate measured in this run by bumpcosmology_b200/bump_peak "
                                    f"({pk.get('fp64_tflops', 'n/a')} TFLOP/s)" if "dfma_per_s" in pk else
                                    "fallback: DFMA rate measured on this pool's B200s in round 1 (bump_peak failed)",
                     "algorithmic_inst_per_sample": ALGO_FP64_INST_PER_SAMPLE,
                     # the same against the FP64 issue rate AT THE SM CLOCK SAMPLED UNDER THIS LOAD: the peak above is
                     # a short burst at the maximum clock, while an O5-size pass runs into the board's power cap
                     "frac_at_sampled_clock": (inst / (fp64_rate * clocks["sm_mhz"] / clocks["sm_max_mhz"])
                                               if clocks and clocks.get("sm_mhz") and clocks.get("sm_max_mhz") else None),
                     "hbm": {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                             "algorithmic_bytes_per_sample": ALGO_BYTES_PER_SAMPLE, "peak_source": peak_src},
                     "note": "the path is transcendental-bound (8 exp + 4 reciprocals per sample, no contraction): the "
                             "binding roofline is the FP64 pipe, as BASELINE.json's north star allows; achieved = "
                             "217.5 FP64-pipe instructions per sample (frozen algorithmic count, DESIGN.md section 4) x "
                             "this rank's samples / the kernel's launch time; traffic = the 7 resident fp64 columns "
                             "of the padded samples, which the kernel reads exactly once (ncu: profiles/)"},
        "result_check": result_check,
    }
    if configs is not None:
        line["configs"] = configs
    cpu_gwtc3 = None
    if world == 1 and not args.no_cpu_baseline and not args.wa:
        cb = cpu_arm(cat, 6, 1)
        line["cpu_baseline"] = {"value": cb["value"], "unit": UNIT, "cores": cb["cores"], "kind": "port",
                                "sample": cb["sample"], "s_per_eval_median": cb["s_per_eval_median"],
                                "host_cpus": cb["host_cpus"], "engine": cb["engine"]}
        if not args.no_nuts:
            from bumpcosmology_b200.catalogs import make_catalog
            g = time_cpu_port(make_catalog("gwtc3"), 20, 3)
            cpu_gwtc3 = g["s_per_eval_median"] if g else None
    if world == 1 and not args.no_nuts and not args.wa and args.workload == "o5":
        try:
            line["nuts"] = {
                "gwtc3_nuts": nuts_record("gwtc3_nuts", 1000, 1000, 4, local_rank, cpu_gwtc3),
                # the standard catalog keeps ~3 % of its samples at the model's hard cut m >= 5 (intensity_models.py:
                # 149): logL jumps as (h, Om, w) move samples across it and NUTS stalls, so
                # this run is bounded to 4 x (60 + 60) and reported as found
                "gwtc3": nuts_record("gwtc3", 60, 60, 4, local_rank, cpu_gwtc3),
                "metric": "min bulk-ESS over the 14 likelihood sites / wall seconds (warm-up included) and / "
                          "sampling seconds; rank-normalised split-chain ESS (arviz formula)"}
        except Exception as e:  # noqa: BLE001
            line["nuts"] = {"error": str(e)}
    out.emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if not multirank_ok:
        sys.exit(3)


if __name__ == "__main__":
    main()
