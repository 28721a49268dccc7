#!/usr/bin/env python
"""bench.py — hyperlikelihood logL+grad evaluations per second on the O5-shaped mock (BASELINE.json metric).

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the CPU path (oracle port) on the box's host cores

One "step" = one evaluation of the hot path (theta -> logL parts, both 14-parameter gradients, Neff's) over the
whole catalog, which is resident in HBM (uploaded once, exactly as the reference's jitted model closes over its
data; NUTS only ever changes theta).  `value` times K back-to-back evaluations with CUDA events on the stream
they are launched on; `e2e` times the public host call (host theta in, host result out) by wall clock.
Prints ONE JSON line on rank 0.
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

METRIC = "hyperlikelihood logL+grad evals/sec (O5 mock)"
UNIT = "evals/s"
ALGO_BYTES_PER_SAMPLE = 32           # m1_det, q, d_L, pdraw as fp64
ALGO_FP64_INST_PER_SAMPLE = 700      # forward + 14-parameter gradient, FP64-pipe instructions
# Measured with ncu on the same command (profiles/r01b_o5_stream_kernel_opmix.txt / _ncu.txt): FP64-pipe warp
# instructions the streaming kernel actually executes per 32 samples, DRAM bytes it reads per (real) sample, and the
# issue cycles its instruction mix needs per warp-sample under the measured B200 issue rules
# (profiles/r01_fp64_issue_microbench.txt: an FP64 instruction occupies its sub-partition for max(2, distinct
# vector-register operands) cycles, every other instruction for ~1; nothing hides in the FP64 pipe's second cycle).
EXEC_FP64_INST_PER_SAMPLE = 217.5
EXEC_OTHER_INST_PER_SAMPLE = 182.4
EXEC_FP64_3REG_PER_SAMPLE = 86.5
DRAM_BYTES_PER_SAMPLE = 56.3


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


def _sample_of(cat, frac):
    ne = max(1, int(round(cat.nobs * frac)))
    ns = max(1, int(round(cat.nsel * frac)))
    data = (cat.m1s_det[:ne], cat.qs[:ne], cat.dls[:ne], cat.pdraw[:ne], cat.m1s_det_sel[:ns], cat.qs_sel[:ns],
            cat.dls_sel[:ns], cat.pdraw_sel[:ns], cat.Ndraw)
    n_sample = ne * cat.nsamp + ns
    desc = (f"{ne} of {cat.nobs} events x {cat.nsamp} samples + {ns} of {cat.nsel} injections "
            f"({100 * n_sample / cat.n_elements:.2f}% of the workload's elements)")
    return data, n_sample / cat.n_elements, desc


def cpu_port_cpp(cat, frac, evals, warm):
    """oracle/bump_cpu.cpp: fused single-pass C++/OpenMP port, all host threads.  Returns (evals/s extrapolated
    linearly to the full workload, per-eval seconds on the sample, sample description, threads) or None."""
    try:
        from oracle import bump_cpu
        data, scale, desc = _sample_of(cat, frac)
        port = bump_cpu.CpuPort(*data)
    except Exception as e:  # noqa: BLE001  (no compiler on the box, ...)
        print(f"[bench] C++ CPU port unavailable: {e}", file=sys.stderr)
        return None
    from bumpcosmology_b200.catalogs import THETA_DEFAULT, draw_prior_thetas
    thetas = np.vstack([THETA_DEFAULT, draw_prior_thetas(7, seed=5)])
    for i in range(warm):
        port.evaluate(thetas[i % len(thetas)])
    ts = []
    for i in range(evals):
        t0 = time.perf_counter()
        port.evaluate(thetas[i % len(thetas)])
        ts.append(time.perf_counter() - t0)
    t = float(np.median(ts))
    threads = port.threads
    port.close()
    return scale / t, t, desc, threads


def cpu_port_torch(cat, frac, evals, warm):
    """oracle/bump_oracle.py: the parity oracle (eager torch fp64 + autograd), all host threads."""
    import torch

    from bumpcosmology_b200.catalogs import THETA_DEFAULT, draw_prior_thetas
    from oracle import bump_oracle as bo
    data, scale, desc = _sample_of(cat, frac)
    thetas = np.vstack([THETA_DEFAULT, draw_prior_thetas(7, seed=5)])
    chunk = max(1, 2_000_000 // max(cat.nsamp, 1))
    for i in range(warm):
        bo.evaluate(thetas[i % len(thetas)], data, grad=True, event_chunk=chunk)
    ts = []
    for i in range(evals):
        t0 = time.perf_counter()
        bo.evaluate(thetas[i % len(thetas)], data, grad=True, event_chunk=chunk)
        ts.append(time.perf_counter() - t0)
    t = float(np.median(ts))
    return scale / t, t, desc, torch.get_num_threads()


def cpu_baseline(cat, evals=8, warm=2):
    """Both CPU stand-ins for "the reference's JAX on the host cores"  on bounded samples of
    the workload; the FASTER one is the reported baseline."""
    out = {}
    cpp = cpu_port_cpp(cat, min(1.0, 6_000_000 / cat.n_elements), evals, warm)
    if cpp:
        out["cpp_openmp"] = {"value": cpp[0], "sample_s_per_eval": cpp[1], "sample": cpp[2], "cores": cpp[3]}
    tv = cpu_port_torch(cat, min(1.0, 900_000 / cat.n_elements), max(3, evals // 2), 1)
    out["torch_eager"] = {"value": tv[0], "sample_s_per_eval": tv[1], "sample": tv[2], "cores": tv[3]}
    best = max(out, key=lambda k: out[k]["value"])
    b = out[best]
    return {"value": b["value"], "unit": UNIT, "cores": b["cores"], "kind": "port", "engine": best,
            "sample": b["sample"] + f", {evals} evals after {warm} warm-ups, median, extrapolated linearly in "
                                    "element count",
            "sample_s_per_eval": b["sample_s_per_eval"], "host_cpus": os.cpu_count(), "engines": out}


def run_reference(args, rank, out):
    """The reference arm: the reference's algorithm on the host cores.  The reference itself (JAX/numpyro) is not
    installable in this image, so this times the faster CPU port of oracle/ (normally the fused C++/OpenMP one)."""
    if rank != 0:
        return
    cat, _ = workload_catalog(args.workload)
    frac = min(1.0, 6_000_000 / cat.n_elements)      # 10 % of O5 per step: the whole arm stays within a few minutes
    cpp = cpu_port_cpp(cat, frac, args.steps, args.warmup)
    if cpp:
        v, t_step, sample, threads = cpp
        engine = "cpp_openmp (oracle/bump_cpu.cpp: fused single pass, log space, libm)"
    else:
        v, t_step, sample, threads = cpu_port_torch(cat, min(1.0, 900_000 / cat.n_elements), args.steps, args.warmup)
        engine = "torch_eager (oracle/bump_oracle.py)"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {cat.nobs} events x {cat.nsamp} samples + {cat.nsel} injections",
                   "note": "reference JAX stack is not installable here; CPU arm = " + engine + ", all host threads; "
                           "each step is one logL+grad evaluation of a bounded sample, extrapolated linearly"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "sample_s_per_eval": t_step, "host_cpus": os.cpu_count(), "engine": engine},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    out.emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--workload", default=os.environ.get("BUMP_BENCH_WORKLOAD", "o5"))
    ap.add_argument("--exchange", default=os.environ.get("BUMP_EXCHANGE", "p2p"), choices=("torch", "nccl", "p2p"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    from bumpcosmology_b200.likelihood import Hyperlikelihood, ShardedHyperlikelihood, shard_catalog

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
        t = torch.tensor([ms_local], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
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
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
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
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    clock_note = None
    if sampler is not None and len(sampler.lines) < 5:
        # short timed regions (small catalogs) end before nvidia-smi delivers its first samples: keep the same load
        # running, untimed, until a handful of samples exist
        t_end = time.perf_counter() + 2.0
        while len(sampler.lines) < 5 and time.perf_counter() < t_end:
            local.time_evals(THETA_DEFAULT, 50) if world == 1 else time.sleep(0.05)
        clock_note = "timed region shorter than the sampling period: sampled under the same load right after it"
    clocks = sampler.stop() if sampler else None
    if clocks is not None and clock_note:
        clocks["note"] = clock_note
    # ---- dominant kernel alone (events around each launch on its stream)
    _, ker_ms = local.time_evals(THETA_DEFAULT, max(3, min(K, 10)), kernel=True)
    ker_ms /= max(3, min(K, 10))
    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    hbm_peak, peak_src = measured_peaks()
    algo_bytes = ALGO_BYTES_PER_SAMPLE * n_local
    achieved = algo_bytes / (ker_ms * 1e-3) / 1e9
    pk = fp64_peak(local_rank)
    fp64 = None
    if "dfma_per_s" in pk:
        inst = EXEC_FP64_INST_PER_SAMPLE * n_local / (ker_ms * 1e-3)
        algo = ALGO_FP64_INST_PER_SAMPLE * n_local / (ker_ms * 1e-3)
        fp64 = {"bound": "fp64 pipe", "achieved": inst, "peak": pk["dfma_per_s"], "unit": "fp64 inst/s",
                "frac": inst / pk["dfma_per_s"], "peak_fp64_tflops": pk["fp64_tflops"],
                "inst_per_sample_executed": EXEC_FP64_INST_PER_SAMPLE,
                "algorithmic_inst_per_sample": ALGO_FP64_INST_PER_SAMPLE, "algorithmic_frac": algo / pk["dfma_per_s"],
                "note": "frac = FP64-pipe instructions the kernel executes (ncu, profiles/) x samples / kernel time, "
                        "over the DFMA issue rate measured by bump_peak in this run; algorithmic_frac uses SURVEY.md "
                        "8d's 700 instructions/sample and exceeds 1 because the linear-space kernel needs 3.2x fewer"}
        # issue ceiling of this instruction mix: 2 cycles per FP64 instruction, +1 for each with three distinct
        # register operands, +1 per other instruction, per warp-sample and sub-partition (4 per SM)
        cyc_model = 2 * EXEC_FP64_INST_PER_SAMPLE + EXEC_FP64_3REG_PER_SAMPLE + EXEC_OTHER_INST_PER_SAMPLE
        sms = pk.get("sms", 148)
        clk = ((clocks or {}).get("sm_mhz") or pk.get("max_clock_mhz", 1965)) * 1e6   # SM clock sampled under load
        cyc_meas = ker_ms * 1e-3 * clk * sms * 4 / (n_local / 32)
        fp64["issue_model"] = {"cycles_per_warp_sample": cyc_model, "measured_cycles_per_warp_sample": cyc_meas,
                               "frac": cyc_model / cyc_meas,
                               "fp64_frac_at_model": 2 * EXEC_FP64_INST_PER_SAMPLE / cyc_model,
                               "note": "what the kernel's own instruction mix allows under the issue rules measured by "
                                       "tools/micro/fp64_operands.cu (profiles/r01_fp64_issue_microbench.txt); "
                                       "shared-memory wavefronts run at 82 % of peak beside it"}
    value = K / (ms_total * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {cat.nobs} events x {cat.nsamp} samples + {cat.nsel} injections"
                               + (", w0-wa dark energy (15 parameters)" if args.wa else ""),
                   "elements": cat.n_elements, "sharding": f"events and injections over {world} rank(s)",
                   "exchange": getattr(like, "exchange", "none") if world > 1 else "none",
                   "l2": "per-rank resident columns %.2f GB > 126 MB L2" % (56 * n_local / 1e9)
                         if 56 * n_local > 2.5e8 else "columns fit in L2 (NUTS re-reads the same data)",
                   "plan": local.plan(), "catalog_gen_s": round(gen_s, 1), "upload_s": round(upload_s, 2)},
        "e2e": {"value": K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 8 * local.ntheta,
                "d2h_bytes_per_step": 8 * (_lib.OUT_HEADER + local.nobs),
                "note": "public host call Hyperlikelihood.__call__(theta): theta from host memory, result to "
                        "host memory, wall clock; the catalog is uploaded once (upload_s), as the reference's "
                        "jitted model closes over its data"},
        "gpu_launches": K * local.launches_per_eval,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved / hbm_peak, "traffic": DRAM_BYTES_PER_SAMPLE * n_local,
                     "traffic_note": "dram__bytes_read+write of one launch from ncu --set full (profiles/), per sample "
                                     "x this rank's samples: 7 resident fp64 columns incl. hoisted logs",
                     "kernel": "stream_kernel",
                     "kernel_ms": ker_ms, "peak_source": peak_src,
                     "note": "algorithmic 32 B/sample; the fp64 path is FP64-pipe bound, see fp64_pipe"},
        "fp64_pipe": fp64,
        "result_check": {"logl": res.logl, "neff_sel": res.neff_sel},
    }
    if world == 1 and not args.no_cpu_baseline and not args.wa:
        line["cpu_baseline"] = cpu_baseline(cat)
    out.emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
