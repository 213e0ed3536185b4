#!/usr/bin/env python
"""bench.py -- EM iterations/s on the 10M-read store; bootstrap replicates/s at N GPUs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C3|C2|small] [--impl reference]

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for the definitions:

  value   step = one complete EM (em::em_par semantics, stop rule niter > 1) on the store
          resident in HBM; value = E+M iterations / s.  At N > 1 (torchrun, one process
          per GPU) rank 0 generates the store and uploads it, ONE NCCL broadcast
          distributes it (timed, "bcast_ms") and every rank runs the same K EMs on its
          copy (a single EM does not shard: replicas only); value = total iterations/s
          over all ranks with the max-over-ranks time.
  bootstrap  the path that shards: max(K,8) replicates per rank, global replicate g on
          rank g mod N, no collective on the data path; "bootstrap.replicates_per_sec"
          = all replicates / max-over-ranks time, reported at every N.
  e2e     the same through the public API with HOST (pinned) buffers: store upload +
          layout + EM + download of the counts inside the timed region.
  --impl reference   the CPU restatement of the reference's rayon em_par
          (oracle/em_par_port.c) on all host cores, bounded sample per step.
"""
from __future__ import annotations

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
    "C3": "C3: synthetic 10M reads x 200k transcripts, avg 8 aln/read (nnz~80M), f32 probs, f64 counts",
    "C2": "C2: synthetic 1M reads x 50k transcripts, avg 6 aln/read (nnz~6M), f32 probs, f64 counts",
    "small": "small: synthetic 50k reads x 5k transcripts (smoke only)",
}


def algorithmic_bytes(n_reads, nnz, n_txps):
    """B_iter = 8*nnz + 4*(N+1) + 24*M."""
    return 8 * nnz + 4 * (n_reads + 1) + 24 * n_txps


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 6:
                self.rows.append(f)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for f in self.rows:
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def gen_store(workload, pinned):
    from oarfish_b200 import synth
    return synth.make_config(workload, pinned=pinned)


# ---------------------------------------------------------------------------------------------
# reference arm: the CPU restatement of em_par on the host cores
# ---------------------------------------------------------------------------------------------

def cpu_em_par_sample(s, sweeps_per_step):
    """One bounded sample: `sweeps_per_step` loop sweeps + the final one of em_par."""
    from oracle import oracle
    ps = oracle.PortStore(s.row_ptr, s.txp_id, s.prob, s.n_txps)
    try:
        t0 = time.perf_counter()
        _, niter, _, sweeps = ps.em_par(max_iter=sweeps_per_step, conv_thresh=1e-3)
        dt = time.perf_counter() - t0
    finally:
        ps.close()
    return (sweeps + 1) / dt, sweeps + 1, dt, oracle.num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle
    s = gen_store(args.workload, pinned=False)
    ps = oracle.PortStore(s.row_ptr, s.txp_id, s.prob, s.n_txps)
    sweeps_per_step = 50 if args.workload == "C3" else 200
    times, iters = [], 0
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        _, _, _, sweeps = ps.em_par(max_iter=sweeps_per_step, conv_thresh=1e-3)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt); iters += sweeps + 1
    ps.close()
    total = sum(times)
    value = iters / total
    cores = oracle.num_threads()
    sample = (f"{sweeps_per_step} loop sweeps + final sweep of em_par per step on the full {args.workload} store, "
              f"{cores} OpenMP threads; restated reference (C), not the Rust binary")
    line = {
        "impl": "reference", "metric": "em_iterations_per_sec", "value": value, "unit": "iterations/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload], "n_reads": s.n_reads, "nnz": s.nnz, "n_txps": s.n_txps},
        "cpu_baseline": {"value": value, "unit": "iterations/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------

def run_gpu(args):
    import torch
    from oarfish_b200 import DeviceStore
    from oarfish_b200 import dist as odist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torchrun (one process per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    multi = world > 1
    if multi:
        import torch.distributed as dist
        # stdout carries exactly one JSON line: NCCL's own banner / debug lines (NCCL_DEBUG=VERSION|INFO on some
        # boxes) go to a file instead
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/oarfish_bench_nccl.%h.%p.log")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if multi:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if not multi:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if not multi:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    peak, peak_src = measured_peak_gbs()
    K, W = args.steps, args.warmup
    seed = 4
    s = gen_store(args.workload, pinned=True) if rank == 0 else None
    bcast_ms = None
    store_bytes = None
    if multi:
        barrier()
        t0 = time.perf_counter()
        rp, tx, pr, ax, n_txps = odist.broadcast_store(*( (s.row_ptr, s.txp_id, s.prob, s.n_txps) if rank == 0 else (None, None, None, 0)),
                                                       src=0, device=dev)
        barrier()
        bcast_ms = 1e3 * max_over_ranks(time.perf_counter() - t0)
        n_reads, nnz = rp.numel() - 1, tx.numel()
        t0 = time.perf_counter()
        ds = DeviceStore(rp, tx, pr, n_txps, device=local_rank)
        barrier()
        build_ms = 1e3 * max_over_ranks(time.perf_counter() - t0)
    else:
        n_reads, nnz, n_txps = s.n_reads, s.nnz, s.n_txps
        t0 = time.perf_counter()
        ds = DeviceStore(s.row_ptr, s.txp_id, s.prob, n_txps, device=local_rank)
        torch.cuda.synchronize()
        build_ms = 1e3 * (time.perf_counter() - t0)
    store_bytes = 8 * (n_reads + 1) + 8 * nnz
    layout = ds.layout_info()

    sampler = ClockSampler(local_rank) if rank == 0 else None

    # ---- timed region: K steps ------------------------------------------------------------------
    out_host = np.empty(n_txps, dtype=np.float64)
    launches = 0
    iters = 0

    def step(i):
        nonlocal launches, iters
        # every rank runs the same complete EM on its resident copy of the store (a single EM does not
        # shard: replicas only); the sharded bootstrap path is timed separately below
        ds.em(max_iter=1000, conv_thresh=1e-3, min_iter=1, out=out_host)
        c = ds.counters()
        launches += c["launches"]; iters += c["sweeps"]

    for i in range(W):
        step(i)
    launches = 0; iters = 0
    barrier()
    if sampler:
        sampler.start()
    t0 = time.perf_counter()
    for i in range(W, W + K):
        step(i)
    barrier()
    elapsed = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop() if sampler else None
    total_iters = sum_over_ranks(iters)
    total_launches = sum_over_ranks(launches)
    value = total_iters / elapsed

    # ---- roofline of the dominant kernel (fused E+M sweep), CUDA events on the launching stream --
    roof = None
    if rank == 0:
        prev = torch.full((n_txps,), n_reads / n_txps, dtype=torch.float64, device=dev)
        curr = torch.zeros(n_txps, dtype=torch.float64, device=dev)
        ds.sweep_timed(prev, curr, 5)
        reps = 50
        ms = ds.sweep_timed(prev, curr, reps) / reps
        alg = algorithmic_bytes(n_reads, nnz, n_txps)
        achieved = alg / (ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": {2: "em_sweep_tiled", 3: "em_sweep_lane"}.get(layout["kernel"], "em_sweep_rowgroup"),
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg, "us_per_launch": ms * 1e3,
                "frac_of_nominal_8TBs": achieved / 8000.0}
        tr = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tr):
            try:
                roof["traffic"] = json.load(open(tr)).get(args.workload)
            except Exception:
                pass

    # ---- bootstrap replicates/s: global replicate g runs on rank g mod N (no data-path collective) ----
    B = max(K, 8)                       # replicates per rank (several, so that unequal iteration counts average out)
    ds.bootstrap(1, seed, first_replicate=100000 + rank)     # warm-up (weights buffers, weighted graph)
    barrier()
    t0 = time.perf_counter()
    if args.boot_schedule == "dynamic" and multi:
        # opt-in: ranks pull global replicate ids from a shared counter (oarfish_b200.dist.ReplicateQueue)
        from oarfish_b200 import dist as odist
        _ids, res = odist.run_replicates_dynamic(lambda g: ds.bootstrap(1, seed, first_replicate=g)[1], world * B)
        nit = np.concatenate(res) if res else np.zeros(0, dtype=np.uint32)
    else:
        _, nit = ds.bootstrap(B, seed, first_replicate=rank, replicate_stride=world)
    boot_launches = ds.counters()["launches"]
    barrier()
    dtb = max_over_ranks(time.perf_counter() - t0)
    boot_iters = sum_over_ranks(float(nit.sum() + 2 * len(nit)))
    boot = {"replicates_per_sec": world * B / dtb, "replicates": world * B, "iterations_per_sec": boot_iters / dtb,
            "niter_rank0": [int(x) for x in nit], "min_iter": 50, "schedule": args.boot_schedule if multi else "static", "bcast_ms": bcast_ms, "gpu_launches_rank0": int(boot_launches)}

    # ---- e2e: host (pinned) buffers through the public API ------------------------------------------
    e2e = None
    if not multi:
        ds.close()
        e_iters = 0
        parts = [0.0, 0.0, 0.0]   # create (upload + validation + layout), EM + counts download, destroy
        for i in range(1 + max(1, K // 2)):
            if i == 1:
                torch.cuda.synchronize(); t0 = time.perf_counter(); e_iters = 0; parts = [0.0, 0.0, 0.0]
            ta = time.perf_counter()
            d2 = DeviceStore(s.row_ptr, s.txp_id, s.prob, n_txps, device=local_rank)
            tb = time.perf_counter()
            d2.em(max_iter=1000, conv_thresh=1e-3, min_iter=1, out=out_host)
            e_iters += d2.counters()["sweeps"]
            tc = time.perf_counter()
            d2.close()
            td = time.perf_counter()
            parts[0] += tb - ta; parts[1] += tc - tb; parts[2] += td - tc
        torch.cuda.synchronize()
        dte = time.perf_counter() - t0
        n_e = max(1, K // 2)
        e2e = {"value": e_iters / dte, "unit": "iterations/s", "h2d_bytes_per_step": store_bytes,
               "d2h_bytes_per_step": 8 * n_txps, "ms_per_step": 1e3 * dte / n_e,
               "ms_create_em_destroy": [round(1e3 * x / n_e, 2) for x in parts],
               "includes": "pinned-host store upload + layout build + EM to convergence + counts download"}
    else:
        # the store crosses PCIe once on rank 0 and NVLink once per rank; amortise that over the K steps
        dte = elapsed + (bcast_ms + build_ms) * 1e-3
        e2e = {"value": total_iters / dte, "unit": "iterations/s",
               "h2d_bytes_per_step": store_bytes / K, "d2h_bytes_per_step": 8 * n_txps,
               "includes": "store upload on rank 0 + NCCL broadcast + per-rank layout build (once) + K EMs per rank with counts download"}

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload ----------------------
    cpu = None
    if rank == 0 and not multi and not args.no_cpu_baseline:
        sweeps = 250 if args.workload == "C3" else 1000
        v, n_it, dt, cores = cpu_em_par_sample(s, sweeps)
        cpu = {"value": v, "unit": "iterations/s", "cores": cores, "kind": "port",
               "sample": f"{n_it} E+M sweeps of em_par on the full {args.workload} store in {dt:.1f} s; restated reference (C), not the Rust binary"}

    if rank == 0:
        line = {
            "metric": "em_iterations_per_sec", "value": value, "unit": "iterations/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": 1e3 * elapsed / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload], "n_reads": n_reads, "nnz": nnz, "n_txps": n_txps,
                       "step": "one EM to convergence per rank (em_par rule, min_iter 1, thr 1e-3); bootstrap replicates timed separately",
                       "l2": "inputs (0.66 GB/sweep) exceed the 126 MB L2; no flush needed",
                       "parallelism": f"{world} GPU(s): EM replicas for `value`, bootstrap replicates sharded g mod N; no data-path collective",
                       "layout": layout},
            "iterations_per_step": total_iters / (K * world),
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(total_launches),
            "bootstrap": boot, "clocks": clocks, "store_build_ms": build_ms,
        }
        print(json.dumps(line), flush=True)
    if multi:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="C3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--boot-schedule", default="static", choices=["static", "dynamic"],
                    help="bootstrap leg at N > 1: replicate g on rank g mod N (default) or pulled from a shared counter")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
