#!/usr/bin/env python
"""bench.py -- EM iterations/s on the 10M-read store as a sharded bootstrap job (BASELINE config 4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C3|C2|small] [--impl reference]

One JSON line on stdout (rank 0).  Definitions (DESIGN.md "Measurement"):

  step    one batch of R = --boot-batch (default 5) bootstrap replicates (em.rs:273-314): multinomial
          read weights + one weighted EM to convergence (do_em rule, min_iter 50, thr 1e-3) + the
          replicate's counts copied to the host.  K steps = K*R replicates IN TOTAL at every N --
          the driver's --steps 20 is BASELINE config 4 (100 replicates) -- so the job is
          strong-scaled: ranks pull global replicate ids from a shared counter (dynamic schedule;
          results are keyed by (seed, id), i.e. independent of N and of who ran what), no
          collective on the data path.
  value   E+M iterations (sweeps) of all replicates on all ranks / max-over-ranks time, store
          resident in HBM.  "replicates_per_sec" is the same job in BASELINE's other unit.
  em_single_gpu   the non-bootstrap EM (em_par rule) on one GPU: iterations/s, ms per EM (rank 0).
  e2e     N = 1: the same metric through the public API from HOST (pinned) buffers, per step:
          store upload + validation + layout build + R replicates + counts download + teardown.
          N > 1: the store crosses PCIe once on rank 0 and NVLink once per rank (NCCL broadcast);
          that and the per-rank layout build are added to the timed region.
  roofline  the dominant kernel of the timed region (the bootstrap-weighted fused E+M sweep), CUDA
          events on the store's stream; "roofline_plain_em" the same for the unweighted sweep.
  --impl reference   the CPU restatement of the reference's bootstrap (oracle/em_par_port.c: T
          concurrent sequential do_em's on a pool of all host cores), bounded sample per step.
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
    "C3": "C4 on the C3 store: synthetic 10M reads x 200k transcripts, avg 8 aln/read (nnz~80M), f32 probs, f64 counts; bootstrap replicates sharded over the GPUs",
    "C2": "C2 store: synthetic 1M reads x 50k transcripts, avg 6 aln/read (nnz~6M), f32 probs, f64 counts; bootstrap replicates sharded over the GPUs",
    "small": "small: synthetic 50k reads x 5k transcripts (smoke only)",
}
SEED = 4
THR = 1e-3


_REAL_STDOUT = None


def claim_stdout():
    """stdout carries exactly ONE line, the JSON record: everything libraries print (NCCL's version banner on the first
    communicator, torchrun notices) is sent to stderr by pointing fd 1 at fd 2 for the life of the process."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def algorithmic_bytes(n_reads, nnz, n_txps, weighted=False):
    """B_iter = 8*nnz + 4*(N+1) + 24*M (+ 4*N for u32 bootstrap weights)."""
    return 8 * nnz + 4 * (n_reads + 1) + 24 * n_txps + (4 * n_reads if weighted else 0)


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy bandwidth)"
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


def config_block(workload, n_reads, nnz, n_txps):
    """Identical in both arms (the driver compares it)."""
    return {"workload": WORKLOADS[workload], "n_reads": int(n_reads), "nnz": int(nnz), "n_txps": int(n_txps)}


# ---------------------------------------------------------------------------------------------
# CPU side: the restated reference on the host cores (cpu_baseline leg and --impl reference)
# ---------------------------------------------------------------------------------------------

def host_threads():
    """All host cores: torchrun exports OMP_NUM_THREADS=1 to its workers, so the pool is sized explicitly."""
    from oracle import oracle
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    return oracle.set_num_threads(n)


def cpu_bootstrap_sample(ps, threads, max_iter, seed):
    """One bounded sample of em::bootstrap (em.rs:292-314): `threads` replicates run side by side, each a sequential
    do_em capped at `max_iter` loop sweeps.  Returns (sweeps, seconds of the EM part, wall seconds)."""
    t0 = time.perf_counter()
    _, nit, secs = ps.bootstrap_timed(threads, seed, max_iter=max_iter, conv_thresh=THR, nthreads=threads)
    wall = time.perf_counter() - t0
    sweeps = int(nit.sum()) + len(nit)            # + the final sweep of every replicate (em.rs:245-252)
    return sweeps, float(secs.max()), wall


CPU_SAMPLE_TEXT = ("{T} replicates side by side on {T} threads (one sequential do_em each, as em::bootstrap's pool runs them), "
                   "{m} loop sweeps + final sweep per replicate on the full {w} store; iterations/s = all sweeps / slowest "
                   "replicate's do_em time (drawing + sorting the N-index sample, {setup:.1f} s per replicate, is excluded: a "
                   "full run amortises it over ~600 sweeps); restated reference (C), not the Rust binary")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle
    threads = host_threads()
    s = gen_store(args.workload, pinned=False)
    ps = oracle.PortStore(s.row_ptr, s.txp_id, s.prob, s.n_txps)
    m = 6 if args.workload == "C3" else 60
    tot_sweeps, tot_em, tot_wall = 0, 0.0, 0.0
    for i in range(args.warmup + args.steps):
        if i < args.warmup and i >= 1:
            continue                              # one warm-up sample is enough for a CPU loop; the rest would only burn lease time
        sw, em_s, wall = cpu_bootstrap_sample(ps, threads, m, SEED + i)
        if i >= args.warmup:
            tot_sweeps += sw; tot_em += em_s; tot_wall += wall
    ps.close()
    value = tot_sweeps / tot_em
    sample = CPU_SAMPLE_TEXT.format(T=threads, m=m, w=args.workload, setup=(tot_wall - tot_em) / max(args.steps, 1))
    line = {
        "impl": "reference", "metric": "em_iterations_per_sec", "value": value, "unit": "iterations/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_wall / max(args.steps, 1),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_block(args.workload, s.n_reads, s.nnz, s.n_txps),
        "cpu_baseline": {"value": value, "unit": "iterations/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------

C5_READS, C5_TXPS, C5_AVG, C5_SEED, C5_EXPRESSED = 50_000, 200_000, 6.0, 5, 5_000


def c5_block(args, rank, world, local_rank, barrier, max_over_ranks, sum_over_ranks):
    """BASELINE config 5 scaled to --c5-cells cells x 50k reads x 200k transcripts (the full 5k cells are 1.5 G alignments:
    minutes of host-side generation): one em::em per cell (single_cell.rs:150) batched per rank; cells are independent
    units, rank r owns a contiguous range (strong scaling, no collective).  Timed per rank: upload of the rank's rows
    from pinned host memory + layout + all EMs + sparse results to the host; cells/s = all cells / max-over-ranks time."""
    import torch
    from oarfish_b200 import DeviceStore, synth
    n_total = args.c5_cells
    c0, c1 = rank * n_total // world, (rank + 1) * n_total // world
    st, crp = synth.make_cells([C5_READS] * (c1 - c0), C5_TXPS, C5_AVG, C5_SEED * 7919 + c0, expressed=C5_EXPRESSED)
    if c1 > c0:   # untimed warm-up pass over the same cells (module load, growth of the device memory pool to its working size)
        with DeviceStore(st.row_ptr, st.txp_id, st.prob, C5_TXPS, device=local_rank) as w:
            w.em_batched(crp, conv_thresh=THR)
    barrier()
    t0 = time.perf_counter()
    em_ms, nit, launches = 0.0, np.zeros(0, np.uint32), 0
    if c1 > c0:
        with DeviceStore(st.row_ptr, st.txp_id, st.prob, C5_TXPS, device=local_rank) as d5:
            _, _, _, nit = d5.em_batched(crp, conv_thresh=THR)
            em_ms = d5.timings_ms()["em"]
            launches = d5.counters()["launches"]
    torch.cuda.synchronize()
    dt_local = time.perf_counter() - t0
    barrier()
    dt = max_over_ranks(dt_local)
    em_ms = max_over_ranks(em_ms)
    iters = sum_over_ranks(float(nit.astype(np.float64).sum() + 2 * len(nit)))
    nnz = sum_over_ranks(float(st.nnz))
    launches = sum_over_ranks(float(launches))
    niter_max = max_over_ranks(float(nit.max()) if len(nit) else 0.0)
    return {"workload": f"config 5 scaled: {n_total} cells x {C5_READS} reads x {C5_TXPS} transcripts (about {C5_EXPRESSED} expressed per cell), avg {C5_AVG:g} aln/read, one em::em per cell (do_em rule, thr {THR})",
            "cells": n_total, "n_gpus": world, "cells_per_sec": n_total / dt, "ms": 1e3 * dt, "em_ms_max_rank": em_ms,
            "nnz": int(nnz), "cell_iterations": int(iters), "niter_mean": iters / max(n_total, 1) - 2, "niter_max": int(niter_max),
            "gpu_launches": int(launches),
            "includes": "per rank: upload of its cells' rows (pinned host) + layout + all per-cell EMs + sparse counts to the host"}


def run_gpu(args):
    import torch
    from oarfish_b200 import DeviceStore
    from oarfish_b200 import dist as odist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("--gpus N > 1 must be launched with torchrun (one process per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    multi = world > 1
    if multi:
        import torch.distributed as dist
        # stdout carries exactly one JSON line: NCCL's own banner / debug lines go to a file instead
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/oarfish_bench_nccl.%h.%p.log")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if multi:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(x, op):
        if not multi:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(x):
        return reduce(x, dist.ReduceOp.MAX) if multi else x

    def sum_over_ranks(x):
        return reduce(x, dist.ReduceOp.SUM) if multi else x

    peak, peak_src = measured_peak_gbs()
    K, W, R = args.steps, args.warmup, args.boot_batch
    s = gen_store(args.workload, pinned=True) if rank == 0 else None
    bcast_ms = 0.0
    if multi:
        barrier()
        t0 = time.perf_counter()
        rp, tx, pr, ax, n_txps = odist.broadcast_store(*((s.row_ptr, s.txp_id, s.prob, s.n_txps) if rank == 0 else (None, None, None, 0)),
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

    # ---- the sharded bootstrap job: global replicate ids pulled from a shared counter ------------------
    n_warm, n_timed = W * R, K * R
    out_host = torch.empty((max(n_timed, n_warm, 1), n_txps), dtype=torch.float64).pin_memory()
    stats = {"launches": 0, "sweeps": 0, "replicates": 0, "niter": []}

    def run_range(first, count):
        """Replicates first .. first+count-1, shared over the ranks; this rank's results land in out_host rows."""
        row = 0

        def one(g):
            nonlocal row
            _, nit = ds.bootstrap(1, SEED, first_replicate=first + g, conv_thresh=THR, out=out_host[row])
            row += 1
            c = ds.counters()
            stats["launches"] += c["launches"]; stats["sweeps"] += c["sweeps"]; stats["replicates"] += 1
            stats["niter"].append(int(nit[0]))

        if multi and args.boot_schedule == "dynamic":
            odist.run_replicates_dynamic(one, count)
        elif multi:
            for g in range(rank, count, world):
                one(g)
        else:
            for g in range(count):
                one(g)

    run_range(1_000_000, n_warm)                     # W untimed warm-up steps (other replicate ids than the timed ones)
    for k in stats:
        stats[k] = [] if k == "niter" else 0
    barrier()
    if sampler:
        sampler.start()
    t0 = time.perf_counter()
    run_range(0, n_timed)                            # EXACTLY K steps = K*R replicates in total
    barrier()
    elapsed = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop() if sampler else None
    total_iters = sum_over_ranks(stats["sweeps"])
    total_launches = sum_over_ranks(stats["launches"])
    reps_by_rank = None
    if multi:
        t = torch.zeros(world, dtype=torch.float64, device=dev); t[rank] = stats["replicates"]
        dist.all_reduce(t); reps_by_rank = [int(x) for x in t.tolist()]
    value = total_iters / elapsed

    # ---- roofline of the dominant kernel (the weighted fused E+M sweep), CUDA events on the launching stream ------
    roof = roof_plain = em_single = None
    if rank == 0:
        prev = torch.from_numpy(out_host[0].numpy().copy()).to(dev) if stats["replicates"] else torch.full((n_txps,), n_reads / n_txps, dtype=torch.float64, device=dev)
        prev.clamp_(min=1e-3)
        curr = torch.zeros(n_txps, dtype=torch.float64, device=dev)
        wts = torch.from_numpy(ds.sample_weights(SEED, 0).view(np.int32)).to(dev)
        kname = {2: "em_sweep_tiled"}.get(layout["kernel"], "em_sweep_rowgroup")
        traffic = None
        tr = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tr):
            try:
                traffic = json.load(open(tr))
            except Exception:
                traffic = None

        # The peak this is compared with is a BURST figure (MEASURED_PEAKS.json: copy kernel timed alone), and the job above
        # leaves the GPU under its software power cap (SM clock ~5 % down): time the kernel alone as well, after a
        # second of idle.  What the sustained job achieves per iteration is reported next to it (job.us_per_iteration).
        time.sleep(1.0)

        def roofline(weights, tag):
            ds.sweep_timed(prev, curr, 5, weights)
            reps = 50
            ms = ds.sweep_timed(prev, curr, reps, weights) / reps
            alg = algorithmic_bytes(n_reads, nnz, n_txps, weighted=weights is not None)
            ach = alg / (ms * 1e-3) / 1e9
            r = {"bound": "hbm", "kernel": kname + ("<weighted>" if weights is not None else "<plain>"), "achieved": ach, "peak": peak,
                 "unit": "GB/s", "frac": ach / peak, "traffic": None, "peak_source": peak_src,
                 "algorithmic_bytes_per_launch": alg, "us_per_launch": ms * 1e3, "frac_of_nominal_8TBs": ach / 8000.0,
                 "timing": f"{reps} back-to-back launches, CUDA events on the store's stream, kernel alone after 1 s of idle (burst, like the peak)"}
            if traffic and traffic.get(f"{args.workload}{tag}") is not None:
                r["traffic"] = traffic[f"{args.workload}{tag}"]
                r["traffic_source"] = "static: " + str(traffic.get("source", "profiles/traffic.json")) + " -- not measured by this run"
            return r

        roof = roofline(wts, "_weighted")
        roof_plain = roofline(None, "")
        # the non-bootstrap EM on one GPU (BASELINE's first metric): em_par rule
        one = np.empty(n_txps, dtype=np.float64)
        ds.em(max_iter=1000, conv_thresh=THR, min_iter=1, out=one)
        torch.cuda.synchronize(); t1 = time.perf_counter(); it1 = 0; n1 = 3
        for _ in range(n1):
            r = ds.em(max_iter=1000, conv_thresh=THR, min_iter=1, out=one)
            it1 += ds.counters()["sweeps"]
        dt1 = time.perf_counter() - t1
        em_single = {"iterations_per_sec": it1 / dt1, "ms_per_em": 1e3 * dt1 / n1, "iterations_per_em": it1 / n1, "niter": r.niter,
                     "rule": "em_par (min_iter 1, thr 1e-3), store resident, counts downloaded"}

    # ---- e2e: host (pinned) buffers through the public API ---------------------------------------------------------
    e2e = e2e_single = None
    if not multi:
        ds.close()
        n_e = max(1, K // 4)

        def e2e_steps(kind):
            its, parts = 0, [0.0, 0.0, 0.0]     # create (upload + validation + layout), compute + download, destroy
            t_begin = None
            for i in range(1 + n_e):
                if i == 1:
                    torch.cuda.synchronize(); t_begin = time.perf_counter(); its = 0; parts = [0.0, 0.0, 0.0]
                ta = time.perf_counter()
                d2 = DeviceStore(s.row_ptr, s.txp_id, s.prob, n_txps, device=local_rank)
                tb = time.perf_counter()
                if kind == "bootstrap":
                    d2.bootstrap(R, SEED, first_replicate=2_000_000 + i * R, conv_thresh=THR, out=out_host[:R])
                else:
                    d2.em(max_iter=1000, conv_thresh=THR, min_iter=1, out=out_host[0])
                its += d2.counters()["sweeps"]
                tc = time.perf_counter()
                d2.close()
                td = time.perf_counter()
                parts[0] += tb - ta; parts[1] += tc - tb; parts[2] += td - tc
            torch.cuda.synchronize()
            dt = time.perf_counter() - t_begin
            return its, dt, [round(1e3 * x / n_e, 2) for x in parts]

        its, dte, parts = e2e_steps("bootstrap")
        e2e = {"value": its / dte, "unit": "iterations/s", "h2d_bytes_per_step": store_bytes, "d2h_bytes_per_step": 8 * n_txps * R,
               "ms_per_step": 1e3 * dte / n_e, "steps": n_e, "ms_create_compute_destroy": parts,
               "includes": f"per step: pinned-host store upload + validation + layout build + {R} bootstrap replicates + counts download + teardown"}
        its, dts, parts = e2e_steps("em")
        e2e_single = {"value": its / dts, "unit": "iterations/s", "h2d_bytes_per_step": store_bytes, "d2h_bytes_per_step": 8 * n_txps,
                      "ms_per_step": 1e3 * dts / n_e, "ms_create_em_destroy": parts,
                      "includes": "per step: pinned-host store upload + layout build + ONE EM to convergence (em_par rule) + counts download + teardown"}
    else:
        dte = elapsed + (bcast_ms + build_ms) * 1e-3
        e2e = {"value": total_iters / dte, "unit": "iterations/s", "h2d_bytes_per_step": store_bytes / K,
               "d2h_bytes_per_step": 8 * n_txps * R,
               "includes": "store upload on rank 0 + NCCL broadcast + per-rank layout build (once per job) + all K steps with counts download"}
        ds.close()

    # ---- CPU baseline (rank 0, N = 1 only): bounded samples of the same workload ------------------------------------
    cpu = cpu_par = None
    if rank == 0 and not multi and not args.no_cpu_baseline:
        from oracle import oracle
        threads = host_threads()
        ps = oracle.PortStore(s.row_ptr, s.txp_id, s.prob, s.n_txps)
        m = 24 if args.workload == "C3" else 120
        sw, em_s, wall = cpu_bootstrap_sample(ps, threads, m, SEED)
        cpu = {"value": sw / em_s, "unit": "iterations/s", "cores": threads, "kind": "port",
               "sample": CPU_SAMPLE_TEXT.format(T=threads, m=m, w=args.workload, setup=wall - em_s)}
        n_sw = 100 if args.workload == "C3" else 400
        t1 = time.perf_counter()
        _, _, _, sweeps = ps.em_par(max_iter=n_sw, conv_thresh=THR)
        dtp = time.perf_counter() - t1
        ps.close()
        cpu_par = {"value": (sweeps + 1) / dtp, "unit": "iterations/s", "cores": threads, "kind": "port",
                   "sample": f"{sweeps + 1} E+M sweeps of em_par (rayon-style parallel sweep, CAS f64 atomics) on the full {args.workload} store in {dtp:.1f} s; compare with em_single_gpu"}

    # ---- BASELINE config 2 sub-record (rank 0, N = 1) ---------------------------------------------------------------
    c2 = None
    if rank == 0 and not multi and args.workload == "C3" and not args.no_c2:
        s2 = gen_store("C2", pinned=True)
        with DeviceStore(s2.row_ptr, s2.txp_id, s2.prob, s2.n_txps, device=local_rank) as d2:
            o2 = np.empty(s2.n_txps)
            d2.em(max_iter=1000, conv_thresh=THR, min_iter=1, out=o2)
            torch.cuda.synchronize(); t1 = time.perf_counter(); it2 = 0
            for _ in range(5):
                r2 = d2.em(max_iter=1000, conv_thresh=THR, min_iter=1, out=o2)
                it2 += d2.counters()["sweeps"]
            dt2 = time.perf_counter() - t1
            p2 = torch.from_numpy(o2).to(dev).clamp_(min=1e-3); c2b = torch.zeros_like(p2)
            d2.sweep_timed(p2, c2b, 10)
            ms2 = d2.sweep_timed(p2, c2b, 200) / 200
            alg2 = algorithmic_bytes(s2.n_reads, s2.nnz, s2.n_txps)
            c2 = {"workload": WORKLOADS["C2"].split(";")[0], "em_iterations_per_sec": it2 / dt2, "niter": r2.niter, "ms_per_em": 1e3 * dt2 / 5,
                  "us_per_sweep": ms2 * 1e3, "roofline_frac": alg2 / (ms2 * 1e-3) / 1e9 / peak,
                  "note": "53 MB per sweep fits the 126 MB L2: back-to-back sweeps are L2-resident, the fraction is against the HBM peak all the same"}

    # ---- BASELINE config 5, scaled (single-cell mode): per-cell EMs batched, cells sharded over the ranks ---------------
    c5 = None
    if not args.no_c5:
        c5 = c5_block(args, rank, world, local_rank, barrier, max_over_ranks, sum_over_ranks)

    if rank == 0:
        line = {
            "metric": "em_iterations_per_sec", "value": value, "unit": "iterations/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": 1e3 * elapsed / K, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_block(args.workload, n_reads, nnz, n_txps),
            "job": {"step": f"{R} bootstrap replicates (weights + weighted EM to convergence, do_em rule min_iter 50, thr {THR}) + counts to the host",
                    "replicates": n_timed, "schedule": (args.boot_schedule if multi else "single rank"),
                    "us_per_iteration": 1e6 * elapsed * world / max(total_iters, 1),   # sustained, per GPU: sweep + bookkeeping + weights + copies, under whatever power cap the job runs into
                    "replicates_by_rank": reps_by_rank, "niter_rank0": stats["niter"][:32],
                    "l2": "inputs (0.72 GB per sweep) exceed the 126 MB L2; no flush needed",
                    "parallelism": f"{world} GPU(s); replicates sharded, store broadcast once over NCCL, no data-path collective",
                    "layout": layout},
            "replicates_per_sec": n_timed / elapsed, "iterations_per_replicate": total_iters / max(n_timed, 1),
            "roofline": roof, "roofline_plain_em": roof_plain, "em_single_gpu": em_single,
            "cpu_baseline": cpu, "cpu_baseline_em_par": cpu_par, "e2e": e2e, "e2e_single_em": e2e_single,
            "gpu_launches": int(total_launches), "c2": c2, "c5": c5, "clocks": clocks,
            "store_build_ms": build_ms, "bcast_ms": bcast_ms if multi else None,
        }
        emit(line)
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
    ap.add_argument("--boot-batch", type=int, default=5, help="bootstrap replicates per step (K steps = K * this many replicates in total)")
    ap.add_argument("--boot-schedule", default="dynamic", choices=["static", "dynamic"],
                    help="at N > 1: ranks pull replicate ids from a shared counter (default) or replicate g runs on rank g mod N")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c2", action="store_true")
    ap.add_argument("--no-c5", action="store_true")
    ap.add_argument("--c5-cells", type=int, default=256, help="cells of the scaled config-5 sub-record (in total, sharded over the ranks)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    claim_stdout()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
