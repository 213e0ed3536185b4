#!/usr/bin/env python
"""Robustness of the sweep to store shape: us/sweep and ns per alignment on
   C3            the headline store (gene-contiguous transcript ids, rows <= 100 alignments)
   C3-permuted   the same store with transcript ids shuffled (isoforms of a gene are no longer id-neighbours)
   C3-longrows   5 % of the reads replaced by reads with 128-400 alignments (beyond --best-n's default of 100)
each checked against the CSR kernel on the same inputs.   python tools/bench_robust.py [C3|C2|small]  -> JSON lines"""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oarfish_b200 import DeviceStore, synth

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
base = synth.make_config(cfg)
variants = [(cfg, base), (cfg + "-permuted", synth.permute_ids(base, 99)), (cfg + "-longrows", synth.with_long_rows(base, 0.05, 128, 400, 98))]
for name, s in variants:
    M = s.n_txps
    t0 = time.perf_counter()
    ds = DeviceStore(s.row_ptr, s.txp_id, s.prob, M)
    create_ms = (time.perf_counter() - t0) * 1e3
    prev = torch.full((M,), s.n_reads / M, dtype=torch.float64, device="cuda")
    curr = torch.zeros(M, dtype=torch.float64, device="cuda")
    r = ds.em(min_iter=1, max_iter=30)
    prev.copy_(torch.from_numpy(r.counts))
    ds.sweep_timed(prev, curr, 3)
    us = ds.sweep_timed(prev, curr, 20) / 20 * 1e3
    ds.sweep(prev, curr); c = curr.cpu().numpy()
    ds.set_kernel(1); ds.sweep(prev, curr); ref = curr.cpu().numpy()
    us_csr = ds.sweep_timed(prev, curr, 3) / 3 * 1e3
    ds.set_kernel(2)
    err = float((np.abs(c - ref) / np.maximum(ref, 1e-300))[ref > 1e-6].max())
    li = ds.layout_info()
    ds.close()
    print(json.dumps({"workload": name, "n_reads": s.n_reads, "nnz": s.nnz, "us_per_sweep": us, "ns_per_alignment": us * 1e3 / s.nnz,
                      "us_per_sweep_csr_kernel": us_csr, "max_rel_err_vs_csr_kernel": err, "create_ms": create_ms, "tiles": li["n_tiles"],
                      "fallback_rows": li["fallback_rows"], "distinct_per_tile": li["sum_distinct"] / max(li["n_tiles"], 1),
                      "items_per_tile": li["sum_units"] / max(li["n_tiles"], 1)}), flush=True)
