#!/usr/bin/env python
"""Config 5 (scaled): batched per-cell EM, cells/s on one GPU.
   python tools/bench_cells.py [n_cells] [reads_per_cell] [n_txps] [expressed_per_cell]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oarfish_b200 import DeviceStore, synth

n_cells = int(sys.argv[1]) if len(sys.argv) > 1 else 296
reads = int(sys.argv[2]) if len(sys.argv) > 2 else 50_000
M = int(sys.argv[3]) if len(sys.argv) > 3 else 200_000
t0 = time.time()
expressed = int(sys.argv[4]) if len(sys.argv) > 4 else 5000      # about 5 k expressed transcripts per cell (0: all)
s, crp = synth.make_cells([reads] * n_cells, M, 6.0, seed=5, expressed=expressed or None)
gen_s = time.time() - t0
ds = DeviceStore(s.row_ptr, s.txp_id, s.prob, M)
ds.em_batched(crp[:3].copy() if False else crp)  # warm-up (allocations, module load)
t0 = time.time()
cell_ptr, txp, val, niter = ds.em_batched(crp)
dt = time.time() - t0
tm = ds.timings_ms()
line = {"metric": "cells_per_sec", "value": n_cells / dt, "unit": "cells/s", "n_cells": n_cells, "reads_per_cell": reads,
        "n_txps": M, "nnz": int(s.nnz), "em_ms": tm["em"], "download_ms": tm["download"], "wall_s": dt,
        "niter_mean": float(niter.mean()), "niter_max": int(niter.max()),
        "alignment_updates_per_sec": float(s.nnz / n_cells * (niter + 2).sum() / (tm["em"] * 1e-3)), "gen_s": gen_s}
# (per-cell parity against the oracle: tests/test_gpu_parity.py::test_batched_cells_match_per_cell_oracle)
print(json.dumps(line))
