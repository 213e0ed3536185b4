"""Dev script for ncu: a handful of raw sweeps on a config (no EM driver, so every matching launch is a full sweep)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oarfish_b200 import synth, DeviceStore
s = synth.make_config(sys.argv[1] if len(sys.argv) > 1 else "C3")
M = s.n_txps
ds = DeviceStore(s.row_ptr, s.txp_id, s.prob, M)
rng = np.random.default_rng(0)
prev = torch.from_numpy(rng.random(M) * 2 * s.n_reads / M + 1e-3).cuda()
curr = torch.zeros(M, dtype=torch.float64, device="cuda")
for _ in range(6):
    ds.sweep(prev, curr, sync=True)
print(ds.sweep_timed(prev, curr, 20) / 20 * 1e3, "us")
