import sys, os, numpy as np, torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))))
from oarfish_b200 import synth, DeviceStore
s = synth.make_config("C3"); M = s.n_txps
ds = DeviceStore(s.row_ptr, s.txp_id, s.prob, M)
prev = torch.full((M,), s.n_reads / M, dtype=torch.float64, device="cuda")
curr = torch.zeros(M, dtype=torch.float64, device="cuda")
for ab in (0, 64, 128, 1|4|32, 1|4|32|64, 1|4|32|64|128):
    os.environ["OAR_ABLATE"] = str(ab)
    ds.sweep_timed(prev, curr, 3)
    print("ablate", ab, "us", round(ds.sweep_timed(prev, curr, 20) / 20 * 1e3, 1), flush=True)
