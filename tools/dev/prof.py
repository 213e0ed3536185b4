import sys, numpy as np, torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))))
from oarfish_b200 import synth, DeviceStore
s = synth.make_config(sys.argv[1] if len(sys.argv) > 1 else "C3")
M = s.n_txps
ds = DeviceStore(s.row_ptr, s.txp_id, s.prob, M)
prev = torch.full((M,), s.n_reads / M, dtype=torch.float64, device="cuda")
curr = torch.zeros(M, dtype=torch.float64, device="cuda")
r = ds.em(min_iter=1, max_iter=30)
prev.copy_(torch.from_numpy(r.counts))
for _ in range(6): ds.sweep(prev, curr, sync=True)
torch.cuda.synchronize(); torch.cuda.profiler.start()   # ncu --profile-from-start off: only the steady-state sweeps below
print(ds.sweep_timed(prev, curr, 20) / 20 * 1e3, "us")
