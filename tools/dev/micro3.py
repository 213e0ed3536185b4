import sys, time, os, numpy as np, torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))))
from oarfish_b200 import synth, DeviceStore
s = synth.make_config("C3"); M = s.n_txps
ds = DeviceStore(s.row_ptr, s.txp_id, s.prob, M)
prev = torch.full((M,), s.n_reads / M, dtype=torch.float64, device="cuda")
curr = torch.zeros(M, dtype=torch.float64, device="cuda")
w = torch.from_numpy(ds.sample_weights(4, 0).view(np.int32)).cuda()
ds.sweep_timed(prev, curr, 5)
print("unweighted us", ds.sweep_timed(prev, curr, 30) / 30 * 1e3)
ds.sweep_timed(prev, curr, 5, weights_dev=w)
print("weighted us", ds.sweep_timed(prev, curr, 30, weights_dev=w) / 30 * 1e3)
t = time.time(); out, nit = ds.bootstrap(4, 4); dt = time.time() - t
print("4 replicates", nit, dt, "s ->", 4 / dt, "rep/s;", (nit.sum() + 8) / dt, "it/s")
