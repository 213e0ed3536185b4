"""Dev helper (GPU box): the raw sweep timed as plain stream launches / with the EM's early-exit test / as graph nodes (OAR_TIMED_MODE)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oarfish_b200 import synth, DeviceStore
s = synth.make_config("C3", pinned=True); M = s.n_txps
ds = DeviceStore(s.row_ptr, s.txp_id, s.prob, M)
prev = torch.full((M,), s.n_reads / M, dtype=torch.float64, device="cuda"); curr = torch.zeros(M, dtype=torch.float64, device="cuda")
r = ds.em(min_iter=1, max_iter=30); prev.copy_(torch.from_numpy(r.counts))
wts = torch.from_numpy(ds.sample_weights(7, 0).astype(np.int32)).cuda()
ds.sweep_timed(prev, curr, 20); ds.sweep_timed(prev, curr, 20, wts)
for reps in (36, 360):
    print(os.environ.get("OAR_TIMED_MODE", "stream"), reps, "sweeps: plain %.1f us, weighted %.1f us" % (ds.sweep_timed(prev, curr, reps) / reps * 1e3, ds.sweep_timed(prev, curr, reps, wts) / reps * 1e3), flush=True)
