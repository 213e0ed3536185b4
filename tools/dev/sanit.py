import sys, numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))))
from oarfish_b200 import synth, DeviceStore
s = synth.make_store(6000, 700, 6.0, 33)
ds = DeviceStore(s.row_ptr, s.txp_id, s.prob, s.n_txps)
r = ds.em(min_iter=1, max_iter=6)
out, nit = ds.bootstrap(1, 3, max_iter=4)
crp = np.array([0, 1000, 1000, 6000], dtype=np.uint64)
ds.em_batched(crp, max_iter=3)
ds.posteriors(r.counts, 0.01); ds.aux_counts()
print("ok", r.niter, r.counts.sum())
