"""Dev helper (GPU box): GPU time per EM iteration (CUDA events around the EM loop) for a fixed number of iterations."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oarfish_b200 import synth, DeviceStore
s = synth.make_config("C3", pinned=True)
ds = DeviceStore(s.row_ptr, s.txp_id, s.prob, s.n_txps)
ds.em(min_iter=1, max_iter=40)
for mi in (90, 396):
    r = ds.em(min_iter=1000, max_iter=mi, conv_thresh=0.0); tm = ds.timings_ms()
    print(sys.argv[1], f"{r.niter} iterations: {tm['em'] * 1e3 / (r.niter + 1):.1f} us per iteration (EM loop, events)", flush=True)
