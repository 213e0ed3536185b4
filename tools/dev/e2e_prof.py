"""Dev script: where an end-to-end step (create from pinned host buffers -> EM -> counts on host -> destroy) spends its time."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oarfish_b200 import synth, DeviceStore
s = synth.make_config(sys.argv[1] if len(sys.argv) > 1 else "C3", pinned=True)
out = np.empty(s.n_txps)
for i in range(4):
    t0 = time.perf_counter(); ds = DeviceStore(s.row_ptr, s.txp_id, s.prob, s.n_txps)
    t1 = time.perf_counter(); r = ds.em(min_iter=1, out=out)
    t2 = time.perf_counter(); r = ds.em(min_iter=1, out=out)
    t3 = time.perf_counter(); ds.close()
    t4 = time.perf_counter()
    print(f"create {1e3*(t1-t0):.1f} ms | first em {1e3*(t2-t1):.1f} ms | second em {1e3*(t3-t2):.1f} ms (device {ds_t if False else 0}) | close {1e3*(t4-t3):.1f} ms | niter {r.niter}", flush=True)
