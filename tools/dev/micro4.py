import sys, time, os, numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))))
from oarfish_b200 import synth, DeviceStore
s = synth.make_config("C3"); M = s.n_txps
for pair in ("1", "0"):
    os.environ["OAR_BOOT_PAIR"] = pair
    ds = DeviceStore(s.row_ptr, s.txp_id, s.prob, M)
    ds.bootstrap(2, 4, max_iter=20)
    t = time.time(); out, nit = ds.bootstrap(2, 4, max_iter=200); dt = time.time() - t
    print("pair", pair, "2 replicates x 201 sweeps:", round(dt*1e3,1), "ms ->", round(dt*1e6/402,1), "us per replicate-sweep", ds.timings_ms()["em"], flush=True)
    ds.close()
