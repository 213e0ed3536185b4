import sys, time, numpy as np
sys.path.insert(0, "/root/repo")
from oarfish_b200 import synth, DeviceStore
s = synth.make_config("C3", pinned=True)
for i in range(3):
    t = time.time(); ds = DeviceStore(s.row_ptr, s.txp_id, s.prob, s.n_txps); dt = time.time() - t
    print("create wall ms", dt * 1e3, ds.timings_ms()["upload"]); ds.close()
