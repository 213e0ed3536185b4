"""Dev helper (GPU box): layout words of the first tiles of a config + a slice from the middle, for offline bank analysis."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oarfish_b200 import synth, DeviceStore
cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
s = synth.make_config(cfg)
ds = DeviceStore(s.row_ptr, s.txp_id, s.prob, s.n_txps)
li = ds.layout_info()
n = min(1500, li["n_tiles"] // 2)
a = ds.layout_lpos(0, n); b = ds.layout_lpos(li["n_tiles"] // 2, n)
np.save(f"gpurun_out/lpos_{cfg}.npy", np.concatenate([a, b]))
print(li)
