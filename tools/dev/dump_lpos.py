"""Dev helper (GPU box): layout words (+ trash offsets) of the first tiles of a config and of a slice from its middle, for the offline
bank analysis of tools/layout_model.py.   python tools/dev/dump_lpos.py <config> [tiles per slice]  -> gpurun_out/lpos_<config>.npy / .npz"""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oarfish_b200 import synth, DeviceStore
cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
s = synth.make_config(cfg)
ds = DeviceStore(s.row_ptr, s.txp_id, s.prob, s.n_txps)
li = ds.layout_info()
n = min(int(sys.argv[2]) if len(sys.argv) > 2 else 1500, li["n_tiles"] // 2)
a, ta = ds.layout_lpos(0, n, with_trash=True); b, tb = ds.layout_lpos(li["n_tiles"] // 2, n, with_trash=True)
np.save(f"gpurun_out/lpos_{cfg}.npy", np.concatenate([a, b]))
np.savez_compressed(f"gpurun_out/lpos_{cfg}.npz", words=np.concatenate([a, b]), trash=np.concatenate([ta, tb]))
print(li)
