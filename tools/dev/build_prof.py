"""Dev script: wall time of DeviceStore creation from pinned host buffers (upload + validation + layout build)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oarfish_b200 import synth, DeviceStore
s = synth.make_config(sys.argv[1] if len(sys.argv) > 1 else "C3", pinned=True)
for i in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
    t = time.time(); ds = DeviceStore(s.row_ptr, s.txp_id, s.prob, s.n_txps); dt = time.time() - t
    t2 = time.time(); ds.close(); dc = time.time() - t2
    print(f"create wall {dt*1e3:.1f} ms (stream events upload+layout {ds.timings_ms()['upload'] if False else 0}) close {dc*1e3:.1f} ms", flush=True)
