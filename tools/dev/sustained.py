"""Dev helper (GPU box): a SUSTAINED bootstrap job (seconds of back-to-back sweeps, as bench.py's timed region) for one library
variant / mode, with the SM clock and board power sampled during it.  Short A/B runs (tools/dev/ab.py) finish before the power
cap bites; the bench's long job does not.   usage: python tools/dev/sustained.py <label> [replicates] [max_iter]"""
import os, subprocess, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oarfish_b200 import DeviceStore, synth
label = sys.argv[1]; R = int(sys.argv[2]) if len(sys.argv) > 2 else 12; mi = int(sys.argv[3]) if len(sys.argv) > 3 else 400
s = synth.make_config("C3", pinned=True)
ds = DeviceStore(s.row_ptr, s.txp_id, s.prob, s.n_txps)
ds.bootstrap(1, 3, max_iter=60)   # warm: graphs, pools
samples = []; stop = False
def sampler():
    while not stop:
        try:
            o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True, timeout=5).stdout
            c, p = o.strip().split(","); samples.append((float(c), float(p)))
        except Exception:
            pass
        time.sleep(0.1)
th = threading.Thread(target=sampler); th.start()
out = np.empty((R, s.n_txps))
t0 = time.perf_counter(); _, niter = ds.bootstrap(R, 11, max_iter=mi, out=out); dt = time.perf_counter() - t0
stop = True; th.join()
sweeps = int((niter + 1).sum())
mid = samples[len(samples) // 4:] or [(0, 0)]
print(f"{label}: {R} replicates x {mi} iterations in {dt:.2f} s = {sweeps / dt:.0f} it/s ({dt / sweeps * 1e6:.1f} us per iteration); "
      f"SM clock median {np.median([c for c, _ in mid]):.0f} MHz, power median {np.median([p for _, p in mid]):.0f} W ({len(samples)} samples)", flush=True)
ds.close()
