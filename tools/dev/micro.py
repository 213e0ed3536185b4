import sys, time, numpy as np, torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))))
from oarfish_b200 import synth, DeviceStore
cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
t=time.time(); s = synth.make_config(cfg); print("gen", time.time()-t, s.n_reads, s.nnz, flush=True)
M = s.n_txps
def timeit(ds, label, reps=20):
    st = torch.cuda.ExternalStream(ds.stream)
    prev = torch.full((M,), s.n_reads / M, dtype=torch.float64, device="cuda")
    curr = torch.zeros(M, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    for _ in range(3): ds.sweep(prev, curr, sync=True)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record(st)
        for _ in range(reps): ds.sweep(prev, curr, sync=False)
        e1.record(st)
    e1.synchronize()
    ms = e0.elapsed_time(e1) / reps
    bytes_alg = 8*s.nnz + 4*(s.n_reads+1) + 24*M
    print(f"{label}: {ms*1e3:.1f} us/sweep  {bytes_alg/ms/1e6:.1f} GB/s alg  frac {bytes_alg/ms/1e6/6533.2:.3f}", flush=True)
    return curr.cpu().numpy()
t=time.time(); ds = DeviceStore(s.row_ptr, s.txp_id, s.prob, M); print("create wall", time.time()-t, ds.timings_ms(), ds.layout_info(), flush=True)
c1 = timeit(ds, "tiled")
ds.set_kernel(1)
c0 = timeit(ds, "rowgroup")
print("maxdiff tiled vs rowgroup", np.abs(c0-c1).max(), (np.abs(c0-c1)/np.maximum(c0,1e-300))[c0>1e-6].max(), c0.sum(), c1.sum())
ds.set_kernel(2)
for mi in (1, 50):
    t=time.time(); r = ds.em(min_iter=mi); print("EM tiled min_iter", mi, r.niter, r.rel_diff, ds.timings_ms(), ds.counters(), "wall", time.time()-t, flush=True)
t=time.time(); out, nit = ds.bootstrap(4, 4); print("boot 4 reps", nit, ds.timings_ms(), "wall", time.time()-t)
