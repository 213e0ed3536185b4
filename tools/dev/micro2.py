import sys, time, os, numpy as np, torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))))
from oarfish_b200 import synth, DeviceStore
cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
s = synth.make_config(cfg); M = s.n_txps
bytes_alg = 8*s.nnz + 4*(s.n_reads+1) + 24*M
ref = None
for cps in [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "4").split(",")]:
    os.environ["OAR_CTAS_PER_SM"] = str(cps)
    ds = DeviceStore(s.row_ptr, s.txp_id, s.prob, M)
    prev = torch.full((M,), s.n_reads / M, dtype=torch.float64, device="cuda")
    curr = torch.zeros(M, dtype=torch.float64, device="cuda")
    r = ds.em(min_iter=1, max_iter=30); prev.copy_(torch.from_numpy(r.counts))
    ds.sweep_timed(prev, curr, 5)
    ms = ds.sweep_timed(prev, curr, 30) / 30
    wts = torch.from_numpy(ds.sample_weights(7, 0).astype(np.int32)).cuda()
    ds.sweep_timed(prev, curr, 3, wts)
    msw = ds.sweep_timed(prev, curr, 30, wts) / 30
    ds.sweep(prev, curr); c = curr.cpu().numpy()
    if ref is None:
        ds.set_kernel(1); ds.sweep(prev, curr); ref = curr.cpu().numpy(); ds.set_kernel(2)
    err = (np.abs(c-ref)/np.maximum(ref,1e-300))[ref>1e-6].max()
    t=time.time(); r = ds.em(min_iter=1); wall=time.time()-t
    li = ds.layout_info()
    print(f"fallback {li.get('fallback_rows')} tiles {li.get('n_tiles')}", end=" | ")
    print(f"ctas/SM={cps}: {ms*1e3:.1f} us/sweep (weighted {msw*1e3:.1f}) frac {bytes_alg/ms/1e6/6533.2:.3f} relerr {err:.2e} | EM niter {r.niter} {ds.timings_ms()['em']:.1f} ms -> {ds.counters()['sweeps']/wall:.0f} it/s", flush=True)
    ds.close()
