import sys, time, numpy as np
sys.path.insert(0, "/root/repo")
from oarfish_b200 import synth, DeviceStore
from oracle import oracle
s = synth.make_config("small")
ds = DeviceStore(s.row_ptr, s.txp_id, s.prob, s.n_txps)
for mi in (50, 1):
    r = ds.em(min_iter=mi)
    c, niter, rel, sw = oracle.do_em(s.row_ptr, s.txp_id, s.prob, s.n_txps, min_iter=mi)
    m = c > 1e-8
    print("min_iter", mi, "gpu niter", r.niter, "oracle", niter, "rel", r.rel_diff, rel, "maxrel", np.abs(r.counts[m]-c[m]).max() if m.any() else 0, (np.abs(r.counts-c)/np.maximum(c,1e-300))[m].max(), ds.timings_ms(), ds.counters())
w = ds.sample_weights(7, 3)
print("weights sum", w.sum(), w.max(), (w==0).mean())
out, nit = ds.bootstrap_weights(w)
c, niter, rel, sw = oracle.do_em(s.row_ptr, s.txp_id, s.prob, s.n_txps, min_iter=50, wts=w)
m = c > 1e-8
print("boot niter", nit, niter, (np.abs(out[0]-c)/np.maximum(c,1e-300))[m].max())
out2, nit2 = ds.bootstrap(4, 7)
print("boot seeded niter", nit2, np.abs(out2[3]-out[0]).max())
