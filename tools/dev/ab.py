"""Dev helper (GPU box): A/B timing of sweep variants in ONE process (the synthetic store is generated once).

usage: python tools/dev/ab.py <config> "<label>:<ctas/SM>[:<OAR_TILE_SPAN>]" ...
The library is the product one unless OAR_EM_LIB points at a variant (tools/dev/build_variants.sh).
Per variant: mean of 30 back-to-back sweeps (CUDA events on the store's stream), plain and bootstrap-weighted,
max relative error of one sweep against the CSR row-group kernel, and a complete EM (em_par rule)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oarfish_b200 import DeviceStore, synth  # noqa: E402

cfg = sys.argv[1]
s = synth.make_config(cfg, pinned=True)
M = s.n_txps
alg = 8 * s.nnz + 4 * (s.n_reads + 1) + 24 * M
peak = 6551.0
ref = None
for spec in sys.argv[2:]:
    f = spec.split(":")
    os.environ["OAR_CTAS_PER_SM"] = f[1]
    if len(f) > 2 and f[2]:
        os.environ["OAR_TILE_SPAN"] = f[2]
    else:
        os.environ.pop("OAR_TILE_SPAN", None)
    t0 = time.perf_counter()
    ds = DeviceStore(s.row_ptr, s.txp_id, s.prob, M)
    t_create = time.perf_counter() - t0
    prev = torch.full((M,), s.n_reads / M, dtype=torch.float64, device="cuda")
    curr = torch.zeros(M, dtype=torch.float64, device="cuda")
    r = ds.em(min_iter=1, max_iter=30)
    prev.copy_(torch.from_numpy(r.counts))
    ds.sweep_timed(prev, curr, 5)
    ms = ds.sweep_timed(prev, curr, 30) / 30
    wts = torch.from_numpy(ds.sample_weights(7, 0).astype(np.int32)).cuda()
    ds.sweep_timed(prev, curr, 3, wts)
    msw = ds.sweep_timed(prev, curr, 30, wts) / 30
    ds.sweep(prev, curr)
    c = curr.cpu().numpy()
    if ref is None:
        ds.set_kernel(1); ds.sweep(prev, curr); ref = curr.cpu().numpy(); ds.set_kernel(2)
    err = (np.abs(c - ref) / np.maximum(ref, 1e-300))[ref > 1e-6].max()
    t = time.perf_counter(); r = ds.em(min_iter=1); wall = time.perf_counter() - t
    li = ds.layout_info()
    sweeps = ds.counters()["sweeps"]
    ds.bootstrap(1, 5)
    t = time.perf_counter(); ds.bootstrap(2, 5); wallb = time.perf_counter() - t
    sweepsb = ds.counters()["sweeps"]
    t = time.perf_counter(); ds.close(); t_close = time.perf_counter() - t
    print(f"{cfg} {os.path.basename(os.environ.get('OAR_EM_LIB', 'product'))} {f[0]} ctas/SM={f[1]} span={li['span']} fb={li['fallback_rows']}: {ms*1e3:.1f} us/sweep (weighted {msw*1e3:.1f}) "
          f"frac {alg/ms/1e6/peak:.3f} relerr {err:.1e} | EM niter {r.niter}: {sweeps/wall:.0f} it/s, 2 replicates {sweepsb/wallb:.0f} it/s | create {t_create*1e3:.0f} ms close {t_close*1e3:.1f} ms",
          flush=True)
