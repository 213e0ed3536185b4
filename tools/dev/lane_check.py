"""Dev script (not a test): row-per-lane sweep against the CSR kernel on small and edge stores, then C3 timing.
Usage: python tools/dev/lane_check.py [check|time] [config] [ctas list]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oarfish_b200 import synth, DeviceStore  # noqa: E402


def csr(rows):
    rp = np.zeros(len(rows) + 1, dtype=np.uint64)
    rp[1:] = np.cumsum([len(r) for r in rows])
    tx = np.array([t for r in rows for t, _ in r], dtype=np.uint32)
    pr = np.array([p for r in rows for _, p in r], dtype=np.float32)
    return rp, tx, pr


def one_sweep(ds, M, prev_np, wts=None):
    prev = torch.from_numpy(prev_np).cuda()
    curr = torch.zeros(M, dtype=torch.float64, device="cuda")
    w = torch.from_numpy(wts.astype(np.int32)).cuda() if wts is not None else None
    ds.sweep(prev, curr, w)
    return curr.cpu().numpy()


def check(name, rp, tx, pr, M, aux=None):
    rng = np.random.default_rng(1)
    prev = rng.random(M) * 10 + 0.1
    N = len(rp) - 1
    wts = rng.integers(0, 4, size=N).astype(np.uint32)
    ds = DeviceStore(rp, tx, pr, M, aux=aux)
    info = ds.layout_info()
    out = {}
    for k in (1, info["kernel"]):
        ds.set_kernel(k)
        out[k] = (one_sweep(ds, M, prev), one_sweep(ds, M, prev, wts))
    ds.close()
    a, b = out[1], out[info["kernel"]]
    ok = True
    for which, x, y in (("plain", a[0], b[0]), ("weighted", a[1], b[1])):
        den = np.maximum(np.abs(x), 1e-300)
        err = np.abs(x - y) / den
        err[(x == 0) & (y == 0)] = 0
        m = float(err.max()) if len(err) else 0.0
        good = m < 1e-10
        ok &= good
        print(f"  {name:24s} {which:8s} kernel {info['kernel']} tiles {info['n_tiles']:6d} fb {info['fallback_rows']:4d} "
              f"max rel err {m:.2e} sum {y.sum():.6f}/{x.sum():.6f} {'OK' if good else 'FAIL'}", flush=True)
    return ok


def run_checks():
    ok = True
    cases = {
        "single_read": [[(0, 1.0)]],
        "all_unique": [[(i % 7, 0.5)] for i in range(300)],
        "all_distinct": [[(i, 0.5)] for i in range(5000)],
        "zero_prob_row": [[(0, 1.0)], [(1, 0.0), (2, 0.0)], [(2, 1.0), (1, 0.25)]],
        "duplicate_txp_in_row": [[(3, 0.5), (3, 0.25), (1, 1.0)]] * 40,
        "ragged": [[(j % 11, 1.0 / (1 + j)) for j in range(1 + (i * 7) % 23)] for i in range(400)],
        "long_rows": [[(j % 300, 0.9 ** (j % 17)) for j in range(n)] for n in (129, 500, 128, 127, 1, 300)] * 3,
        "rows_127": [[(j % 200, 0.9 ** (j % 17)) for j in range(127)] for _ in range(70)],
        "len_13_20": [[((i + j * 3) % 40, 1.0 / (1 + j)) for j in range(13 + i % 8)] for i in range(900)],
        "chunk_exact": [[(j % 16, 1.0) for j in range(16)]] * 64,
    }
    for name, rows in cases.items():
        rp, tx, pr = csr(rows)
        ok &= check(name, rp, tx, pr, int(tx.max()) + 3)
    for cfg in ("tiny", "small"):
        s = synth.make_config(cfg)
        ok &= check(cfg, s.row_ptr, s.txp_id, s.prob, s.n_txps)
        aux = np.random.default_rng(5).random(s.nnz) + 0.5
        ok &= check(cfg + "+aux", s.row_ptr, s.txp_id, s.prob, s.n_txps, aux=aux)
    print("CHECKS", "PASSED" if ok else "FAILED", flush=True)
    return ok


def run_timing(cfg, ctas):
    s = synth.make_config(cfg)
    M = s.n_txps
    bytes_alg = 8 * s.nnz + 4 * (s.n_reads + 1) + 24 * M
    ref = None
    for cps in ctas:
        os.environ["OAR_CTAS_PER_SM"] = str(cps)
        t = time.time()
        ds = DeviceStore(s.row_ptr, s.txp_id, s.prob, M)
        info = ds.layout_info()
        build = time.time() - t
        prev = torch.full((M,), s.n_reads / M, dtype=torch.float64, device="cuda")
        curr = torch.zeros(M, dtype=torch.float64, device="cuda")
        r = ds.em(min_iter=1, max_iter=30)
        prev.copy_(torch.from_numpy(r.counts))
        ds.sweep_timed(prev, curr, 5)
        ms = ds.sweep_timed(prev, curr, 30) / 30
        ds.sweep(prev, curr)
        c = curr.cpu().numpy()
        if ref is None:
            ds.set_kernel(1); ds.sweep(prev, curr); ref = curr.cpu().numpy(); ds.set_kernel(info["kernel"])
        err = (np.abs(c - ref) / np.maximum(ref, 1e-300))[ref > 1e-6].max()
        w = torch.ones(s.n_reads, dtype=torch.int32, device="cuda")
        msw = ds.sweep_timed(prev, curr, 20, w) / 20
        t = time.time(); r = ds.em(min_iter=1); wall = time.time() - t
        print(f"{cfg} kernel {info['kernel']} span {info['span']} tiles {info['n_tiles']} slots {info['slots']} ctas/SM={cps}: "
              f"{ms*1e3:.1f} us/sweep frac {bytes_alg/ms/1e6/6533.2:.3f} (weighted {msw*1e3:.1f} us) relerr {err:.2e} | "
              f"EM niter {r.niter} {ds.timings_ms()['em']:.1f} ms -> {ds.counters()['sweeps']/wall:.0f} it/s | create {build*1e3:.0f} ms "
              f"sumD {info['sum_distinct']} sumU {info['sum_units']}", flush=True)
        ds.close()


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "check"
    if mode == "check":
        sys.exit(0 if run_checks() else 1)
    cfg = sys.argv[2] if len(sys.argv) > 2 else "C3"
    ctas = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "3").split(",")]
    run_timing(cfg, ctas)
