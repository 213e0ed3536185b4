"""ctypes binding of the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.  Parity is UNPINNED by the
reference (it ships no EM tests); see oracle/em_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboarfish_oracle.so")
_lib = None
_vp = C.c_void_p


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not found: run `make oracle`")
        L = C.CDLL(LIB_PATH)
        L.oracle_m_step.restype = None
        L.oracle_m_step.argtypes = [_vp, _vp, _vp, _vp, _vp, C.c_uint64, _vp, _vp, _vp]
        L.oracle_do_em.restype = C.c_uint32
        L.oracle_do_em.argtypes = [_vp, _vp, _vp, _vp, C.c_uint64, C.c_uint32, _vp, C.c_uint32, C.c_double,
                                   C.c_uint32, _vp, C.c_uint64, _vp, _vp, _vp, _vp]
        L.oracle_get_sample_inds.restype = None
        L.oracle_get_sample_inds.argtypes = [C.c_uint64, C.c_uint64, _vp]
        L.oracle_inds_to_weights.restype = None
        L.oracle_inds_to_weights.argtypes = [_vp, C.c_uint64, C.c_uint64, _vp]
        L.oracle_posteriors.restype = None
        L.oracle_posteriors.argtypes = [_vp, _vp, _vp, _vp, C.c_uint64, _vp, C.c_double, _vp, _vp]
        L.oracle_aux_counts.restype = None
        L.oracle_aux_counts.argtypes = [_vp, _vp, C.c_uint64, C.c_uint32, _vp, _vp]
        L.oracle_coverage_model_binomial.restype = None
        L.oracle_coverage_model_binomial.argtypes = [_vp, _vp, _vp, _vp, C.c_uint64, C.c_uint64, _vp, C.c_uint32, C.c_uint32, _vp]
        L.oracle_statrs_ln_gamma.restype = C.c_double
        L.oracle_statrs_ln_gamma.argtypes = [C.c_double]
        L.oracle_coverage_model.restype = None
        L.oracle_coverage_model.argtypes = [_vp, _vp, _vp, _vp, C.c_uint64, C.c_uint64, _vp, C.c_uint32, C.c_uint32, C.c_double, _vp]
        L.port_num_threads.restype = C.c_int
        L.port_set_num_threads.restype = None
        L.port_set_num_threads.argtypes = [C.c_int]
        L.port_store_create.restype = _vp
        L.port_store_create.argtypes = [_vp, _vp, _vp, _vp, C.c_uint64, C.c_uint64, C.c_uint32]
        L.port_store_destroy.restype = None
        L.port_store_destroy.argtypes = [_vp]
        L.port_em_par.restype = C.c_uint32
        L.port_em_par.argtypes = [_vp, C.c_int, _vp, C.c_uint32, C.c_double, _vp, _vp, _vp]
        L.port_bootstrap.restype = None
        L.port_bootstrap.argtypes = [_vp, C.c_int, C.c_uint32, C.c_uint64, C.c_uint32, C.c_double, C.c_int, _vp, _vp]
        L.port_bootstrap_timed.restype = None
        L.port_bootstrap_timed.argtypes = [_vp, C.c_int, C.c_uint32, C.c_uint64, C.c_uint32, C.c_double, C.c_int, _vp, _vp, _vp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data


def _c(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


def m_step(row_ptr, txp, prob, prev, cov=None, inds=None, wts=None):
    """One sweep (em.rs:87-133); returns curr."""
    row_ptr = _c(row_ptr, np.uint64); txp = _c(txp, np.uint32); prob = _c(prob, np.float32)
    cov = _c(cov, np.float64); inds = _c(inds, np.uint64); wts = _c(wts, np.uint32)
    prev = _c(prev, np.float64)
    curr = np.zeros_like(prev)
    n_rows = len(inds) if inds is not None else len(row_ptr) - 1
    lib().oracle_m_step(_p(row_ptr), _p(txp), _p(prob), _p(cov), _p(inds), n_rows, _p(wts), _p(prev), _p(curr))
    return curr


def do_em(row_ptr, txp, prob, n_txps, max_iter=1000, conv_thresh=1e-3, min_iter=50, cov=None, init=None,
          inds=None, wts=None):
    """do_em (em.rs:144-255) / em_par stop rule with min_iter=1.  Returns (counts, niter, rel_diff, sweeps)."""
    row_ptr = _c(row_ptr, np.uint64); txp = _c(txp, np.uint32); prob = _c(prob, np.float32)
    cov = _c(cov, np.float64); inds = _c(inds, np.uint64); wts = _c(wts, np.uint32); init = _c(init, np.float64)
    out = np.zeros(n_txps, dtype=np.float64)
    niter = C.c_uint32(0); rel = C.c_double(0.0)
    sweeps = lib().oracle_do_em(_p(row_ptr), _p(txp), _p(prob), _p(cov), len(row_ptr) - 1, n_txps, _p(init),
                                max_iter, conv_thresh, min_iter, _p(inds), 0 if inds is None else len(inds),
                                _p(wts), _p(out), C.byref(niter), C.byref(rel))
    return out, int(niter.value), float(rel.value), int(sweeps)


def get_sample_inds(n, seed):
    """bootstrap::get_sample_inds (bootstrap.rs:7-16) with the oracle's seeded generator."""
    out = np.empty(n, dtype=np.uint64)
    lib().oracle_get_sample_inds(n, seed, _p(out))
    return out


def inds_to_weights(inds, n_rows):
    inds = _c(inds, np.uint64)
    w = np.zeros(n_rows, dtype=np.uint32)
    lib().oracle_inds_to_weights(_p(inds), len(inds), n_rows, _p(w))
    return w


def posteriors(row_ptr, txp, prob, counts, display_thresh=0.0, cov=None):
    """write_out_prob inner loop (write_function.rs:283-332)."""
    row_ptr = _c(row_ptr, np.uint64); txp = _c(txp, np.uint32); prob = _c(prob, np.float32); cov = _c(cov, np.float64)
    counts = _c(counts, np.float64)
    out = np.zeros(len(txp), dtype=np.float64); kept = np.zeros(len(row_ptr) - 1, dtype=np.uint32)
    lib().oracle_posteriors(_p(row_ptr), _p(txp), _p(prob), _p(cov), len(row_ptr) - 1, _p(counts), display_thresh, _p(out), _p(kept))
    return out, kept


def aux_counts(row_ptr, txp, n_txps):
    """get_aux_counts (aux_counts.rs:23-50)."""
    row_ptr = _c(row_ptr, np.uint64); txp = _c(txp, np.uint32)
    u = np.zeros(n_txps, dtype=np.uint32); t = np.zeros(n_txps, dtype=np.uint32)
    lib().oracle_aux_counts(_p(row_ptr), _p(txp), len(row_ptr) - 1, n_txps, _p(u), _p(t))
    return u, t


def coverage_model(row_ptr, txp, start, end, txp_len, bin_width=100, growth_rate=2.0):
    """--model-coverage stage (bulk.rs:103-108) -> coverage_probabilities f64[nnz]."""
    row_ptr = _c(row_ptr, np.uint64); txp = _c(txp, np.uint32); start = _c(start, np.uint32); end = _c(end, np.uint32)
    txp_len = _c(txp_len, np.uint32)
    out = np.zeros(len(txp), dtype=np.float64)
    lib().oracle_coverage_model(_p(row_ptr), _p(txp), _p(start), _p(end), len(row_ptr) - 1, len(txp), _p(txp_len), len(txp_len),
                                bin_width, growth_rate, _p(out))
    return out


def coverage_model_binomial(row_ptr, txp, start, end, txp_len, bin_width=100):
    """The single-cell driver's coverage stage (single_cell.rs:132-137, binomial_probability.rs) -> f64[nnz]."""
    row_ptr = _c(row_ptr, np.uint64); txp = _c(txp, np.uint32); start = _c(start, np.uint32); end = _c(end, np.uint32)
    txp_len = _c(txp_len, np.uint32)
    out = np.zeros(len(txp), dtype=np.float64)
    lib().oracle_coverage_model_binomial(_p(row_ptr), _p(txp), _p(start), _p(end), len(row_ptr) - 1, len(txp), _p(txp_len),
                                         len(txp_len), bin_width, _p(out))
    return out


def statrs_ln_gamma(x):
    return float(lib().oracle_statrs_ln_gamma(float(x)))


class _FilterOpts(C.Structure):
    _fields_ = [("which_strand", C.c_int32), ("min_aligned_len", C.c_uint32), ("three_prime_clip", C.c_int64),
                ("five_prime_clip", C.c_uint32), ("min_aligned_fraction", C.c_float), ("score_threshold", C.c_float),
                ("score_prob_denom", C.c_float)]


DISCARD_NAMES = ("discard_5p", "discard_3p", "discard_score", "discard_aln_frac", "discard_aln_len", "discard_ori", "discard_supp",
                 "no_mapping", "no_valid_aln", "valid_best_aln")


def filter_records(group_ptr, ref_id, aln_start, aln_end, aln_span, score, flags, seq_len, txp_len, *, which_strand=0,
                   min_aligned_len=50, three_prime_clip=2**31 - 1, five_prime_clip=2**32 - 1, min_aligned_fraction=0.5,
                   score_threshold=0.95, score_prob_denom=5.0):
    """AlignmentFilters::filter + add_filtered_group (oarfish_types.rs:955-1130, :718-738):
    -> (row_ptr u64, txp u32, prob f32, src u32, group u32, discard dict)."""
    gp = _c(group_ptr, np.uint64)
    cols = [_c(a, np.uint32) for a in (ref_id, aln_start, aln_end, aln_span)]
    sc = _c(score, np.int32); fl = _c(flags, np.uint8); sl = _c(seq_len, np.uint32); tl = _c(txp_len, np.uint32)
    G, R = len(gp) - 1, len(sc)
    rp = np.zeros(G + 1, dtype=np.uint64); tx = np.zeros(max(R, 1), dtype=np.uint32); pr = np.zeros(max(R, 1), dtype=np.float32)
    src = np.zeros(max(R, 1), dtype=np.uint32); grp = np.zeros(max(G, 1), dtype=np.uint32)
    rows = C.c_uint64(0); disc = (C.c_uint64 * 10)()
    o = _FilterOpts(which_strand, min_aligned_len, three_prime_clip, five_prime_clip, min_aligned_fraction, score_threshold, score_prob_denom)
    L = lib()
    L.oracle_filter.restype = C.c_uint64
    L.oracle_filter.argtypes = [_vp] * 8 + [C.c_uint64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
    nnz = L.oracle_filter(_p(gp), _p(cols[0]), _p(cols[1]), _p(cols[2]), _p(cols[3]), _p(sc), _p(fl), _p(sl), G, _p(tl), C.byref(o),
                          _p(rp), _p(tx), _p(pr), _p(src), _p(grp), C.byref(rows), disc)
    n = int(rows.value)
    return rp[:n + 1].copy(), tx[:nnz].copy(), pr[:nnz].copy(), src[:nnz].copy(), grp[:n].copy(), {k: int(v) for k, v in zip(DISCARD_NAMES, disc)}


class PortStore:
    """AoS copy of a store laid out like the Rust InMemoryAlignmentStore (CPU baseline timing)."""

    def __init__(self, row_ptr, txp, prob, n_txps, cov=None):
        row_ptr = _c(row_ptr, np.uint64); txp = _c(txp, np.uint32); prob = _c(prob, np.float32)
        cov = _c(cov, np.float64)
        self.n_txps = n_txps
        self.model_coverage = cov is not None
        self._h = lib().port_store_create(_p(row_ptr), _p(txp), _p(prob), _p(cov), len(row_ptr) - 1, len(txp), n_txps)

    def em_par(self, max_iter=1000, conv_thresh=1e-3, init=None):
        init = _c(init, np.float64)
        out = np.zeros(self.n_txps, dtype=np.float64)
        niter = C.c_uint32(0); rel = C.c_double(0.0)
        sweeps = lib().port_em_par(self._h, int(self.model_coverage), _p(init), max_iter, conv_thresh, _p(out),
                                   C.byref(niter), C.byref(rel))
        return out, int(niter.value), float(rel.value), int(sweeps)

    def bootstrap(self, num_boot, seed, max_iter=1000, conv_thresh=1e-3, nthreads=0):
        out = np.zeros((num_boot, self.n_txps), dtype=np.float64)
        niter = np.zeros(max(num_boot, 1), dtype=np.uint32)
        lib().port_bootstrap(self._h, int(self.model_coverage), num_boot, seed, max_iter, conv_thresh, nthreads,
                             _p(out), _p(niter))
        return out, niter[:num_boot]

    def bootstrap_timed(self, num_boot, seed, max_iter=1000, conv_thresh=1e-3, nthreads=0):
        """bootstrap() that also returns the seconds each replicate spent inside do_em (sample drawing and sorting excluded)."""
        out = np.zeros((num_boot, self.n_txps), dtype=np.float64)
        niter = np.zeros(max(num_boot, 1), dtype=np.uint32)
        secs = np.zeros(max(num_boot, 1), dtype=np.float64)
        lib().port_bootstrap_timed(self._h, int(self.model_coverage), num_boot, seed, max_iter, conv_thresh, nthreads,
                                   _p(out), _p(niter), _p(secs))
        return out, niter[:num_boot], secs[:num_boot]

    def close(self):
        if self._h:
            lib().port_store_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def num_threads() -> int:
    return int(lib().port_num_threads())


def set_num_threads(n: int) -> int:
    """Size the OpenMP pool explicitly (torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    lib().port_set_num_threads(int(n))
    return num_threads()
