"""Thin object wrapper over the C ABI (include/oarfish_em.h).

`DeviceStore` owns one `oar_store` handle: the alignment store uploaded to HBM.
Arrays may be numpy arrays (host), or torch tensors (pinned host or CUDA) --
only their addresses cross the boundary.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import check


def _addr(x, dtype: np.dtype, name: str) -> Tuple[int, int, object]:
    """(address, n_elements, keepalive) of a contiguous array of `dtype`."""
    if x is None:
        return 0, 0, None
    if isinstance(x, np.ndarray):
        if x.dtype != dtype:
            raise TypeError(f"{name}: expected dtype {dtype}, got {x.dtype}")
        if not x.flags["C_CONTIGUOUS"]:
            raise ValueError(f"{name}: array must be C-contiguous")
        return x.ctypes.data, x.size, x
    # torch tensor (duck-typed so that torch stays an optional import here)
    if hasattr(x, "data_ptr") and hasattr(x, "is_contiguous"):
        import torch

        want = {np.dtype(np.uint64): (torch.uint64, torch.int64), np.dtype(np.uint32): (torch.uint32, torch.int32),
                np.dtype(np.float32): (torch.float32,), np.dtype(np.float64): (torch.float64,)}[np.dtype(dtype)]
        if x.dtype not in want:
            raise TypeError(f"{name}: expected torch dtype in {want}, got {x.dtype}")
        if not x.is_contiguous():
            raise ValueError(f"{name}: tensor must be contiguous")
        return x.data_ptr(), x.numel(), x
    raise TypeError(f"{name}: unsupported array type {type(x)}")


class EMResult:
    __slots__ = ("counts", "niter", "rel_diff")

    def __init__(self, counts, niter: int, rel_diff: float):
        self.counts = counts
        self.niter = niter
        self.rel_diff = rel_diff


class DeviceStore:
    """An InMemoryAlignmentStore (src/util/oarfish_types.rs:547-558) resident in HBM as CSR."""

    def __init__(self, row_ptr, txp_id, prob, n_txps: int, aux=None, device: int = 0):
        lib = _lib.load_em_lib()
        rp, n_rp, k0 = _addr(row_ptr, np.dtype(np.uint64), "row_ptr")
        tp, n_t, k1 = _addr(txp_id, np.dtype(np.uint32), "txp_id")
        pp, n_p, k2 = _addr(prob, np.dtype(np.float32), "prob")
        ap, n_a, k3 = _addr(aux, np.dtype(np.float64), "aux")
        if n_rp < 1:
            raise ValueError("row_ptr must have at least one element")
        if n_t != n_p or (aux is not None and n_a != n_t):
            raise ValueError("txp_id, prob and aux must have the same length")
        self.n_reads = n_rp - 1
        self.nnz = n_t
        self.n_txps = int(n_txps)
        self.device = int(device)
        self._lib = lib
        self._h = C.c_void_p()
        check(lib.oar_store_create(rp, tp, pp, ap, self.n_reads, self.nnz, self.n_txps, self.device, C.byref(self._h)))

    @classmethod
    def from_records(cls, group_ptr, ref_id, aln_start, aln_end, aln_span, score, flags, seq_len, txp_len, *,
                     which_strand: int = 0, min_aligned_len: int = 50, three_prime_clip: int = 2**31 - 1,
                     five_prime_clip: int = 2**32 - 1, min_aligned_fraction: float = 0.5, score_threshold: float = 0.95,
                     score_prob_denom: float = 5.0, device: int = 0, want_index: bool = False):
        """AlignmentFilters::filter (oarfish_types.rs:955-1130) over all read groups on the device: returns
        (store, discard table dict[, (src record of every retained alignment, group of every retained read)])."""
        lib = _lib.load_em_lib()
        gp = np.ascontiguousarray(group_ptr, dtype=np.uint64)
        cols = [np.ascontiguousarray(a, dtype=np.uint32) for a in (ref_id, aln_start, aln_end, aln_span)]
        sc = np.ascontiguousarray(score, dtype=np.int32)
        fl = np.ascontiguousarray(flags, dtype=np.uint8)
        sl = np.ascontiguousarray(seq_len, dtype=np.uint32)
        tl = np.ascontiguousarray(txp_len, dtype=np.uint32)
        n_groups, n_rec = len(gp) - 1, len(sc)
        opts = _lib.FilterOpts(int(which_strand), int(min_aligned_len), int(three_prime_clip), int(five_prime_clip),
                               float(min_aligned_fraction), float(score_threshold), float(score_prob_denom))
        disc = (C.c_uint64 * 10)()
        src = np.zeros(max(n_rec, 1), dtype=np.uint32) if want_index else None
        grp = np.zeros(max(n_groups, 1), dtype=np.uint32) if want_index else None
        self = cls.__new__(cls)
        self._lib = lib
        self._h = C.c_void_p()
        check(lib.oar_store_create_filtered(gp.ctypes.data, cols[0].ctypes.data, cols[1].ctypes.data, cols[2].ctypes.data,
                                            cols[3].ctypes.data, sc.ctypes.data, fl.ctypes.data, sl.ctypes.data, n_groups, n_rec,
                                            tl.ctypes.data, len(tl), C.byref(opts), int(device), C.byref(self._h), disc,
                                            src.ctypes.data if want_index else None, grp.ctypes.data if want_index else None))
        nr, nz, nt, dv = C.c_uint64(0), C.c_uint64(0), C.c_uint32(0), C.c_int(0)
        check(lib.oar_store_info(self._h, C.byref(nr), C.byref(nz), C.byref(nt), C.byref(dv)))
        self.n_reads, self.nnz, self.n_txps, self.device = int(nr.value), int(nz.value), int(nt.value), int(dv.value)
        names = ("discard_5p", "discard_3p", "discard_score", "discard_aln_frac", "discard_aln_len", "discard_ori", "discard_supp",
                 "no_mapping", "no_valid_aln", "valid_best_aln")
        table = {k: int(v) for k, v in zip(names, disc)}
        if want_index:
            return self, table, (src[:self.nnz].copy(), grp[:self.n_reads].copy())
        return self, table

    def export_csr(self):
        """(row_ptr u64, txp_id u32, prob f32) of the store as it sits in HBM."""
        rp = np.zeros(self.n_reads + 1, dtype=np.uint64)
        tx = np.zeros(max(self.nnz, 1), dtype=np.uint32)
        pr = np.zeros(max(self.nnz, 1), dtype=np.float32)
        check(self._lib.oar_store_export(self._h, rp.ctypes.data, tx.ctypes.data, pr.ctypes.data))
        return rp, tx[:self.nnz], pr[:self.nnz]

    # -- lifetime ------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.oar_store_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- configuration ---------------------------------------------------------
    def set_kernel(self, kernel: int) -> None:
        check(self._lib.oar_store_set_kernel(self._h, int(kernel)))

    def set_progress(self, fn=None) -> None:
        """fn(niter, rel_diff) after every polled batch of EM iterations (the reference's em.rs:219-233 log lines)."""
        if fn is None:
            self._progress = None
            check(self._lib.oar_store_set_progress(self._h, None, None))
            return
        self._progress = _lib.PROGRESS_FN(lambda niter, rel, _user: fn(int(niter), float(rel)))
        check(self._lib.oar_store_set_progress(self._h, C.cast(self._progress, C.c_void_p), None))

    @property
    def stream(self) -> int:
        return int(self._lib.oar_store_stream(self._h) or 0)

    def timings_ms(self):
        out = (C.c_double * 4)()
        check(self._lib.oar_store_timings(self._h, out))
        return {"upload": out[0], "em": out[1], "download": out[2], "weights": out[3]}

    def layout_info(self):
        out = (C.c_uint64 * 8)()
        check(self._lib.oar_store_layout_info(self._h, out))
        keys = ("tiled", "n_tiles", "slots", "fallback_rows", "sum_distinct", "sum_units", "span", "kernel")
        return {k: int(v) for k, v in zip(keys, out)}

    def layout_lpos(self, first_tile: int, n_tiles: int, with_trash: bool = False):
        """Per-slot layout words of a range of tiles [and each tile's trash offset in doubles] (inspection; see oar_store_layout_lpos)."""
        out = np.empty((n_tiles, 1024), dtype=np.uint32)
        trash = np.empty(n_tiles, dtype=np.uint32)
        check(self._lib.oar_store_layout_lpos(self._h, first_tile, n_tiles, out.ctypes.data_as(C.c_void_p),
                                              trash.ctypes.data_as(C.c_void_p) if with_trash else None))
        return (out, trash // 8) if with_trash else out

    def counters(self):
        out = (C.c_uint64 * 2)()
        check(self._lib.oar_store_counters(self._h, out))
        return {"launches": int(out[0]), "sweeps": int(out[1])}

    # -- compute -----------------------------------------------------------------
    def em(self, max_iter: int = 1000, conv_thresh: float = 1e-3, min_iter: int = 50,
           init=None, out=None) -> EMResult:
        """One EM to convergence (em::em with min_iter=50, em::em_par with min_iter=1)."""
        if out is None:
            out = np.empty(self.n_txps, dtype=np.float64)
        op, n_o, _ = _addr(out, np.dtype(np.float64), "out")
        if n_o != self.n_txps:
            raise ValueError("out must have n_txps elements")
        ip, n_i, _k = _addr(init, np.dtype(np.float64), "init")
        if init is not None and n_i != self.n_txps:
            raise ValueError("init must have n_txps elements")
        niter = C.c_uint32(0)
        rel = C.c_double(0.0)
        check(self._lib.oar_em(self._h, ip, int(max_iter), float(conv_thresh), int(min_iter), op,
                               C.byref(niter), C.byref(rel)))
        return EMResult(out, int(niter.value), float(rel.value))

    def bootstrap(self, num_boot: int, seed: int, max_iter: int = 1000, conv_thresh: float = 1e-3,
                  first_replicate: int = 0, replicate_stride: int = 1, out=None):
        """em::bootstrap (em.rs:292): returns (num_boot x M counts, niter per replicate)."""
        if out is None:
            out = np.empty((num_boot, self.n_txps), dtype=np.float64)
        op, n_o, _ = _addr(out, np.dtype(np.float64), "out")
        if n_o != num_boot * self.n_txps:
            raise ValueError("out must have num_boot * n_txps elements")
        niter = np.zeros(max(num_boot, 1), dtype=np.uint32)
        check(self._lib.oar_bootstrap(self._h, int(num_boot), int(seed), int(first_replicate), int(replicate_stride),
                                      int(max_iter), float(conv_thresh), op, niter.ctypes.data))
        return out, niter[:num_boot]

    def bootstrap_weights(self, weights, max_iter: int = 1000, conv_thresh: float = 1e-3, min_iter: int = 50,
                          out=None):
        wp, n_w, _ = _addr(weights, np.dtype(np.uint32), "weights")
        if self.n_reads == 0 or n_w % self.n_reads != 0:
            raise ValueError("weights must hold R * n_reads elements")
        R = n_w // self.n_reads
        if out is None:
            out = np.empty((R, self.n_txps), dtype=np.float64)
        op, n_o, _ = _addr(out, np.dtype(np.float64), "out")
        niter = np.zeros(max(R, 1), dtype=np.uint32)
        check(self._lib.oar_bootstrap_weights(self._h, wp, R, int(max_iter), float(conv_thresh), int(min_iter), op,
                                              niter.ctypes.data))
        return out, niter[:R]

    def sample_weights(self, seed: int, replicate: int, out=None):
        if out is None:
            out = np.empty(self.n_reads, dtype=np.uint32)
        op, n_o, _ = _addr(out, np.dtype(np.uint32), "out")
        check(self._lib.oar_bootstrap_sample_weights(self._h, int(seed), int(replicate), op))
        return out

    def em_batched(self, cell_row_ptr, max_iter: int = 1000, conv_thresh: float = 1e-3, min_iter: int = 50):
        """Per-cell EMs (single_cell.rs:150).  Returns (cell_ptr u64[C+1], txp u32, val f64, niter u32[C]):
        cell c's transcripts with an alignment, ascending, and their counts."""
        crp = np.ascontiguousarray(cell_row_ptr, dtype=np.uint64)
        n_cells = len(crp) - 1
        cell_ptr = np.zeros(n_cells + 1, dtype=np.uint64)
        cap = self.nnz
        txp = np.empty(max(cap, 1), dtype=np.uint32)
        val = np.empty(max(cap, 1), dtype=np.float64)
        niter = np.zeros(max(n_cells, 1), dtype=np.uint32)
        nnz = C.c_uint64(0)
        check(self._lib.oar_em_batched(self._h, crp.ctypes.data, n_cells, int(max_iter), float(conv_thresh), int(min_iter),
                                       cell_ptr.ctypes.data, txp.ctypes.data, val.ctypes.data, cap, C.byref(nnz),
                                       niter.ctypes.data))
        n = int(nnz.value)
        return cell_ptr, txp[:n].copy(), val[:n].copy(), niter[:n_cells]

    def coverage_model(self, aln_start, aln_end, txp_len, bin_width: int = 100, growth_rate: float = 2.0, model: str = "logistic"):
        """--model-coverage (bulk.rs:103-108): computes coverage_probabilities on the device, installs them as the
        store's aux factor and returns them (f64[nnz]).  model="binomial" is the single-cell driver's bin model
        (single_cell.rs:132-137, binomial_probability.rs)."""
        sp, n_s, _ = _addr(np.ascontiguousarray(aln_start, dtype=np.uint32), np.dtype(np.uint32), "aln_start")
        ep, n_e, _k = _addr(np.ascontiguousarray(aln_end, dtype=np.uint32), np.dtype(np.uint32), "aln_end")
        lp, n_l, _k2 = _addr(np.ascontiguousarray(txp_len, dtype=np.uint32), np.dtype(np.uint32), "txp_len")
        if n_s != self.nnz or n_e != self.nnz or n_l != self.n_txps:
            raise ValueError("aln_start/aln_end need nnz elements, txp_len n_txps")
        out = np.zeros(max(self.nnz, 1), dtype=np.float64)
        if model == "binomial":
            check(self._lib.oar_store_coverage_model_binomial(self._h, sp, ep, lp, int(bin_width), out.ctypes.data))
        elif model == "logistic":
            check(self._lib.oar_store_coverage_model(self._h, sp, ep, lp, int(bin_width), float(growth_rate), out.ctypes.data))
        else:
            raise ValueError(f"unknown coverage model {model!r}")
        return out[:self.nnz]

    def posteriors(self, counts, display_thresh: float = 0.0):
        """write_out_prob's inner loop (write_function.rs:283-332): (per-alignment probs f64[nnz], kept u32[N])."""
        cp, n_c, _ = _addr(np.ascontiguousarray(counts, dtype=np.float64), np.dtype(np.float64), "counts")
        if n_c != self.n_txps:
            raise ValueError("counts must have n_txps elements")
        out = np.zeros(max(self.nnz, 1), dtype=np.float64)
        kept = np.zeros(max(self.n_reads, 1), dtype=np.uint32)
        check(self._lib.oar_posteriors(self._h, cp, float(display_thresh), out.ctypes.data, kept.ctypes.data))
        return out[:self.nnz], kept[:self.n_reads]

    def aux_counts(self):
        """aux_counts::get_aux_counts (aux_counts.rs:23-50): (unique u32[M], total u32[M])."""
        uniq = np.zeros(self.n_txps, dtype=np.uint32)
        tot = np.zeros(self.n_txps, dtype=np.uint32)
        check(self._lib.oar_aux_counts(self._h, uniq.ctypes.data, tot.ctypes.data))
        return uniq, tot

    def sweep(self, prev_dev, curr_dev, weights_dev=None, sync: bool = True) -> None:
        """One raw fused E+M sweep on device buffers (torch CUDA tensors)."""
        pp, n_p, _ = _addr(prev_dev, np.dtype(np.float64), "prev")
        cp, n_c, _ = _addr(curr_dev, np.dtype(np.float64), "curr")
        wp, _n, _k = _addr(weights_dev, np.dtype(np.uint32), "weights")
        if n_p != self.n_txps or n_c != self.n_txps:
            raise ValueError("prev/curr must have n_txps elements")
        check(self._lib.oar_sweep(self._h, pp, cp, wp, 1 if sync else 0))


    def sweep_timed(self, prev_dev, curr_dev, reps: int, weights_dev=None) -> float:
        pp, n_p, _ = _addr(prev_dev, np.dtype(np.float64), "prev")
        cp, n_c, _ = _addr(curr_dev, np.dtype(np.float64), "curr")
        wp, _n, _k = _addr(weights_dev, np.dtype(np.uint32), "weights")
        ms = C.c_float(0.0)
        check(self._lib.oar_sweep_timed(self._h, pp, cp, wp, int(reps), C.byref(ms)))
        return float(ms.value)


class _BorrowedStore(DeviceStore):
    """A DeviceStore view of a store owned by a MultiStore (never destroyed from here)."""

    def __init__(self, lib, handle, n_reads, nnz, n_txps, device):
        self._lib = lib
        self._h = C.c_void_p(handle)
        self.n_reads, self.nnz, self.n_txps, self.device = n_reads, nnz, n_txps, device

    def close(self) -> None:
        self._h = C.c_void_p()


class MultiStore:
    """One resident copy of a store on each of `devices`, driven from one host thread (oar_multi_*):
    em::bootstrap's internal fan-out (em.rs:292-314) with devices in place of pool threads."""

    def __init__(self, row_ptr, txp_id, prob, n_txps: int, aux=None, devices: Optional[Sequence[int]] = None):
        lib = _lib.load_em_lib()
        if devices is None:
            devices = list(range(device_count()))
        rp, n_rp, k0 = _addr(row_ptr, np.dtype(np.uint64), "row_ptr")
        tp, n_t, k1 = _addr(txp_id, np.dtype(np.uint32), "txp_id")
        pp, n_p, k2 = _addr(prob, np.dtype(np.float32), "prob")
        ap, n_a, k3 = _addr(aux, np.dtype(np.float64), "aux")
        self.n_reads, self.nnz, self.n_txps = n_rp - 1, n_t, int(n_txps)
        self.devices = [int(d) for d in devices]
        dv = (C.c_int * len(self.devices))(*self.devices)
        self._lib = lib
        self._h = C.c_void_p()
        check(lib.oar_multi_create(rp, tp, pp, ap, self.n_reads, self.nnz, self.n_txps, dv, len(self.devices), C.byref(self._h)))

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.oar_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def store(self, i: int = 0) -> DeviceStore:
        h = self._lib.oar_multi_store(self._h, int(i))
        if not h:
            raise IndexError(i)
        return _BorrowedStore(self._lib, h, self.n_reads, self.nnz, self.n_txps, self.devices[i])

    def info(self):
        n = C.c_int(0)
        ms = (C.c_double * 2)()
        per = np.zeros(len(self.devices), dtype=np.uint32)
        check(self._lib.oar_multi_info(self._h, C.byref(n), ms, per.ctypes.data))
        return {"n_devices": int(n.value), "upload_ms": ms[0], "replicate_ms": ms[1], "last_per_device": per.tolist()}

    def bootstrap(self, num_boot: int, seed: int, max_iter: int = 1000, conv_thresh: float = 1e-3, out=None):
        """em::bootstrap over all devices: (num_boot x M counts, niter per replicate); row g is replicate (seed, g)."""
        if out is None:
            out = np.empty((num_boot, self.n_txps), dtype=np.float64)
        op, n_o, _ = _addr(out, np.dtype(np.float64), "out")
        if n_o != num_boot * self.n_txps:
            raise ValueError("out must have num_boot * n_txps elements")
        niter = np.zeros(max(num_boot, 1), dtype=np.uint32)
        check(self._lib.oar_multi_bootstrap(self._h, int(num_boot), int(seed), int(max_iter), float(conv_thresh), op,
                                            niter.ctypes.data))
        return out, niter[:num_boot]


def em_batched_multi(row_ptr, txp_id, prob, n_txps: int, cell_row_ptr, devices: Optional[Sequence[int]] = None, aux=None,
                     max_iter: int = 1000, conv_thresh: float = 1e-3, min_iter: int = 50):
    """Per-cell EMs sharded over devices (oar_em_batched_multi).  Host arrays only.
    Returns (cell_ptr u64[C+1], txp u32, val f64, niter u32[C], cells per device)."""
    lib = _lib.load_em_lib()
    if devices is None:
        devices = list(range(device_count()))
    rp, n_rp, k0 = _addr(row_ptr, np.dtype(np.uint64), "row_ptr")
    tp, n_t, k1 = _addr(txp_id, np.dtype(np.uint32), "txp_id")
    pp, n_p, k2 = _addr(prob, np.dtype(np.float32), "prob")
    ap, n_a, k3 = _addr(aux, np.dtype(np.float64), "aux")
    crp = np.ascontiguousarray(cell_row_ptr, dtype=np.uint64)
    n_cells = len(crp) - 1
    dv = (C.c_int * len(devices))(*[int(d) for d in devices])
    cell_ptr = np.zeros(n_cells + 1, dtype=np.uint64)
    cap = n_t
    txp = np.empty(max(cap, 1), dtype=np.uint32)
    val = np.empty(max(cap, 1), dtype=np.float64)
    niter = np.zeros(max(n_cells, 1), dtype=np.uint32)
    per = np.zeros(len(devices), dtype=np.uint32)
    nnz = C.c_uint64(0)
    check(lib.oar_em_batched_multi(rp, tp, pp, ap, n_rp - 1, n_t, int(n_txps), crp.ctypes.data, n_cells, dv, len(devices),
                                   int(max_iter), float(conv_thresh), int(min_iter), cell_ptr.ctypes.data, txp.ctypes.data,
                                   val.ctypes.data, cap, C.byref(nnz), niter.ctypes.data, per.ctypes.data))
    n = int(nnz.value)
    return cell_ptr, txp[:n].copy(), val[:n].copy(), niter[:n_cells], per.tolist()


def device_count() -> int:
    n = _lib.load_em_lib().oar_device_count()
    if n < 0:
        check(n)
    return n
