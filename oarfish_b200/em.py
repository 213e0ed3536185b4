"""Host-side mirror of the reference's inference interface (src/em.rs).

Same names, argument meaning and return shapes as the Rust functions the
oarfish drivers call, so that the parity tests read like tests of the
reference:

    em(em_info, nthreads)            <- em::em        (em.rs:262)
    em_par(em_info, nthreads)        <- em::em_par    (em.rs:320)
    bootstrap(em_info, n, nthreads)  <- em::bootstrap (em.rs:292)

`InMemoryAlignmentStore`, `AlnInfo` and `EMInfo` mirror
src/util/oarfish_types.rs:330-344, :408-428 and :547-738 as far as the EM reads
them.  The shim flattens the store exactly as the Rust shim in
INTEGRATION.md does (boundaries -> row_ptr, AlnInfo.ref_id -> txp_id,
as_probabilities -> prob, coverage_probabilities -> aux when model_coverage)
and calls the C ABI.  Everything numeric happens on the GPU.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Iterator, List, Optional, Tuple

import numpy as np

from .engine import DeviceStore

# AlnInfo (oarfish_types.rs:330-337).  rustc reorders the fields of a default-repr struct by
# alignment (f64 first), giving 24 bytes; the field order below reproduces that size.
ALN_INFO_DTYPE = np.dtype(
    [("prob", np.float64), ("ref_id", np.uint32), ("start", np.uint32), ("end", np.uint32), ("strand", np.uint8)],
    align=True,
)


@dataclass
class AlignmentFilters:
    """Only the field of AlignmentFilters (oarfish_types.rs:763-800) that the EM reads."""
    model_coverage: bool = False


@dataclass
class TranscriptInfo:
    """TranscriptInfo (oarfish_types.rs:430-437); the EM reads only lenf (for the KDE hook)."""
    len: int = 1
    lenf: float = 1.0


class InMemoryAlignmentStore:
    """CSR-like container of per-read alignment groups (oarfish_types.rs:547-558)."""

    def __init__(self, filter_opts: Optional[AlignmentFilters] = None):
        self.filter_opts = filter_opts or AlignmentFilters()
        self.alignments = np.zeros(0, dtype=ALN_INFO_DTYPE)
        self.as_probabilities = np.zeros(0, dtype=np.float32)
        self.coverage_probabilities = np.zeros(0, dtype=np.float64)
        self._boundaries = np.zeros(1, dtype=np.uint64)  # private in the reference (:555)
        self._pending: List[Tuple[np.ndarray, np.ndarray]] = []
        self._device_store: Optional[DeviceStore] = None
        self._device_key = None

    # -- construction ---------------------------------------------------------
    @classmethod
    def from_csr(cls, row_ptr, txp_id, prob, coverage=None, model_coverage: bool = False,
                 start=None, end=None) -> "InMemoryAlignmentStore":
        s = cls(AlignmentFilters(model_coverage=model_coverage))
        n = len(txp_id)
        s.alignments = np.zeros(n, dtype=ALN_INFO_DTYPE)
        s.alignments["ref_id"] = txp_id
        s.alignments["start"] = 0 if start is None else start
        s.alignments["end"] = 1 if end is None else end
        s.as_probabilities = np.ascontiguousarray(prob, dtype=np.float32)
        s.coverage_probabilities = (np.zeros(n, dtype=np.float64) if coverage is None
                                    else np.ascontiguousarray(coverage, dtype=np.float64))
        s._boundaries = np.ascontiguousarray(row_ptr, dtype=np.uint64)
        return s

    def add_filtered_group(self, alns: np.ndarray, as_probs: np.ndarray) -> bool:
        """oarfish_types.rs:718-738: append one read's alignments; empty groups are dropped."""
        if len(alns) == 0:
            return False
        self._pending.append((np.asarray(alns, dtype=ALN_INFO_DTYPE), np.asarray(as_probs, dtype=np.float32)))
        return True

    def _flush(self) -> None:
        if not self._pending:
            return
        alns = [self.alignments] + [a for a, _ in self._pending]
        probs = [self.as_probabilities] + [p for _, p in self._pending]
        lens = np.array([len(a) for a, _ in self._pending], dtype=np.uint64)
        self.alignments = np.concatenate(alns)
        self.as_probabilities = np.concatenate(probs)
        self.coverage_probabilities = np.concatenate(
            [self.coverage_probabilities, np.zeros(int(lens.sum()), dtype=np.float64)])
        self._boundaries = np.concatenate([self._boundaries, self._boundaries[-1] + np.cumsum(lens)])
        self._pending = []
        self._device_store = None

    # -- reference accessors ----------------------------------------------------
    def len(self) -> int:  # oarfish_types.rs:562-564
        self._flush()
        return max(len(self._boundaries) - 1, 0)

    __len__ = len

    def num_aligned_reads(self) -> int:  # :745-747
        return self.len()

    def total_len(self) -> int:  # :740-742
        self._flush()
        return len(self.alignments)

    def iter(self) -> Iterator[Tuple[np.ndarray, np.ndarray, np.ndarray]]:  # :651-656
        self._flush()
        b = self._boundaries
        for i in range(len(b) - 1):
            s, e = int(b[i]), int(b[i + 1])
            yield self.alignments[s:e], self.as_probabilities[s:e], self.coverage_probabilities[s:e]

    def random_sampling_iter(self, inds):  # :658-669
        self._flush()
        b = self._boundaries
        for i in inds:
            s, e = int(b[i]), int(b[i + 1])
            yield self.alignments[s:e], self.as_probabilities[s:e], self.coverage_probabilities[s:e]

    # -- the shim: flatten and upload --------------------------------------------
    def csr(self):
        """(row_ptr u64, txp_id u32, prob f32, aux f64|None) exactly as handed to oar_store_create."""
        self._flush()
        txp = np.ascontiguousarray(self.alignments["ref_id"])
        aux = np.ascontiguousarray(self.coverage_probabilities) if self.filter_opts.model_coverage else None
        return np.ascontiguousarray(self._boundaries), txp, np.ascontiguousarray(self.as_probabilities), aux

    def invalidate_device(self) -> None:
        """Drop the HBM copy: the next em / em_par / bootstrap uploads the store again."""
        if self._device_store is not None:
            self._device_store.close()
        self._device_store = None
        self._device_key = None

    def device_store(self, n_txps: int, device: int = 0) -> DeviceStore:
        """The HBM copy of the store (created on first use, reused by em then bootstrap like bulk.rs:155-179).  It is
        a snapshot keyed on the store's shape AND on the identity of its arrays: replacing `alignments`,
        `as_probabilities` or `coverage_probabilities` (normalize_read_probs does that once the store is built) or
        changing `filter_opts.model_coverage` uploads again; after mutating an array IN PLACE call invalidate_device()."""
        self._flush()
        key = (int(n_txps), int(device), len(self.alignments), self.filter_opts.model_coverage, id(self.alignments),
               id(self.as_probabilities), id(self.coverage_probabilities))
        if self._device_store is None or self._device_key != key:
            if self._device_store is not None:
                self._device_store.close()
            rp, txp, prob, aux = self.csr()
            self._device_store = DeviceStore(rp, txp, prob, n_txps, aux=aux, device=device)
            self._device_key = key
        return self._device_store


@dataclass
class EMInfo:
    """EMInfo (oarfish_types.rs:408-428)."""
    eq_map: InMemoryAlignmentStore
    txp_info: List[TranscriptInfo]
    max_iter: int = 1000           # --max-em-iter default, prog_opts.rs:532
    convergence_thresh: float = 1e-3  # --convergence-thresh default, prog_opts.rs:536
    init_abundances: Optional[np.ndarray] = None
    kde_model: Optional[object] = None
    device: int = field(default_factory=lambda: int(os.environ.get("OARFISH_EM_DEVICE", "0")))


def _run(em_info: EMInfo, min_iter: int) -> np.ndarray:
    if em_info.kde_model is not None:
        # em.rs:173-178 reads kde_model[(txp_len, aln_span)] from the un-vendored `kders` crate; the
        # factor is iteration-invariant and must be folded into coverage_probabilities by the caller.
        raise NotImplementedError("kde_model: fold the density into coverage_probabilities and set model_coverage")
    n_txps = len(em_info.txp_info)
    store = em_info.eq_map.device_store(n_txps, em_info.device)
    init = None
    if em_info.init_abundances is not None:
        init = np.ascontiguousarray(em_info.init_abundances, dtype=np.float64)
    return store.em(max_iter=em_info.max_iter, conv_thresh=em_info.convergence_thresh, min_iter=min_iter,
                    init=init).counts


def em(em_info: EMInfo, _nthreads: int = 1) -> np.ndarray:
    """em::em (em.rs:262-271): do_em with the `niter > 50` stop rule (em.rs:212)."""
    return _run(em_info, 50)


def em_par(em_info: EMInfo, nthreads: int = 1) -> np.ndarray:
    """em::em_par (em.rs:320-447): same EM, stop rule `niter > 1` (em.rs:399)."""
    return _run(em_info, 1)


def bootstrap(em_info: EMInfo, num_boot: int, nthreads: int = 1, seed: Optional[int] = None) -> List[np.ndarray]:
    """em::bootstrap (em.rs:292-314): `num_boot` resampled EMs, result[replicate][transcript].

    The reference draws from the unseeded thread RNG (em.rs:274); here the
    resampling is keyed by `seed` (entropy from the OS when None), so runs are
    reproducible on request."""
    if seed is None:
        seed = int.from_bytes(os.urandom(8), "little")
    n_txps = len(em_info.txp_info)
    store = em_info.eq_map.device_store(n_txps, em_info.device)
    out, _ = store.bootstrap(num_boot, seed, max_iter=em_info.max_iter, conv_thresh=em_info.convergence_thresh)
    return [out[i] for i in range(num_boot)]
