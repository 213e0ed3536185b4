"""Seeded synthetic alignment stores; see csrc/synth.c."""
from __future__ import annotations

from typing import NamedTuple, Optional

import numpy as np

from . import _lib


class SynthStore(NamedTuple):
    row_ptr: np.ndarray   # N+1 u64
    txp_id: np.ndarray    # nnz u32
    prob: np.ndarray      # nnz f32
    n_txps: int
    true_txp: Optional[np.ndarray]
    abund: Optional[np.ndarray]

    @property
    def n_reads(self) -> int:
        return len(self.row_ptr) - 1

    @property
    def nnz(self) -> int:
        return len(self.txp_id)


# the configurations BASELINE.md section 4 names
CONFIGS = {
    "tiny": dict(n_reads=2_000, n_txps=300, avg_aln=4.0, seed=11),
    "small": dict(n_reads=50_000, n_txps=5_000, avg_aln=6.0, seed=12),
    "C2": dict(n_reads=1_000_000, n_txps=50_000, avg_aln=6.0, seed=2),
    "C3": dict(n_reads=10_000_000, n_txps=200_000, avg_aln=8.0, seed=3),
}


def make_store(n_reads: int, n_txps: int, avg_aln: float, seed: int, want_truth: bool = False,
               pinned: bool = False) -> SynthStore:
    """Generate a store on the host.  With pinned=True the arrays live in
    page-locked memory (torch) so that uploads run at full PCIe/C2C speed."""
    lib = _lib.load_synth_lib()

    def alloc(n, dtype):
        if pinned:
            import torch
            tdt = {np.uint64: torch.int64, np.uint32: torch.int32, np.float32: torch.float32,
                   np.float64: torch.float64}[dtype]
            t = torch.empty(max(n, 1), dtype=tdt, pin_memory=True)
            return t.numpy().view(dtype)[:n]
        return np.empty(n, dtype=dtype)

    row_ptr = alloc(n_reads + 1, np.uint64)
    rc = lib.oar_synth_plan(n_reads, n_txps, float(avg_aln), seed, row_ptr.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"oar_synth_plan failed ({rc})")
    nnz = int(row_ptr[-1])
    txp = alloc(nnz, np.uint32)
    prob = alloc(nnz, np.float32)
    true_txp = np.empty(n_reads, dtype=np.uint32) if want_truth else None
    abund = np.empty(n_txps, dtype=np.float64) if want_truth else None
    rc = lib.oar_synth_fill(n_reads, n_txps, float(avg_aln), seed, row_ptr.ctypes.data, txp.ctypes.data,
                            prob.ctypes.data, true_txp.ctypes.data if want_truth else None,
                            abund.ctypes.data if want_truth else None)
    if rc != 0:
        raise RuntimeError(f"oar_synth_fill failed ({rc})")
    return SynthStore(row_ptr, txp, prob, n_txps, true_txp, abund)


def make_config(name: str, **kw) -> SynthStore:
    return make_store(**CONFIGS[name], **kw)


def permute_ids(store: SynthStore, seed: int) -> SynthStore:
    """The same store with transcript ids shuffled: isoforms of a gene are no longer neighbours in id space (a
    reference whose sequences are not grouped by gene).  Robustness workload; the EM result is the permuted one."""
    rng = np.random.default_rng(seed)
    perm = rng.permutation(store.n_txps).astype(np.uint32)
    return SynthStore(store.row_ptr, perm[store.txp_id], store.prob, store.n_txps, None, None)


def with_long_rows(store: SynthStore, frac: float, lo: int, hi: int, seed: int) -> SynthStore:
    """Replace a fraction of the reads by reads with lo..hi alignments (beyond --best-n's default of 100): a window of
    neighbouring transcripts starting at the read's first target.  Robustness workload for the layout's row-length cliff."""
    rng = np.random.default_rng(seed)
    n = store.n_reads
    lens = np.diff(store.row_ptr).astype(np.int64)
    pick = rng.random(n) < frac
    new_lens = lens.copy()
    new_lens[pick] = rng.integers(lo, hi + 1, size=int(pick.sum()))
    new_lens = np.minimum(new_lens, store.n_txps)
    rp = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(new_lens, out=rp[1:])
    nnz = int(rp[-1])
    txp = np.empty(nnz, dtype=np.uint32)
    prob = np.empty(nnz, dtype=np.float32)
    keep = ~pick
    # unchanged rows: copy their alignments
    src_idx = np.repeat(store.row_ptr[:-1][keep].astype(np.int64), lens[keep]) + (np.arange(int(lens[keep].sum())) - np.repeat(np.cumsum(lens[keep]) - lens[keep], lens[keep]))
    dst_idx = np.repeat(rp[:-1][keep].astype(np.int64), lens[keep]) + (np.arange(int(lens[keep].sum())) - np.repeat(np.cumsum(lens[keep]) - lens[keep], lens[keep]))
    txp[dst_idx] = store.txp_id[src_idx]
    prob[dst_idx] = store.prob[src_idx]
    # long rows: first target of the old read, then the following transcripts (wrapping), probabilities from the generator's table
    L = new_lens[pick]
    first = store.txp_id[store.row_ptr[:-1][pick].astype(np.int64)].astype(np.int64)
    off = np.arange(int(L.sum())) - np.repeat(np.cumsum(L) - L, L)
    d_idx = np.repeat(rp[:-1][pick].astype(np.int64), L) + off
    txp[d_idx] = ((np.repeat(first, L) + off) % store.n_txps).astype(np.uint32)
    ptab = np.exp(-np.arange(61, dtype=np.float32) / np.float32(5.0)).astype(np.float32)
    pr = ptab[rng.integers(0, 61, size=len(d_idx))]
    pr[off == 0] = 1.0
    prob[d_idx] = pr
    return SynthStore(rp, txp, prob, store.n_txps, None, None)


def make_cells(cell_reads, n_txps: int, avg_aln: float, seed: int, expressed: Optional[int] = None):
    """Concatenated store of several cells (single-cell mode): cell c has cell_reads[c] reads drawn
    from its own abundance vector.  With `expressed` (about 5 k expressed
    transcripts per cell) a cell's reads come from its own sorted random subset of that many transcripts, picked as
    runs of neighbouring ids (genes).  Returns (SynthStore, cell_row_ptr u64[C+1])."""
    rps, txs, prs = [np.zeros(1, dtype=np.uint64)], [], []
    cell_row_ptr = np.zeros(len(cell_reads) + 1, dtype=np.uint64)
    off = 0
    for c, n in enumerate(cell_reads):
        cell_row_ptr[c + 1] = cell_row_ptr[c] + n
        if n == 0:
            continue
        if expressed is not None and expressed < n_txps:
            rng = np.random.default_rng(seed * 7919 + c)
            starts = rng.choice(n_txps // 8, size=max(1, expressed // 8), replace=False).astype(np.int64) * 8
            subset = np.unique((starts[:, None] + np.arange(8)[None, :]).ravel())
            subset = subset[subset < n_txps].astype(np.uint32)
            st = make_store(int(n), len(subset), avg_aln, seed * 1000003 + c)
            st = SynthStore(st.row_ptr, subset[st.txp_id], st.prob, n_txps, None, None)
        else:
            st = make_store(int(n), n_txps, avg_aln, seed * 1000003 + c)
        rps.append(st.row_ptr[1:] + np.uint64(off))
        txs.append(st.txp_id)
        prs.append(st.prob)
        off += st.nnz
    row_ptr = np.concatenate(rps)
    txp = np.concatenate(txs) if txs else np.zeros(0, np.uint32)
    prob = np.concatenate(prs) if prs else np.zeros(0, np.float32)
    return SynthStore(row_ptr, txp, prob, n_txps, None, None), cell_row_ptr


def make_coordinates(store: SynthStore, seed: int):
    """Synthetic transcript lengths and alignment intervals for the coverage model (AlnInfo.start/.end,
    TranscriptInfo.len): lengths in [300, 6000], each alignment a random sub-interval covering 30-100 %."""
    rng = np.random.default_rng(seed)
    txp_len = rng.integers(300, 6000, size=store.n_txps, dtype=np.int64)
    L = txp_len[store.txp_id.astype(np.int64)]
    span = np.maximum((rng.uniform(0.3, 1.0, size=store.nnz) * L).astype(np.int64), 1)
    start = (rng.uniform(0.0, 1.0, size=store.nnz) * (L - span + 1)).astype(np.int64)
    end = np.minimum(start + span, L)
    return start.astype(np.uint32), end.astype(np.uint32), txp_len.astype(np.uint32)


def make_records(n_groups: int, n_txps: int, seed: int, mean_records: float = 6.0):
    """Raw alignment records for the filter stage (AlignmentFilters::filter, oarfish_types.rs:955-1130): columns as
    oar_store_create_filtered takes them, with every reason for a discard represented (unmapped, supplementary and
    wrong-strand records, short spans, 3' / 5' clipping, low scores, groups with nothing valid, non-positive best
    scores, groups whose sequence length sits on a later record or on none).
    Returns dict(group_ptr, ref_id, aln_start, aln_end, aln_span, score, flags, seq_len, txp_len)."""
    rng = np.random.default_rng(seed)
    k = 1 + rng.poisson(mean_records - 1.0, size=n_groups)
    k[rng.random(n_groups) < 0.02] = 0                                  # reads without any record
    gp = np.zeros(n_groups + 1, dtype=np.uint64); np.cumsum(k, out=gp[1:])
    R = int(gp[-1])
    grp = np.repeat(np.arange(n_groups), k)
    txp_len = rng.integers(300, 6000, size=n_txps).astype(np.uint32)
    base_t = rng.integers(0, n_txps, size=n_groups)
    ref = ((base_t[grp] + rng.integers(0, 12, size=R)) % n_txps).astype(np.uint32)
    L = txp_len[ref].astype(np.int64)
    span = np.maximum((rng.uniform(0.02, 1.0, size=R) * L).astype(np.int64), 1)
    start = (rng.uniform(0, 1, size=R) * (L - span + 1)).astype(np.int64) + 1
    end = start + span - 1
    best = rng.integers(-20, 400, size=n_groups)
    score = (best[grp] - rng.integers(0, 40, size=R) * (rng.random(R) < 0.7)).astype(np.int32)
    flags = np.zeros(R, dtype=np.uint8)
    flags[rng.random(R) < 0.05] |= 1
    flags[rng.random(R) < 0.30] |= 2
    flags[rng.random(R) < 0.05] |= 4
    read_len = rng.integers(200, 4000, size=n_groups)
    seq = np.zeros(R, dtype=np.uint32)
    first = gp[:-1][k > 0].astype(np.int64)
    carrier = first + (rng.integers(0, 3, size=len(first)) % k[k > 0])   # the record that carries the sequence
    seq[carrier] = read_len[k > 0]
    seq[carrier[rng.random(len(carrier)) < 0.03]] = 0                    # some reads carry no sequence at all
    return dict(group_ptr=gp, ref_id=ref, aln_start=start.astype(np.uint32), aln_end=end.astype(np.uint32),
                aln_span=span.astype(np.uint32), score=score, flags=flags, seq_len=seq, txp_len=txp_len)
