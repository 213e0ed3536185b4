"""`.oarstore`: on-disk interchange of an alignment store.

Lets the Rust CLI (built elsewhere) dump the InMemoryAlignmentStore it hands to the EM
(src/util/oarfish_types.rs:547-558) so that it can be quantified, tested or benchmarked here.
Little-endian, 64-byte header, then the arrays exactly as `oar_store_create` takes them:

    0   8s   magic  b"OARSTORE"
    8   u32  version (1)
    12  u32  flags   bit 0: aux array present (coverage_probabilities * density, f64)
                     bit 1: the EM's answer for this store follows the arrays (counts f64[n_txps]) and the
                            EM parameters below are valid -- a dump of the reference's own run, the route to
                            pinned parity (INTEGRATION.md has the Rust side)
                     bit 2: that answer comes from this repo's restated oracle, NOT from the Rust binary
    16  u64  n_reads
    24  u64  nnz
    32  u64  n_txps
    40  u32  min_iter       50 = em::em / do_em, 1 = em_par          (flags & 2)
    44  u32  max_iter                                                 (flags & 2)
    48  f64  convergence_thresh                                       (flags & 2)
    56  u32  niter          loop counter at exit, 0xFFFFFFFF unknown  (flags & 2)
    60  4x   reserved (0)
    64  u64  row_ptr[n_reads + 1]      == boundaries
        u32  txp_id[nnz]               == AlnInfo.ref_id
        f32  prob[nnz]                 == as_probabilities
        (pad to 8 bytes)
        f64  aux[nnz]                  only if flags & 1
        (pad to 8 bytes)
        f64  counts[n_txps]            only if flags & 2
"""
from __future__ import annotations

import struct
from typing import Optional, Tuple

import numpy as np

MAGIC = b"OARSTORE"
VERSION = 1
_HDR = struct.Struct("<8sIIQQQIIdI4x")
NITER_UNKNOWN = 0xFFFFFFFF


def write_store(path: str, row_ptr, txp_id, prob, n_txps: int, aux=None, counts=None, min_iter: int = 0, max_iter: int = 0,
                conv_thresh: float = 0.0, niter: int = NITER_UNKNOWN, from_oracle: bool = False) -> None:
    row_ptr = np.ascontiguousarray(row_ptr, dtype="<u8")
    txp_id = np.ascontiguousarray(txp_id, dtype="<u4")
    prob = np.ascontiguousarray(prob, dtype="<f4")
    if len(txp_id) != len(prob) or int(row_ptr[-1]) != len(txp_id) or row_ptr[0] != 0:
        raise ValueError("inconsistent store arrays")
    with open(path, "wb") as f:
        flags = (1 if aux is not None else 0) | (2 if counts is not None else 0) | (4 if counts is not None and from_oracle else 0)
        f.write(_HDR.pack(MAGIC, VERSION, flags, len(row_ptr) - 1, len(txp_id), int(n_txps), int(min_iter), int(max_iter),
                          float(conv_thresh), int(niter)))
        f.write(row_ptr.tobytes()); f.write(txp_id.tobytes()); f.write(prob.tobytes())
        if aux is not None:
            aux = np.ascontiguousarray(aux, dtype="<f8")
            if len(aux) != len(txp_id):
                raise ValueError("aux must have nnz elements")
            f.write(b"\0" * ((-f.tell()) % 8))
            f.write(aux.tobytes())
        if counts is not None:
            counts = np.ascontiguousarray(counts, dtype="<f8")
            if len(counts) != int(n_txps):
                raise ValueError("counts must have n_txps elements")
            f.write(b"\0" * ((-f.tell()) % 8))
            f.write(counts.tobytes())


def read_store(path: str) -> Tuple[np.ndarray, np.ndarray, np.ndarray, int, Optional[np.ndarray]]:
    """-> (row_ptr u64, txp_id u32, prob f32, n_txps, aux f64 | None), memory-mapped."""
    return read_store_full(path)[:5]


def read_store_full(path: str):
    """-> (row_ptr, txp_id, prob, n_txps, aux | None, reference | None); reference = dict(counts, min_iter, max_iter,
    conv_thresh, niter | None, from_oracle) when the file carries the EM's answer."""
    with open(path, "rb") as f:
        magic, version, flags, n_reads, nnz, n_txps, min_iter, max_iter, thr, niter = _HDR.unpack(f.read(_HDR.size))
    if magic != MAGIC or version != VERSION:
        raise ValueError(f"{path}: not an .oarstore v{VERSION} file")
    off = _HDR.size
    row_ptr = np.memmap(path, dtype="<u8", mode="r", offset=off, shape=(n_reads + 1,)); off += 8 * (n_reads + 1)
    txp_id = np.memmap(path, dtype="<u4", mode="r", offset=off, shape=(nnz,)); off += 4 * nnz
    prob = np.memmap(path, dtype="<f4", mode="r", offset=off, shape=(nnz,)); off += 4 * nnz
    aux = None
    if flags & 1:
        off += (-off) % 8
        aux = np.memmap(path, dtype="<f8", mode="r", offset=off, shape=(nnz,)); off += 8 * nnz
    ref = None
    if flags & 2:
        off += (-off) % 8
        ref = {"counts": np.memmap(path, dtype="<f8", mode="r", offset=off, shape=(n_txps,)), "min_iter": int(min_iter),
               "max_iter": int(max_iter), "conv_thresh": float(thr), "niter": None if niter == NITER_UNKNOWN else int(niter),
               "from_oracle": bool(flags & 4)}
    if int(row_ptr[-1]) != nnz:
        raise ValueError(f"{path}: row_ptr does not end at nnz")
    return row_ptr, txp_id, prob, int(n_txps), aux, ref
