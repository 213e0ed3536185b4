"""Multi-GPU plumbing for the bootstrap path (one process per GPU, torch.distributed).

Bootstrap replicates are independent units (em.rs:298-313 fans them out on a
rayon pool): global replicate g runs on rank g mod G and its resampling
weights are a pure function of (seed, g), so results do not depend on G.  The
alignment store is broadcast once from rank 0 (NCCL over NVLink on the GPU box,
gloo in the CPU tests); there is no collective on the EM data path.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np


def shard_replicates(num_boot: int, rank: int, world: int) -> Tuple[int, int, int]:
    """(first_replicate, stride, count) of the replicates `rank` owns."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    count = (num_boot - rank + world - 1) // world if num_boot > rank else 0
    return rank, world, count


def replicate_owner_table(num_boot: int, world: int) -> List[List[int]]:
    return [list(range(r, num_boot, world)) for r in range(world)]


def broadcast_store(row_ptr, txp_id, prob, n_txps: int, aux=None, src: int = 0, device=None):
    """Broadcast the CSR arrays of the store from `src` to every rank.

    On `src` the arguments are the arrays (numpy or torch, host or device); on
    the other ranks they are ignored (pass None).  Returns torch tensors on
    `device` (a torch.device; CPU for gloo)."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank()
    device = torch.device("cpu") if device is None else torch.device(device)

    def as_tensor(a, dtype):
        if a is None:
            return None
        if isinstance(a, np.ndarray):
            if a.dtype == np.uint64:
                a = a.view(np.int64)
            elif a.dtype == np.uint32:
                a = a.view(np.int32)
            a = torch.from_numpy(a)
        return a.to(device=device, dtype=dtype, non_blocking=True)

    hdr = torch.zeros(4, dtype=torch.int64, device=device)
    if rank == src:
        hdr[0] = len(row_ptr) - 1
        hdr[1] = len(txp_id)
        hdr[2] = int(n_txps)
        hdr[3] = 1 if aux is not None else 0
    dist.broadcast(hdr, src)
    n_reads, nnz, n_txps, has_aux = (int(x) for x in hdr.tolist())
    if rank == src:
        t_rp = as_tensor(row_ptr, torch.int64)
        t_tx = as_tensor(txp_id, torch.int32)
        t_pr = as_tensor(prob, torch.float32)
        t_ax = as_tensor(aux, torch.float64) if has_aux else None
    else:
        t_rp = torch.empty(n_reads + 1, dtype=torch.int64, device=device)
        t_tx = torch.empty(nnz, dtype=torch.int32, device=device)
        t_pr = torch.empty(nnz, dtype=torch.float32, device=device)
        t_ax = torch.empty(nnz, dtype=torch.float64, device=device) if has_aux else None
    for t in (t_rp, t_tx, t_pr, t_ax):
        if t is not None:
            dist.broadcast(t, src)
    return t_rp, t_tx, t_pr, t_ax, n_txps


def gather_replicates(local: "np.ndarray", num_boot: int, world: int, rank: int, n_txps: int) -> Optional[np.ndarray]:
    """Reassemble [replicate][transcript] on rank 0 from the per-rank shards (host tensors)."""
    import torch
    import torch.distributed as dist

    shards = [None] * world if rank == 0 else None
    dist.gather_object(np.ascontiguousarray(local), shards, dst=0)
    if rank != 0:
        return None
    out = np.empty((num_boot, n_txps), dtype=np.float64)
    for r, sh in enumerate(shards):
        idx = list(range(r, num_boot, world))
        if idx:
            out[idx] = np.asarray(sh).reshape(len(idx), n_txps)
    return out


# ---------------------------------------------------------------------------
# Dynamic replicate scheduling (opt-in)
# ---------------------------------------------------------------------------
# Replicates differ in length (486..1000 EM iterations on C3), so the static g mod G split leaves ranks idle at the end
# (8 GPUs: 7.1x one GPU).  Because a replicate's weights depend only on (seed, g), WHICH rank runs it is free: ranks can
# pull the next global replicate id from a shared counter instead.  The counter lives in the process group's key-value
# store (TCPStore.add is atomic); one round trip per replicate (~0.1 ms against ~100 ms of EM), still no collective on
# the EM data path.

_queue_serial = 0


class ReplicateQueue:
    """Shared work queue over the global replicate ids 0 .. num_boot-1 (collective constructor: every rank must create
    its queues in the same order)."""

    def __init__(self, num_boot: int, store=None):
        global _queue_serial
        import torch.distributed as dist

        if store is None:
            from torch.distributed import distributed_c10d as c10d
            store = c10d._get_default_store()
        self._store = store
        self._num_boot = int(num_boot)
        self._key = f"oarfish_b200/replicate_queue/{_queue_serial}"
        _queue_serial += 1
        if dist.get_rank() == 0:
            store.add(self._key, 0)      # create the counter before anybody pulls from it
        dist.barrier()

    def next(self) -> Optional[int]:
        """The next replicate id nobody has taken yet, or None when all are handed out."""
        g = int(self._store.add(self._key, 1)) - 1
        return g if g < self._num_boot else None


def run_replicates_dynamic(run_one, num_boot: int, store=None) -> Tuple[List[int], list]:
    """Pull replicate ids from a ReplicateQueue until it is empty; `run_one(g)` computes global replicate g.
    Returns (ids this rank ran, their results in the same order)."""
    q = ReplicateQueue(num_boot, store)
    ids, results = [], []
    while True:
        g = q.next()
        if g is None:
            break
        ids.append(g)
        results.append(run_one(g))
    return ids, results


def gather_replicates_by_id(ids: List[int], local: "np.ndarray", num_boot: int, rank: int, n_txps: int) -> Optional[np.ndarray]:
    """Reassemble [replicate][transcript] on rank 0 when every rank ran an arbitrary subset `ids` (rows of `local`)."""
    import torch.distributed as dist

    world = dist.get_world_size()
    shards = [None] * world if rank == 0 else None
    dist.gather_object((list(ids), np.ascontiguousarray(local)), shards, dst=0)
    if rank != 0:
        return None
    out = np.full((num_boot, n_txps), np.nan, dtype=np.float64)
    seen = []
    for sh_ids, sh in shards:
        if sh_ids:
            out[sh_ids] = np.asarray(sh).reshape(len(sh_ids), n_txps)
            seen += sh_ids
    if sorted(seen) != list(range(num_boot)):
        raise RuntimeError("replicate queue: the ranks did not cover every replicate exactly once")
    return out
