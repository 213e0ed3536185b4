"""world_size-2 gloo test of the bootstrap sharding plumbing (CPU only)."""
import os
import socket

import numpy as np
import pytest

from oarfish_b200 import dist as odist


def test_shard_replicates_cover_everything_once():
    for num_boot in (0, 1, 7, 100):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                first, stride, count = odist.shard_replicates(num_boot, r, world)
                mine = [first + i * stride for i in range(count)]
                assert mine == odist.replicate_owner_table(num_boot, world)[r]
                seen += mine
            assert sorted(seen) == list(range(num_boot))
    # 100 replicates over 8 ranks: 13,13,13,13,12,12,12,12
    assert [odist.shard_replicates(100, r, 8)[2] for r in range(8)] == [13, 13, 13, 13, 12, 12, 12, 12]


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oarfish_b200 import synth
        if rank == 0:
            s = synth.make_config("tiny")
            args = (s.row_ptr, s.txp_id, s.prob, s.n_txps)
        else:
            args = (None, None, None, 0)
        rp, tx, pr, ax, n_txps = odist.broadcast_store(*args, src=0)
        ref = synth.make_config("tiny")
        ok = (np.array_equal(rp.numpy().view(np.uint64), ref.row_ptr) and np.array_equal(tx.numpy().view(np.uint32), ref.txp_id)
              and np.array_equal(pr.numpy(), ref.prob) and n_txps == ref.n_txps and ax is None)
        # each rank fabricates "its" replicates: row g holds the value g
        num_boot = 5
        first, stride, count = odist.shard_replicates(num_boot, rank, world)
        local = np.stack([np.full(n_txps, float(first + i * stride)) for i in range(count)]) if count else np.zeros((0, n_txps))
        out = odist.gather_replicates(local, num_boot, world, rank, n_txps)
        if rank == 0:
            ok = ok and out.shape == (num_boot, n_txps) and all(np.all(out[g] == g) for g in range(num_boot))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_store_broadcast_and_gather_world2():
    import torch.multiprocessing as mp
    sock = socket.socket(); sock.bind(("127.0.0.1", 0)); port = sock.getsockname()[1]; sock.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def _dyn_worker(rank, world, port, q):
    import time
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_txps, num_boot = 7, 11

        def run_one(g):
            time.sleep(0.05 if rank == 0 else 0.005)        # rank 0 is ten times slower: the others take over its share
            return np.full(n_txps, float(g))                # "replicate g" is a pure function of g, whoever runs it

        ids, res = odist.run_replicates_dynamic(run_one, num_boot)
        local = np.stack(res) if res else np.zeros((0, n_txps))
        out = odist.gather_replicates_by_id(ids, local, num_boot, rank, n_txps)
        ok = True
        if rank == 0:
            ok = out.shape == (num_boot, n_txps) and all(np.all(out[g] == g) for g in range(num_boot))
        # a second queue in the same process group starts from zero again
        ids2, _ = odist.run_replicates_dynamic(lambda g: g, 3)
        q.put((rank, bool(ok), len(ids), ids2))
    finally:
        dist.destroy_process_group()


def test_dynamic_replicate_queue_world2():
    """Replicates pulled from the shared counter: every id runs exactly once, the faster rank takes more of them."""
    import torch.multiprocessing as mp
    sock = socket.socket(); sock.bind(("127.0.0.1", 0)); port = sock.getsockname()[1]; sock.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dyn_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res)
    n0, n1 = res[0][2], res[1][2]
    assert n0 + n1 == 11 and n1 > n0                      # the fast rank did more
    assert sorted(res[0][3] + res[1][3]) == [0, 1, 2]
