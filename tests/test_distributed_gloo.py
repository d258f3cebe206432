"""World-size-2 gloo test of the rotation sharding / final gather (the N>1 path of bench.py and Docker)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deeplocalproteindocking_b200 import distributed as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_rot, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = D.shard_indices(n_rot, rank, world)
    # fake "search": score depends only on the rotation index
    local = torch.stack([-(idx.float() % 7), idx.float(), idx.float() * 2, idx.float() * 3], dim=1)
    full = D.gather_best(local, n_rot, rank, world)
    top = D.merge_top_lists([(int(i), 1, 2, 3, float(-(i % 7))) for i in idx.tolist()], K=5)
    if rank == 0:
        q.put((full.tolist(), top))
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_process():
    n_rot, world = 11, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_rot, q)) for r in range(world)]
    for p in procs:
        p.start()
    full, top = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    i = torch.arange(n_rot).float()
    want = torch.stack([-(i % 7), i, i * 2, i * 3], dim=1)
    assert full == want.tolist()
    single = sorted([(int(k), 1, 2, 3, float(-(k % 7))) for k in range(n_rot)], key=lambda t: t[4])[:5]
    assert top == single


def test_shards_partition_the_rotation_set():
    for n, w in ((1854, 8), (29025, 8), (5, 8), (14868, 4)):
        parts = [D.shard_indices(n, r, w) for r in range(w)]
        assert sorted(torch.cat(parts).tolist()) == list(range(n))
        assert max(p.numel() for p in parts) == D.padded_shard_size(n, w)
