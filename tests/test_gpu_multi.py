"""Two-rank NCCL run of the sharded search (one process per GPU): identical to the single-GPU result.  Skipped on boxes
with fewer than two GPUs (the world-size-2 gloo test in test_distributed_gloo.py covers the host logic on CPU)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model(dev):
    from deeplocalproteindocking_b200 import E3MultiResRepr4x4, GlobalDockingModel, SimpleFilter
    torch.manual_seed(0)
    rep = E3MultiResRepr4x4(multiplier=8)
    return GlobalDockingModel(rep, SimpleFilter(rep.get_num_outputs()), threshold_clash=5.0).to(dev)


def _run(rank, world, port, q, angle_inc):
    from deeplocalproteindocking_b200 import Docker, synthetic
    from deeplocalproteindocking_b200 import distributed as D
    dev = "cuda:%d" % rank
    torch.cuda.set_device(rank)
    if world > 1:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    rec, lig = synthetic.config_pair(20, seed=3)
    docker = Docker(_model(dev), angle_inc=angle_inc, box_size=20, resolution=1.25, max_conf=40, rot_batch=16,
                    clash_source="rotate", device=dev)
    docker.rot.R = docker.rot.R[:101]                       # a short, odd-sized rotation set
    best, top = D.dock_sharded(docker, rec, lig)
    if rank == 0:
        q.put((best.tolist(), top))          # plain lists: a tensor in the queue is shared through an fd the child may close first
    if world > 1:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_sharded_dock_equals_single_gpu():
    ctx = mp.get_context("spawn")
    out = []
    for world in (1, 2):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_run, args=(r, world, port, q, 20)) for r in range(world)]
        for p in procs:
            p.start()
        out.append(q.get(timeout=300))
        for p in procs:
            p.join(timeout=120)
            assert p.exitcode == 0
    (b1, t1), (b2, t2) = out
    assert b1 == b2
    assert t1 == t2
