"""Multi-GPU sharding of the rotation set: one process per GPU, rank r scores rotations
{i : i mod W = r}; the only exchange is one all_gather of the per-rotation best poses at the end.

The reference has no in-process parallelism (its 5-node run split *targets* across jobs, README.md:71), so there is
nothing to mirror here beyond keeping the result identical to the single-GPU one.
Works with any torch.distributed backend (nccl on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_indices(n_rot: int, rank: int, world: int) -> torch.Tensor:
    """Interleaved shard: uniform cost per rank and file order preserved within a rank."""
    if rank >= n_rot:
        return torch.empty(0, dtype=torch.int64)
    return torch.arange(rank, n_rot, world)


def padded_shard_size(n_rot: int, world: int) -> int:
    return (n_rot + world - 1) // world


def gather_best(local_best: torch.Tensor, n_rot: int, rank: int, world: int, group=None) -> torch.Tensor:
    """local_best [len(shard),4] (score,x,y,z) -> [n_rot,4] on every rank, rows in rotation order.

    Shards are padded to ceil(n/W) rows with score=+inf so that one all_gather_into_tensor suffices."""
    if world == 1:
        return local_best
    m = padded_shard_size(n_rot, world)
    buf = torch.full((m, 4), float("inf"), device=local_best.device, dtype=local_best.dtype)
    buf[:local_best.shape[0]] = local_best
    out = torch.empty(world * m, 4, device=local_best.device, dtype=local_best.dtype)
    dist.all_gather_into_tensor(out, buf, group=group)
    # rank r, local row j  ->  rotation j*W + r
    full = out.reshape(world, m, 4).transpose(0, 1).reshape(world * m, 4)
    return full[:n_rot].contiguous()


def merge_top_lists(local_top: Sequence[Tuple[int, int, int, int, float]], K: int, group=None) -> List[tuple]:
    """All-gather every rank's top list and reproduce the reference ordering (Docker.py:100-105): ascending score,
    ties in rotation-file order (the order a single process would have appended them in).

    One fixed-size tensor collective (`all_gather_into_tensor` of [K,5] float64 rows (rot, x, y, z, score), +inf padded) --
    no pickling, same call on nccl and gloo."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return list(local_top)[:K]
    world = dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    buf = torch.full((K, 5), float("inf"), dtype=torch.float64)
    rows = sorted(local_top, key=lambda t: (t[4], t[0]))[:K]          # a rank can contribute at most K of the global K
    if rows:
        buf[:len(rows)] = torch.tensor(rows, dtype=torch.float64)
    buf = buf.to(dev)
    out = torch.empty(world * K, 5, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, buf, group=group)
    out = out.cpu()
    merged = [(int(r[0]), int(r[1]), int(r[2]), int(r[3]), float(r[4])) for r in out.tolist() if r[0] != float("inf")]
    merged.sort(key=lambda t: (t[4], t[0]))
    return merged[:K]


def dock_sharded(docker, ureceptor, uligand, group=None):
    """`Docker.dockSE3` with the rotation set sharded over the ranks of `group` (one process per GPU): every rank scores
    rotations {i : i mod W = rank}, then one all_gather of the per-rotation best poses and one of the per-rank top lists
    reproduce the single-process result; rank 0 writes the pose file (Docker.py:107-133).  Returns (best [N_rot,4], top)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        docker.dockSE3(ureceptor, uligand)
        return docker.best_per_rotation, docker.top_list
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n_rot = len(docker.rot)
    docker.dockSE3(ureceptor, uligand, rotation_subset=shard_indices(n_rot, rank, world), write=False)
    local = docker.best_per_rotation.to(docker.device)
    best = gather_best(local, n_rot, rank, world, group=group).cpu()
    docker.best_per_rotation = best
    if docker.max_conf > 1:
        docker.top_list = merge_top_lists(docker.top_list, docker.max_conf, group=group)
    else:
        items = [(i, int(best[i, 1]), int(best[i, 2]), int(best[i, 3]), float(best[i, 0])) for i in range(n_rot)]
        items.sort(key=lambda t: t[4])
        docker.top_list = items[:docker.max_conf]
    if rank == 0:
        docker.write_conformations()
    return best, docker.top_list
