"""Multi-GPU plumbing: self-play shards by environment, one process per GPU, no collective
inside a search.  torch.distributed (NCCL over NVLink on GPUs; gloo in the CPU tests) is used only for
(a) the parameter / hash-bitset broadcast after a learner update and (b) the all-gather of compact
trajectories into the replay buffer, plus (c) the OR-merge of the per-rank hash bitsets after a learner update."""
from __future__ import annotations


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous split of `total` envs; the first `total % world` ranks get one extra."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def broadcast_params(net, src: int = 0) -> None:
    """device_put_replicated equivalent (main.py:212): every leaf of the FC params + the hash bitset."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    for h in range(4):
        for l in range(3):
            dist.broadcast(net.w[h][l], src)
            dist.broadcast(net.b[h][l], src)
    dist.broadcast(net.binary_set, src)


def pack_trajectory(action, reward, terminated, compact_state):
    """[B,4] int32 rows: action, reward bits, terminated, first 4 bytes of the compact state (DeepSea: the whole state)."""
    import torch

    cs = compact_state.contiguous().view(torch.uint8).reshape(compact_state.shape[0], -1)[:, :4].contiguous().view(torch.int32).reshape(-1)
    return torch.stack([action.to(torch.int32), reward.reshape(-1).contiguous().view(torch.int32), terminated.to(torch.int32), cs], 1).contiguous()


def all_gather_trajectory(traj, out=None):
    """Concatenate every rank's [B_r,4] trajectory block in rank order (the pmap output concatenation of
    main.py:383-385).  Ranks may hold different B_r (uneven shards)."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return traj
    world = dist.get_world_size()
    sizes = [torch.zeros(1, dtype=torch.int64, device=traj.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([traj.shape[0]], dtype=torch.int64, device=traj.device))
    sizes = [int(s.item()) for s in sizes]
    if len(set(sizes)) == 1:
        out = out if out is not None else torch.empty((world * sizes[0], traj.shape[1]), dtype=traj.dtype, device=traj.device)
        dist.all_gather_into_tensor(out.view(-1), traj.view(-1))
        return out
    m = max(sizes)  # uneven shards: pad to the largest block, gather, then drop the padding
    padded = torch.zeros((m, traj.shape[1]), dtype=traj.dtype, device=traj.device)
    padded[: traj.shape[0]] = traj
    buf = torch.empty((world * m, traj.shape[1]), dtype=traj.dtype, device=traj.device)
    dist.all_gather_into_tensor(buf.view(-1), padded.view(-1))
    return torch.cat([buf[r * m : r * m + s] for r, s in enumerate(sizes)], 0)


def merge_hash_sets(binary_set):
    """OR-merge of every rank's hash bitset, in place.  BaseHash.update (hashes.py:45-50) sets bits for
    the observations a device trained on; under the reference's pmap each device's `binary_set` then diverges and only
    device 0's copy is ever read back (train.py:90-96, main.py:467-469).  Here every rank ends with the union.  NCCL has no
    bitwise-OR reduction, so the 2 MiB sets are all-gathered (world x 2 MiB over NVLink) and OR-ed locally."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return binary_set
    world = dist.get_world_size()
    buf = torch.empty((world, binary_set.numel()), dtype=binary_set.dtype, device=binary_set.device)
    dist.all_gather_into_tensor(buf.view(-1), binary_set.contiguous().view(-1))
    merged = buf[0]
    for r in range(1, world):
        merged = torch.bitwise_or(merged, buf[r])
    binary_set.view(-1).copy_(merged)
    return binary_set
