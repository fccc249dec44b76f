"""Multi-GPU plumbing for the STA path: the pair batch is embarrassingly parallel (no cross-pair op anywhere in
sta_model.py:247-291), so rank r simply owns a contiguous block of pairs; weights are replicated with ONE broadcast
of the packed arena; there is no collective in steady state.  Only the tiny pose outputs are optionally gathered.

Works with any torch.distributed backend (nccl on the GPU box, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def pair_shard(total_pairs, rank, world):
    """Contiguous block partition [lo, hi) of `total_pairs` for `rank`; the first `total % world` ranks get one extra."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(int(total_pairs), world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def broadcast_arena(arena, src=0, group=None):
    """Broadcast the packed weight arena (uint8 tensor aliasing sta_weight_arena) from `src`."""
    dist.broadcast(arena, src=src, group=group)
    return arena


def gather_poses(pose, conf, total_pairs, group=None):
    """All-gather the per-rank relative poses (B_r,4,4) and confidences (B_r,) into global-order tensors.
    68 bytes per pair: negligible next to the pointmaps, which stay rank-local."""
    world = dist.get_world_size(group)
    sizes = [pair_shard(total_pairs, r, world) for r in range(world)]
    maxn = max(hi - lo for lo, hi in sizes)
    pad_pose = pose.new_zeros((maxn, 4, 4))
    pad_conf = conf.new_zeros((maxn,))
    pad_pose[: pose.shape[0]] = pose
    pad_conf[: conf.shape[0]] = conf
    poses = [torch.empty_like(pad_pose) for _ in range(world)]
    confs = [torch.empty_like(pad_conf) for _ in range(world)]
    dist.all_gather(poses, pad_pose, group=group)
    dist.all_gather(confs, pad_conf, group=group)
    out_pose = torch.cat([p[: hi - lo] for p, (lo, hi) in zip(poses, sizes)], dim=0)
    out_conf = torch.cat([c[: hi - lo] for c, (lo, hi) in zip(confs, sizes)], dim=0)
    return out_pose, out_conf
