"""Multi-GPU partitioning of the path: independent MPC instances are sharded contiguously over ranks; the only
collective is a gather/reduce of per-instance statistics."""
import numpy as np


def shard_range(total, rank, world):
    """Contiguous block [lo, hi) of instances owned by `rank`; blocks differ by at most one instance."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_instance_stats(cost, iters, status, group=None):
    """All-gather per-instance (cost, iterations, status) over the process group (NCCL on GPUs, gloo in the CPU
    tests). Inputs are 1-D numpy arrays of the local shard; returns global arrays ordered by rank."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return np.asarray(cost), np.asarray(iters), np.asarray(status)
    world = dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    local = torch.tensor(np.stack([cost, iters.astype(np.float64), status.astype(np.float64)], axis=1), dtype=torch.float64,
                         device=dev)
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    mx = int(max(int(s.item()) for s in sizes))
    pad = torch.zeros((mx, 3), dtype=torch.float64, device=dev)
    pad[:local.shape[0]] = local
    outs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    full = torch.cat([o[:int(s.item())] for o, s in zip(outs, sizes)]).cpu().numpy()
    return full[:, 0], full[:, 1].astype(np.int64), full[:, 2].astype(np.int64)
