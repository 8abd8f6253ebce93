"""Multi-GPU plumbing: scenes x rollouts shard with NO data-path collective; the only exchange is the final gather of
trajectories / validity (reference: torchmetrics `dist_reduce_fx="cat"` states all-gathered at the end of each test
step, src/utils/submission.py:45-46,169-170 and src/pl_modules/waymo_motion.py:894-909). One process per GPU,
`torch.distributed` (NCCL over NVLink on the box, gloo in the CPU tests)."""
from typing import Dict, List, Tuple

import torch
import torch.distributed as dist
from torch import Tensor


def shard_range(n_scenes: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block partition of scenes (keeps the 32 rollouts of a scene, and therefore its map / traffic-light
    K/V tables, on one GPU). Ranks [0, n % world) get one extra scene."""
    base, rem = divmod(n_scenes, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(batch: Dict[str, Tensor], world: int, rank: int) -> Dict[str, Tensor]:
    n = next(iter(batch.values())).shape[0]
    lo, hi = shard_range(n, world, rank)
    return {k: v[lo:hi] for k, v in batch.items()}


def gather_scenes(local: Tensor, n_scenes: int) -> Tensor:
    """All-gather per-scene results [n_local, ...] -> [n_scenes, ...] in global scene order (uneven shards are padded
    to the largest shard for the collective and trimmed afterwards)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_range(n_scenes, world, r)[1] - shard_range(n_scenes, world, r)[0] for r in range(world)]
    m = max(sizes)
    pad = local
    if local.shape[0] < m:
        pad = torch.cat([local, local.new_zeros((m - local.shape[0],) + tuple(local.shape[1:]))], 0)
    out: List[Tensor] = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad.contiguous())
    return torch.cat([o[:s] for o, s in zip(out, sizes)], 0)
