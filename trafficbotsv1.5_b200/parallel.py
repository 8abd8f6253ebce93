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


def scene_checksums(t: Tensor) -> Tensor:
    """Order-sensitive per-scene checksum of a result tensor [n, ...] (fp32 / bool / int): the raw 32-bit words weighted
    by their position, summed in int64. Used to verify that a gather put every scene's bytes in the right place."""
    x = t.contiguous()
    if x.dtype == torch.bool:
        x = x.to(torch.int32)
    w = x.view(torch.int32).reshape(x.shape[0], -1).to(torch.int64)
    pos = torch.arange(1, w.shape[1] + 1, device=w.device, dtype=torch.int64)
    return (w * (pos % 65521)).sum(1)


class GradBucket:
    """Data-parallel training (the reference trains with Lightning DDP, configs/trainer/default.yaml): every rank
    evaluates `training.TrainStep` on its own scenes; the ONE exchange of a step is the average of the gradients. All
    gradients live in one flat fp32 buffer (each `param.grad` is a view of it, so autograd accumulates straight into
    the bucket) and are reduced with a single all-reduce (NCCL over NVLink on the box: 32 MB for the 8 M parameters of
    the training configuration; gloo in the CPU tests)."""

    def __init__(self, params: Dict[str, Tensor]):
        self.params = params
        self.names = sorted(params)
        n = sum(params[k].numel() for k in self.names)
        any_p = params[self.names[0]]
        self.flat = torch.zeros(n, dtype=any_p.dtype, device=any_p.device)
        self.views, off = {}, 0
        for k in self.names:
            p = params[k]
            self.views[k] = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.zero()

    def zero(self) -> None:
        """Zero the bucket and (re-)attach its views as the parameters' .grad."""
        self.flat.zero_()
        for k in self.names:
            self.params[k].grad = self.views[k]

    def all_reduce(self) -> None:
        """Average the gradients over the ranks (no-op in a single process). Gradients autograd created outside the
        bucket (a parameter whose .grad was None at backward time) are folded in first."""
        for k in self.names:
            g = self.params[k].grad
            if g is not None and g.data_ptr() != self.views[k].data_ptr():
                self.views[k].copy_(g)
                self.params[k].grad = self.views[k]
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(dist.get_world_size())


class OverlappedGather:
    """All-gather of per-batch results on a side stream, overlapped with the next batch's compute (SURVEY 5 last row;
    the reference gathers its `cat` metric states after every test step, submission.py:45-46,169-170). `submit` copies
    the local result out of the engine's buffers (they are overwritten by the next rollout) into one of two staging
    buffers on the current stream, then the side stream all-gathers it into slot `i` of the result store
    [n_batches, world, ...]. On CPU / gloo (tests) the same calls run synchronously."""

    def __init__(self, n_batches: int, local_shape, dtype, device, n_stage: int = 2):
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.cuda = torch.device(device).type == "cuda"
        self.store = torch.empty((n_batches, self.world) + tuple(local_shape), dtype=dtype, device=device)
        self.stage = [torch.empty(tuple(local_shape), dtype=dtype, device=device) for _ in range(n_stage)]
        self.side = torch.cuda.Stream(device=device) if self.cuda else None
        self.free = [None] * n_stage  # event: the gather that last read this staging buffer has finished
        self.n = 0

    def submit(self, local: Tensor) -> None:
        i, slot = self.n, self.n % len(self.stage)
        self.n += 1
        buf = self.stage[slot]
        if self.cuda:
            cur = torch.cuda.current_stream()
            if self.free[slot] is not None:
                cur.wait_event(self.free[slot])
            buf.copy_(local)
            ready = torch.cuda.Event()
            ready.record(cur)
            with torch.cuda.stream(self.side):
                self.side.wait_event(ready)
                self._gather(i, buf)
                self.free[slot] = torch.cuda.Event()
                self.free[slot].record(self.side)
        else:
            buf.copy_(local)
            self._gather(i, buf)

    def _gather(self, i: int, buf: Tensor) -> None:
        if self.world == 1:
            self.store[i, 0].copy_(buf)
        else:
            dist.all_gather_into_tensor(self.store[i].view(-1), buf.view(-1))

    def wait(self) -> Tensor:
        """Join the side stream; returns the store [n_batches, world, ...] (global scene g of a contiguous-block shard
        lives at [batch of g within its rank, rank of g])."""
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.side)
        self.n = 0
        return self.store
