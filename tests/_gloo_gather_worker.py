"""Worker for tests/test_parallel_cpu.py, launched by torch.distributed.run (2 ranks, gloo, CPU)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tbpkg  # noqa: E402,F401
from trafficbotsv1_5_b200 import parallel  # noqa: E402


def fake_rollout(scene_ids, R=3, A=4, T=5):
    g = torch.arange(R * A * T * 3, dtype=torch.float32).view(1, R, A, T, 3)
    return scene_ids.view(-1, 1, 1, 1, 1).float() * 1000.0 + g


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    for n_scenes in (4, 5, 1):  # even, uneven, fewer scenes than ranks
        ids = torch.arange(n_scenes)
        lo, hi = parallel.shard_range(n_scenes, world, rank)
        full = parallel.gather_scenes(fake_rollout(ids[lo:hi]), n_scenes)
        assert torch.equal(full, fake_rollout(ids)), (rank, n_scenes)
        batch = {"a": torch.arange(n_scenes * 2).view(n_scenes, 2)}
        assert parallel.shard_batch(batch, world, rank)["a"].shape[0] == hi - lo
    # batched sweep: every rank walks its shard in batches, each batch is gathered (OverlappedGather; synchronous on
    # CPU) and rank 0 verifies the gathered bytes against per-scene checksums gathered separately
    n_scenes, per_batch = 8, 2
    lo, hi = parallel.shard_range(n_scenes, world, rank)
    n_batches = (hi - lo) // per_batch
    og = parallel.OverlappedGather(n_batches, (per_batch, 3, 4, 5, 3), torch.float32, "cpu")
    sums = []
    for b in range(n_batches):
        ids = torch.arange(lo + b * per_batch, lo + (b + 1) * per_batch)
        local = fake_rollout(ids)
        sums.append(parallel.scene_checksums(local))
        og.submit(local)
    store = og.wait()
    all_sums = [torch.empty(n_batches, per_batch, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(all_sums, torch.stack(sums))
    for r in range(world):
        rlo, _ = parallel.shard_range(n_scenes, world, r)
        for b in range(n_batches):
            ids = torch.arange(rlo + b * per_batch, rlo + (b + 1) * per_batch)
            assert torch.equal(store[b, r], fake_rollout(ids)), (rank, r, b)
            assert torch.equal(parallel.scene_checksums(store[b, r]), all_sums[r][b])
    assert not torch.equal(parallel.scene_checksums(store[0, 0]), parallel.scene_checksums(store[0, 0].flip(1)))
    # data-parallel training: gradients accumulate into one flat bucket through autograd, one all-reduce averages them
    torch.manual_seed(0)
    params = {"b.w": torch.randn(3, 4, requires_grad=True), "a.bias": torch.randn(5, requires_grad=True),
              "unused": torch.randn(2, requires_grad=True)}
    bucket = parallel.GradBucket(params)
    x = torch.full((4,), float(rank + 1))
    loss = (params["b.w"] @ x).sum() + (params["a.bias"] * (rank + 1)).sum()
    loss.backward()
    bucket.all_reduce()
    mean = sum(range(1, world + 1)) / world
    assert torch.allclose(params["b.w"].grad, torch.full((3, 4), mean)) and params["b.w"].grad.data_ptr() == bucket.views["b.w"].data_ptr()
    assert torch.allclose(params["a.bias"].grad, torch.full((5,), mean))
    assert float(params["unused"].grad.abs().max()) == 0.0
    bucket.zero()
    assert float(bucket.flat.abs().max()) == 0.0
    dist.barrier()
    if rank == 0:
        print("GLOO_GATHER_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
