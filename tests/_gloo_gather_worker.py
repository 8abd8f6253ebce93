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
    dist.barrier()
    if rank == 0:
        print("GLOO_GATHER_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
