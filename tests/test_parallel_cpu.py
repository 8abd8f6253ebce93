"""CPU, world_size 2 over gloo: scene sharding + final trajectory gather reproduce the single-process order."""
import os
import socket
import subprocess
import sys

from trafficbotsv1_5_b200 import parallel


def test_shard_range_partitions():
    for n in (1, 2, 5, 16, 513):
        for w in (1, 2, 4, 8):
            rs = [parallel.shard_range(n, w, r) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rs[:-1], rs[1:]))
            sizes = [hi - lo for lo, hi in rs]
            assert max(sizes) - min(sizes) <= 1


def test_gather_world2_gloo():
    """torchrun-style launch (as bench.py is launched on the box), 2 ranks, gloo backend, 127.0.0.1 rendezvous."""
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_gloo_gather_worker.py")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), worker],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "GLOO_GATHER_OK" in out.stdout
