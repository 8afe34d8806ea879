"""CPU: the N>1 host logic (channel sharding, scatter from rank 0, gather of rows) with gloo, world 2."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

from supersdr_b200 import sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_channel_shard_partitions():
    for total in (1, 7, 8, 4096, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            sh = sharding.all_shards(total, world)
            assert sh[0][0] == 0 and sum(c for _, c in sh) == total
            assert all(sh[i][0] + sh[i][1] == sh[i + 1][0] for i in range(world - 1))
            assert max(c for _, c in sh) - min(c for _, c in sh) <= 1
    assert sharding.channel_shard(65536, 3, 8) == (24576, 8192)      # BASELINE config 4: 8192 ch/GPU


def test_scatter_gather_world2_gloo(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent("""
        import os, sys
        sys.path.insert(0, %r)
        import numpy as np, torch, torch.distributed as dist
        from supersdr_b200 import sharding
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        total, n, N = 5, 2, 64
        root = torch.arange(total * n * N * 2, dtype=torch.float32).reshape(total, n, N, 2) if rank == 0 else None
        mine = sharding.scatter_from_root(root, total, (n, N, 2), torch.float32)
        first, count = sharding.channel_shard(total, rank, world)
        want = torch.arange(total * n * N * 2, dtype=torch.float32).reshape(total, n, N, 2)[first:first + count]
        assert torch.equal(mine, want)
        rows = mine.sum(dim=(1, 3))                       # stand-in for the per-channel kernel
        full = sharding.gather_rows_to_root(rows, total)
        if rank == 0:
            assert torch.equal(full, torch.arange(total * n * N * 2, dtype=torch.float32).reshape(total, n, N, 2).sum(dim=(1, 3)))
            print("OK")
        dist.destroy_process_group()
    """ % ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "OK" in outs[0]
