"""CPU: the N>1 host logic -- channel sharding, the torch-free rendezvous (world 2, two processes), and the shard byte
ranges checked against a world-size-2 gloo scatter/gather of the same batch (torch only here, in the test)."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from supersdr_b200 import sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_world2(tmp_path, body, port):
    script = tmp_path / "w.py"
    script.write_text("import os, sys\nsys.path.insert(0, %r)\n" % ROOT + textwrap.dedent(body))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    return outs


def test_channel_shard_partitions():
    for total in (1, 7, 8, 4096, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            sh = sharding.all_shards(total, world)
            assert sh[0][0] == 0 and sum(c for _, c in sh) == total
            assert all(sh[i][0] + sh[i][1] == sh[i + 1][0] for i in range(world - 1))
            assert max(c for _, c in sh) - min(c for _, c in sh) <= 1
    assert sharding.channel_shard(65536, 3, 8) == (24576, 8192)      # BASELINE config 4: 8192 ch/GPU
    off, cnt = sharding.shard_bytes(5, 1000, 2)
    assert off == [0, 3000] and cnt == [3000, 2000]
    with pytest.raises(ValueError):
        sharding.channel_shard(8, 2, 2)


def test_product_package_has_no_torch():
    """north_star: no PyTorch in the framework -- the package's sources never import torch."""
    import re
    pkg = os.path.join(ROOT, "supersdr_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            assert not re.search(r"^\s*(import|from)\s+torch", open(os.path.join(pkg, f)).read(), flags=re.M), f


def test_rendezvous_world2(tmp_path):
    """Rank 0's payload (the 128-byte NCCL id in production) reaches rank 1 over the socket rendezvous, twice in a row
    (a second communicator), with a foreign listener squatting on the first port."""
    port = _free_port()
    squat = socket.socket()
    squat.bind(("127.0.0.1", port + 1))          # the default rendezvous port (MASTER_PORT + 1) is taken
    squat.listen(1)
    try:
        outs = _run_world2(tmp_path, """
            from supersdr_b200 import sharding
            rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
            for rnd in range(2):
                payload = bytes(range(128)) if rank == 0 else b""
                if rnd:
                    payload = payload[::-1]
                got = sharding.rendezvous(payload, rank, world, timeout=60)
                want = bytes(range(128))[::-1] if rnd else bytes(range(128))
                assert got == want, (rank, rnd)
            print("OK", rank)
        """, port)
    finally:
        squat.close()
    assert "OK 0" in outs[0] and "OK 1" in outs[1]


def test_shard_ranges_match_gloo_scatter_gather_world2(tmp_path):
    """The byte ranges ssdr_nccl_scatter / ssdr_nccl_gather are given (sharding.shard_bytes) are exactly the blocks a
    reference scatter / gather of the same [channels][n_avg][nfft] batch moves -- world size 2, gloo, CPU."""
    outs = _run_world2(tmp_path, """
        import numpy as np, torch, torch.distributed as dist
        from supersdr_b200 import sharding
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        total, n, N = 5, 2, 64
        per = n * N * 8                                           # bytes per channel (complex64)
        full = torch.arange(total * n * N * 2, dtype=torch.float32).reshape(total, n * N * 2)
        off, cnt = sharding.shard_bytes(total, per, world)
        first, count = sharding.channel_shard(total, rank, world)
        assert off[rank] == first * per and cnt[rank] == count * per
        # scatter: root sends each rank the block [off, off + cnt) of its flat byte buffer
        mine = torch.empty(count, n * N * 2)
        if rank == 0:
            flat = full.numpy().view(np.uint8).reshape(-1)
            for r in range(1, world):
                blk = torch.from_numpy(flat[off[r]:off[r] + cnt[r]].copy())
                dist.send(blk, r)
            mine = torch.from_numpy(flat[off[0]:off[0] + cnt[0]].copy().view(np.float32).reshape(count, -1))
        else:
            buf = torch.empty(cnt[rank], dtype=torch.uint8)
            dist.recv(buf, 0)
            mine = torch.from_numpy(buf.numpy().view(np.float32).reshape(count, -1))
        assert torch.equal(mine, full[first:first + count])
        # gather of per-channel rows (stand-in for the pixel rows, one row of N bytes per channel)
        rows = (mine.reshape(count, n, N, 2).sum(dim=(1, 3)) % 251).to(torch.uint8)
        roff, rcnt = sharding.shard_bytes(total, N, world)
        if rank == 0:
            out = torch.zeros(total * N, dtype=torch.uint8)
            out[roff[0]:roff[0] + rcnt[0]] = rows.reshape(-1)
            for r in range(1, world):
                buf = torch.empty(rcnt[r], dtype=torch.uint8)
                dist.recv(buf, r)
                out[roff[r]:roff[r] + rcnt[r]] = buf
            want = (full.reshape(total, n, N, 2).sum(dim=(1, 3)) % 251).to(torch.uint8)
            assert torch.equal(out.reshape(total, N), want)
            print("OK")
        else:
            dist.send(rows.reshape(-1).contiguous(), 0)
        dist.destroy_process_group()
    """, _free_port())
    assert "OK" in outs[0]
