"""Multi-GPU plumbing: channels are fully independent, so a batch shards across ranks as
contiguous channel blocks with NO collective in the math.  One process per GPU; torch.distributed is
used only as plumbing (rendezvous, optional NCCL scatter of an input batch from rank 0 over NVLink,
optional gather of the small pixel rows)."""
import numpy as np


def channel_shard(total, rank, world):
    """Contiguous block of channels owned by ``rank``: (first, count).  Remainders go to the first
    ranks, so counts differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    base, rem = divmod(int(total), int(world))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def all_shards(total, world):
    return [channel_shard(total, r, world) for r in range(world)]


def scatter_from_root(root_tensor, total_channels, per_channel_shape, dtype, group=None, device=None):
    """Scatter a [total_channels, *per_channel_shape] batch held by rank 0 to all ranks (NCCL over
    NVLink when the tensors are CUDA tensors; gloo on CPU in the tests).  Returns this rank's shard as
    a torch tensor; pass ``.data_ptr()`` to the ``*_dev`` entry points."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    shards = all_shards(total_channels, world)
    first, count = shards[rank]
    out = torch.empty((count,) + tuple(per_channel_shape), dtype=dtype, device=device)
    if rank == 0:
        ops = []
        for r, (f, c) in enumerate(shards):
            if r == 0:
                out.copy_(root_tensor[f:f + c])
            elif c:
                ops.append(dist.P2POp(dist.isend, root_tensor[f:f + c].contiguous(), r, group))
        reqs = dist.batch_isend_irecv(ops) if ops else []
    else:
        reqs = dist.batch_isend_irecv([dist.P2POp(dist.irecv, out, 0, group)]) if count else []
    for q in reqs:
        q.wait()
    return out


def gather_rows_to_root(local_rows, total_channels, group=None):
    """Gather per-rank [count, W] rows (numpy or torch) into [total_channels, W] on rank 0."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    t = local_rows if isinstance(local_rows, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(local_rows))
    shards = all_shards(total_channels, world)
    if rank == 0:
        full = torch.empty((total_channels,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        f, c = shards[0]
        full[f:f + c] = t
        ops = [dist.P2POp(dist.irecv, full[f:f + c], r, group) for r, (f, c) in enumerate(shards) if r and c]
        for q in (dist.batch_isend_irecv(ops) if ops else []):
            q.wait()
        return full
    if shards[rank][1]:
        for q in dist.batch_isend_irecv([dist.P2POp(dist.isend, t.contiguous(), 0, group)]):
            q.wait()
    return None


def export_device_buffer(dev_ptr):
    """64-byte IPC handle of a device allocation of this process (ssdr_ipc_export): send it to the other ranks (e.g.
    ``torch.distributed.broadcast_object_list``) so that their kernels can read the buffer in place over NVLink."""
    import ctypes as C
    from . import _lib
    h = (C.c_ubyte * 64)()
    _lib.check(_lib.lib.ssdr_ipc_export(C.c_void_p(dev_ptr), h))
    return bytes(h)


def open_peer_buffer(handle64):
    """Map a peer rank's exported allocation; returns the device pointer valid in this process (close with
    ``close_peer_buffer``)."""
    import ctypes as C
    from . import _lib
    p = C.c_void_p()
    buf = (C.c_ubyte * 64).from_buffer_copy(handle64)
    _lib.check(_lib.lib.ssdr_ipc_open(buf, C.byref(p)))
    return p.value


def close_peer_buffer(dev_ptr):
    import ctypes as C
    from . import _lib
    _lib.check(_lib.lib.ssdr_ipc_close(C.c_void_p(dev_ptr)))
