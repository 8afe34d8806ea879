"""Multi-GPU plumbing: channels are fully independent, so a batch shards across ranks as
contiguous channel blocks with NO collective in the math.  One process per GPU.  Everything here goes through the C ABI
(``ssdr_nccl_*`` = grouped ncclSend/ncclRecv over NVLink, ``ssdr_ipc_*`` = CUDA IPC peer mapping) -- no PyTorch:

* ``channel_shard`` / ``all_shards``  -- which channels a rank owns;
* ``rendezvous``                      -- hands rank 0's 128-byte NCCL id to every rank over a TCP socket
                                         (MASTER_ADDR / MASTER_PORT of the launcher, next port up);
* ``Comm``                            -- scatter of an input batch from a root rank's HBM to the owning ranks, gather of the
                                         pixel rows back, barrier, max over ranks (timing);
* ``export_device_buffer`` / ``open_peer_buffer`` -- the root-ingest deployment WITHOUT a scatter: every rank's kernel
                                         reads its shard in place from the root's HBM (peer loads over NVLink).
The reference is a one-receiver client and has no counterpart (utils_supersdr.py opens one W/F and one SND socket)."""
import ctypes as C
import os
import socket
import struct
import time

from . import _lib

NCCL_ID_BYTES = 128
_MAGIC = b"SSDRNCCL"


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


def shard_bytes(total_channels, bytes_per_channel, world):
    """(offsets, counts) in bytes of every rank's shard inside a [total_channels][bytes_per_channel] buffer."""
    sh = all_shards(total_channels, world)
    return [f * bytes_per_channel for f, _ in sh], [c * bytes_per_channel for _, c in sh]


# ---------------------------------------------------------------------------------------------------
# rendezvous: rank 0 serves a small payload (the NCCL unique id) to the other ranks
# ---------------------------------------------------------------------------------------------------
def _recv_exact(conn, n):
    buf = b""
    while len(buf) < n:
        part = conn.recv(n - len(buf))
        if not part:
            raise ConnectionError("peer closed during rendezvous")
        buf += part
    return buf


def rendezvous(payload, rank, world, addr=None, port=None, timeout=120.0, tries=16):
    """Every rank returns rank 0's ``payload`` (bytes).  Rank 0 listens on ``addr:port`` (default MASTER_ADDR and
    MASTER_PORT + 1: torchrun's own store owns MASTER_PORT), trying up to ``tries`` consecutive ports; the other ranks
    scan the same ports and accept only an answer that carries this job's cookie (world size + base port), so a
    foreign listener on one of the ports is skipped."""
    addr = addr or os.environ.get("MASTER_ADDR", "127.0.0.1")
    port = int(port if port is not None else int(os.environ.get("MASTER_PORT", "29500")) + 1)
    cookie = _MAGIC + struct.pack("<II", int(world), port)
    if world == 1:
        return bytes(payload)
    if rank == 0:
        srv = None
        for k in range(tries):
            try:
                srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
                srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
                srv.bind((addr, port + k))
                break
            except OSError:
                srv.close()
                srv = None
        if srv is None:
            raise OSError("rendezvous: no free port in %d..%d" % (port, port + tries - 1))
        srv.listen(world)
        srv.settimeout(timeout)
        served = 0
        msg = cookie + struct.pack("<I", len(payload)) + bytes(payload)
        try:
            while served < world - 1:
                conn, _ = srv.accept()
                with conn:
                    conn.settimeout(timeout)
                    if _recv_exact(conn, len(cookie)) == cookie:
                        conn.sendall(msg)
                        served += 1
        finally:
            srv.close()
        return bytes(payload)
    deadline = time.time() + timeout
    while time.time() < deadline:
        for k in range(tries):
            try:
                with socket.create_connection((addr, port + k), timeout=2.0) as conn:
                    conn.settimeout(10.0)
                    conn.sendall(cookie)
                    if _recv_exact(conn, len(cookie)) != cookie:
                        continue
                    (n,) = struct.unpack("<I", _recv_exact(conn, 4))
                    return _recv_exact(conn, n)
            except (OSError, ConnectionError):
                continue
        time.sleep(0.05)
    raise TimeoutError("rendezvous with rank 0 at %s:%d.. timed out" % (addr, port))


# ---------------------------------------------------------------------------------------------------
# NCCL communicator over the C ABI
# ---------------------------------------------------------------------------------------------------
class Comm:
    """One NCCL communicator per process (one process per GPU).  ``Comm.from_env()`` reads RANK / WORLD_SIZE /
    MASTER_ADDR / MASTER_PORT as torchrun sets them."""

    def __init__(self, rank, world, addr=None, port=None):
        _lib.init()
        if not _lib.lib.ssdr_nccl_available():
            raise _lib.SsdrError("NCCL is not available: " + _lib.last_error())
        self.rank, self.world = int(rank), int(world)
        ident = (C.c_ubyte * NCCL_ID_BYTES)()
        if self.rank == 0:
            _lib.check(_lib.lib.ssdr_nccl_unique_id(ident))
        blob = rendezvous(bytes(ident), self.rank, self.world, addr, port)
        ident = (C.c_ubyte * NCCL_ID_BYTES).from_buffer_copy(blob)
        h = C.c_void_p()
        _lib.check(_lib.lib.ssdr_nccl_init(C.byref(h), ident, self.rank, self.world))
        self._h = h

    @classmethod
    def from_env(cls):
        return cls(int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))

    def _arrays(self, total_channels, bytes_per_channel):
        off, cnt = shard_bytes(total_channels, bytes_per_channel, self.world)
        A = C.c_size_t * self.world
        return A(*off), A(*cnt), cnt[self.rank]

    def scatter_from_root(self, root_dev_ptr, total_channels, bytes_per_channel, recv_dev_ptr, root=0, sync=True):
        """Rank ``root`` holds [total_channels][bytes_per_channel] in its HBM; every rank receives its contiguous channel
        block into ``recv_dev_ptr`` (grouped ncclSend/ncclRecv over NVLink).  Returns this rank's byte count."""
        off, cnt, mine = self._arrays(total_channels, bytes_per_channel)
        _lib.check(_lib.lib.ssdr_nccl_scatter(self._h, C.c_void_p(root_dev_ptr or 0), off, cnt, C.c_void_p(recv_dev_ptr or 0), int(root)))
        if sync:
            self.sync()
        return mine

    def gather_rows_to_root(self, send_dev_ptr, total_channels, bytes_per_channel, root_dev_ptr, root=0, sync=True):
        """The reverse for the per-channel result rows (pixels): rank r's rows land at its channel offset of the
        root's [total_channels][bytes_per_channel] buffer."""
        off, cnt, mine = self._arrays(total_channels, bytes_per_channel)
        _lib.check(_lib.lib.ssdr_nccl_gather(self._h, C.c_void_p(send_dev_ptr or 0), C.c_void_p(root_dev_ptr or 0), off, cnt, int(root)))
        if sync:
            self.sync()
        return mine

    def max_over_ranks(self, value):
        v = C.c_double(float(value))
        _lib.check(_lib.lib.ssdr_nccl_allreduce_max_f64(self._h, C.byref(v)))
        return v.value

    def barrier(self):
        _lib.check(_lib.lib.ssdr_nccl_barrier(self._h))

    def sync(self):
        _lib.check(_lib.lib.ssdr_nccl_sync(self._h))

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib.ssdr_nccl_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------------
# CUDA IPC peer ingest
# ---------------------------------------------------------------------------------------------------
def export_device_buffer(dev_ptr):
    """64-byte IPC handle of a device allocation of this process (ssdr_ipc_export): send it to the other ranks (e.g.
    with ``rendezvous``) so that their kernels can read the buffer in place over NVLink."""
    h = (C.c_ubyte * 64)()
    _lib.check(_lib.lib.ssdr_ipc_export(C.c_void_p(dev_ptr), h))
    return bytes(h)


def open_peer_buffer(handle64):
    """Map a peer rank's exported allocation; returns the device pointer valid in this process (close with
    ``close_peer_buffer``)."""
    p = C.c_void_p()
    buf = (C.c_ubyte * 64).from_buffer_copy(handle64)
    _lib.check(_lib.lib.ssdr_ipc_open(buf, C.byref(p)))
    return p.value


def close_peer_buffer(dev_ptr):
    _lib.check(_lib.lib.ssdr_ipc_close(C.c_void_p(dev_ptr)))
