"""Minimal TCP rendezvous for one-process-per-GPU runs (torchrun-style env: RANK, WORLD_SIZE, MASTER_ADDR,
MASTER_PORT): rank 0 creates the NCCL unique id and hands it to the other ranks.  Nothing else travels over
this channel -- all data-path communication is the single ncclAllGather inside libmogp_b200."""
import os
import socket
import struct
import time

_MAGIC = b"MOGPB200"


def env_rank_world():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", os.environ.get("RANK", "0"))))


def _recv_exact(conn, n):
    buf = b""
    while len(buf) < n:
        chunk = conn.recv(n - len(buf))
        if not chunk:
            raise ConnectionError("rendezvous peer closed the connection")
        buf += chunk
    return buf


def broadcast_bytes(payload, rank, world, tag=0, timeout=120.0):
    """Rank 0's ``payload`` (bytes) is returned on every rank."""
    if world == 1:
        return payload
    addr = os.environ.get("MASTER_ADDR", "127.0.0.1")
    base = int(os.environ.get("MASTER_PORT", "29500")) + 101 + 7 * int(tag)
    ports = [base + k for k in range(8)]
    if rank == 0:
        srv = None
        for port in ports:
            try:
                srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
                srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
                srv.bind((addr, port))
                break
            except OSError:
                srv.close()
                srv = None
        if srv is None:
            raise RuntimeError("rendezvous: no free port near %d" % base)
        srv.listen(world)
        srv.settimeout(timeout)
        msg = _MAGIC + struct.pack("<I", len(payload)) + payload
        served = 0
        conns = []
        while served < world - 1:
            conn, _ = srv.accept()
            hello = _recv_exact(conn, len(_MAGIC))
            if hello != _MAGIC:
                conn.close()
                continue
            conn.sendall(msg)
            conns.append(conn)
            served += 1
        for c in conns:
            c.close()
        srv.close()
        return payload
    deadline = time.time() + timeout
    while time.time() < deadline:
        for port in ports:
            try:
                conn = socket.create_connection((addr, port), timeout=2.0)
            except OSError:
                continue
            try:
                conn.sendall(_MAGIC)
                head = _recv_exact(conn, len(_MAGIC) + 4)
                if head[:len(_MAGIC)] != _MAGIC:
                    continue
                (n,) = struct.unpack("<I", head[len(_MAGIC):])
                return _recv_exact(conn, n)
            except (OSError, ConnectionError):
                continue
            finally:
                conn.close()
        time.sleep(0.2)
    raise TimeoutError("rendezvous: could not reach rank 0 at %s:%s" % (addr, ports))


def init_comm(device=None):
    """-> libmogp.Comm for this rank (None when WORLD_SIZE == 1)."""
    from . import libmogp
    rank, world, local_rank = env_rank_world()
    if world == 1:
        return None
    uid = libmogp.Comm.unique_id() if rank == 0 else None
    uid = broadcast_bytes(uid, rank, world)
    return libmogp.Comm(uid, rank, world, local_rank if device is None else device)
