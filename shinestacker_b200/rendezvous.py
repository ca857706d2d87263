"""Bootstrap for the multi-GPU exchange: a tiny all-gather over TCP (standard library only).

The product's multi-GPU data path needs no collective library: after set-up, the ranks talk through NVLink peer
memory and epoch flags only (include/shinestack_b200.h, "fused peer exchange").  Set-up itself is one all-gather of
384-byte blobs (CUDA IPC handles) per native handle; this module provides it for one-process-per-GPU launches
(`torchrun`-style environment: RANK, WORLD_SIZE, MASTER_ADDR, MASTER_PORT).  Rank 0 listens on MASTER_PORT + offset,
the others connect (with retries) and keep the connection for later rounds.
"""
import os
import socket
import struct
import time

PORT_OFFSET = 29


def _recv_exact(sock, n):
    buf = bytearray()
    while len(buf) < n:
        chunk = sock.recv(n - len(buf))
        if not chunk:
            raise ConnectionError("rendezvous peer closed the connection")
        buf += chunk
    return bytes(buf)


def _send_msg(sock, payload):
    sock.sendall(struct.pack("<I", len(payload)) + payload)


def _recv_msg(sock):
    (n,) = struct.unpack("<I", _recv_exact(sock, 4))
    return _recv_exact(sock, n)


class Single:
    """world == 1."""
    rank, world = 0, 1

    def all_gather(self, payload):
        return [bytes(payload)]

    def barrier(self):
        pass

    def close(self):
        pass


class TcpRendezvous:
    def __init__(self, rank, world, addr="127.0.0.1", port=29500 + PORT_OFFSET, timeout=120.0):
        self.rank, self.world = int(rank), int(world)
        self._peers = []      # rank 0: sockets of ranks 1..world-1 (by rank); others: [socket to rank 0]
        deadline = time.time() + timeout
        if self.rank == 0:
            srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
            srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            srv.bind((addr, port))
            srv.listen(self.world)
            srv.settimeout(timeout)
            by_rank = {}
            try:
                while len(by_rank) < self.world - 1:
                    conn, _ = srv.accept()
                    conn.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
                    conn.settimeout(timeout)
                    (r,) = struct.unpack("<I", _recv_exact(conn, 4))
                    by_rank[r] = conn
            finally:
                srv.close()
            self._peers = [by_rank[r] for r in range(1, self.world)]
        else:
            while True:
                try:
                    s = socket.create_connection((addr, port), timeout=5.0)
                    break
                except OSError:
                    if time.time() > deadline:
                        raise
                    time.sleep(0.05)
            s.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
            s.settimeout(timeout)
            s.sendall(struct.pack("<I", self.rank))
            self._peers = [s]

    @classmethod
    def from_env(cls, offset=0):
        """RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT as set by torchrun; world 1 needs no sockets."""
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if world <= 1:
            return Single()
        return cls(int(os.environ.get("RANK", "0")), world, os.environ.get("MASTER_ADDR", "127.0.0.1"),
                   int(os.environ.get("MASTER_PORT", "29500")) + PORT_OFFSET + offset)

    def all_gather(self, payload):
        """Every rank contributes ``payload`` (bytes); returns the list of all contributions in rank order."""
        payload = bytes(payload)
        if self.rank == 0:
            parts = [payload] + [_recv_msg(s) for s in self._peers]
            blob = b"".join(struct.pack("<I", len(p)) + p for p in parts)
            for s in self._peers:
                _send_msg(s, blob)
            return parts
        s = self._peers[0]
        _send_msg(s, payload)
        blob = _recv_msg(s)
        parts, off = [], 0
        for _ in range(self.world):
            (n,) = struct.unpack_from("<I", blob, off)
            parts.append(blob[off + 4:off + 4 + n])
            off += 4 + n
        return parts

    def barrier(self):
        self.all_gather(b"")

    def close(self):
        for s in self._peers:
            try:
                s.close()
            except OSError:
                pass
        self._peers = []
