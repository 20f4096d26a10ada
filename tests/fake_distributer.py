"""Python fake of the C# Distributer's worker-facing protocol, for loopback tests.

Restates (byte for byte) DistributedMandelbrot/Distributer.cs:230-270 (accept loop / dispatch
on the first byte), :358-392 (HandleWorkloadRequest), :397-458 (HandleWorkloadResponse) and
DistributerWorkload.cs:53-100 (16-byte little-endian Workload).  Single-threaded, one client at
a time, like the reference.  Test infrastructure only.
"""
from __future__ import annotations

import socket
import struct
import threading
import time

REQUEST, RESPONSE = 0x00, 0x01
AVAILABLE, NOT_AVAILABLE = 0x10, 0x11
ACCEPT, REJECT = 0x20, 0x21
DATA_CHUNK_SIZE = 4096 * 4096          # DataChunk.cs:20,27


class FakeDistributer:
    def __init__(self, workloads, lease_seconds: float = 3600.0, reject=()):
        """workloads: list of (level, mrd, ir, ii); reject: workloads whose lease is dropped
        before the response arrives (forces the 0x21 path)."""
        self.pending = list(workloads)
        self.leases = {}                 # workload -> expiry
        self.completed = {}              # workload -> bytes
        self.reject = set(reject)
        self.lease_seconds = lease_seconds
        self.log = []
        self.srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        self.srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        self.srv.bind(("127.0.0.1", 0))
        self.srv.listen(16)              # Distributer.cs:221
        self.port = self.srv.getsockname()[1]
        self._stop = False
        self.error = None
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _recv(self, c, n):
        buf = bytearray()
        while len(buf) < n:
            part = c.recv(n - len(buf))
            if not part:
                raise ConnectionError("short read")
            buf += part
        return bytes(buf)

    def _run(self):
        self.srv.settimeout(0.2)
        while not self._stop:
            try:
                c, _ = self.srv.accept()
            except socket.timeout:
                continue
            except OSError:
                break
            try:
                c.settimeout(10.0)
                code = self._recv(c, 1)[0]
                if code == REQUEST:
                    self._handle_request(c)
                elif code == RESPONSE:
                    self._handle_response(c)
                else:
                    self.log.append(("unknown", code))
            except Exception as e:  # noqa: BLE001 -- surfaced to the test through .error
                self.error = e
            finally:
                c.close()

    def _handle_request(self, c):
        # Distributer.cs:358-392
        if self.pending:
            w = self.pending.pop(0)
            c.sendall(bytes([AVAILABLE]))
            for field in w:              # four separate 4-byte sends (DistributerWorkload.cs:61-75)
                c.sendall(struct.pack("<I", field))
            if w not in self.reject:
                self.leases[w] = time.monotonic() + self.lease_seconds
            self.log.append(("lease", w))
        else:
            c.sendall(bytes([NOT_AVAILABLE]))
            self.log.append(("none",))

    def _handle_response(self, c):
        # Distributer.cs:397-458
        w = struct.unpack("<IIII", self._recv(c, 16))
        exp = self.leases.get(w)
        if exp is not None and exp > time.monotonic():
            c.sendall(bytes([ACCEPT]))
            data = self._recv(c, DATA_CHUNK_SIZE)
            # the reference closes right after reading dataChunkSize bytes; anything more is a bug
            c.settimeout(0.2)
            try:
                extra = c.recv(1)
            except socket.timeout:
                extra = b""
            assert extra == b"", "worker sent more than dataChunkSize bytes"
            del self.leases[w]
            self.completed[w] = data
            self.log.append(("accept", w))
        else:
            c.sendall(bytes([REJECT]))
            self.log.append(("reject", w))

    def close(self):
        self._stop = True
        try:
            self.srv.close()
        finally:
            self.thread.join(timeout=5)
