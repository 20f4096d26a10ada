"""Wire protocol of the drop-in worker against a fake Distributer on a loopback socket
(restating Distributer.cs:230-270,358-458; DistributerWorkload.cs:53-100).

CPU run: process_workload is stubbed with the CPU oracle (a test stand-in, not a fallback) so
the protocol logic is exercised without a GPU.  GPU run (-m gpu): the real CUDA path."""
import struct

import numpy as np
import pytest

from fake_distributer import DATA_CHUNK_SIZE, FakeDistributer


def _run_worker(worker, port):
    n = 0
    while worker.do_workload_single("127.0.0.1", port, verbose=False):
        n += 1
    return n


def _check_session(oracle, fd, n_done, workloads, rejected, pipelined=False):
    assert fd.error is None, fd.error
    assert n_done == len(workloads)
    if pipelined:
        assert fd.log.count(("none",)) == 1                          # leases run ahead of the responses
    else:
        assert fd.log[-1] == ("none",)                               # 0x11 ended the loop
    for w in workloads:
        if w in rejected:
            assert ("reject", w) in fd.log and w not in fd.completed
            continue
        level, mrd, ir, ii = w
        data = np.frombuffer(fd.completed[w], dtype=np.uint8)
        assert data.size == DATA_CHUNK_SIZE
        assert np.array_equal(data, oracle.process_workload(level, mrd, ir, ii, 4096))


WORKLOADS = [(4, 8, 0, 0), (4, 8, 3, 0), (4, 6, 0, 3)]            # corner chunks: cheap on the oracle too


def test_protocol_with_oracle_stub(oracle, monkeypatch):
    from distributedmandelbrot_b200 import worker
    monkeypatch.setattr(worker, "process_workload",
                        lambda level, mrd, ir, ii: oracle.process_workload(level, mrd, ir, ii, 4096))
    fd = FakeDistributer(WORKLOADS, reject=[WORKLOADS[1]])
    try:
        n = _run_worker(worker, fd.port)
    finally:
        fd.close()
    _check_session(oracle, fd, n, WORKLOADS, {WORKLOADS[1]})


@pytest.mark.gpu
def test_protocol_end_to_end_on_gpu(gpu, oracle):
    from distributedmandelbrot_b200 import worker
    wl = WORKLOADS + [(4, 64, 1, 2)]
    fd = FakeDistributer(wl, reject=[wl[0]])
    try:
        n = _run_worker(worker, fd.port)
    finally:
        fd.close()
    _check_session(oracle, fd, n, wl, {wl[0]})


@pytest.mark.gpu
def test_pipelined_worker_on_gpu(gpu, oracle):
    """Several leases in flight (row f4): same bytes on the wire, every workload exactly once."""
    from distributedmandelbrot_b200 import worker
    wl = [(4, 8, 0, 0), (4, 8, 3, 0), (4, 6, 0, 3), (4, 64, 1, 2), (4, 8, 3, 3), (4, 9, 0, 1), (4, 7, 2, 0)]
    fd = FakeDistributer(wl, reject=[wl[2]])
    try:
        n = worker.run_pipelined("127.0.0.1", fd.port, depth=3, verbose=False)
    finally:
        fd.close()
    _check_session(oracle, fd, n, wl, {wl[2]}, pipelined=True)


def test_workload_record_is_16_le_bytes():
    """SURVEY.md Appendix B.1 example: (level=4, mrd=256, ir=1, ii=2)."""
    import socket
    from distributedmandelbrot_b200 import worker
    a, b = socket.socketpair()
    try:
        a.sendall(bytes.fromhex("04000000" "00010000" "01000000" "02000000"))
        assert worker.receive_workload(b) == (4, 256, 1, 2)
        # fragmented delivery (the reference's bare recv(4) would mis-parse this)
        for chunk in (b"\x0a\x00", b"\x00\x00\x00", b"\x04\x00\x00\x03\x00\x00", b"\x00\x05\x00\x00\x00"):
            a.sendall(chunk)
        assert worker.receive_workload(b) == (10, 1024, 3, 5)
    finally:
        a.close()
        b.close()


def test_unknown_codes_raise(monkeypatch):
    import socket
    import threading
    from distributedmandelbrot_b200 import worker
    srv = socket.socket()
    srv.bind(("127.0.0.1", 0))
    srv.listen(1)
    port = srv.getsockname()[1]

    def serve():
        c, _ = srv.accept()
        c.recv(1)
        c.sendall(b"\x7f")
        c.close()

    t = threading.Thread(target=serve, daemon=True)
    t.start()
    with pytest.raises(Exception, match="Unknown response code"):
        worker.do_workload_single("127.0.0.1", port, verbose=False)
    t.join(5)
    srv.close()


def test_constants_match_reference():
    from distributedmandelbrot_b200 import worker
    assert (worker.MIN_AXIS, worker.MAX_AXIS) == (-2, 2)
    assert (worker.REQUEST_CODE, worker.RESPONSE_CODE) == (0x00, 0x01)
    assert (worker.WORKLOAD_AVAILABLE_CODE, worker.WORKLOAD_NOT_AVAILABLE_CODE) == (0x10, 0x11)
    assert (worker.WORKLOAD_ACCEPT_CODE, worker.WORKLOAD_REJECT_CODE) == (0x20, 0x21)
    assert worker.DATA_CHUNK_SIZE == 16_777_216
    assert struct.calcsize("<IIII") == 16


# ---- all GPUs of a box as ONE worker behind the Distributer (worker.run_inbox) ---------------------

def _inbox_rank(rank, world, port, dport, name, wl):
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import oracle
    import torch.distributed as dist
    from distributedmandelbrot_b200 import worker
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    state = {}

    def share(obj):
        box = [obj]
        dist.broadcast_object_list(box, src=0)
        state["leases"] = box[0]
        return box[0]

    cache = {}

    def stub(first, count, frame):
        # stands in for the CUDA dispatch loop: bands of the current leases, from the CPU oracle
        d, rows = 4096, 256
        bpc = d // rows
        for u in range(first, first + count):
            w = state["leases"][u // bpc]
            if w not in cache:
                cache[w] = oracle.process_workload(*w, d)
            lo = (u % bpc) * rows * d
            frame[u * rows * d:(u + 1) * rows * d] = cache[w][lo:lo + rows * d]

    sent = worker.run_inbox("127.0.0.1", dport, rank=rank, world=world, barrier=dist.barrier, share=share,
                            launch=stub, verbose=False, name=name, batch=4)
    counts = [None] * world
    dist.all_gather_object(counts, sent)
    assert sum(counts) == len(wl)
    dist.destroy_process_group()


def test_inbox_worker_behind_the_distributer(oracle, tmp_path):
    """Two ranks act as ONE worker: rank 0 leases batches of 4 (then a short last batch of 1), the bands of
    every chunk are delivered by both ranks into the shared frame, and whichever rank completes a chunk
    returns its 16 MiB payload over the reference protocol."""
    import os
    import socket
    import torch.multiprocessing as mp
    wl = [(4, 8, 0, 0), (4, 8, 3, 0), (4, 6, 0, 3), (4, 7, 3, 3), (8, 5, 0, 0)]
    fd = FakeDistributer(wl)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    try:
        mp.spawn(_inbox_rank, args=(2, port, fd.port, f"/mbw_test_{os.getpid()}", wl), nprocs=2, join=True)
    finally:
        fd.close()
    assert fd.error is None, fd.error
    assert sorted(fd.completed) == sorted(wl) and not fd.leases
    for w in wl:
        data = np.frombuffer(fd.completed[w], dtype=np.uint8)
        assert np.array_equal(data, oracle.process_workload(*w, 4096)), w
    assert [e for e in fd.log if e[0] == "lease"] == [("lease", w) for w in wl]


@pytest.mark.gpu
def test_inbox_worker_on_gpu_behind_the_distributer(gpu, oracle):
    """run_inbox on the real CUDA path (one rank): batches of leases rendered band by band into the shared
    page-locked frame, payloads returned over the protocol, a shorter last batch retargets the same frame."""
    import os
    from distributedmandelbrot_b200 import worker
    wl = WORKLOADS + [(4, 64, 1, 2), (4, 9, 3, 3)]
    fd = FakeDistributer(wl)
    try:
        sent = worker.run_inbox("127.0.0.1", fd.port, rank=0, world=1, device=0, verbose=False, batch=3,
                                name=f"/mbw_gpu_{os.getpid()}")
    finally:
        fd.close()
    assert fd.error is None, fd.error
    assert sent == len(wl) and sorted(fd.completed) == sorted(wl)
    for w in wl:
        assert np.array_equal(np.frombuffer(fd.completed[w], dtype=np.uint8), oracle.process_workload(*w, 4096)), w
