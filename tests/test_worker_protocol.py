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
