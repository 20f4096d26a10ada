"""In-box work-stealing tile queue (csrc/tile_wsq.cpp): every index handed out exactly once."""
import multiprocessing as mp
import os
import threading

import numpy as np
import pytest


def _name(tag):
    return f"/mbq_test_{os.getpid()}_{tag}"


def test_single_rank_drains_in_order(native):
    with native.TileQueue(_name("a"), 1, 10, create=True) as q:
        got = []
        while (r := q.pop(0, 3)) is not None:
            got.append(r)
        assert got == [(0, 3, False), (3, 3, False), (6, 3, False), (9, 1, False)]
        assert q.pop(0, 1) is None
        q.reset(4)
        assert q.pop(0, 8) == (0, 4, False)


def test_static_split_then_steal(native):
    with native.TileQueue(_name("b"), 4, 100, create=True) as q:
        assert q.pop(2, 5) == (50, 5, False)          # rank 2 owns [50, 75)
        # rank 0 drains its own range, then must steal
        seen = []
        while (r := q.pop(0, 7)) is not None:
            seen.append(r)
        idx = sorted(i for f, g, _ in seen for i in range(f, f + g))
        assert idx == [i for i in range(100) if not 50 <= i < 55]
        assert any(s for _, _, s in seen)
        for r in range(4):
            assert q.pop(r, 1) is None


def test_threads_exactly_once(native):
    n_ranks, total = 8, 20000
    with native.TileQueue(_name("c"), n_ranks, total, create=True) as q:
        hits = np.zeros(total, np.int32)
        steals = [0] * n_ranks

        def work(rank):
            h = native.TileQueue(q.name.decode())
            while (r := h.pop(rank, 1 + rank % 3)) is not None:
                f, g, s = r
                hits[f:f + g] += 1
                steals[rank] += int(s)
            h.close()

        ts = [threading.Thread(target=work, args=(r,)) for r in range(n_ranks)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        assert (hits == 1).all()


def _proc(name, rank, total, ret):
    from distributedmandelbrot_b200 import _native as nat
    q = nat.TileQueue(name)
    mine = []
    while (r := q.pop(rank, 2)) is not None:
        mine.append((r[0], r[1]))
    q.close()
    ret.put((rank, mine))


def test_processes_exactly_once(native):
    n_ranks, total = 4, 5000
    name = _name("d")
    ctx = mp.get_context("spawn")
    with native.TileQueue(name, n_ranks, total, create=True):
        ret = ctx.Queue()
        ps = [ctx.Process(target=_proc, args=(name, r, total, ret)) for r in range(n_ranks)]
        [p.start() for p in ps]
        res = [ret.get(timeout=120) for _ in ps]
        [p.join(timeout=60) for p in ps]
    hits = np.zeros(total, np.int32)
    for _, mine in res:
        for f, g in mine:
            hits[f:f + g] += 1
    assert (hits == 1).all()


def test_errors(native):
    with pytest.raises(native.NativeError):
        native.TileQueue("/mbq_does_not_exist_xyz")
    with pytest.raises(ValueError):
        native.TileQueue("no_leading_slash", 2, 10, create=True)
    with native.TileQueue(_name("e"), 2, 10, create=True) as q:
        with pytest.raises(ValueError):
            q.pop(5, 1)
