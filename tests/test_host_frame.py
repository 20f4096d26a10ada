"""Host logic of the end-to-end in-box path: the shared payload frame (csrc/host_frame.cpp), its
completion accounting, and InboxWorker across world_size-2 gloo processes with a stub in place of
the CUDA dispatch loop -- including a rank that dies after claiming bands."""
import os
import socket

import numpy as np
import pytest


def test_hostframe_accounting(native):
    name = f"/mbf_test_{os.getpid()}"
    f = native.HostFrame(name, payload_bytes=10 * 100, n_units=10, units_per_group=4, create=True)
    try:
        assert (f.nbytes, f.n_units, f.n_groups) == (1000, 10, 3)         # groups of 4, 4, 2 units
        assert [f.group_left(g) for g in range(3)] == [4, 4, 2]
        g = native.HostFrame(name)                                         # a second mapping sees the same bytes
        f.array[:5] = [1, 2, 3, 4, 5]
        assert g.array[:5].tolist() == [1, 2, 3, 4, 5]
        assert f.mark_done(0, 3) == [] and g.group_left(0) == 1
        assert g.mark_done(3, 2) == [0]                                    # unit 3 completes group 0, unit 4 opens group 1
        assert f.mark_done(3, 1) == []                                     # idempotent per unit
        assert f.missing() == [5, 6, 7, 8, 9]
        assert f.mark_done(8, 2) == [2] and f.mark_done(5, 3) == [1] and f.missing() == []
        f.reset()
        assert [g.group_left(k) for k in range(3)] == [4, 4, 2] and len(g.missing()) == 10
        with pytest.raises(ValueError):
            f.mark_done(9, 2)
        g.close()
    finally:
        f.close()
    with pytest.raises(native.NativeError):
        native.HostFrame(name)                                             # unlinked by its creator


def test_workload_bands_layout(native):
    from distributedmandelbrot_b200.frame import workload_bands
    wl = [(10, 300, 3, 5), (4, 64, 0, 0), (16, 500, 5, 8)]
    tiles, bpc = workload_bands(wl, 100, 32, shortcut=True)
    assert bpc == 4 and len(tiles) == 12
    cover = np.zeros(3 * 100 * 100, np.int32)
    for k, t in enumerate(tiles):
        c = k // bpc
        assert (t.mrd, t.definition, t.w, t.out_pitch, t.x0, t.flags) == (wl[c][1], 100, 100, 100, 0, native.TILE_PERIOD_SHORTCUT)
        assert t.y0 == (k % bpc) * 32 and t.out_offset == c * 10000 + t.y0 * 100
        cover[t.out_offset:t.out_offset + t.w * t.h] += 1
    assert (cover == 1).all()
    ref = native.chunk_geometry(16, 5, 8, 100, 500)
    assert (tiles[8].start_r, tiles[8].step_i) == (ref.start_r, ref.step_i)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, name, tmp):
    import time
    import torch.distributed as dist
    from distributedmandelbrot_b200.frame import InboxWorker
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    wl = [(8, 20 + k, 1 + k % 2, 3) for k in range(6)]
    d, rows = 48, 8
    die = {"on": False}

    def stub(first, count, frame):
        # stands in for the CUDA loop: band u is filled with (u % 251) + 1; rank 1 is slow so rank 0 steals
        if die["on"] and rank == 1:
            return False                                        # claimed, never delivered
        for u in range(first, first + count):
            frame[u * rows * d:(u + 1) * rows * d] = (u % 251) + 1
        if rank == 1:
            time.sleep(0.002)

    w = InboxWorker(wl, definition=d, band_rows=rows, rank=rank, world=world, name=name, barrier=dist.barrier,
                    launch=stub, bands_per_claim=2)
    want = np.repeat((np.arange(w.n_bands) % 251 + 1).astype(np.uint8), rows * d)
    seen = []
    for _pass in range(2):
        w.rearm()
        w.frame.array[:] = 0 if rank == 0 else w.frame.array[:]
        dist.barrier()
        got = w.render(consume=lambda wk, p: seen.append((wk, int(p[0]))))
        dist.barrier()
        assert w.missing() == []
        assert np.array_equal(w.frame.array, want), "frame differs"
        every = [None] * world
        dist.all_gather_object(every, got)
        assert sorted(g for e in every for g in e) == list(range(len(wl))), "every payload completed exactly once"
    assert all(p == (wl.index(wk) * w.bands_per_chunk) % 251 + 1 for wk, p in seen)
    # a rank dies after claiming: survivors finish, the hole is detected and repaired by rank 0
    die["on"] = True
    w.rearm()
    if rank == 0:
        w.frame.array[:] = 0
    dist.barrier()
    w.render()
    dist.barrier()
    holes = w.missing()
    if rank == 0:
        assert holes, "rank 1 claimed bands and delivered none"
        assert not np.array_equal(w.frame.array, want)
        fixed = w.repair()
        assert w.missing() == [] and np.array_equal(w.frame.array, want)
        assert fixed, "the repair completed at least one payload"
        with open(os.path.join(tmp, "ok"), "w") as f:
            f.write(f"holes={len(holes)} steals={w.steals}")
    dist.barrier()
    w.close()
    dist.destroy_process_group()


def test_two_rank_inbox_worker_with_a_dead_rank(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_rank_main, args=(2, _free_port(), f"/mbi_test_{os.getpid()}", str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()


# ---- storage writer: lazy, queue-driven dispatch across ranks (no GPU: stub serializer) -------------

class _StubSerializer:
    """Stands in for storage.BatchSerializer: 'renders' a chunk as a constant payload and serialises it
    with the restated CPU codec; records when batches were submitted relative to queue claims."""

    def __init__(self, log, delay=0.0):
        self.log, self.delay = log, delay

    def submit(self, workloads):
        import time
        self.log.append(("submit", list(workloads)))
        time.sleep(self.delay)
        return list(workloads)

    def collect(self, ticket):
        import chunk_codec
        out = []
        for (level, mrd, ir, ii) in ticket:
            payload = np.full(64, (ir * 7 + ii) % 5, np.uint8)
            et = chunk_codec.entry_type(payload)
            out.append(((level, mrd, ir, ii), et, chunk_codec.serialize(payload) if et == 0 else None))
        return out


def _store_rank(rank, world, port, name, tmp):
    import torch.distributed as dist
    from distributedmandelbrot_b200 import _native as nat
    from distributedmandelbrot_b200.storage import DataStore, queue_chunks, render_level_to_store
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    level = 6
    order = [(ir, ii) for ir in range(level) for ii in range(level)]
    q = nat.TileQueue(name, world, len(order), create=True) if rank == 0 else None
    dist.barrier()
    if rank != 0:
        q = nat.TileQueue(name)
    store = DataStore(tmp)
    log = []
    claimed = []

    def mine():
        for c in queue_chunks(q, rank, order, 2):
            claimed.append(c)
            log.append(("claim", c))
            yield c

    n = render_level_to_store(store, level, 50, mine(), definition=8, batch=4,
                              serializer=_StubSerializer(log, delay=0.02 if rank == 1 else 0.0))
    # lazy: never more than one batch (4 chunks) claimed ahead of the batch being submitted
    ahead = 0
    for kind, what in log:
        ahead = ahead + 1 if kind == "claim" else ahead - len(what)
        assert ahead <= 4, "chunks were claimed before the previous batch was submitted"
    counts = [None] * world
    dist.all_gather_object(counts, (n, claimed))
    dist.barrier()
    if rank == 0:
        assert sum(c[0] for c in counts) == len(order)
        assert sorted(x for c in counts for x in c[1]) == sorted(order), "every chunk claimed exactly once"
        assert counts[0][0] > counts[1][0], "the fast rank took more than the slow one (work stealing while rendering)"
        done = DataStore(tmp).completed()
        assert sorted((lv, r, i) for (lv, r, i, _t, _f) in done) == sorted((level, r, i) for r, i in order)
        with open(os.path.join(tmp, "ok"), "w") as f:
            f.write(str([c[0] for c in counts]))
    q.close()
    dist.destroy_process_group()


def test_storage_writer_draws_chunks_lazily_from_the_queue(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_store_rank, args=(2, _free_port(), f"/mbs_test_{os.getpid()}", str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()
