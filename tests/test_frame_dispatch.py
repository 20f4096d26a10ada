"""Host logic of the in-box frame renderer: band geometry, work-stealing dispatch across
world_size-2 processes (gloo on CPU, a stub in place of the kernel launch), final gather."""
import os
import socket

import numpy as np
import pytest


def test_band_geometry_covers_frame_exactly_once(native):
    from distributedmandelbrot_b200.frame import FrameSpec, frame_bands
    spec = FrameSpec("t", level=10, blocks=3, definition=100, ir0=2, ii0=4, mrd=50, band_rows=32)
    tiles, extents = frame_bands(spec)
    assert len(tiles) == spec.n_bands == 9 * 4
    cover = np.zeros(spec.pixels, np.int32)
    for t, (off, cnt) in zip(tiles, extents):
        assert (t.x0, t.w, t.out_pitch, t.definition, t.mrd) == (0, 100, 100, 100, 50)
        assert t.out_offset == off and cnt == t.h * 100
        cover[off:off + cnt] += 1
    assert (cover == 1).all()
    # chunk order: imag-major block index, chunk-contiguous layout
    assert spec.chunk_indices()[:4] == [(2, 4), (3, 4), (4, 4), (2, 5)]
    t0 = native.chunk_geometry(10, 3, 4, 100, 50)
    assert (tiles[4].start_r, tiles[4].start_i) == (t0.start_r, t0.start_i)


def test_baseline_configs_are_pinned():
    from distributedmandelbrot_b200.frame import CONFIGS
    c4, c5, c2, c3 = CONFIGS["cfg4"], CONFIGS["cfg5"], CONFIGS["cfg2"], CONFIGS["cfg3"]
    assert (c4.level, c4.blocks, c4.definition, c4.ir0, c4.ii0, c4.mrd, c4.dtype) == (800, 16, 4096, 242, 412, 10000, "f64")
    assert c4.pixels == 65536 ** 2 and c5.pixels == 32768 ** 2 and c2.pixels == 4096 ** 2 and c3.pixels == 16384 ** 2
    assert (c5.level, c5.ir0, c5.ii0, c5.mrd) == (4_000_000, 1_256_353, 2_131_821, 50000)
    assert c3.dtype == "f32" and c3.mrd == 4096 and c2.mrd == 1024


def test_coalesce():
    from distributedmandelbrot_b200.frame import coalesce
    assert coalesce([(4, 2), (0, 2), (2, 2), (9, 1)]) == [(0, 6), (9, 1)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, qname, tmp):
    import time
    import torch
    import torch.distributed as dist
    from distributedmandelbrot_b200 import _native as nat
    from distributedmandelbrot_b200.frame import FrameRenderer, FrameSpec
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    spec = FrameSpec("t", level=8, blocks=2, definition=64, ir0=1, ii0=3, mrd=20, band_rows=4)

    def stub_launch(first, count, out, _ext=[None]):
        # stands in for mb_render_counts: band b is filled with b + 1; rank 1 is "slow" so rank 0 steals
        for b in range(first, first + count):
            off, cnt = fr.extents[b]
            out[off:off + cnt] = b + 1
        if rank == 1:
            time.sleep(0.002)

    fr = FrameRenderer(spec, rank=rank, world=world, launch=stub_launch,
                       out=torch.zeros(spec.pixels, dtype=torch.int32), bands_per_claim=2)
    if rank == 0:
        q = nat.TileQueue(qname, world, spec.n_bands, create=True)
    dist.barrier()
    if rank != 0:
        q = nat.TileQueue(qname)
    for _pass in range(2):                 # second pass re-arms the same segment
        fr.out.zero_()
        dist.barrier()
        fr.render_stolen(q)
        fr.gather_to_root()
        dist.barrier()
        if rank == 0:
            want = torch.zeros(spec.pixels, dtype=torch.int32)
            for b, (off, cnt) in enumerate(fr.extents):
                want[off:off + cnt] = b + 1
            assert torch.equal(fr.out, want), "gathered frame differs"
            q.reset(spec.n_bands)
        dist.barrier()
    owned = [None] * world
    dist.all_gather_object(owned, fr.claimed)
    if rank == 0:
        bands = sorted(b for o in owned for f, c in o for b in range(f, f + c))
        assert bands == list(range(spec.n_bands)), "every band exactly once"
        with open(os.path.join(tmp, "ok"), "w") as f:
            f.write(f"{[len(o) for o in owned]} steals={fr.steals}")
    q.close()
    dist.destroy_process_group()


def test_two_rank_work_stealing_and_gather(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    qname = f"/mbq_test_frame_{os.getpid()}"
    mp.spawn(_rank_main, args=(2, port, qname, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()
