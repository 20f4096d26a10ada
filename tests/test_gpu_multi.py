"""In-box multi-GPU paths on real hardware.

* one GPU (always runs under -m gpu): InboxWorker -- band-granular dispatch into the shared pinned host
  frame -- returns byte-for-byte what process_workload returns, plain and with the periodicity shortcut,
  including the repair of bands a dead rank left behind;
* two or more GPUs (skipped on a 1-GPU box; `gpurun --gpus 2`): work-stolen bands + the NCCL gather,
  the peer-memory frame buffer (`--gather p2p`) and the InboxWorker, each compared ELEMENT-WISE with
  the single-GPU frame / with process_workload."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

WL = [(800, 3000, 250, 420), (800, 3000, 249, 420), (10, 700, 3, 5), (4, 64, 0, 0), (16, 900, 5, 8), (8, 300, 3, 4)]
D = 320


def test_inbox_worker_single_gpu_equals_process_workload(gpu, oracle):
    from distributedmandelbrot_b200 import worker
    from distributedmandelbrot_b200.frame import InboxWorker
    for shortcut in (False, True):
        w = InboxWorker(WL, definition=D, band_rows=48, name=f"/mbi_gpu1_{os.getpid()}", shortcut=shortcut, bands_per_claim=3, depth=3)
        try:
            seen = []
            for _pass in range(2):
                w.rearm()
                w.frame.array[:] = 0xEE
                done = w.render(consume=lambda wk, p: seen.append(wk))
                assert sorted(done) == list(range(len(WL))) and w.missing() == []
                for k, wk in enumerate(WL):
                    assert np.array_equal(w.payload(k), worker.process_workload(*wk, definition=D)), (shortcut, wk)
            assert sorted(seen) == sorted(WL + WL)
            assert np.array_equal(w.payload(2), oracle.process_workload(*WL[2], D))
            # punch holes: as if a rank had claimed these bands and died; repair renders exactly them
            w.rearm()
            w.frame.array[:] = 0
            w.render()
            want = w.frame.array.copy()
            w.frame.reset()
            w.frame.array[:] = 0x55
            units = list(range(w.n_bands))
            alive = [u for u in units if u % 5 != 3]
            for u in alive:
                w.frame.mark_done(u, 1)
                lo = int(w.tiles[u].out_offset)
                w.frame.array[lo:lo + D * int(w.tiles[u].h)] = want[lo:lo + D * int(w.tiles[u].h)]
            assert w.missing() == [u for u in units if u % 5 == 3]
            fixed = w.repair()
            assert w.missing() == [] and np.array_equal(w.frame.array, want) and fixed
        finally:
            w.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    from distributedmandelbrot_b200 import _native as nat
    from distributedmandelbrot_b200 import worker
    from distributedmandelbrot_b200.frame import CONFIGS, FrameRenderer, InboxWorker, queue_name, shared_frame_buffer
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    nat.init(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    # a 3x3 block of cfg4 around the zoom centre at reduced definition: boundary and interior chunks
    spec = CONFIGS["cfg4"].scaled(name="t", blocks=3, ir0=249, ii0=419, definition=512, mrd=4000, band_rows=32)
    ref = None
    if rank == 0:
        one = FrameRenderer(spec, device=rank)
        one.render_all()
        ref = one.out.clone()
        del one
    qn = queue_name() + "_t"
    q = nat.TileQueue(qn, world, spec.n_bands, create=True) if rank == 0 else None
    dist.barrier()
    if rank != 0:
        q = nat.TileQueue(qn)
    notes = []
    for mode in ("nccl", "p2p"):
        shared, closer = (shared_frame_buffer(spec, rank, rank) if mode == "p2p" else (None, None))
        fr = FrameRenderer(spec, device=rank, rank=rank, world=world, bands_per_claim=2, depth=4, out=shared)
        for _pass in range(2):
            dist.barrier()
            if rank == 0:
                q.reset(spec.n_bands)
                fr.out.fill_(-5)
            dist.barrier()
            fr.render_stolen(q)
            if shared is None:
                fr.gather_to_root()
            torch.cuda.synchronize()
            dist.barrier()
            if rank == 0:
                assert torch.equal(fr.out, ref), f"{mode}: gathered frame differs from the single-GPU frame"
        owned = [None] * world
        dist.all_gather_object(owned, fr.claimed)
        if rank == 0:
            bands = sorted(b for o in owned for f, c in o for b in range(f, f + c))
            assert bands == list(range(spec.n_bands))
            notes.append(f"{mode}: bands per rank {[sum(c for _, c in o) for o in owned]}")
        del fr
        if closer is not None:
            torch.cuda.synchronize()
            dist.barrier()
            closer()
    q.close()
    # end-to-end form: payloads in the shared pinned host frame, byte-identical to process_workload
    for shortcut in (False, True):
        w = InboxWorker(WL, definition=D, band_rows=32, rank=rank, world=world, device=rank, barrier=dist.barrier,
                        shortcut=shortcut, bands_per_claim=2, depth=4)
        for _pass in range(2):
            w.rearm()
            if rank == 0:
                w.frame.array[:] = 0xEE
            dist.barrier()
            done = w.render()
            dist.barrier()
            assert w.missing() == []
            every = [None] * world
            dist.all_gather_object(every, done)
            assert sorted(g for e in every for g in e) == list(range(len(WL)))
            for k, wk in enumerate(WL):
                if k % world == rank:                      # every rank checks its share of the SHARED frame
                    assert np.array_equal(w.payload(k), worker.process_workload(*wk, definition=D)), (shortcut, wk)
        if rank == 0:
            notes.append(f"inbox shortcut={shortcut}: claims {w.claims} steals {w.steals}")
        w.close()
    if rank == 0:
        with open(os.path.join(tmp, "ok"), "w") as f:
            f.write("; ".join(notes))
    dist.destroy_process_group()


def test_multi_gpu_frame_and_inbox_paths(gpu, tmp_path):
    import torch
    import torch.multiprocessing as mp
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    world = min(n, 4)
    mp.spawn(_rank_main, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    print((tmp_path / "ok").read_text())
