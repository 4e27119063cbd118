"""N>1 host logic on CPU: world_size-2 gloo processes shard a frame by scanline exactly like
nrt_set_partition / rowsFor (csrc/nrt.cu) and gather it on rank 0 (distributed.gather_rows).
The renderer stand-in is the test-only emulation of the device code (no GPU here)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import emu_binding as emu
    from nim_raytracer_b200 import api, distributed as D, scenes
    sc, o = scenes.bunny(stride=32), api.Options(64, 37)      # odd height: ragged shards
    fb = api.newFramebuf(o.width, o.height)
    rays = 0
    band = api.bandRows(o)                                    # 1 spp: bands of 16 scanlines (rows of 16 x 16 tiles)
    assert band == 16
    for y in D.rows_of(rank, world, o.height, band=band):
        _, st, _, _ = emu.render(sc, o, fb=fb, y0=y, y1=min(y + band, o.height))
        rays += st.numRays
    t = torch.from_numpy(fb.image().copy())
    full = D.gather_rows(t, rank, world, dist, band=band)
    # the same frame through the shared host framebuffer: every rank writes its own rows, no gather
    shared = D.SharedHostFramebuffer(o.width * o.height * 12, rank, world, dist, register=False)
    img = shared.array.reshape(o.height, o.width, 3)
    D.merge_rows(img, fb.image(), rank, world, band=band)
    dist.barrier()
    if rank == 0:
        assert (img == full.numpy()).all()
    del img
    dist.barrier()
    shared.close()
    tot = torch.tensor([rays], dtype=torch.int64)
    dist.all_reduce(tot)
    ms = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)                 # the max-over-ranks timing reduction
    assert ms.item() == float(world)
    if rank == 0:
        np.savez(out_path, fb=full.numpy(), rays=tot.item())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_scanline_sharding(tmp_path, oracle_mod):
    import emu_binding as emu
    emu.build()
    out = str(tmp_path / "gathered.npz")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    from nim_raytracer_b200 import api, scenes
    sc, o = scenes.bunny(stride=32), api.Options(64, 37)
    rfb, rst, _ = oracle_mod.render(sc, o)
    g = np.load(out)
    assert (g["fb"].reshape(-1) == rfb.data).all()
    assert int(g["rays"]) == rst.numRays


def test_rows_partition_properties():
    from nim_raytracer_b200 import distributed as D
    for world in (1, 2, 3, 4, 8):
        for (h, y0, y1, step) in ((1080, 0, None, 1), (37, 3, 30, 1), (64, 0, None, 4), (5, 0, None, 1)):
            seen = []
            for r in range(world):
                seen += D.rows_of(r, world, h, y0, y1, step)
            want = [y for y in range(max(0, y0), h if y1 is None else min(y1, h)) if (y - y0) % step == 0]
            assert sorted(seen) == want and len(set(seen)) == len(seen)
        # whole-resolution passes: bands of T scanlines (rows of T x T tiles) dealt out round-robin
        for (h, band) in ((2160, 4), (37, 16), (5, 2)):
            starts = sorted(sum((D.rows_of(r, world, h, band=band) for r in range(world)), []))
            assert starts == list(range(0, h, band))
            cover = np.concatenate([D.owned_rows(r, world, h, band) for r in range(world)])
            assert sorted(cover.tolist()) == list(range(h))
            for r in range(world):
                assert set(D.owned_rows(r, world, h, band).tolist()) == {y for s0 in D.rows_of(r, world, h, band=band) for y in range(s0, min(s0 + band, h))}
