"""One-process-per-GPU sharding of a frame (torchrun) — host-side plumbing only.

The reference shards a frame by scanline with zero communication
(src/raytracer.nim:67-70: one WorkMsg per line pulled by worker threads).  Here
rank r of W renders the units — bands of T scanlines (rows of T x T screen-space
tiles) of a whole-resolution pass, rendered scanlines of a progressive pass —
that the serpentine deal nrt_unit_owner gives it (nrt_set_partition; unit_owner()
below is the same rule as rowsFor() in csrc/nrt.cu) and the float32 framebuffer
is assembled on rank 0 in one of two ways:

  * "ipc"  — rank 0 owns the device framebuffer, exports it with CUDA IPC
             (nrt_ipc_export) and every rank's final pixel-store kernel writes its
             rows straight into it over NVLink (nrt_render_device on the opened
             pointer): no gather pass, no NCCL.
  * "gather" — every rank renders into a local buffer and the rows are gathered
             with torch.distributed (NCCL on GPUs, gloo in the CPU tests).

For HOST framebuffers (the reference's Framebuf) SharedHostFramebuffer maps one
POSIX shared-memory buffer into every rank of the node: each rank's nrt_render copies
the scanlines it rendered into it over its own PCIe link, exactly as the reference's
worker threads write disjoint rows of one Framebuf (src/raytracer.nim:67-70) — no
device gather and no single-GPU 100 MB device-to-host copy.

torch.distributed is used for rendezvous, barriers, the max-over-ranks timing
reduction and (only in "gather" mode) the row gather.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np


def unit_owner(unit, world: int):
    """nrt_unit_owner (csrc/nrt.cu: unitOwner): the serpentine deal of units (bands / rendered scanlines) to ranks.
    Works on ints and numpy arrays."""
    r = unit % world
    return np.where((unit // world) % 2 == 1, world - 1 - r, r) if isinstance(unit, np.ndarray) else (world - 1 - r if (unit // world) % 2 else r)


def rows_of(rank: int, world: int, height: int, y0: int = 0, y1: Optional[int] = None, step: int = 1, band: int = 1) -> List[int]:
    """First rows of the units of [y0, y1) owned by `rank` (csrc/nrt.cu: rowsFor; the library's own answer is
    nrt_partition_rows / api.partitionRows, and tests/test_cabi.py holds the two equal).  A unit is a rendered scanline
    of a progressive pass ((y - y0) mod step == 0; band == 1) or a band of `band` scanlines of a whole-resolution
    pass (step == 1; band = api.bandRows(opts): rows of T x T tiles).  Units are numbered from y0 and dealt out
    in the serpentine order of unit_owner, so a progressive pass with step >= world still uses every rank.  """
    y1 = height if y1 is None else y1
    unit = step * band
    return [y for y in range(max(0, y0), min(y1, height)) if (y - y0) % unit == 0 and unit_owner((y - y0) // unit, world) == rank]


def owned_rows(rank: int, world: int, height: int, band: int = 1) -> np.ndarray:
    """Every scanline of a whole frame that `rank` renders: the bands of `band` rows that unit_owner gives it."""
    y = np.arange(height)
    return y[unit_owner(y // band, world) == rank]


def merge_rows(dst: np.ndarray, src: np.ndarray, rank: int, world: int, band: int = 1) -> None:
    """Copies the rows owned by `rank` from a full-size (H, ...) buffer into dst."""
    rows = owned_rows(rank, world, dst.shape[0], band)
    dst[rows] = src[rows]


def gather_rows(local, rank: int, world: int, dist=None, group=None, dst: int = 0, band: int = 1):
    """Gathers the band-interleaved pieces of a (H, W, C) torch tensor on rank `dst`.

    Every rank passes its full-size buffer of which only its own bands (owned_rows) are
    valid.  Works with any torch.distributed backend (nccl for CUDA tensors,
    gloo for CPU tensors).  Returns the assembled frame on `dst`, None elsewhere."""
    import torch

    if world == 1:
        return local
    H = local.shape[0]
    idx = [torch.as_tensor(owned_rows(r, world, H, band), device=local.device) for r in range(world)]
    per = max(int(i.numel()) for i in idx)
    mine = local[idx[rank]]
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: mine.shape[0]] = mine
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, gather_list=bufs, dst=dst, group=group)
    if rank != dst:
        return None
    out = torch.empty_like(local)
    for r in range(world):
        out[idx[r]] = bufs[r][: idx[r].numel()]
    return out


class PeerFramebuffer:
    """Device framebuffer owned by rank 0 and mapped into every rank with CUDA IPC."""

    def __init__(self, nbytes: int, rank: int, world: int, dist=None):
        from . import api

        self.api, self.rank, self.world, self.nbytes = api, rank, world, nbytes
        self.ptr = C.c_void_p()
        self.owner = rank == 0
        L = api.lib()
        if self.owner:
            api.check(L.nrt_device_alloc(nbytes, C.byref(self.ptr)), "nrt_device_alloc")
            api.check(L.nrt_device_memset(self.ptr, 0, nbytes), "nrt_device_memset")
        if world > 1:
            h = api.nrt_ipc_handle()
            if self.owner:
                api.check(L.nrt_ipc_export(self.ptr, C.byref(h)), "nrt_ipc_export")
            box = [bytes(h.bytes)]
            dist.broadcast_object_list(box, src=0)
            if not self.owner:
                h = api.nrt_ipc_handle()
                C.memmove(h.bytes, box[0], 64)
                api.check(L.nrt_ipc_open(C.byref(h), C.byref(self.ptr)), "nrt_ipc_open")

    def to_host(self, out: np.ndarray) -> None:
        self.api.check(self.api.lib().nrt_copy_to_host(out.ctypes.data_as(C.c_void_p), self.ptr, out.nbytes),
                       "nrt_copy_to_host")

    def close(self) -> None:
        L = self.api.lib()
        if self.ptr:
            if self.owner:
                L.nrt_device_free(self.ptr)
            else:
                L.nrt_ipc_close(self.ptr)
            self.ptr = C.c_void_p()


class SharedHostFramebuffer:
    """Host framebuffer (float32, nbytes) shared by the ranks of one node through /dev/shm and
    page-locked in every rank (nrt_host_register), so that each rank's nrt_render() delivers its
    own scanlines with a pinned 2-D copy.  `register=False` skips the page-locking (CPU tests).
    Raises on every rank if any rank failed, so that callers can fall back together."""

    def __init__(self, nbytes: int, rank: int, world: int, dist=None, register: bool = True):
        import mmap
        import os

        self.rank, self.world, self.nbytes = rank, world, nbytes
        self.path, self.mm, self.array, self.registered = None, None, None, False
        self.api = None
        box = [None]
        if rank == 0:
            path = f"/dev/shm/nrt_fb_{os.getpid()}_{id(self) & 0xFFFF:x}"
            try:
                fd = os.open(path, os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o600)
                try:
                    os.posix_fallocate(fd, 0, nbytes)      # ENOSPC here instead of SIGBUS on first touch
                    self.mm = mmap.mmap(fd, nbytes)
                finally:
                    os.close(fd)
                self.path = box[0] = path
            except OSError:
                if os.path.exists(path):
                    os.unlink(path)
        if world > 1:
            dist.broadcast_object_list(box, src=0)
        ok = box[0] is not None
        if ok and rank != 0:
            try:
                fd = os.open(box[0], os.O_RDWR)
                try:
                    self.mm = mmap.mmap(fd, nbytes)
                finally:
                    os.close(fd)
                self.path = box[0]
            except OSError:
                ok = False
        if ok:
            self.array = np.frombuffer(self.mm, dtype=np.float32)
            if register:
                from . import api

                self.api = api
                rc = api.lib().nrt_host_register(C.c_void_p(self.array.ctypes.data), nbytes)
                self.registered = rc == 0
                ok = self.registered
        if world > 1:
            flags = [None] * world
            dist.all_gather_object(flags, bool(ok))
            ok = all(flags)
        if not ok:
            self.close()
            raise RuntimeError("shared host framebuffer unavailable (/dev/shm space or cudaHostRegister)")

    @property
    def ptr(self) -> C.c_void_p:
        return C.c_void_p(self.array.ctypes.data)

    def close(self) -> None:
        import os

        if self.registered and self.array is not None:
            self.api.lib().nrt_host_unregister(C.c_void_p(self.array.ctypes.data))
            self.registered = False
        self.array = None
        if self.mm is not None:
            try:
                self.mm.close()
            except BufferError:      # a caller still holds a view: the mapping goes with the process
                pass
            self.mm = None
        if self.rank == 0 and self.path and os.path.exists(self.path):
            os.unlink(self.path)
        self.path = None
