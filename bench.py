#!/usr/bin/env python
"""bench.py — measures the render hot path on N B200s of one node.

  python bench.py --gpus N --steps K --warmup W            (N=1)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   (N>1)
  python bench.py --impl reference ...                      (the CPU restatement, host cores)

A "step" is one whole frame of the workload (default: BASELINE config 4 — the bunny + spheres
scene at 3840x2160, 16 spp, reflections to depth 8: the configuration BASELINE.json's target and
scaling requirement are quoted on; `--workload config2` selects the 1920x1080 1-spp bunny frame
used for the single-kernel roofline runs under profiles/).  The frame is
strong-scaled: the frame's bands (rows of T x T screen-space tiles, T scanlines high)
are dealt out to the N ranks in serpentine order (nrt_unit_owner) and every rank's
final pixel-store kernel writes its tiles into rank 0's device framebuffer over
NVLink (CUDA IPC peer pointer); `value` = rays traced by all ranks / max-over-ranks
device time.  `e2e` goes through the host-buffer C-ABI call (nrt_scene_update +
nrt_render): scene host->device and framebuffer device->host inside the timed
region.  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Mrays/sec"
UNIT = "Mrays/s"
NOMINAL_FP32_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12   # 74.45 at clocks.max.sm


def workload(name: str):
    from nim_raytracer_b200 import api, scenes
    if name == "config2":
        return scenes.bunny(), api.Options(1920, 1080), "bunny.geom 69,451 tris + plane, 2 distant lights, 1920x1080, 1 spp (BASELINE config 2)"
    if name == "config3":
        o = api.Options(1920, 1080, antialias=api.Antialias(api.akGrid, 4), depthMode=api.NRT_DEPTH_INTENDED, maxRayDepth=8)
        return scenes.bunny_spheres(), o, "bunny + 7 spheres (reflection 0/0.5/1), 1920x1080, 16 spp grid, depth 8 (BASELINE config 3)"
    if name == "config4":
        o = api.Options(3840, 2160, antialias=api.Antialias(api.akGrid, 4), depthMode=api.NRT_DEPTH_INTENDED, maxRayDepth=8)
        return scenes.bunny_spheres(), o, "bunny + 7 spheres, 3840x2160, 16 spp grid, depth 8 (BASELINE config 4)"
    if name == "config5":   # synthetic stress: 1M random triangles as one mesh + 10k spheres (10 % mirrors), 64 spp
        o = api.Options(3840, 2160, antialias=api.Antialias(api.akGrid, 8), depthMode=api.NRT_DEPTH_INTENDED, maxRayDepth=8)
        return scenes.stress(), o, "1M random triangles + 10k spheres + plane, 3840x2160, 64 spp grid, depth 8 (BASELINE config 5)"
    if name == "config5s":  # the same scene at 1/4 linear resolution and 16 spp (1/64 of the samples)
        o = api.Options(960, 540, antialias=api.Antialias(api.akGrid, 4), depthMode=api.NRT_DEPTH_INTENDED, maxRayDepth=8)
        return scenes.stress(), o, "1M random triangles + 10k spheres + plane, 960x540, 16 spp grid, depth 8 (BASELINE config 5 scene, 1/64 of its samples)"
    if name == "config1":
        return scenes.spheres_reflection(), api.Options(640, 480), "spheres-reflection.nim 640x480, 1 spp (BASELINE config 1)"
    if name == "tiny":  # CI-sized
        return scenes.bunny(stride=8), api.Options(320, 180), "decimated bunny 320x180 (smoke)"
    raise SystemExit(f"unknown workload {name}")


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (the clocks line of
    /opt/skills/guides/B200_PROFILING.md) through NVML in a background thread — an
    `nvidia-smi -lms` child process perturbs short frames and buffers its output."""

    def __init__(self, gpu_index: int, period_s: float = 0.01):
        self.gpu, self.period, self.rows, self.stop_flag, self.thread, self.err = gpu_index, period_s, [], False, None, None

    def prepare(self):
        """NVML init + the slow one-off queries, outside the timed region."""
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except Exception:
                    idx = self.gpu
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.err = str(e)
            self.nv = None

    def start(self):
        if getattr(self, "nv", None) is None:
            self.prepare()
        if self.nv is None:
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.rows.append((sm, self.max_mhz, rs, pw))
            except Exception as e:  # noqa: BLE001
                self.err = str(e)
                break
            time.sleep(self.period)

    def stop(self):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=1.0)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"nvml unavailable: {self.err}"], "samples": 0}
        bits = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        reasons = sorted({name for r in self.rows for b, name in bits.items() if r[2] & b})
        return {"sm_mhz": statistics.median(r[0] for r in self.rows), "sm_max_mhz": max(r[1] for r in self.rows),
                "reasons": reasons, "samples": len(self.rows), "power_w_max": max(r[3] for r in self.rows)}


def _cpu_sample(scene_desc, opts, target_s):
    """Bounded, representative CPU sample of the workload: the SAME scene, options and spp at 1/k of
    the linear resolution (every ray type in the same proportion; scanline sub-sampling is not
    representative because mesh rows cost ~1000x sky rows).  k is chosen from a probe so that one
    render takes about `target_s` seconds.  Returns (opts_small, k)."""
    import copy
    import oracle

    def small(k):
        o = copy.copy(opts)
        o.width, o.height = max(8, opts.width // k), max(8, opts.height // k)
        return o

    k = 32
    t = time.time()
    oracle.render(scene_desc, small(k), fast=True)
    dt = max(time.time() - t, 1e-3)
    # time ~ 1/k^2
    k_new = int(max(1, min(64, round(k * (dt / target_s) ** 0.5))))
    return small(k_new), k_new


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path timed on the host cores.
    The reference (Nim) cannot be built here, so this is the oracle port (oracle/ref_cpu.cpp,
    -ffast-math build mirroring src/nim.cfg:2), one scanline per work item over all host threads
    (src/raytracer.nim:61-70, concurrency/workerpool.nim:172)."""
    if rank != 0:
        return
    import oracle
    from nim_raytracer_b200 import api
    scene, opts, desc = workload(args.workload)
    sd = api.SceneDesc(scene)
    cores = oracle.hardware_threads()
    so, k = _cpu_sample(sd, opts, float(os.environ.get("NRT_REF_STEP_SECONDS", "5")))
    for _ in range(args.warmup):
        oracle.render(sd, so, fast=True)
    rays = 0
    t0 = time.time()
    for _ in range(args.steps):
        _, st, _ = oracle.render(sd, so, fast=True)
        rays += st.numRays
    el = time.time() - t0
    v = rays / el / 1e6
    full_rays_est = (rays / args.steps) * (opts.width * opts.height) / (so.width * so.height)
    sample = (f"the same scene/options at 1/{k} linear resolution ({so.width}x{so.height}, same spp) per step, "
              f"all {cores} host threads, float64 -O3 -ffast-math")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "note": "CPU restatement of the reference (Nim toolchain absent); a step renders the bounded sample described in cpu_baseline.sample"},
        "frames_per_s_extrapolated": v * 1e6 / full_rays_est,
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(scene_desc, opts, rays_per_frame, budget_s=12.0):
    import oracle
    cores = oracle.hardware_threads()
    so, k = _cpu_sample(scene_desc, opts, budget_s)
    t = time.time()
    _, st, _ = oracle.render(scene_desc, so, fast=True)
    el = time.time() - t
    v = st.numRays / el / 1e6
    return {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"the same scene/options at 1/{k} linear resolution ({so.width}x{so.height}, same spp), {el:.1f} s, "
                      f"oracle -O3 -ffast-math float64, all host threads",
            "frames_per_s_extrapolated": v * 1e6 / max(rays_per_frame, 1)}


# Modelled algorithmic bytes per sample a kernel family's launch works on (float64 SoA state, DESIGN.md section 6;
# two lights, one mesh object).  FusedBounce takes every active sample of a bounce; the wavefront kernels take the
# samples with a mesh ray (nrt_profile.wavefront_samples).
STATE_BYTES_PER_SAMPLE = {
    "FusedBounce": 25,     # W accum 24 + flag 1 (a sample handed to the wavefront: W ray 32 + flag 1; a continuing one: + ray,
                           # origin, weight 72) - the kernel is COMPUTE bound (float64 pipe / issue slots), see `compute`
    "gen+gate": 34,        # NRT_PATH=0 only: W rayD 32 + active 1 + gate code 1
    "Shade": 102,          # R rayD 32 + active 1 + code 1;  W hitObj 4 + hitW 32 + nrm 32 (hit) | accum 24 (miss)
    "k_gate_flags": 70,    # R hitObj 4 + hitW 32 + nrm 32;  W 2 codes (one per light)
    "ShadowTrace": 72,     # R hitObj 4 + hitW 32 + nrm 32 + 2 codes;  W 2 occlusion flags   (NRT_FUSE_RESOLVE=0)
    "ShadowResolve": 95,   # R hitObj 4 + hitW 32 + nrm 32 + 2 codes;  W accum 24 + active 1  (ShadowTrace + Resolve in one launch)
    "Resolve": 63,         # R hitObj 4 + nrm 32 + 2 flags;  W accum 24 + active 1
    "Finalize": 25,        # R accum 24;  W 12 bytes per pixel (16 samples)
}
WAVEFRONT_FAMILIES = ("Shade", "k_gate_flags", "ShadowTrace", "ShadowResolve", "Resolve")


def _hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"])
    except Exception:  # noqa: BLE001
        return 6459.0   # the pool's measured copy bandwidth (B200_PROFILING.md fallback)


def _traffic(workload_name, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, from the committed
    `ncu --set full` capture of this workload (profiles/kernel_traffic.json); None if not captured."""
    try:
        with open(os.path.join(ROOT, "profiles", "kernel_traffic.json")) as f:
            tab = json.load(f).get(workload_name, {})
        return next((v.get("dram_bytes_per_launch") for k, v in tab.items() if kernel.startswith(k)), None)
    except Exception:  # noqa: BLE001
        return None


def _traffic_entry(workload_name, kernel):
    """The committed ncu figures of one launch of `kernel` (profiles/kernel_traffic.json) or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "kernel_traffic.json")) as f:
            tab = json.load(f).get(workload_name, {})
        return next((v for k, v in tab.items() if kernel.startswith(k)), None)
    except Exception:  # noqa: BLE001
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    # Default workload = BASELINE config 4 (the configuration the north-star target and the 1->8 GPU
    # scaling requirement are quoted on); it fits one GPU, so every N renders the same frame.
    ap.add_argument("--workload", default=os.environ.get("NRT_WORKLOAD", "config4"))
    ap.add_argument("--gather", default="ipc", choices=["ipc", "gather"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    # one process driving every GPU (nrt_init(0) + host framebuffer): the way a single-process caller such as the
    # reference's front-end would use N GPUs; run WITHOUT torchrun:  python bench.py --gpus N --inproc
    ap.add_argument("--inproc", action="store_true")
    # L2 between timed steps: "stream" = none needed, every frame streams its per-sample state (GBs, see
    # config.l2 in the output) through HBM; "flush" = additionally write 256 MiB (> 126 MB L2) before every step.
    ap.add_argument("--l2", default=os.environ.get("NRT_BENCH_L2", "stream"), choices=["stream", "flush"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    from nim_raytracer_b200 import api, distributed as D

    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allreduce(v, op):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    L = api.lib()
    if args.inproc:
        if world != 1:
            raise SystemExit("--inproc is a single process: run it without torchrun")
        api.initRenderer(devices=list(range(args.gpus)))
    else:
        api.initRenderer(devices=[local_rank])
        # one process per GPU: run on the CPUs close to it (the render chain is dependent launches with host round
        # trips; NRT_BENCH_PIN=0 leaves the placement to the OS).  The library pins its own lane threads itself.
        if os.environ.get("NRT_BENCH_PIN", "1") != "0":
            api.pinToDevice(0)
    api.setPartition(rank, world)
    scene, opts, desc = workload(args.workload)
    ds = api.DeviceScene(scene)
    W, H = opts.width, opts.height
    fb_bytes = W * H * 3 * 4
    co = opts.to_c()

    peer = None
    local_t = None
    if args.gather == "ipc":
        try:
            peer = D.PeerFramebuffer(fb_bytes, rank, world, dist)
        except Exception as e:  # CUDA IPC unavailable in this container: gather with NCCL instead
            if rank == 0:
                print(f"[bench] CUDA IPC unavailable ({e}); falling back to --gather gather", file=sys.stderr)
            peer = None
    if peer is None:
        local_t = torch.zeros((H, W, 3), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    cs = api.nrt_stats()

    def step_resident():
        target = peer.ptr if peer is not None else C.c_void_p(local_t.data_ptr())
        api.check(L.nrt_render_device(ds.handle, C.byref(co), 0, H, 1, 1, target, C.byref(cs), None), "nrt_render_device")
        if peer is None and world > 1:
            return D.gather_rows(local_t, rank, world, dist, band=api.bandRows(opts))
        return None

    # ---- device-resident throughput ------------------------------------------------
    # W untimed warm-up steps, then keep warming until the GPU has been busy for ~0.4 s: a B200
    # that was idle runs its first tens of milliseconds below the sustained clock.
    for _ in range(args.warmup):
        flush.fill_(1)          # also warms torch's fill kernel (its first launch costs ~20 ms)
        step_resident()
    tw = time.perf_counter()
    extra_warm = 0
    while time.perf_counter() - tw < 0.4:
        step_resident()
        extra_warm += 1
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.prepare()
        barrier_free = True  # noqa: F841 (NVML init done before the timed region starts)
        sampler.start()
    prof_acc = {"mesh_ms": 0.0, "flops": 0.0, "launches": 0, "klaunches": 0, "tests": 0, "tests_ref": 0, "cand": 0, "pre": 0, "rays_mesh": 0, "frame_ms": 0.0}
    barrier()
    t0 = time.perf_counter()
    api.check(L.nrt_timer_begin(), "nrt_timer_begin")
    rays = 0
    dbg = []
    for _ in range(args.steps):
        ta = time.perf_counter()
        if args.l2 == "flush":
            flush.fill_(1)                  # L2 flush between timed iterations
            torch.cuda.synchronize()
        tb = time.perf_counter()
        step_resident()
        tc = time.perf_counter()
        dbg.append((tb - ta, tc - tb))
        rays += cs.num_rays
        p = ds.profile()
        prof_acc["mesh_ms"] += p.mesh_filter_ms; prof_acc["flops"] += p.fp32_flops
        prof_acc["launches"] += p.mesh_filter_launches; prof_acc["klaunches"] += p.kernel_launches
        prof_acc["tests"] += p.mesh_tests; prof_acc["tests_ref"] += p.mesh_tests_ref; prof_acc["cand"] += p.candidates; prof_acc["pre"] += p.pre_candidates
        prof_acc["rays_mesh"] += p.mesh_rays; prof_acc["frame_ms"] += p.total_ms
    ms_dev = C.c_double()
    api.check(L.nrt_timer_end(C.byref(ms_dev)), "nrt_timer_end")
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    if os.environ.get("NRT_BENCH_DEBUG") and rank == 0:
        print("[bench debug] per-step (flush+sync ms, render ms):", [(round(a * 1e3, 2), round(b * 1e3, 2)) for a, b in dbg[:12]],
              "frame_ms avg", prof_acc["frame_ms"] / args.steps, file=sys.stderr)
    OP = None
    if dist is not None:
        OP = dist.ReduceOp
    t_ms = allreduce(max(wall_ms, ms_dev.value), OP.MAX if OP else None)
    # per-rank view of the timed region (diagnostic: which rank the MAX comes from, host vs device time)
    rank_times = [{"rank": rank, "wall_ms_per_step": wall_ms / args.steps, "device_ms_per_step": ms_dev.value / args.steps}]
    if dist is not None:
        gathered = [None] * world
        dist.all_gather_object(gathered, rank_times[0])
        rank_times = gathered
    total_rays = allreduce(float(rays), OP.SUM if OP else None)
    klaunches = allreduce(float(prof_acc["klaunches"]), OP.SUM if OP else None)
    value = total_rays / (t_ms * 1e-3) / 1e6

    # ---- end to end: host scene -> device, framebuffer -> pinned host, every step ---
    hp = C.c_void_p()
    api.check(L.nrt_host_alloc_pinned(fb_bytes, C.byref(hp)), "nrt_host_alloc_pinned")
    host_fb = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_float)), shape=(W * H * 3,))
    h2d = sum(int(g.vertices.nbytes + g.normals.nbytes + g.vertexIdx.nbytes + g.normalIdx.nbytes)
              for g in {id(o.geometry): o.geometry for o in scene.objects if o.geometry.kind == api.NRT_GEOM_MESH}.values())
    h2d += len(scene.objects) * C.sizeof(api.nrt_object) + len(scene.lights) * C.sizeof(api.nrt_light) + 512

    pinned_inputs = api.pinSceneArrays(scene)   # the step's inputs are copied from pinned host memory
    # N > 1: one host framebuffer shared by the ranks of the node (POSIX shared memory, page-locked in every
    # rank); each rank's nrt_render copies the scanlines it rendered into it over its own PCIe link.  Fallback
    # (no /dev/shm space, registration refused): peer stores into rank 0's device buffer + one copy from there.
    shared = None
    if world > 1:
        try:
            shared = D.SharedHostFramebuffer(fb_bytes, rank, world, dist)
            host_fb = shared.array
        except Exception as e:  # noqa: BLE001
            if rank == 0:
                print(f"[bench] shared host framebuffer unavailable ({e}); rank 0 copies the gathered frame", file=sys.stderr)

    def step_e2e():
        ds.update()                                           # scene: host -> device (+ device-side precompute)
        if world == 1:
            api.check(L.nrt_render(ds.handle, C.byref(co), 0, H, 1, 1, hp, C.byref(cs), None), "nrt_render")
        elif shared is not None:
            api.check(L.nrt_render(ds.handle, C.byref(co), 0, H, 1, 1, shared.ptr, C.byref(cs), None), "nrt_render")
            barrier()
        else:
            out = step_resident()
            barrier()
            if rank == 0:
                if peer is not None:
                    peer.to_host(host_fb)
                else:
                    host_fb[:] = out.reshape(-1).cpu().numpy()

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    rays_e = 0
    for _ in range(args.steps):
        step_e2e()
        rays_e += cs.num_rays
    barrier()
    e_ms = allreduce((time.perf_counter() - t0) * 1e3, OP.MAX if OP else None)
    rays_e = allreduce(float(rays_e), OP.SUM if OP else None)
    e2e = {"value": rays_e / (e_ms * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": fb_bytes + 64,
           "ms_per_step": e_ms / args.steps, "frames_per_s": args.steps / (e_ms * 1e-3),
           "host_inputs": f"{len(pinned_inputs)} mesh arrays page-locked in place (nrt_host_register), small structs pageable",
           "host_framebuffer": ("pinned, one process" if world == 1 else
                                "shared by the ranks (POSIX shared memory, page-locked): every rank copies its own scanlines" if shared is not None
                                else "rank 0 copies the frame gathered on its device")}
    checksum = float(np.asarray(host_fb, dtype=np.float64).sum()) if rank == 0 else 0.0
    # the timed frame is the verified frame: sha256 of the float32 framebuffer the e2e steps delivered against the
    # CPU oracle's frame of the same workload (tests/golden/, generated by make_goldens.py)
    golden = {"config4": "config4_bunny_spheres_3840x2160_g4", "config3": "config3_bunny_spheres_1920x1080_g4",
              "config2": "config2_bunny_1920x1080", "config1": "config1_spheres_640x480"}.get(args.workload)
    parity = None
    if rank == 0 and golden and os.path.exists(os.path.join(ROOT, "tests", "golden", golden + ".npz")):
        import hashlib
        want = str(np.load(os.path.join(ROOT, "tests", "golden", golden + ".npz"))["fb_sha256"])
        got = hashlib.sha256(np.ascontiguousarray(host_fb).tobytes()).hexdigest()
        parity = {"golden": f"tests/golden/{golden}.npz", "fb_sha256": got, "matches_oracle": got == want}
        if got != want:
            raise SystemExit(f"bench: the rendered frame differs from the oracle's ({got} != {want})")

    # ---- per-kernel-family shares: one untimed frame with every launch bracketed by CUDA events ----
    api.setKernelTiming(True)
    step_resident()
    barrier()
    ktimes = ds.kernelTimes()
    kframe_ms = ds.profile().total_ms
    api.setKernelTiming(False)

    if rank == 0:
        peak = C.c_double(); clk = C.c_double(); peak64 = C.c_double()
        api.check(L.nrt_measure_fp32_peak(C.byref(peak), C.byref(clk)), "nrt_measure_fp32_peak")
        L.nrt_measure_fp64_peak.argtypes = [C.POINTER(C.c_double)]
        api.check(L.nrt_measure_fp64_peak(C.byref(peak64)), "nrt_measure_fp64_peak")
        # --inproc: the library's profile sums tests / flops / samples over the process's devices while its times are one
        # device's (the slowest): roofline figures are PER DEVICE, so the sums are divided by the device count
        ndev = float(args.gpus) if args.inproc else 1.0
        achieved = prof_acc["flops"] / ndev / max(prof_acc["mesh_ms"] * 1e-3, 1e-12) / 1e12
        intersection = {
            "bound": "fp32", "kernel": "k_mesh_prefilter", "achieved": achieved, "peak": peak.value, "unit": "TFLOP/s",
            "frac": achieved / peak.value if peak.value > 0 else None,
            "peak_source": "FFMA micro-kernel measured in this run (MEASURED_PEAKS.json has no FP32 entry)",
            "peak_nominal": NOMINAL_FP32_TFLOPS, "frac_of_nominal": achieved / NOMINAL_FP32_TFLOPS,
            "traffic": _traffic(args.workload, "k_mesh_prefilter"),
            "flops_per_test": prof_acc["flops"] / max(prof_acc["tests"], 1), "tests_per_step": prof_acc["tests"] / args.steps,
            "ref_tests_per_step": prof_acc["tests_ref"] / args.steps,
            "avg_launch_ms": prof_acc["mesh_ms"] / max(prof_acc["launches"], 1),
            "gtests_per_s": prof_acc["tests"] / ndev / max(prof_acc["mesh_ms"] * 1e-3, 1e-12) / 1e9,
            "candidates_per_step": prof_acc["cand"] / args.steps, "pre_candidates_per_step": prof_acc["pre"] / args.steps,
            "mesh_rays_per_step": prof_acc["rays_mesh"] / args.steps,
            "note": "achieved = executed float32 flops of the prefilter launches (FFMA = 2) / their summed CUDA-event time (the lanes' "
                    "launches overlap in the timed frames, so this understates the kernel alone; `kernels` below times one lane); "
                    "ref_tests = rays x all faces, what geom.nim:346 would evaluate (reported, never used for the fraction)",
        }
        ksum = sum(ms for ms, _ in ktimes.values()) or 1.0
        samples = float(cs.num_primary_rays) / ndev   # primary samples this rank (--inproc: one device) rendered in the last frame
        hbm_peak = _hbm_peak()
        kprof = ds.profile()                   # of the kernel-timing frame (one lane)
        wf0 = float(kprof.wavefront_samples[0]) / ndev
        kernels = []
        frame_bytes = 0.0
        for name, (ms, nl) in sorted(ktimes.items(), key=lambda kv: -kv[1][0]):
            k = {"kernel": name, "ms": ms, "share": ms / ksum, "launches": nl, "longest_launch_ms": ds.kernelMaxMs.get(name)}
            key = next((key for key in STATE_BYTES_PER_SAMPLE if name.startswith(key)), None)
            if key is not None:   # streaming kernels over the per-sample state: modelled algorithmic bytes
                bps = STATE_BYTES_PER_SAMPLE[key]
                # samples of the family's LONGEST launch (bounce 0): every sample for FusedBounce / Finalize, the samples with a
                # mesh ray for the wavefront kernels; all its launches: the active / wavefront samples of every bounce
                n1 = wf0 if key in WAVEFRONT_FAMILIES else samples
                nall = (sum(kprof.wavefront_samples) / ndev if key in WAVEFRONT_FAMILIES else
                        sum(kprof.active_samples) / ndev if key == "FusedBounce" else samples)
                lms = k["longest_launch_ms"] or ms
                gbs = bps * n1 / max(lms * 1e-3, 1e-12) / 1e9
                k.update({"bound": "hbm", "algorithmic_bytes_longest_launch": bps * n1, "achieved_GBps": gbs, "frac_of_hbm_peak": gbs / hbm_peak,
                          "samples_longest_launch": n1})
                frame_bytes += bps * nall
            kernels.append(k)
        pre_ms = ktimes.get("k_mesh_prefilter", (0.0, 0))[0]
        intersection["kernel_share_of_step"] = pre_ms / ksum
        intersection["share_source"] = "CUDA events around every launch of one untimed single-lane frame (nrt_set_kernel_timing); 'kernels' lists every family"
        if pre_ms > 0 and kprof.mesh_filter_ms > 0:
            # the kernel ALONE (the single-lane kernel-timing frame: nothing else in flight) is the roofline figure — in the timed
            # frames the lanes' launches overlap each other and other kernels, so their summed event time counts shared time twice
            alone = kprof.fp32_flops / ndev / (kprof.mesh_filter_ms * 1e-3) / 1e12
            intersection.update({"achieved_overlapped_sum": intersection["achieved"], "achieved": alone,
                                 "frac": alone / peak.value if peak.value > 0 else None, "frac_of_nominal": alone / NOMINAL_FP32_TFLOPS,
                                 "avg_launch_ms": kprof.mesh_filter_ms / max(kprof.mesh_filter_launches, 1),
                                 "gtests_per_s": kprof.mesh_tests / ndev / (kprof.mesh_filter_ms * 1e-3) / 1e9})
            if ndev > 1:
                intersection["per_device"] = "tests and flops of the process's devices / device count (times are the slowest device's)"
        # `roofline` = the dominant kernel family of the step by measured share; per LAUNCH (its longest = bounce-0 launch)
        dom = kernels[0]
        if dom["kernel"].startswith("k_mesh_prefilter"):
            roofline = dict(intersection)
        else:
            tr = _traffic_entry(args.workload, dom["kernel"])
            roofline = {
                "bound": "hbm", "kernel": dom["kernel"], "achieved": dom.get("achieved_GBps"), "peak": hbm_peak, "unit": "GB/s",
                "frac": dom.get("frac_of_hbm_peak"), "traffic": tr.get("dram_bytes_per_launch") if tr else None,
                "algorithmic_bytes_per_launch": dom.get("algorithmic_bytes_longest_launch"),
                "launch": "the family's longest launch of the frame = its bounce-0 launch over every primary sample",
                "launch_ms": dom.get("longest_launch_ms"), "kernel_share_of_step": dom["share"],
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy bandwidth measured on this pool)",
            }
            if dom["kernel"].startswith("FusedBounce"):
                # The kernel keeps a sample's whole bounce in registers: it writes 25 bytes per sample and is bound by the
                # float64 pipe and issue slots (the reference's float64 arithmetic, operation by operation).  Its compute
                # roofline: float64 thread-instructions per launch (ncu smsp__inst_executed_pipe_fp64 of the same launch,
                # profiles/kernel_traffic.json) / launch time, against the DFMA micro-kernel measured in this run.
                comp = {"bound": "fp64 pipe", "peak": peak64.value / 2.0, "unit": "T float64 instr/s (thread level)",
                        "peak_source": "register-resident DFMA micro-kernel measured in this run (TFLOP/s / 2)"}
                if tr and tr.get("fp64_thread_inst_per_launch") and dom.get("longest_launch_ms"):
                    scale = samples / float(tr.get("samples_per_launch", samples))
                    ach = tr["fp64_thread_inst_per_launch"] * scale / (dom["longest_launch_ms"] * 1e-3) / 1e12
                    comp.update({"achieved": ach, "frac": ach / comp["peak"] if comp["peak"] > 0 else None,
                                 "fp64_thread_inst_per_sample": tr["fp64_thread_inst_per_launch"] / float(tr.get("samples_per_launch", samples)),
                                 "issue_slot_frac_ncu": tr.get("issue_active_frac"), "fp64_pipe_frac_ncu": tr.get("fp64_pipe_frac")})
                roofline["compute"] = comp
                roofline["note"] = ("FusedBounce replaces four HBM-streaming kernels (r01: 43 GB of per-sample state per frame) by one "
                                    "register-resident kernel: it is NOT memory bound (frac = its 25 algorithmic bytes per sample against "
                                    "the HBM peak); what bounds it is `compute` - the float64 pipe and issue slots")
        l2_note = ("256 MiB device fill between timed steps (inside the timed region)" if args.l2 == "flush" else
                   f"inputs larger than L2: every frame streams {frame_bytes / 1e9:.1f} GB of per-sample state per GPU through HBM "
                   f"(>= {frame_bytes / 126e6:.0f}x the 126 MB L2), so nothing a step reads survives from the previous one except the "
                   "scene records (L2-resident within a step as well); --l2 flush adds a 256 MiB fill per step")
        band = api.bandRows(opts)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus if args.inproc else world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64 (reference order) behind f32 first looks", "data": "synthetic",
            "config": {"workload": desc,
                       "parallelism": (f"bands of {band} scanlines (rows of {band}x{band} screen-space tiles) dealt out in serpentine order x"
                                       f"{args.gpus if args.inproc else world} GPUs, up to 4 concurrent lanes per GPU; "
                                       + ("one process, peer stores into device 0" if args.inproc else
                                          ("CUDA-IPC peer stores" if peer is not None else "NCCL row gather") + " to rank 0")),
                       "l2": l2_note,
                       "warmup_extra_steps": extra_warm},
            "frames_per_s": args.steps / (t_ms * 1e-3), "rays_per_frame": total_rays / args.steps,
            "device_ms_per_step": ms_dev.value / args.steps,
            "ranks": rank_times,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(klaunches),
            "roofline": roofline, "intersection_kernel": intersection, "kernels": kernels, "kernel_timing_frame_ms": kframe_ms,
            "frame_hbm_bytes_model": frame_bytes,
            "fused_path": {"active_samples_per_bounce": list(kprof.active_samples), "wavefront_samples_per_bounce": list(kprof.wavefront_samples),
                           "tail_samples": int(kprof.tail_samples)},
            "fb_checksum": checksum, "parity": parity,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(ds.desc, opts, total_rays / args.steps)
        print(json.dumps(line), flush=True)

    api.unpinSceneArrays(pinned_inputs)
    del host_fb
    if shared is not None:
        barrier()
        shared.close()
    L.nrt_host_free_pinned(hp)
    if peer is not None:
        barrier()
        peer.close()
    ds.close()
    api.shutdown()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
