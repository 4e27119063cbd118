"""Random-scene parity fuzzer (CPU; --gpu: the C ABI on a B200): the device code's per-element bodies and host orchestration (through the
TEST-ONLY emulation build, tests/emu) against the oracle, bit for bit — ids, t_hit, float32 framebuffer, Stats.

    python tools/fuzz_emu.py [--seeds A:B] [--jobs N] [--big] [--gpu]
    NRT_FUZZ_PATH=2 ...            every seed on one NRT_PATH (2 = PathMega, which the seeds never draw)
    NRT_FUZZ_EXTRA=K=V,K=V ...     further environment for every seed
    NRT_FUZZ_LOG=file ...          "start seed" / "done seed" lines (a seed that never returns is the one without "done")

Seeds below 2^20 are ordinary scenes; seed 2^20 + s is the scene of s with degenerate parts worked in (_degenerate).
Every case draws a scene (spheres / planes / boxes / meshes under random affine transforms, distant + point lights,
a random camera), render options (resolution, antialiasing kind, depth mode, maxRayDepth, bias) and a setting of the
library's path knobs (NRT_PATH, NRT_HARD_TAIL_BELOW, NRT_TILE, NRT_CHUNK_SAMPLES, NRT_CAND_CAP ...) from its seed.
A failing seed is printed with what differed; `case(seed)` rebuilds it (tests/test_emu_parity.py replays a fixed set).
Test infrastructure: nothing here is on a product path.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from nim_raytracer_b200 import api, linalg as L, scenes  # noqa: E402
from nim_raytracer_b200.api import (DistantLight, Material, Object, PointLight, Scene, initBox, initPlane,  # noqa: E402
                                    initSphere, point, vec, vec3)
from nim_raytracer_b200.loaders import trianglesToMesh  # noqa: E402

AXES = (L.X_AXIS, L.Y_AXIS, L.Z_AXIS)


def _xf(rng, centre, spread, affine_only_translate=False):
    t = centre + rng.uniform(-1.0, 1.0, 3) * spread
    m = L.translate(L.mat4(1.0), vec3(*t))
    if affine_only_translate:
        return m
    for _ in range(int(rng.integers(0, 3))):
        m = L.rotate(m, AXES[int(rng.integers(0, 3))], L.deg_to_rad(float(rng.uniform(-180.0, 180.0))))
    if rng.random() < 0.5:
        m = L.scale(m, tuple(float(x) for x in rng.uniform(0.4, 1.8, 3)))
    return m


def _mesh(rng, centre):
    kind = rng.random()
    if kind < 0.5:       # a decimated bunny (flipped or native winding)
        stride = int(rng.choice([16, 32, 64, 128, 256]))
        if BIG:
            stride = max(2, stride // 8)
        tri = scenes.bunny_triangles(flip_winding=bool(rng.random() < 0.7), stride=stride)
        spread = np.array([3.0, 0.5, 3.0])
    elif kind < 0.8:     # a soup of small random triangles
        n = int(rng.integers(1, 1500))
        c = rng.uniform(-2.0, 2.0, (n, 1, 3))
        tri = c + rng.uniform(-0.3, 0.3, (n, 3, 3))
        spread = np.array([3.0, 1.0, 3.0])
    else:                # a few large triangles (slivers included)
        n = int(rng.integers(1, 12))
        tri = rng.uniform(-3.0, 3.0, (n, 3, 3))
        if rng.random() < 0.3:
            tri[:, 2] = tri[:, 1] + rng.uniform(-1e-3, 1e-3, (n, 3))
        spread = np.array([3.0, 1.0, 3.0])
    g = trianglesToMesh(np.ascontiguousarray(tri, dtype=np.float64))
    g.objectToWorld = _xf(rng, centre, spread, affine_only_translate=bool(rng.random() < 0.5))
    g.worldToObject = L.inverse(g.objectToWorld)
    return g


BIG = os.environ.get("NRT_FUZZ_BIG") == "1"   # `--big`: 3x the resolution, 8x denser bunnies (seconds per seed)
DEGENERATE = 1 << 20   # seeds from here on: the scene of seed - DEGENERATE with degenerate parts worked in (_degenerate)


def _degenerate(rng, sc):
    """Inputs whose arithmetic leaves the comfortable range — the reference has no validation, so IEEE rules define the
    result (NaN compares false, inf propagates) and the float32 first looks must hand over without changing a bit."""
    from nim_raytracer_b200.api import NRT_GEOM_BOX, NRT_GEOM_MESH, NRT_GEOM_SPHERE
    for _ in range(int(rng.integers(1, 4))):
        k = int(rng.integers(0, 7))
        want = {0: NRT_GEOM_SPHERE, 1: NRT_GEOM_BOX, 2: NRT_GEOM_MESH}.get(k)
        pool = [ob for ob in sc.objects if want is None or ob.geometry.kind == want] or sc.objects
        o = pool[int(rng.integers(0, len(pool)))]
        g = o.geometry
        if k == 0 and g.kind == NRT_GEOM_SPHERE:
            g.r = float(rng.choice([0.0, -0.7, 1e-160, 1e160, 1e30]))
        elif k == 1 and g.kind == NRT_GEOM_BOX:
            a = int(rng.integers(0, 3))
            if rng.random() < 0.5:
                g.vmin[a], g.vmax[a] = g.vmax[a], g.vmin[a]          # inverted slab
            else:
                g.vmax[a] = g.vmin[a]                                # zero thickness
        elif k == 2 and g.kind == NRT_GEOM_MESH:
            v = g.vertices
            n = v.shape[0]
            i = rng.integers(0, n, max(1, n // 50))
            m = int(rng.integers(0, 4))
            if m == 0:
                v[i, :3] = v[(i + 1) % n, :3]                        # zero-area faces
            elif m == 1:
                v[i, int(rng.integers(0, 3))] = float(rng.choice([1e30, -1e30, 1e38, 1e200]))
            elif m == 2:
                v[i, :3] *= 1e-30
            else:
                v[i, int(rng.integers(0, 3))] = float(rng.choice([np.inf, -np.inf, np.nan]))
        elif k == 3:
            m = np.array(g.objectToWorld, dtype=np.float64, copy=True)
            a = int(rng.integers(0, 3))
            m[:3, a] *= float(rng.choice([0.0, 1e-200, 1e150, -1.0]))   # singular / tiny / huge / mirrored
            g.objectToWorld = m
            with np.errstate(all="ignore"):
                g.worldToObject = L.inverse(m)
        elif k == 4:
            # (inside [0, 1] or negative = off.  Outside that range the recursion's (1-k)*local + k*refl and the device's
            # path weights are no longer sums of same-signed terms: cancellation magnifies their float64 rounding
            # difference into visible bits, and products of two huge k overflow differently)
            o.material.reflection = float(rng.choice([-0.5, 1.0, 1e-9, 0.999999]))   # (1e-300: seed 2^20 + 10287 — two of them underflow the weight to 0, and 0 * an infinite albedo is nan where the recursion keeps inf)
        elif k == 5:
            o.material.albedo = vec3(*rng.choice([0.0, -1.0, 1e30, np.inf], 3))
        elif k == 6 and sc.lights:
            li = sc.lights[int(rng.integers(0, len(sc.lights)))]
            if isinstance(li, DistantLight):
                li.dir = vec(*rng.choice([0.0, 0.0, 1.0, -1.0, 1e-200], 3))   # unnormalised, possibly zero
            else:
                li.pos = point(*sc.objects[0].geometry.objectToWorld[:3, 3])  # a light inside / on an object
            li.intensity = float(rng.choice([li.intensity, 0.0, -1.0, 1e300]))
    if rng.random() < 0.2:
        sc.fov = float(rng.choice([1e-6, 179.999, 180.0, 0.0, 360.0]))
    return sc


def case(seed: int):
    """(scene, options, env) of one seed."""
    if seed >= DEGENERATE:
        sc, opts, env = case(seed - DEGENERATE)
        return _degenerate(np.random.default_rng(seed), sc), opts, env
    rng = np.random.default_rng(seed)
    centre = np.array([0.0, 1.5, -11.0])
    nobj = int(rng.choice([1, 2, 3, 5, 8, 12, 20, 33, 40, 70]))
    nmesh_max = int(rng.choice([0, 1, 1, 2, 4]))
    objects, nmesh = [], 0
    for i in range(nobj):
        k = rng.random()
        mat = Material(albedo=vec3(*rng.uniform(0.1, 1.0, 3)),
                       reflection=float(rng.choice([0.0, 0.0, 0.3, 0.7, 1.0])))
        if nmesh < nmesh_max and k < 0.25:
            objects.append(Object(f"m{i}", _mesh(rng, centre), mat))
            nmesh += 1
        elif k < 0.6:
            objects.append(Object(f"s{i}", initSphere(r=float(rng.uniform(0.2, 1.6)),
                                                      objectToWorld=_xf(rng, centre, np.array([6.0, 2.0, 6.0]),
                                                                        bool(rng.random() < 0.6))), mat))
        elif k < 0.75:
            m = L.translate(L.mat4(1.0), vec3(0.0, float(rng.uniform(-1.0, 0.2)), 0.0))
            if rng.random() < 0.4:
                m = L.rotate(m, AXES[int(rng.choice([0, 2]))], L.deg_to_rad(float(rng.uniform(-8.0, 8.0))))
            objects.append(Object(f"p{i}", initPlane(objectToWorld=m), mat))
        else:
            lo = -rng.uniform(0.2, 1.2, 3)
            hi = rng.uniform(0.2, 1.2, 3)
            objects.append(Object(f"b{i}", initBox(vmin=vec(*lo), vmax=vec(*hi),
                                                   objectToWorld=_xf(rng, centre, np.array([6.0, 2.0, 6.0]))), mat))
    nl = int(rng.choice([0, 1, 1, 2, 2, 3, 5, 33]))
    lights = []
    for _ in range(nl):
        col = vec3(*rng.uniform(0.2, 1.0, 3))
        if rng.random() < 0.6:
            d = vec(float(rng.uniform(-1.0, 1.0)), float(rng.uniform(-1.5, -0.2)), float(rng.uniform(-1.0, 1.0)))
            lights.append(DistantLight(color=col, intensity=float(rng.uniform(0.2, 2.0)), dir=L.normalize(d)))
        else:
            p = centre + np.array([rng.uniform(-6, 6), rng.uniform(2, 9), rng.uniform(-6, 6)])
            lights.append(PointLight(color=col, intensity=float(rng.uniform(200.0, 3000.0)), pos=point(*p)))
    cam = L.mat4(1.0)
    cam = L.translate(cam, vec3(float(rng.uniform(-3, 3)), float(rng.uniform(1.0, 6.0)), float(rng.uniform(0.0, 4.0))))
    cam = L.rotate(cam, L.Y_AXIS, L.deg_to_rad(float(rng.uniform(-15.0, 15.0))))
    cam = L.rotate(cam, L.X_AXIS, L.deg_to_rad(float(rng.uniform(-25.0, 5.0))))
    sc = Scene(objects=objects, lights=lights, fov=float(rng.uniform(30.0, 80.0)), cameraToWorld=cam,
               bgColor=vec3(*rng.uniform(0.0, 0.4, 3)))
    w, h = int(rng.integers(8, 97)), int(rng.integers(4, 65))
    if BIG:       # (frames whose waves exceed the tail thresholds, so that the library's DEFAULT routing is what runs)
        w, h = 3 * w + 64, 3 * h + 36
    if nl > 32:   # (34 rays per sample and bounce through the single-threaded emulation: keep these frames small)
        w, h = min(w, 40), min(h, 30)
    aa_kind = int(rng.choice([api.akNone, api.akNone, api.akGrid, api.akGrid, api.akJittered, api.akMultiJittered,
                              api.akCorrelatedMultiJittered]))
    grid = 1 if aa_kind == api.akNone else int(rng.integers(1, 4))
    opts = api.Options(w, h, antialias=api.Antialias(aa_kind, grid),
                       bias=float(rng.choice([1e-8, 1e-8, 1e-4, 0.0])),
                       maxRayDepth=int(rng.choice([0, 1, 2, 5, 8])),
                       depthMode=int(rng.choice([api.NRT_DEPTH_REFBUG, api.NRT_DEPTH_INTENDED])),
                       bounceCap=int(rng.choice([64, 64, 64, 3])), seed=int(rng.integers(0, 1 << 30)))
    env = {"NRT_HARD_TAIL_BELOW": str(int(rng.choice([0, 0, 16384, 64])))}
    if rng.random() < 0.6:
        env["NRT_PATH"] = str(int(rng.choice([0, 1])))
    if rng.random() < 0.3:
        env["NRT_TAIL_BELOW"] = str(int(rng.choice([0, 256, 32768])))
    if rng.random() < 0.3:
        env["NRT_TILE"] = str(int(rng.choice([1, 2, 4, 8])))
    if rng.random() < 0.3:
        env["NRT_CHUNK_SAMPLES"] = str(int(rng.choice([256, 1000, 4096])))
    if rng.random() < 0.15:
        env["NRT_CAND_CAP"] = str(int(rng.choice([1, 64, 1024])))
    if rng.random() < 0.15:
        env["NRT_PAIR_CAP"] = str(int(rng.choice([1, 16, 256])))
    if rng.random() < 0.1 and nl <= 32:   # (every face for each of 34 rays per sample and bounce: minutes in the emulation)
        env["NRT_FORCE_EXACT"] = "1"
    if rng.random() < 0.15:
        env["NRT_PREFILTER_CULL"] = "0"
    if rng.random() < 0.15:
        env["NRT_HOT_HEADER"] = "0"
    if rng.random() < 0.15:
        env["NRT_FUSE_RESOLVE"] = "0"
    if rng.random() < 0.1:
        env["NRT_SHADOW_GATE_PER_SAMPLE"] = "0"
    if rng.random() < 0.1:
        env["NRT_SHADOW_TRACE_PER_SAMPLE"] = "0"
    return sc, opts, env


KNOBS = ("NRT_HARD_TAIL_BELOW", "NRT_PATH", "NRT_TAIL_BELOW", "NRT_TILE", "NRT_CHUNK_SAMPLES", "NRT_CAND_CAP", "NRT_PAIR_CAP",
         "NRT_FORCE_EXACT", "NRT_PREFILTER_CULL", "NRT_HOT_HEADER", "NRT_FUSE_RESOLVE", "NRT_SHADOW_GATE_PER_SAMPLE",
         "NRT_SHADOW_TRACE_PER_SAMPLE")


def _passes_and_ranges(seed, sc, opts, whole):
    """renderLine*'s step / maxStep passes (renderer.nim:162-211) from maxStep down to 1 into one framebuffer, and the
    frame as ragged line ranges: after every pass the emulation's framebuffer equals the oracle's; the last pass and the
    union of the ranges equal the whole-frame render."""
    import emu_binding as emu
    import oracle
    rng = np.random.default_rng(seed + (1 << 40))
    bad = []
    ms = int(rng.choice([2, 4, 8]))
    fb, rfb = api.newFramebuf(opts.width, opts.height), api.newFramebuf(opts.width, opts.height)
    step = ms
    while step >= 1:
        emu.render(sc, opts, fb=fb, step=step, maxStep=ms)
        oracle.render(sc, opts, fb=rfb, step=step, maxStep=ms)
        if not (fb.data.view(np.uint32) == rfb.data.view(np.uint32)).all():
            bad.append(f"pass step {step}/{ms}")
        step //= 2
    if not (fb.data.view(np.uint32) == whole.data.view(np.uint32)).all():
        bad.append(f"passes {ms}..1 != whole frame")
    fb2 = api.newFramebuf(opts.width, opts.height)
    y = 0
    while y < opts.height:
        n = int(rng.integers(1, 9))
        emu.render(sc, opts, fb=fb2, y0=y, y1=min(y + n, opts.height))
        y += n
    if not (fb2.data.view(np.uint32) == whole.data.view(np.uint32)).all():
        bad.append("line ranges != whole frame")
    return bad


def run(seed: int, render=None) -> str | None:
    """None when the emulated device path and the oracle agree on every bit; else what differed.
    `render(scene, opts, aov)` -> (fb, stats) replaces the emulation (the -m gpu replay passes the C-ABI call)."""
    import emu_binding as emu
    import oracle
    sc, opts, env = case(seed)
    saved = {k: os.environ.get(k) for k in KNOBS}
    try:
        for k in KNOBS:
            os.environ.pop(k, None)
        os.environ.update(env)
        if os.environ.get("NRT_FUZZ_PATH"):   # (every seed on one path, e.g. 2 = PathMega, which the seeds themselves never draw)
            os.environ["NRT_PATH"] = os.environ["NRT_FUZZ_PATH"]
        for kv in filter(None, os.environ.get("NRT_FUZZ_EXTRA", "").split(",")):   # e.g. NRT_FUZZ_EXTRA=NRT_FORK_MIN=1
            k, v = kv.split("=", 1)
            os.environ[k] = v
        a1, a2 = api.Aov(opts.width, opts.height), api.Aov(opts.width, opts.height)
        rfb, rst, _ = oracle.render(sc, opts, aov=a1)
        try:
            if render is None:
                fb, st, _, _ = emu.render(sc, opts, aov=a2)
            else:
                fb, st = render(sc, opts, a2)
        except Exception as e:  # noqa: BLE001
            return f"render failed: {e}"
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    bad = []
    if render is None and seed % 3 == 0:
        os.environ.update(env)
        try:
            bad += _passes_and_ranges(seed, sc, opts, rfb)
        finally:
            for k in env:
                os.environ.pop(k, None)
            os.environ.update({k: v for k, v in saved.items() if v is not None})
    if not (a1.obj_id == a2.obj_id).all():
        bad.append(f"obj_id x{int((a1.obj_id != a2.obj_id).sum())}")
    if not (a1.tri_id == a2.tri_id).all():
        bad.append(f"tri_id x{int((a1.tri_id != a2.tri_id).sum())}")
    if not (a1.t_hit.view(np.uint64) == a2.t_hit.view(np.uint64)).all() and not (a1.t_hit == a2.t_hit).all():
        bad.append("t_hit")
    if not (fb.data.view(np.uint32) == rfb.data.view(np.uint32)).all():
        bad.append(f"fb x{int((fb.data.view(np.uint32) != rfb.data.view(np.uint32)).sum())}")
    if st != rst:
        bad.append(f"stats {st} != {rst}")
    return ", ".join(bad) if bad else None


def _gpu_render(sc, opts, aov):
    """The product path: nrt_render through the C ABI on cuda:0 (needs a B200; `--gpu`)."""
    fb = api.newFramebuf(opts.width, opts.height)
    ds = api.DeviceScene(sc)
    try:
        st = api.renderFrame(ds, opts, fb, aov=aov)
    finally:
        ds.close()
    return fb, st


def _job(seed):
    log = os.environ.get("NRT_FUZZ_LOG")   # (a seed that never returns is the one with "start" and no "done")
    if log:
        with open(log, "a") as f:
            f.write(f"start {seed}\n")
    try:
        return seed, run(seed, render=_gpu_render if os.environ.get("NRT_FUZZ_GPU") == "1" else None)
    except Exception as e:  # noqa: BLE001
        return seed, f"exception {type(e).__name__}: {e}"
    finally:
        if log:
            with open(log, "a") as f:
                f.write(f"done {seed}\n")


class _NoPool:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    @staticmethod
    def imap_unordered(fn, it):
        return map(fn, it)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", default="0:200")
    ap.add_argument("--jobs", type=int, default=max(1, (os.cpu_count() or 2) // 2))
    ap.add_argument("--gpu", action="store_true", help="the CUDA path through the C ABI instead of the emulation (one process)")
    ap.add_argument("--big", action="store_true", help="3x the resolution and 8x denser bunnies (set before the workers start)")
    a = ap.parse_args()
    if a.big and not BIG:
        os.environ["NRT_FUZZ_BIG"] = "1"
        os.execv(sys.executable, [sys.executable] + sys.argv)
    if a.gpu:
        os.environ["NRT_FUZZ_GPU"] = "1"
        a.jobs = 1
        api.initRenderer(1)
    lo, hi = (int(x) for x in a.seeds.split(":"))
    import emu_binding as emu
    import oracle
    emu.build()
    oracle.build()
    from multiprocessing import Pool
    fails = 0
    with Pool(a.jobs) if not a.gpu else _NoPool() as pool:   # (--gpu: in this process — a CUDA context does not survive fork)
        for seed, msg in pool.imap_unordered(_job, range(lo, hi)):
            if msg:
                fails += 1
                sc, o, env = case(seed)
                print(f"seed {seed}: {msg}   [{len(sc.objects)} objects, {len(sc.lights)} lights, {o.width}x{o.height}, "
                      f"aa {o.antialias.kind}/{o.antialias.gridSize}, depth {o.maxRayDepth} mode {o.depthMode}, env {env}]", flush=True)
    print(f"{hi - lo} cases, {fails} failed")
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
