"""Binding for the TEST-ONLY host emulation (tests/emu/nrt_emu.cpp).  Not a product path."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

from nim_raytracer_b200 import api

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_lib = None


def build(force: bool = False) -> str:
    out = os.path.join(_HERE, "libnrt_emu.so")
    srcs = [os.path.join(_HERE, "nrt_emu.cpp")] + [
        os.path.join(_ROOT, "nim_raytracer_b200", "csrc", n) for n in ("nrt_core.h", "nrt_filter.h", "nrt_pipeline.h", "nrt_renderer.h")
    ] + [os.path.join(_ROOT, "include", "nrt.h")]
    if force or not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        subprocess.run(["g++", "-std=c++17", "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-shared", "-fPIC",
                        "-I" + os.path.join(_ROOT, "include"), "-o", out, srcs[0]], check=True)
    return out


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.emu_render.argtypes = [C.POINTER(api.nrt_scene_desc), C.POINTER(api.nrt_options), C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_void_p, C.POINTER(api.nrt_stats), C.POINTER(api.nrt_aov),
                                    C.POINTER(C.c_int64)]
    return _lib


def render(scene, opts: api.Options, fb=None, aov=None, step=1, maxStep=1, y0=0, y1=None):
    desc = scene if isinstance(scene, api.SceneDesc) else api.SceneDesc(scene)
    fb = fb or api.newFramebuf(opts.width, opts.height)
    co, cs = opts.to_c(), api.nrt_stats()
    ca = aov.to_c() if aov is not None else None
    prof = (C.c_int64 * 7)()
    rc = lib().emu_render(desc.ref(), C.byref(co), y0, opts.height if y1 is None else y1, step, maxStep,
                          fb.data.ctypes.data_as(C.c_void_p), C.byref(cs), C.byref(ca) if ca is not None else None, prof)
    if rc != 0:
        raise RuntimeError(f"emu_render failed: {rc}")
    names = ("mesh_rays", "mesh_tests", "candidates", "exact_rays", "launches", "filter_tests", "pre_candidates")
    return fb, api.Stats.from_c(cs), aov, dict(zip(names, list(prof)))
