"""CPU oracle binding — TEST INFRASTRUCTURE ONLY (see oracle/ref_cpu.cpp header).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from nim_raytracer_b200 import api

_HERE = os.path.dirname(os.path.abspath(__file__))
_libs = {}


def build(force: bool = False) -> None:
    """Compiles liboracle.so / liboracle_fast.so (+ _ref when /root/reference exists)."""
    need = force or not all(os.path.exists(os.path.join(_HERE, n)) for n in ("liboracle.so", "liboracle_fast.so"))
    src = os.path.join(_HERE, "ref_cpu.cpp")
    for n in ("liboracle.so", "liboracle_fast.so"):
        p = os.path.join(_HERE, n)
        if os.path.exists(p) and os.path.getmtime(p) < os.path.getmtime(src):
            need = True
    if need:
        subprocess.run(["make", "-C", _HERE, "-B", "all"], check=True, capture_output=True)


def lib(fast: bool = False) -> C.CDLL:
    name = "liboracle_fast.so" if fast else "liboracle.so"
    if name not in _libs:
        build()
        L = C.CDLL(os.path.join(_HERE, name))
        dp = C.POINTER(C.c_double)
        L.oracle_render.argtypes = [C.POINTER(api.nrt_scene_desc), C.POINTER(api.nrt_options), C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_void_p, C.POINTER(api.nrt_stats),
                                    C.POINTER(api.nrt_aov), C.c_int]
        L.oracle_solve_quadratic.argtypes = [C.c_double] * 3 + [dp, dp]
        L.oracle_solve_quadratic.restype = None
        for fn, n in (("oracle_aabb_intersect", 4), ("oracle_plane_intersect", 2), ("oracle_ray_triangle", 5)):
            getattr(L, fn).argtypes = [dp] * n
            getattr(L, fn).restype = C.c_double
        L.oracle_sphere_intersect.argtypes = [C.c_double, dp, dp]
        L.oracle_sphere_intersect.restype = C.c_double
        L.oracle_cast_primary_ray.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, dp, dp]
        L.oracle_cast_primary_ray.restype = None
        L.oracle_samples.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int, dp]
        L.oracle_samples.restype = None
        _libs[name] = L
    return _libs[name]


def ref_geomtest():
    """The reference's own C++ AABB test compiled into oracle/_ref (or None)."""
    p = os.path.join(_HERE, "_ref", "libgeomtest_ref.so")
    if not os.path.exists(p):
        return None
    L = C.CDLL(p)
    dp = C.POINTER(C.c_double)
    L.geomtest_ref_aabb_intersect.argtypes = [dp] * 4
    L.geomtest_ref_aabb_intersect.restype = C.c_double
    return L


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def render(scene, opts: api.Options, fb: api.Framebuf = None, aov: api.Aov = None, step: int = 1,
           maxStep: int = 1, y0: int = 0, y1: int = None, nthreads: int = 0, fast: bool = False):
    """oracle_render on a Scene (or a prebuilt api.SceneDesc).  Returns (fb, stats, aov)."""
    desc = scene if isinstance(scene, api.SceneDesc) else api.SceneDesc(scene)
    fb = fb or api.newFramebuf(opts.width, opts.height)
    co, cs = opts.to_c(), api.nrt_stats()
    ca = aov.to_c() if aov is not None else None
    rc = lib(fast).oracle_render(desc.ref(), C.byref(co), y0, opts.height if y1 is None else y1, step, maxStep,
                                 fb.data.ctypes.data_as(C.c_void_p), C.byref(cs),
                                 C.byref(ca) if ca is not None else None, nthreads)
    if rc != 0:
        raise RuntimeError(f"oracle_render failed: {rc}")
    return fb, api.Stats.from_c(cs), aov


def render_rows(scene, opts: api.Options, rows, fb: api.Framebuf = None, nthreads: int = 0, fast: bool = False):
    """oracle_render_rows: the worker-pool loop over an explicit scanline list."""
    desc = scene if isinstance(scene, api.SceneDesc) else api.SceneDesc(scene)
    fb = fb or api.newFramebuf(opts.width, opts.height)
    co, cs = opts.to_c(), api.nrt_stats()
    r = np.ascontiguousarray(rows, dtype=np.int32)
    L = lib(fast)
    L.oracle_render_rows.argtypes = [C.POINTER(api.nrt_scene_desc), C.POINTER(api.nrt_options), C.c_void_p, C.c_int,
                                     C.c_void_p, C.POINTER(api.nrt_stats), C.c_int]
    rc = L.oracle_render_rows(desc.ref(), C.byref(co), r.ctypes.data_as(C.c_void_p), len(r),
                              fb.data.ctypes.data_as(C.c_void_p), C.byref(cs), nthreads)
    if rc != 0:
        raise RuntimeError(f"oracle_render_rows failed: {rc}")
    return fb, api.Stats.from_c(cs)


def hardware_threads() -> int:
    return int(lib().oracle_hardware_threads())


def solve_quadratic(a, b, c):
    t1, t2 = C.c_double(), C.c_double()
    lib().oracle_solve_quadratic(a, b, c, C.byref(t1), C.byref(t2))
    return t1.value, t2.value


def aabb_intersect(vmin, vmax, orig, direction):
    (_, a), (_, b), (_, c), (_, d) = _d(vmin), _d(vmax), _d(orig), _d(direction)
    return lib().oracle_aabb_intersect(a, b, c, d)


def sphere_intersect(r, orig, direction):
    (k1, c), (k2, d) = _d(orig), _d(direction)
    return lib().oracle_sphere_intersect(r, c, d)


def plane_intersect(orig, direction):
    (k1, c), (k2, d) = _d(orig), _d(direction)
    return lib().oracle_plane_intersect(c, d)


def ray_triangle(orig, direction, v0, v1, v2):
    keep = [_d(x) for x in (orig, direction, v0, v1, v2)]
    return lib().oracle_ray_triangle(*[k[1] for k in keep])


def cast_primary_ray(w, h, x, y, fov, c2w):
    from nim_raytracer_b200 import linalg
    m, mp = _d(linalg.to_c(c2w))
    out = np.zeros(8)
    lib().oracle_cast_primary_ray(w, h, x, y, fov, mp, out.ctypes.data_as(C.POINTER(C.c_double)))
    return out[:4].copy(), out[4:].copy()


def samples(kind, m, seed=0, width=1, x=0, y=0):
    out = np.zeros(2 * m * m)
    lib().oracle_samples(kind, m, seed, width, x, y, out.ctypes.data_as(C.POINTER(C.c_double)))
    return out.reshape(m * m, 2)


def outvalues(fb_data: np.ndarray, bits: int = 8, sRGB: bool = True) -> np.ndarray:
    """writePpm's samples (utils/framebuf.nim:74-89) of float32 components: uint8, or big-endian uint16 above 8 bits."""
    a = np.ascontiguousarray(fb_data, dtype=np.float32).reshape(-1)
    out = np.zeros(a.size * (1 if bits <= 8 else 2), dtype=np.uint8)
    L = lib()
    L.oracle_outvalues.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]
    L.oracle_outvalues.restype = None
    L.oracle_outvalues(a.ctypes.data_as(C.c_void_p), a.size, bits, int(sRGB), out.ctypes.data_as(C.c_void_p))
    return out if bits <= 8 else out.view(">u2")


def rgba8(fb_data: np.ndarray, alpha: int = 0xFF) -> np.ndarray:
    """ImageRGBA.copyFrom (utils/image.nim:45-54)."""
    a = np.ascontiguousarray(fb_data, dtype=np.float32).reshape(-1)
    out = np.zeros(a.size // 3 * 4, dtype=np.uint8)
    L = lib()
    L.oracle_rgba8.argtypes = [C.c_void_p, C.c_int64, C.c_ubyte, C.c_void_p]
    L.oracle_rgba8.restype = None
    L.oracle_rgba8(a.ctypes.data_as(C.c_void_p), a.size // 3, alpha, out.ctypes.data_as(C.c_void_p))
    return out


def write_ppm(fb, filename: str, bits: int = 8, sRGB: bool = True) -> bool:
    """utils/framebuf.nim:55-93: "P6 w h maxval " + the samples."""
    with open(filename, "wb") as f:
        f.write(f"P6 {fb.w} {fb.h} {(1 << bits) - 1} ".encode())
        f.write(outvalues(fb.data, bits, sRGB).tobytes())
    return True
