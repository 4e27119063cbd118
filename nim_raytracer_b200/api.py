"""Python host layer over the C ABI (include/nrt.h) of the B200 render path.

Mirrors the reference's renderer interface — same names and argument meaning —
so tests read like the reference's own callers:

  reference (Nim)                                   here
  ------------------------------------------------  ---------------------------
  Options / Antialias   renderer.nim:10-28          Options / Antialias
  Scene / Object        scene.nim:7-18              Scene / Object
  Material              material.nim:4-7            Material
  DistantLight/PointLight light.nim:8-17            DistantLight / PointLight
  initSphere/initPlane/initBox/initTriangleMesh
                        geom.nim:159-198            same names
  Framebuf / newFramebuf  utils/framebuf.nim:7-19   Framebuf / newFramebuf
  Stats                 stats.nim:4-13              Stats
  initRenderer()        renderer.nim:214            initRenderer(ngpu)
  renderLine(scene, opts, fb, y, step, maxStep)
                        renderer.nim:162-211        renderLine(...) (one line)
  (raytracer.nim:67-109 queue-all-lines loop)       renderFrame(...)

There is NO CPU fallback: if libnrt.so (the CUDA build) is missing or no GPU is
usable, calls raise.  The CPU oracle lives in oracle/ and is never imported here.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import linalg

_HERE = os.path.dirname(os.path.abspath(__file__))
# NRT_LIB: another build of the SAME CUDA library (occupancy / flag experiments under tools/); never a fallback
LIB_PATH = os.environ.get("NRT_LIB") or os.path.join(_HERE, "csrc", "libnrt.so")

# ----------------------------------------------------------------- enums ----
NRT_GEOM_SPHERE, NRT_GEOM_PLANE, NRT_GEOM_BOX, NRT_GEOM_MESH = 0, 1, 2, 3
NRT_LIGHT_DISTANT, NRT_LIGHT_POINT = 0, 1
akNone, akGrid, akJittered, akMultiJittered, akCorrelatedMultiJittered = 0, 1, 2, 3, 4
NRT_DEPTH_REFBUG, NRT_DEPTH_INTENDED = 0, 1

NRT_ERRORS = {
    0: "NRT_OK", -1: "NRT_ERR_INVALID", -2: "NRT_ERR_CUDA", -3: "NRT_ERR_NO_DEVICE",
    -4: "NRT_ERR_NOT_INIT", -5: "NRT_ERR_OVERFLOW", -6: "NRT_ERR_UNSUPPORTED",
}


# ------------------------------------------------- ctypes mirror of nrt.h ----
class nrt_object(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("mesh", C.c_int32),
        ("object_to_world", C.c_double * 16), ("world_to_object", C.c_double * 16),
        ("radius", C.c_double), ("vmin", C.c_double * 4), ("vmax", C.c_double * 4),
        ("albedo", C.c_double * 3), ("reflection", C.c_double),
    ]


class nrt_mesh(C.Structure):
    _fields_ = [
        ("nverts", C.c_int64), ("vertices", C.POINTER(C.c_double)),
        ("nnormals", C.c_int64), ("normals", C.POINTER(C.c_double)),
        ("nfaces", C.c_int64), ("vertex_idx", C.POINTER(C.c_int64)),
        ("normal_idx", C.POINTER(C.c_int64)),
    ]


class nrt_light(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("_pad", C.c_int32), ("color", C.c_double * 3),
        ("intensity", C.c_double), ("dir", C.c_double * 4), ("pos", C.c_double * 4),
    ]


class nrt_scene_desc(C.Structure):
    _fields_ = [
        ("nobjects", C.c_int32), ("nlights", C.c_int32), ("nmeshes", C.c_int32), ("_pad", C.c_int32),
        ("objects", C.POINTER(nrt_object)), ("lights", C.POINTER(nrt_light)),
        ("meshes", C.POINTER(nrt_mesh)),
        ("fov", C.c_double), ("camera_to_world", C.c_double * 16), ("bg_color", C.c_double * 3),
    ]


class nrt_options(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32), ("aa_kind", C.c_int32), ("grid_size", C.c_int32),
        ("bias", C.c_double), ("max_ray_depth", C.c_int32), ("depth_mode", C.c_int32),
        ("bounce_cap", C.c_int32), ("_pad", C.c_int32), ("seed", C.c_uint64),
    ]


class nrt_stats(C.Structure):
    _fields_ = [
        ("num_primary_rays", C.c_int64), ("num_intersection_tests", C.c_int64),
        ("num_intersection_hits", C.c_int64), ("num_rays", C.c_int64),
        ("num_capped_samples", C.c_int64),
    ]


class nrt_aov(C.Structure):
    _fields_ = [
        ("obj_id", C.POINTER(C.c_int32)), ("tri_id", C.POINTER(C.c_int32)),
        ("t_hit", C.POINTER(C.c_double)),
    ]


NRT_KERNEL_CATEGORIES = 24


class nrt_kernel_times(C.Structure):
    _fields_ = [("ms", C.c_double * NRT_KERNEL_CATEGORIES), ("launches", C.c_int64 * NRT_KERNEL_CATEGORIES),
                ("max_ms", C.c_double * NRT_KERNEL_CATEGORIES)]


class nrt_profile(C.Structure):
    _fields_ = [
        ("total_ms", C.c_double), ("mesh_filter_ms", C.c_double),
        ("mesh_filter_launches", C.c_int64), ("mesh_tests", C.c_int64),
        ("mesh_tests_ref", C.c_int64), ("mesh_rays", C.c_int64), ("candidates", C.c_int64),
        ("pre_candidates", C.c_int64),
        ("kernel_launches", C.c_int64), ("fp32_flops", C.c_double),
        ("mesh_tests_by_mode", C.c_int64 * 4), ("mesh_ms_by_mode", C.c_double * 4),
        ("active_samples", C.c_int64 * 8), ("wavefront_samples", C.c_int64 * 8), ("tail_samples", C.c_int64), ("lanes", C.c_int64),
    ]


class nrt_ipc_handle(C.Structure):
    _fields_ = [("bytes", C.c_ubyte * 64)]


class NrtError(RuntimeError):
    pass


_lib = None


def lib() -> C.CDLL:
    """Loads csrc/libnrt.so (built by __graft_entry__.build()).  Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NrtError(
            f"{LIB_PATH} not found: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback."
        )
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    L.nrt_init.argtypes = [i32, C.POINTER(C.c_int)]
    L.nrt_shutdown.restype = None
    L.nrt_last_error.restype = C.c_char_p
    L.nrt_set_partition.argtypes = [i32, i32]
    L.nrt_scene_create.argtypes = [C.POINTER(nrt_scene_desc), C.POINTER(vp)]
    L.nrt_scene_update.argtypes = [vp, C.POINTER(nrt_scene_desc)]
    L.nrt_scene_destroy.argtypes = [vp]
    L.nrt_scene_destroy.restype = None
    L.nrt_render.argtypes = [vp, C.POINTER(nrt_options), i32, i32, i32, i32, vp,
                             C.POINTER(nrt_stats), C.POINTER(nrt_aov)]
    L.nrt_render_device.argtypes = L.nrt_render.argtypes
    L.nrt_framebuf_to_srgb8.argtypes = [vp, i32, i32, i32, vp]
    L.nrt_framebuf_quantize.argtypes = [vp, i32, i32, i32, i32, vp]
    L.nrt_framebuf_to_rgba8.argtypes = [vp, i32, i32, C.c_ubyte, vp]
    L.nrt_get_profile.argtypes = [vp, C.POINTER(nrt_profile)]
    L.nrt_set_kernel_timing.argtypes = [i32]
    L.nrt_get_kernel_times.argtypes = [vp, C.POINTER(nrt_kernel_times)]
    L.nrt_kernel_category_name.argtypes = [i32]
    L.nrt_kernel_category_name.restype = C.c_char_p
    L.nrt_device_alloc.argtypes = [i64, C.POINTER(vp)]
    L.nrt_device_free.argtypes = [vp]
    L.nrt_device_memset.argtypes = [vp, i32, i64]
    L.nrt_copy_to_host.argtypes = [vp, vp, i64]
    L.nrt_ipc_export.argtypes = [vp, C.POINTER(nrt_ipc_handle)]
    L.nrt_ipc_open.argtypes = [C.POINTER(nrt_ipc_handle), C.POINTER(vp)]
    L.nrt_ipc_close.argtypes = [vp]
    L.nrt_measure_fp32_peak.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.nrt_timer_end.argtypes = [C.POINTER(C.c_double)]
    L.nrt_host_alloc_pinned.argtypes = [i64, C.POINTER(vp)]
    L.nrt_host_free_pinned.argtypes = [vp]
    L.nrt_host_register.argtypes = [vp, i64]
    L.nrt_host_unregister.argtypes = [vp]
    _lib = L
    return L


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().nrt_last_error()
        raise NrtError(f"{what}: {NRT_ERRORS.get(rc, rc)}: {msg.decode() if msg else ''}")


# ------------------------------------------------ reference-shaped types ----
def vec(x, y, z) -> np.ndarray:  # geom.nim:11
    return np.array([x, y, z, 0.0], dtype=np.float64)


def point(x, y, z) -> np.ndarray:  # geom.nim:14
    return np.array([x, y, z, 1.0], dtype=np.float64)


def vec3(x, y=None, z=None) -> np.ndarray:
    if y is None:
        return np.array([x, x, x], dtype=np.float64)
    return np.array([x, y, z], dtype=np.float64)


@dataclass
class Material:  # material.nim:4-7
    albedo: np.ndarray = field(default_factory=lambda: vec3(0.0))
    reflection: float = 0.0


@dataclass
class Geometry:  # geom.nim:137-155
    kind: int
    objectToWorld: np.ndarray
    worldToObject: np.ndarray
    r: float = 0.0
    vmin: np.ndarray = field(default_factory=lambda: np.zeros(4))
    vmax: np.ndarray = field(default_factory=lambda: np.zeros(4))
    # TriangleMesh fields
    vertices: Optional[np.ndarray] = None   # (nverts, 4) float64
    normals: Optional[np.ndarray] = None    # (nnormals, 4) float64
    vertexIdx: Optional[np.ndarray] = None  # (nfaces, 3) int64
    normalIdx: Optional[np.ndarray] = None  # (nfaces, 3) int64


def initSphere(r: float, objectToWorld: np.ndarray) -> Geometry:  # geom.nim:159-162
    return Geometry(NRT_GEOM_SPHERE, objectToWorld, linalg.inverse(objectToWorld), r=float(r))


def initPlane(objectToWorld: np.ndarray) -> Geometry:  # geom.nim:165-167
    return Geometry(NRT_GEOM_PLANE, objectToWorld, linalg.inverse(objectToWorld))


def initBox(vmin, vmax, objectToWorld: np.ndarray) -> Geometry:  # geom.nim:169-172
    return Geometry(NRT_GEOM_BOX, objectToWorld, linalg.inverse(objectToWorld),
                    vmin=np.asarray(vmin, dtype=np.float64), vmax=np.asarray(vmax, dtype=np.float64))


def initTriangleMesh(vertices, normals, vertexIdx, normalIdx, objectToWorld: np.ndarray) -> Geometry:
    """geom.nim:190-198; faces are given as two (nfaces,3) index arrays."""
    return Geometry(
        NRT_GEOM_MESH, objectToWorld, linalg.inverse(objectToWorld),
        vertices=np.ascontiguousarray(vertices, dtype=np.float64).reshape(-1, 4),
        normals=np.ascontiguousarray(normals, dtype=np.float64).reshape(-1, 4),
        vertexIdx=np.ascontiguousarray(vertexIdx, dtype=np.int64).reshape(-1, 3),
        normalIdx=np.ascontiguousarray(normalIdx, dtype=np.int64).reshape(-1, 3),
    )


@dataclass
class Object:  # scene.nim:7-10
    name: str
    geometry: Geometry
    material: Material


@dataclass
class DistantLight:  # light.nim:12-13
    color: np.ndarray
    intensity: float
    dir: np.ndarray


@dataclass
class PointLight:  # light.nim:15-16
    color: np.ndarray
    intensity: float
    pos: np.ndarray


@dataclass
class Scene:  # scene.nim:13-18
    objects: List[Object]
    lights: list
    fov: float
    cameraToWorld: np.ndarray
    bgColor: np.ndarray


@dataclass
class Antialias:  # renderer.nim:13-22
    kind: int = akNone
    gridSize: int = 1


@dataclass
class Options:  # renderer.nim:24-28 (+ build-specific depth_mode / bounce_cap / seed)
    width: int
    height: int
    antialias: Antialias = field(default_factory=Antialias)
    bias: float = 1e-8           # raytracer.nim:50
    maxRayDepth: int = 5         # raytracer.nim:51
    depthMode: int = NRT_DEPTH_REFBUG
    bounceCap: int = 64
    seed: int = 0

    def to_c(self) -> nrt_options:
        return nrt_options(self.width, self.height, self.antialias.kind, self.antialias.gridSize,
                           self.bias, self.maxRayDepth, self.depthMode, self.bounceCap, 0, self.seed)


@dataclass
class Stats:  # stats.nim:4-13
    numPrimaryRays: int = 0
    numIntersectionTests: int = 0
    numIntersectionHits: int = 0
    numRays: int = 0
    numCappedSamples: int = 0

    def __iadd__(self, o: "Stats") -> "Stats":
        self.numPrimaryRays += o.numPrimaryRays
        self.numIntersectionTests += o.numIntersectionTests
        self.numIntersectionHits += o.numIntersectionHits
        self.numRays += o.numRays
        self.numCappedSamples += o.numCappedSamples
        return self

    @staticmethod
    def from_c(s: nrt_stats) -> "Stats":
        return Stats(s.num_primary_rays, s.num_intersection_tests, s.num_intersection_hits,
                     s.num_rays, s.num_capped_samples)


class Framebuf:  # utils/framebuf.nim:7-28
    def __init__(self, w: int, h: int):
        self.w, self.h = w, h
        self.data = np.zeros(w * h * 3, dtype=np.float32)

    def image(self) -> np.ndarray:
        return self.data.reshape(self.h, self.w, 3)

    def __getitem__(self, xy):
        x, y = xy
        assert x < self.w and y < self.h
        o = (y * self.w + x) * 3
        return self.data[o:o + 3]


def newFramebuf(w: int, h: int) -> Framebuf:
    return Framebuf(w, h)


class Aov:
    """Per-pixel debug outputs of the first sample's primary ray (nrt_aov)."""

    def __init__(self, w: int, h: int):
        self.obj_id = np.full(w * h, -2, dtype=np.int32)
        self.tri_id = np.full(w * h, -2, dtype=np.int32)
        self.t_hit = np.full(w * h, np.nan, dtype=np.float64)

    def to_c(self) -> nrt_aov:
        return nrt_aov(self.obj_id.ctypes.data_as(C.POINTER(C.c_int32)),
                       self.tri_id.ctypes.data_as(C.POINTER(C.c_int32)),
                       self.t_hit.ctypes.data_as(C.POINTER(C.c_double)))


# -------------------------------------------------- Scene -> nrt_scene_desc --
class SceneDesc:
    """Flattens a Scene into the POD description of include/nrt.h and keeps the
    backing arrays alive.  Used for both libnrt.so and (in tests) the oracle."""

    def __init__(self, scene: Scene):
        self._keep = []
        meshes, mesh_index = [], {}
        objs = (nrt_object * len(scene.objects))()
        for i, o in enumerate(scene.objects):
            g = o.geometry
            co = objs[i]
            co.kind = g.kind
            co.mesh = -1
            co.object_to_world = (C.c_double * 16)(*linalg.to_c(g.objectToWorld))
            co.world_to_object = (C.c_double * 16)(*linalg.to_c(g.worldToObject))
            co.radius = g.r
            co.vmin = (C.c_double * 4)(*[float(v) for v in g.vmin])
            co.vmax = (C.c_double * 4)(*[float(v) for v in g.vmax])
            co.albedo = (C.c_double * 3)(*[float(v) for v in o.material.albedo])
            co.reflection = float(o.material.reflection)
            if g.kind == NRT_GEOM_MESH:
                if id(g) not in mesh_index:
                    mesh_index[id(g)] = len(meshes)
                    meshes.append(g)
                co.mesh = mesh_index[id(g)]
        cm = (nrt_mesh * max(1, len(meshes)))()
        for i, g in enumerate(meshes):
            self._keep += [g.vertices, g.normals, g.vertexIdx, g.normalIdx]
            cm[i].nverts = g.vertices.shape[0]
            cm[i].vertices = g.vertices.ctypes.data_as(C.POINTER(C.c_double))
            cm[i].nnormals = g.normals.shape[0]
            cm[i].normals = g.normals.ctypes.data_as(C.POINTER(C.c_double))
            cm[i].nfaces = g.vertexIdx.shape[0]
            cm[i].vertex_idx = g.vertexIdx.ctypes.data_as(C.POINTER(C.c_int64))
            cm[i].normal_idx = g.normalIdx.ctypes.data_as(C.POINTER(C.c_int64))
        cl = (nrt_light * max(1, len(scene.lights)))()
        for i, l in enumerate(scene.lights):
            cl[i].color = (C.c_double * 3)(*[float(v) for v in l.color])
            cl[i].intensity = float(l.intensity)
            if isinstance(l, DistantLight):
                cl[i].kind = NRT_LIGHT_DISTANT
                cl[i].dir = (C.c_double * 4)(*[float(v) for v in l.dir])
            else:
                cl[i].kind = NRT_LIGHT_POINT
                cl[i].pos = (C.c_double * 4)(*[float(v) for v in l.pos])
        d = nrt_scene_desc()
        d.nobjects, d.nlights, d.nmeshes = len(scene.objects), len(scene.lights), len(meshes)
        d.objects, d.lights, d.meshes = objs, cl, cm
        d.fov = float(scene.fov)
        d.camera_to_world = (C.c_double * 16)(*linalg.to_c(scene.cameraToWorld))
        d.bg_color = (C.c_double * 3)(*[float(v) for v in scene.bgColor])
        self._keep += [objs, cl, cm]
        self.c = d
        self.n_mesh_faces = [int(g.vertexIdx.shape[0]) for g in meshes]

    def ref(self):
        return C.byref(self.c)


# ------------------------------------------------------------ entry points --
_initialised = False


def initRenderer(ngpu: int = 1, devices: Optional[Sequence[int]] = None) -> None:
    """renderer.nim:214 + raytracer.nim:61-65: selects this process' GPUs."""
    global _initialised
    L = lib()
    if devices is not None:
        arr = (C.c_int * len(devices))(*devices)
        check(L.nrt_init(len(devices), arr), "nrt_init")
    else:
        check(L.nrt_init(ngpu, None), "nrt_init")
    _initialised = True


def deviceLocalCpus(index: int = 0) -> List[int]:
    """CPUs close to selected device `index` (nrt_device_local_cpus); [] when unknown."""
    buf = C.create_string_buffer(4096)
    L = lib()
    L.nrt_device_local_cpus.argtypes = [C.c_int, C.c_char_p, C.c_int]
    check(L.nrt_device_local_cpus(index, buf, len(buf)), "nrt_device_local_cpus")
    cpus: List[int] = []
    for part in buf.value.decode().split(","):
        part = part.strip()
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.extend(range(int(a), int(b or a) + 1))
    return cpus


def pinToDevice(index: int = 0) -> bool:
    """Runs the calling process' threads on the CPUs close to device `index` (threads created later inherit it)."""
    cpus = deviceLocalCpus(index)
    if not cpus:
        return False
    try:
        os.sched_setaffinity(0, set(cpus) & os.sched_getaffinity(0) or set(cpus))
        return True
    except OSError:
        return False


def shutdown() -> None:
    global _initialised
    if _lib is not None:
        _lib.nrt_shutdown()
    _initialised = False


def setKernelTiming(enable: bool) -> None:
    check(lib().nrt_set_kernel_timing(1 if enable else 0), "nrt_set_kernel_timing")


def setPartition(index: int, count: int) -> None:
    check(lib().nrt_set_partition(index, count), "nrt_set_partition")


class DeviceScene:
    """A Scene resident on the GPU(s) (nrt_scene)."""

    def __init__(self, scene: Scene):
        if not _initialised:
            initRenderer()
        self.desc = SceneDesc(scene)
        self.handle = C.c_void_p()
        check(lib().nrt_scene_create(self.desc.ref(), C.byref(self.handle)), "nrt_scene_create")

    def update(self, scene: Optional[Scene] = None) -> None:
        if scene is not None:
            self.desc = SceneDesc(scene)
        check(lib().nrt_scene_update(self.handle, self.desc.ref()), "nrt_scene_update")

    def profile(self) -> nrt_profile:
        p = nrt_profile()
        check(lib().nrt_get_profile(self.handle, C.byref(p)), "nrt_get_profile")
        return p

    def kernelTimes(self) -> dict:
        """{kernel family: (ms, launches)} of the last frame rendered with setKernelTiming(True)."""
        t = nrt_kernel_times()
        check(lib().nrt_get_kernel_times(self.handle, C.byref(t)), "nrt_get_kernel_times")
        out = {}
        for i in range(NRT_KERNEL_CATEGORIES):
            name = lib().nrt_kernel_category_name(i).decode()
            if name and t.launches[i]:
                out[name] = (t.ms[i], int(t.launches[i]))
        self.kernelMaxMs = {lib().nrt_kernel_category_name(i).decode(): t.max_ms[i] for i in range(NRT_KERNEL_CATEGORIES) if t.launches[i]}
        return out

    def close(self) -> None:
        if self.handle:
            lib().nrt_scene_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pinSceneArrays(scene: Scene) -> list:
    """Page-locks the mesh arrays of `scene` in place (nrt_host_register) so that DeviceScene /
    update() upload them at full PCIe speed; returns the arrays that were registered (pass them to
    unpinSceneArrays before they are freed).  Arrays the driver refuses stay pageable."""
    done, seen = [], set()
    for o in scene.objects:
        g = o.geometry
        if g.kind != NRT_GEOM_MESH or id(g) in seen:
            continue
        seen.add(id(g))
        for a in (g.vertices, g.normals, g.vertexIdx, g.normalIdx):
            if a.nbytes >= (1 << 16) and lib().nrt_host_register(a.ctypes.data_as(C.c_void_p), a.nbytes) == 0:
                done.append(a)
    return done


def unpinSceneArrays(arrays: list) -> None:
    for a in arrays:
        lib().nrt_host_unregister(a.ctypes.data_as(C.c_void_p))


def _as_device_scene(scene) -> DeviceScene:
    return scene if isinstance(scene, DeviceScene) else DeviceScene(scene)


def renderFrame(scene, opts: Options, fb: Framebuf, step: int = 1, maxStep: int = 1,
                aov: Optional[Aov] = None, y0: int = 0, y1: Optional[int] = None) -> Stats:
    """All lines of raytracer.nim:67-70 in one call (nrt_render)."""
    ds = _as_device_scene(scene)
    co, cs = opts.to_c(), nrt_stats()
    ca = aov.to_c() if aov is not None else None
    check(lib().nrt_render(ds.handle, C.byref(co), y0, opts.height if y1 is None else y1, step, maxStep,
                           fb.data.ctypes.data_as(C.c_void_p), C.byref(cs),
                           C.byref(ca) if ca is not None else None), "nrt_render")
    return Stats.from_c(cs)


def renderLine(scene, opts: Options, fb: Framebuf, y: int, step: int = 1, maxStep: int = 1) -> Stats:
    """renderer.nim:162-211 — one scanline (kept for the worker-pool style caller)."""
    if step <= 0 or (step & (step - 1)) or maxStep <= 0 or (maxStep & (maxStep - 1)) or maxStep < step:
        raise AssertionError("isPowerOfTwo(step) and isPowerOfTwo(maxStep) and maxStep >= step")
    ds = _as_device_scene(scene)
    co, cs = opts.to_c(), nrt_stats()
    # a line that the caller queued even though y mod step != 0 is still rendered
    check(lib().nrt_render(ds.handle, C.byref(co), y, y + 1, step, maxStep,
                           fb.data.ctypes.data_as(C.c_void_p), C.byref(cs), None), "nrt_render")
    return Stats.from_c(cs)


def framebufToSrgb8(fb: Framebuf, sRGB: bool = True) -> np.ndarray:
    """GPU version of writePpm's outvalue (utils/framebuf.nim:74-78)."""
    out = np.zeros(fb.w * fb.h * 3, dtype=np.uint8)
    check(lib().nrt_framebuf_to_srgb8(fb.data.ctypes.data_as(C.c_void_p), fb.w, fb.h, int(sRGB),
                                      out.ctypes.data_as(C.c_void_p)), "nrt_framebuf_to_srgb8")
    return out.reshape(fb.h, fb.w, 3)


def framebufQuantize(fb: Framebuf, bits: int = 8, sRGB: bool = True) -> np.ndarray:
    """writePpm's samples (utils/framebuf.nim:55-93) for any bits in 1..16: uint8, or big-endian uint16 above 8 bits."""
    if not 1 <= bits <= 16:
        raise ValueError("bits must be in 1..16 (framebuf.nim:56)")
    out = np.zeros(fb.w * fb.h * 3, dtype=np.uint8 if bits <= 8 else ">u2")
    check(lib().nrt_framebuf_quantize(fb.data.ctypes.data_as(C.c_void_p), fb.w, fb.h, bits, int(sRGB),
                                      out.ctypes.data_as(C.c_void_p)), "nrt_framebuf_quantize")
    return out.reshape(fb.h, fb.w, 3)


def framebufToRgba8(fb: Framebuf, alpha: int = 0xFF) -> np.ndarray:
    """ImageRGBA.copyFrom (utils/image.nim:45-54)."""
    out = np.zeros(fb.w * fb.h * 4, dtype=np.uint8)
    check(lib().nrt_framebuf_to_rgba8(fb.data.ctypes.data_as(C.c_void_p), fb.w, fb.h, alpha,
                                      out.ctypes.data_as(C.c_void_p)), "nrt_framebuf_to_rgba8")
    return out.reshape(fb.h, fb.w, 4)


NRT_OUT_RGB, NRT_OUT_RGBA8 = 0, 1


def bandRows(opts: "Options", step: int = 1, maxStep: int = 1) -> int:
    """Band height T of a pass (nrt_band_rows_for): whole-resolution passes are dealt out in bands of T scanlines
    (rows of T x T tiles); progressive passes in single scanlines."""
    L = lib()
    L.nrt_band_rows_for.argtypes = [C.POINTER(nrt_options), C.c_int, C.c_int]
    co = opts.to_c()
    return int(L.nrt_band_rows_for(C.byref(co), step, maxStep))


def partitionRows(height: int, index: int, count: int, y0: int = 0, y1: Optional[int] = None, step: int = 1, band: int = 1) -> List[int]:
    """First scanlines of the units of [y0, y1) that partition `index` of `count` renders (nrt_partition_rows: the
    library's own enumeration of raytracer.nim:67-70's work items; distributed.rows_of is its Python mirror)."""
    L = lib()
    L.nrt_partition_rows.argtypes = [C.c_int] * 7 + [C.POINTER(C.c_int), C.c_int]
    y1 = height if y1 is None else y1
    n = int(L.nrt_partition_rows(height, y0, y1, step, band, index, count, None, 0))
    if n < 0:
        raise NrtError("nrt_partition_rows: NRT_ERR_INVALID")
    buf = (C.c_int * max(n, 1))()
    L.nrt_partition_rows(height, y0, y1, step, band, index, count, buf, n)
    return list(buf[:n])


def outputCutPoints(bits: int = 8) -> np.ndarray:
    """The cut-point table of the sRGB pow branch for `bits` (host computation inside libnrt.so; no GPU needed)."""
    out = np.zeros((1 << bits) - 1, dtype=np.float32)
    L = lib()
    L.nrt_output_cut_points.argtypes = [C.c_int, C.c_void_p]
    check(L.nrt_output_cut_points(bits, out.ctypes.data_as(C.c_void_p)), "nrt_output_cut_points")
    return out


def renderFrameQuantized(scene, opts: "Options", bits: int = 8, sRGB: bool = True, rgba: bool = False, alpha: int = 0xFF,
                         step: int = 1, maxStep: int = 1, image: Optional[np.ndarray] = None):
    """renderFrame + writePpm's samples (or ImageRGBA bytes with rgba=True) in one call: the output stage is the
    epilogue of the pixel-store kernel and only the integer image is copied to the host.  Returns (image, Stats);
    image is (h, w, 3) uint8 / big-endian uint16, or (h, w, 4) uint8."""
    ds = _as_device_scene(scene)
    h, w = opts.height, opts.width
    if image is None:
        image = np.zeros((h, w, 4), dtype=np.uint8) if rgba else np.zeros((h, w, 3), dtype=np.uint8 if bits <= 8 else ">u2")
    co, cs = opts.to_c(), nrt_stats()
    L = lib()
    L.nrt_render_quantized.argtypes = [C.c_void_p, C.POINTER(nrt_options), C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(nrt_stats)]
    check(L.nrt_render_quantized(ds.handle, C.byref(co), 0, h, step, maxStep, NRT_OUT_RGBA8 if rgba else NRT_OUT_RGB,
                                 bits, int(sRGB), alpha, image.ctypes.data_as(C.c_void_p), C.byref(cs)), "nrt_render_quantized")
    return image, Stats.from_c(cs)


def writePpm(fb: Framebuf, filename: str, bits: int = 8, sRGB: bool = True) -> bool:
    """utils/framebuf.nim:55-93: P6, maxval 2^bits - 1, 8-bit or big-endian 16-bit samples (GPU output stage)."""
    img = framebufQuantize(fb, bits, sRGB)
    with open(filename, "wb") as f:
        f.write(f"P6 {fb.w} {fb.h} {(1 << bits) - 1} ".encode())
        f.write(img.tobytes())
    return True
