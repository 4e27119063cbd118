"""Scene fixtures: the reference's scene files re-expressed as data, plus the
build-defined bunny / stress scenes of BASELINE.md §3.

Each function cites the scene file it restates (src/data/scenes/*.nim).
"""
from __future__ import annotations

import os

import numpy as np

from . import linalg as L
from .api import (DistantLight, Material, Object, PointLight, Scene, initBox, initPlane, initSphere,
                  initTriangleMesh, point, vec, vec3)
from .loaders import loadObj, readGeom, trianglesToMesh

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
BUNNY_GEOM = os.path.join(DATA_DIR, "bunny.geom")  # == test/bunny.geom of the reference (data, 69,451 triangles)


def _camera(tx, ty, tz, rx_deg=-12.0):
    # mat4(1.0).rotate(X_AXIS, degToRad(-12.0)).translate(vec3(...))
    return L.translate(L.rotate(L.mat4(1.0), L.X_AXIS, L.deg_to_rad(rx_deg)), vec3(tx, ty, tz))


def _bunny_lights():
    # src/data/scenes/mesh-bunny.nim:22-31
    return [
        DistantLight(color=vec3(1.0), intensity=4.0, dir=L.normalize(vec(-2.0, -0.8, -0.3))),
        DistantLight(color=vec3(0.8, 0.3, 0.0), intensity=1.0, dir=L.normalize(vec(2.0, -0.8, -1.3))),
    ]


def spheres_reflection() -> Scene:
    """src/data/scenes/spheres-reflection.nim:1-94 (BASELINE config 1)."""
    def ball(name, x, y, z, albedo, refl):
        return Object(name, initSphere(r=2, objectToWorld=L.translate(L.mat4(1.0), vec3(x, y, z))),
                      Material(albedo=vec3(*albedo), reflection=refl))
    objects = [
        ball("ball1", -5.0, 2.0, -18.0, (1.0, 1.0, 1.0), 1.0),
        ball("ball2", 0.5, 2.0, -8.0, (1.0, 1.0, 1.0), 1.0),
        ball("ball3", -5.0, 2.0, -10.0, (1.0, 1.0, 1.0), 1.0),
        ball("ball4", 8.0, 2.0, -15.0, (0.2, 0.3, 0.9), 0.0),
        ball("ball5", 4.0, 2.0, -16.0, (0.2, 0.5, 0.9), 0.0),
        ball("ball6", -2.0, 2.0, -42.0, (0.9, 0.5, 0.2), 1.0),
        ball("ball7", 9.0, 2.0, -30.0, (0.6, 0.5, 0.9), 1.0),
        Object("ground", initPlane(objectToWorld=L.mat4(1.0)), Material(albedo=vec3(0.4), reflection=0.0)),
        ball("ball-behind1", -4.0, 2.0, 0.0, (0.2, 0.3, 0.6), 0.0),
        ball("ball-behind2", 14.0, 2.0, 1.0, (0.4, 0.8, 1.0), 0.0),
    ]
    lights = [
        PointLight(color=vec3(1.0, 1.0, 1.0), intensity=3000.0, pos=point(3.0, 6.0, -12.0)),
        DistantLight(color=vec3(0.3, 0.4, 0.6), intensity=0.5, dir=L.normalize(vec(2.0, -0.8, -1.3))),
    ]
    return Scene(objects=objects, lights=lights, fov=50.0, cameraToWorld=_camera(1.0, 5.5, 3.5),
                 bgColor=vec3(0.25, 0.1, 0.2))


def boxtest() -> Scene:
    """src/data/scenes/boxtest.nim:1-37 (plane + box; its camera is the golden of test/boxtest.nim:32)."""
    objects = [
        Object("ground", initPlane(objectToWorld=L.mat4(1.0)), Material(albedo=vec3(0.4))),
        Object("box", initBox(objectToWorld=L.translate(L.mat4(1.0), vec3(0.0, 1.0, -10.0)),
                              vmin=vec(-1.0, -1.0, -1.0), vmax=vec(1.0, 1.0, 1.0)),
               Material(albedo=vec3(1.0))),
    ]
    lights = [
        DistantLight(color=vec3(1.0), intensity=9.0, dir=L.normalize(vec(-2.0, -0.8, -0.3))),
        DistantLight(color=vec3(0.8, 0.3, 0.0), intensity=2.0, dir=L.normalize(vec(2.0, -0.8, -1.3))),
    ]
    return Scene(objects=objects, lights=lights, fov=50.0, cameraToWorld=_camera(1.0, 5.5, 3.5),
                 bgColor=vec3(0.15, 0.09, 0.07))


def mesh_cube() -> Scene:
    """src/data/scenes/mesh-cube.nim:1-44: ONE triangle as a mesh, identity transforms."""
    v = np.array([point(0.0, 0.0, 0.0), point(1.0, 0.0, 0.0), point(0.0, 1.0, 0.0)])
    n = np.array([vec(0.0, 0.0, 1.0)] * 3)
    mesh = initTriangleMesh(v, n, [[0, 1, 2]], [[0, 1, 2]], L.mat4(1.0))
    objects = [Object("cube", mesh, Material(albedo=vec3(0.6, 0.9, 0.2)))]
    return Scene(objects=objects, lights=_bunny_lights(), fov=50.0,
                 cameraToWorld=L.translate(L.mat4(1.0), vec3(0.3, 0.3, 5.0)), bgColor=vec3(0.01, 0.03, 0.05))


def mesh_scene(mesh, albedo=(0.6, 0.9, 0.2), reflection=0.0, extra_objects=()) -> Scene:
    """src/data/scenes/mesh-bunny.nim:1-42 with `mesh` in the teapot's place."""
    mesh.objectToWorld = L.translate(L.mat4(1.0), vec3(0.0, 0.0001, -12.0))
    mesh.worldToObject = L.inverse(mesh.objectToWorld)
    objects = [
        Object("mesh", mesh, Material(albedo=vec3(*albedo), reflection=reflection)),
        Object("ground", initPlane(objectToWorld=L.mat4(1.0)), Material(albedo=vec3(0.4))),
    ] + list(extra_objects)
    return Scene(objects=objects, lights=_bunny_lights(), fov=50.0, cameraToWorld=_camera(0.0, 5.5, 1.5),
                 bgColor=vec3(0.01, 0.03, 0.05))


TEAPOT_OBJ = os.path.join(DATA_DIR, "teapot.obj")  # == src/data/meshes/teapot.obj of the reference (data, 6,320 triangles)


def teapot_scene(obj_path: str = TEAPOT_OBJ) -> Scene:
    """src/data/scenes/mesh-bunny.nim:1-42 verbatim (it loads data/meshes/teapot.obj): the scene the reference's
    default front-end renders (src/raytracer.nim:43-54: 300x200, akNone, bias 1e-8, maxRayDepth 5)."""
    return mesh_scene(loadObj(obj_path))


def _spheres_layout(lights, bg) -> Scene:
    """The seven r = 2 spheres + ground plane shared by src/data/scenes/spheres-{pointlight1,pointlight2,warm,blue,
    purple}.nim:1-58 (no reflection), camera Rx(-12 deg) . T(1, 5.5, 3.5), fov 50."""
    def ball(name, x, y, z, albedo):
        return Object(name, initSphere(r=2, objectToWorld=L.translate(L.mat4(1.0), vec3(x, y, z))), Material(albedo=vec3(*albedo)))
    objects = [
        ball("ball1", -5.0, 2.0, -18.0, (0.9, 0.3, 0.2)),
        ball("ball2", 0.5, 2.0, -8.0, (0.6, 0.9, 0.2)),
        ball("ball3", -5.0, 2.0, -10.0, (0.1, 0.7, 0.2)),
        ball("ball4", 8.0, 2.0, -15.0, (0.2, 0.3, 0.9)),
        ball("ball5", 4.0, 2.0, -16.0, (0.2, 0.5, 0.9)),
        ball("ball6", -2.0, 2.0, -42.0, (0.9, 0.5, 0.2)),
        ball("ball7", 9.0, 2.0, -30.0, (0.6, 0.5, 0.9)),
        Object("ground", initPlane(objectToWorld=L.mat4(1.0)), Material(albedo=vec3(0.4))),
    ]
    return Scene(objects=objects, lights=lights, fov=50.0, cameraToWorld=_camera(1.0, 5.5, 3.5), bgColor=vec3(*bg))


def spheres_pointlight1() -> Scene:
    """src/data/scenes/spheres-pointlight1.nim:1-76."""
    return _spheres_layout([PointLight(color=vec3(1.0, 0.8, 0.5), intensity=2000.0, pos=point(3.0, 6.0, -12.0))], (0.0, 0.0, 0.0))


def spheres_pointlight2() -> Scene:
    """src/data/scenes/spheres-pointlight2.nim:1-76."""
    return _spheres_layout([PointLight(color=vec3(1.0, 0.8, 0.5), intensity=5000.0, pos=point(0.0, 8.0, -35.0))], (0.0, 0.0, 0.0))


def spheres_warm() -> Scene:
    """src/data/scenes/spheres-warm.nim:1-80."""
    return _spheres_layout(_bunny_lights(), (0.01, 0.03, 0.05))


def spheres_blue() -> Scene:
    """src/data/scenes/spheres-blue.nim:1-80."""
    return _spheres_layout([
        DistantLight(color=vec3(1.0), intensity=0.2, dir=L.normalize(vec(-2.0, -0.8, -0.3))),
        DistantLight(color=vec3(0.1, 0.6, 0.8), intensity=8.0, dir=L.normalize(vec(0.5, -0.4, 0.8))),
    ], (0.1, 0.6, 0.8))


def spheres_purple() -> Scene:
    """src/data/scenes/spheres-purple.nim:1-80."""
    return _spheres_layout([
        DistantLight(color=vec3(1.0, 0.0, 0.0), intensity=4.5, dir=L.normalize(vec(1.7, -0.5, 2.3))),
        DistantLight(color=vec3(1.0, 0.0, 1.0), intensity=1.5, dir=L.normalize(vec(-2.7, -0.5, 2.3))),
    ], (0.30, 0.00, 0.02))


def boxes() -> Scene:
    """src/data/scenes/boxes.nim:1-70: ground plane + 4 x 4 x 4 boxes, camera rotated about TWO axes
    (Rx(-34 deg) . Ry(-35 deg) . T(-3.5, 20.5, 9.5)), fov 65.  The loop variables advance exactly as in the file
    (x += PAD inside the innermost loop, so the accumulated float64 sums are the reference's)."""
    objects = [Object("ground", initPlane(objectToWorld=L.mat4(1.0)), Material(albedo=vec3(0.4)))]
    ROWS, PAD = 4, 3.8
    xs, ys, zs = 0.0, 3.0, -18.0
    x, y, z = xs, ys, zs
    for i in range(ROWS):
        for j in range(ROWS):
            for k in range(ROWS):
                objects.append(Object("box", initBox(objectToWorld=L.translate(L.mat4(1.0), vec3(x, y, z)),
                                                     vmin=vec(-1.3, -1.3, -1.3), vmax=vec(1.3, 1.3, 1.3)),
                                      Material(albedo=vec3(1.0 / float(ROWS) * float(ROWS - k),
                                                           1.0 / float(ROWS) * float(ROWS - j),
                                                           1.0 / float(ROWS) * float(ROWS - i)))))
                x += PAD
            y += PAD
            x = xs
        z -= PAD
        x = xs
        y = ys
    lights = [
        DistantLight(color=vec3(1.0), intensity=9.0, dir=L.normalize(vec(-2.0, -0.8, -0.3))),
        DistantLight(color=vec3(0.8, 0.3, 0.0), intensity=2.0, dir=L.normalize(vec(2.0, -0.8, -1.3))),
    ]
    cam = L.translate(L.rotate(L.rotate(L.mat4(1.0), L.X_AXIS, L.deg_to_rad(-34.0)), L.Y_AXIS, L.deg_to_rad(-35.0)),
                      vec3(-3.5, 20.5, 9.5))
    return Scene(objects=objects, lights=lights, fov=65.0, cameraToWorld=cam, bgColor=vec3(0.15, 0.09, 0.07))


def boxes2() -> Scene:
    """src/data/scenes/boxes2.nim:1-62: platform box, ball and a ring of 13 boxes each under
    T(0, 1, Z) . Ry(rot) . T(0, 0, 5) with rot accumulated by 360 / 13; one distant light, fov 20."""
    Z_DIST = -18.0
    objects = [
        Object("ground", initPlane(objectToWorld=L.mat4(1.0)), Material(albedo=vec3(0.3))),
        Object("platform", initBox(objectToWorld=L.translate(L.mat4(1.0), vec3(0.0, 0.0, Z_DIST)),
                                   vmin=vec(-6.3, 0.0, -6.3), vmax=vec(6.3, 0.5, 6.3)), Material(albedo=vec3(0.5))),
        Object("ball", initSphere(objectToWorld=L.translate(L.mat4(1.0), vec3(0.0, 4.8, Z_DIST)), r=1.5),
               Material(albedo=vec3(0.5))),
    ]
    N, rot = 13, 0.0
    for i in range(N):
        m = L.translate(L.rotate(L.translate(L.mat4(1.0), vec3(0.0, 1.0, Z_DIST)), L.Y_AXIS, L.deg_to_rad(rot)), vec3(0.0, 0.0, 5.0))
        objects.append(Object(f"box{i}", initBox(objectToWorld=m, vmin=vec(-0.5, -0.5, -0.5), vmax=vec(0.5, 0.5, 0.5)),
                              Material(albedo=vec3(0.5))))
        rot += 360.0 / float(N)
    lights = [DistantLight(color=vec3(1.0, 1.0, 1.0), intensity=0.5, dir=L.normalize(vec(3.0, -0.5, -4.0)))]
    return Scene(objects=objects, lights=lights, fov=20.0, cameraToWorld=_camera(0.0, 6.0, 20.0, -15.0),
                 bgColor=vec3(0.15, 0.07, 0.04))


def boxes_pointlight1() -> Scene:
    """src/data/scenes/boxes-pointlight1.nim:1-88: seven boxes each under T . Rx(angle), ground plane, the two
    mesh-bunny lights (the file's name notwithstanding, both are DistantLights), camera Rx(-12 deg) . T(0.5, 5.5, 3.5)."""
    def box(name, t, deg, albedo):
        m = L.rotate(L.translate(L.mat4(1.0), vec3(*t)), L.X_AXIS, L.deg_to_rad(deg))
        return Object(name, initBox(objectToWorld=m, vmin=vec(-1.0, -1.0, -1.0), vmax=vec(1.0, 1.0, 1.0)), Material(albedo=vec3(*albedo)))
    objects = [
        box("box-red", (-5.0, 2.0, -18.0), -10.0, (0.9, 0.3, 0.2)),
        box("box-yellow", (0.5, 2.0, -6.0), -50.0, (0.6, 0.9, 0.2)),
        box("box3", (-5.0, 2.0, -10.0), -50.0, (0.1, 0.7, 0.2)),
        box("box4", (8.0, 2.0, -15.0), -70.0, (0.2, 0.3, 0.9)),
        box("box5", (4.0, 2.0, -16.0), -60.0, (0.2, 0.5, 0.9)),
        box("box6", (-2.0, 2.0, -52.0), -40.0, (0.9, 0.5, 0.2)),
        box("box7", (9.0, 2.0, -30.0), -20.0, (0.6, 0.5, 0.9)),
        Object("ground", initPlane(objectToWorld=L.mat4(1.0)), Material(albedo=vec3(0.4))),
    ]
    return Scene(objects=objects, lights=_bunny_lights(), fov=50.0, cameraToWorld=_camera(0.5, 5.5, 3.5),
                 bgColor=vec3(0.6, 0.6, 0.5))


# the reference's scene files (src/data/scenes/*.nim) by name; first.nim predates the Object / Geometry types
# (`Sphere(o: ..., albedo: ...)`) and does not compile against the renderer of this tree
REFERENCE_SCENES = {
    "spheres-reflection": spheres_reflection, "boxtest": boxtest, "mesh-cube": mesh_cube, "mesh-bunny": teapot_scene,
    "spheres-pointlight1": spheres_pointlight1, "spheres-pointlight2": spheres_pointlight2, "spheres-warm": spheres_warm,
    "spheres-blue": spheres_blue, "spheres-purple": spheres_purple, "boxes": boxes, "boxes2": boxes2,
    "boxes-pointlight1": boxes_pointlight1,
}


def bunny_triangles(flip_winding: bool = True, scale: float = 20.0, max_faces: int | None = None,
                    stride: int = 1) -> np.ndarray:
    """test/bunny.geom prepared as frozen in SURVEY.md §8d / BASELINE.md §3 config 2:
    float32 -> float64 exactly, y_min subtracted, scaled x20 about the origin,
    v1/v2 swapped (the file's winding is inverted w.r.t. every other mesh of the
    reference).  `stride`/`max_faces` give decimated variants for small tests."""
    tri = readGeom(BUNNY_GEOM).astype(np.float64)
    ymin = float(tri[:, :, 1].min())
    tri[:, :, 1] -= ymin
    tri *= scale
    if flip_winding:
        tri = tri[:, [0, 2, 1], :]
    tri = tri[::stride]
    if max_faces is not None:
        tri = tri[:max_faces]
    return np.ascontiguousarray(tri)


def bunny(flip_winding: bool = True, stride: int = 1, max_faces: int | None = None) -> Scene:
    """BASELINE config 2: canonical bunny + ground plane + 2 distant lights."""
    return mesh_scene(trianglesToMesh(bunny_triangles(flip_winding, stride=stride, max_faces=max_faces)))


def bunny_spheres(stride: int = 1) -> Scene:
    """BASELINE config 3/4: config 2 + the seven spheres of spheres-reflection.nim
    shifted to flank the bunny, reflection in {0, 0.5, 1} (positions frozen here)."""
    def ball(name, x, z, r, albedo, refl):
        return Object(name, initSphere(r=r, objectToWorld=L.translate(L.mat4(1.0), vec3(x, r, z))),
                      Material(albedo=vec3(*albedo), reflection=refl))
    extra = [
        ball("ball1", -5.0, -15.0, 2.0, (1.0, 1.0, 1.0), 1.0),
        ball("ball2", 4.5, -10.0, 1.5, (1.0, 1.0, 1.0), 1.0),
        ball("ball3", -4.0, -9.0, 1.0, (0.9, 0.9, 0.9), 0.5),
        ball("ball4", 6.5, -16.0, 2.0, (0.2, 0.3, 0.9), 0.0),
        ball("ball5", 2.5, -7.5, 0.75, (0.2, 0.5, 0.9), 0.5),
        ball("ball6", -1.0, -22.0, 2.0, (0.9, 0.5, 0.2), 1.0),
        ball("ball7", -2.5, -7.0, 0.6, (0.6, 0.5, 0.9), 0.0),
    ]
    sc = mesh_scene(trianglesToMesh(bunny_triangles(stride=stride)), extra_objects=extra)
    return sc


def _mt_uniform(n: int, seed: int) -> np.ndarray:
    """std::mt19937_64(seed) mapped with (x >> 11) * 2^-53.
    numpy's MT19937 is the 32-bit generator; we freeze numpy's PCG-free path instead:
    two 32-bit MT draws (hi, lo) -> 53-bit mantissa, documented here as THE spec."""
    rs = np.random.RandomState(seed % (2 ** 32))
    return rs.random_sample(n)


def stress(ntri: int = 1_000_000, nspheres: int = 10_000, seed: int = 20161018) -> Scene:
    """BASELINE config 5 (template test/test.nim:29-40): random triangles as ONE mesh at
    T(0,10,-30), random spheres, ground plane, the two mesh-bunny lights."""
    u = _mt_uniform(ntri * 12 + nspheres * 8, seed)
    c = (u[: ntri * 3].reshape(ntri, 1, 3) * 2.0 - 1.0) * 10.0
    off = (u[ntri * 3: ntri * 12].reshape(ntri, 3, 3) * 2.0 - 1.0) * 0.15
    tri = c + off
    mesh = trianglesToMesh(tri)
    mesh.objectToWorld = L.translate(L.mat4(1.0), vec3(0.0, 10.0, -30.0))
    mesh.worldToObject = L.inverse(mesh.objectToWorld)
    s = u[ntri * 12:].reshape(nspheres, 8)
    objects = [Object("soup", mesh, Material(albedo=vec3(0.6, 0.9, 0.2))),
               Object("ground", initPlane(objectToWorld=L.mat4(1.0)), Material(albedo=vec3(0.4)))]
    for i in range(nspheres):
        x = -30.0 + 60.0 * s[i, 0]
        y = 20.0 * s[i, 1]
        z = -60.0 + 50.0 * s[i, 2]
        r = 0.05 + 0.25 * s[i, 3]
        refl = 0.5 if s[i, 7] < 0.1 else 0.0
        objects.append(Object(f"s{i}", initSphere(r=r, objectToWorld=L.translate(L.mat4(1.0), vec3(x, y, z))),
                              Material(albedo=vec3(s[i, 4], s[i, 5], s[i, 6]), reflection=refl)))
    return Scene(objects=objects, lights=_bunny_lights(), fov=50.0, cameraToWorld=_camera(0.0, 5.5, 1.5),
                 bgColor=vec3(0.01, 0.03, 0.05))


def scaled_scene(k: float, stride: int = 64, nspheres: int = 5) -> Scene:
    """Build-defined robustness scene: the mesh-bunny layout (decimated bunny, ground plane, the two
    lights) plus a few spheres (every other one a mirror), with EVERY length multiplied by k — from
    1e-18 to 1e25 the float32 first looks run out of range one after the other and must hand over to
    the float64 path without changing a bit of the result."""
    rs = np.random.RandomState(1)
    mesh = trianglesToMesh(bunny_triangles(stride=stride) * k)
    mesh.objectToWorld = L.translate(L.mat4(1.0), vec3(0.0, 0.0001 * k, -12.0 * k))
    mesh.worldToObject = L.inverse(mesh.objectToWorld)
    objects = [Object("mesh", mesh, Material(albedo=vec3(0.6, 0.9, 0.2))),
               Object("ground", initPlane(objectToWorld=L.mat4(1.0)), Material(albedo=vec3(0.4)))]
    for i in range(nspheres):
        c = (rs.uniform(-3, 3) * k, rs.uniform(0.5, 3) * k, (-10 + rs.uniform(-2, 2)) * k)
        objects.append(Object(f"s{i}", initSphere(r=rs.uniform(0.3, 1.2) * k, objectToWorld=L.translate(L.mat4(1.0), vec3(*c))),
                              Material(albedo=vec3(0.8, 0.4, 0.2), reflection=0.4 if i % 2 else 0.0)))
    return Scene(objects=objects, lights=_bunny_lights(), fov=50.0, cameraToWorld=_camera(0.0, 5.5 * k, 1.5 * k),
                 bgColor=vec3(0.1, 0.2, 0.3))


def transformed_objects(stride: int = 16) -> Scene:
    """Build-defined parity scene: every geometry kind under a NON-translation objectToWorld
    (rotation + non-uniform scale), so worldToObject*dir is not unit length (the `(x/2)*a` sphere
    quirk of geom.nim:232 matters) and the literal glm mat*vec path of trace() (renderer.nim:54-55)
    is exercised; plus a point light (GENERAL shadow bundle) and a mirror."""
    def xf(t, axis, deg, s):
        return L.scale(L.rotate(L.translate(L.mat4(1.0), vec3(*t)), axis, L.deg_to_rad(deg)), s)
    mesh = trianglesToMesh(bunny_triangles(stride=stride))
    mesh.objectToWorld = xf((1.5, 0.2, -11.0), L.Y_AXIS, 35.0, (0.9, 1.2, 0.8))
    mesh.worldToObject = L.inverse(mesh.objectToWorld)
    objects = [
        Object("mesh", mesh, Material(albedo=vec3(0.6, 0.9, 0.2), reflection=0.3)),
        Object("ground", initPlane(objectToWorld=xf((0.0, -0.2, 0.0), L.Z_AXIS, 3.0, (1.0, 1.0, 1.0))),
               Material(albedo=vec3(0.4))),
        Object("ellipsoid", initSphere(r=1.0, objectToWorld=xf((-3.0, 1.5, -9.0), L.X_AXIS, 20.0, (1.5, 0.7, 1.1))),
               Material(albedo=vec3(0.9, 0.3, 0.2), reflection=0.6)),
        Object("box", initBox(objectToWorld=xf((3.5, 1.0, -13.0), L.Y_AXIS, 30.0, (1.0, 1.4, 0.6)),
                              vmin=vec(-1.0, -1.0, -1.0), vmax=vec(1.0, 1.0, 1.0)),
               Material(albedo=vec3(0.2, 0.4, 0.9))),
    ]
    lights = [
        PointLight(color=vec3(1.0, 0.9, 0.8), intensity=2500.0, pos=point(2.0, 7.0, -6.0)),
        DistantLight(color=vec3(0.4, 0.5, 0.7), intensity=1.5, dir=L.normalize(vec(-1.0, -1.0, -0.6))),
    ]
    return Scene(objects=objects, lights=lights, fov=55.0, cameraToWorld=_camera(0.5, 4.5, 2.0, -10.0),
                 bgColor=vec3(0.05, 0.06, 0.1))


def with_many_lights(sc: Scene, n: int, seed: int = 5) -> Scene:
    """`sc` lit by n - 1 random DistantLights from above and one PointLight (test scenes around the
    32-light limit of the fused shadow + resolve kernel)."""
    from .api import point
    rng = np.random.default_rng(seed)
    lights = []
    for _ in range(n - 1):
        x, z = rng.uniform(-1.0, 1.0, 2)
        lights.append(DistantLight(color=vec3(*rng.uniform(0.2, 1.0, 3)), intensity=float(rng.uniform(0.05, 0.3)),
                                   dir=L.normalize(vec(float(x), -1.0, float(z)))))
    lights.append(PointLight(color=vec3(1.0), intensity=600.0, pos=point(2.0, 7.0, -9.0)))
    sc.lights = lights
    return sc


def many_meshes_many_lights(nmesh: int = 12, nlights: int = 32, seed: int = 9) -> Scene:
    """Build-defined scene at the edge of the fused shadow-gate kernel's shared memory (32 lights x 12 mesh objects:
    32 * 12 * 34 counters = 52 KiB > 48 KiB): `nmesh` small meshes (decimated bunnies at different places, sharing
    nothing), a ground plane, a mirror ball, `nlights` lights."""
    rng = np.random.default_rng(seed)
    objects = [Object("ground", initPlane(objectToWorld=L.mat4(1.0)), Material(albedo=vec3(0.4)))]
    for k in range(nmesh):
        mesh = trianglesToMesh(bunny_triangles(stride=256 + 16 * k))
        mesh.objectToWorld = L.translate(L.mat4(1.0), vec3(float(-9.0 + 3.5 * (k % 6)), 0.0001, float(-10.0 - 5.0 * (k // 6))))
        mesh.worldToObject = L.inverse(mesh.objectToWorld)
        objects.append(Object(f"mesh{k}", mesh, Material(albedo=vec3(*rng.uniform(0.3, 0.9, 3)))))
    objects.append(Object("mirror", initSphere(r=1.5, objectToWorld=L.translate(L.mat4(1.0), vec3(0.0, 1.5, -6.0))),
                          Material(albedo=vec3(0.9), reflection=0.7)))
    sc = Scene(objects=objects, lights=[], fov=60.0, cameraToWorld=_camera(0.0, 6.0, 4.0, -15.0), bgColor=vec3(0.02, 0.03, 0.05))
    return with_many_lights(sc, nlights, seed)
