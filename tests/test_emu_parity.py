"""CPU-side (no GPU) checks of the device code: the SAME per-element bodies and the SAME
host orchestration that libnrt.so compiles for sm_100a (csrc/nrt_core.h, nrt_pipeline.h,
nrt_renderer.h) run through the test-only loop backend (tests/emu) and must reproduce the
oracle bit for bit.  This validates the wavefront logic, the float64 arithmetic order and —
most importantly — that the float32 mesh filter never drops a pair the float64 reference
accepts.  The -m gpu tests repeat the same comparisons on the real CUDA path."""
import numpy as np
import pytest

import emu_binding as emu
from nim_raytracer_b200 import api, scenes


def check(scene, opts, oracle_mod, **kw):
    a1, a2 = api.Aov(opts.width, opts.height), api.Aov(opts.width, opts.height)
    rfb, rst, _ = oracle_mod.render(scene, opts, aov=a1, **kw)
    fb, st, _, prof = emu.render(scene, opts, aov=a2, **kw)
    assert (a1.obj_id == a2.obj_id).all() and (a1.tri_id == a2.tri_id).all() and (a1.t_hit == a2.t_hit).all()
    assert (fb.data == rfb.data).all()
    assert st == rst
    return prof


def test_spheres_boxes_no_mesh(oracle_mod):
    check(scenes.spheres_reflection(), api.Options(160, 120), oracle_mod)
    check(scenes.boxtest(), api.Options(96, 64), oracle_mod)


def test_one_triangle_mesh(oracle_mod):
    prof = check(scenes.mesh_cube(), api.Options(64, 64), oracle_mod)
    assert prof["mesh_rays"] > 0


def test_bunny_full_mesh(oracle_mod, monkeypatch):
    # (the per-ray accounting below is the wavefront's: with the fused path only the samples that come near the mesh
    # reach it, and their few rays no longer fill whole 256-ray runs at this tiny resolution)
    monkeypatch.setenv("NRT_PATH", "0")
    monkeypatch.setenv("NRT_PREFILTER_CULL", "0")
    prof = check(scenes.bunny(), api.Options(160, 90), oracle_mod)
    # shared-origin (primary) and shared-direction (distant-light shadow) bundles drop the faces
    # that can never be hit (plane faces away / det < 1e-6): roughly half of the 69,451
    assert prof["exact_rays"] == 0 and 0.4 * 69451 < prof["mesh_tests"] / prof["mesh_rays"] < 0.6 * 69451
    # the filter is selective: ~1 float64 re-evaluation per ray that enters the box
    assert prof["candidates"] < 2 * prof["mesh_rays"]
    # two-level traversal (Morton-ordered 256-record chunks + chunk bounds): same image, ids and
    # stats, same candidates, fewer tests even at this tiny resolution (a 256-ray run spans 1.6 rows)
    monkeypatch.setenv("NRT_PREFILTER_CULL", "1")
    prof2 = check(scenes.bunny(), api.Options(160, 90), oracle_mod)
    assert prof2["candidates"] == prof["candidates"] and prof2["pre_candidates"] == prof["pre_candidates"]
    assert prof2["mesh_tests"] < 0.7 * prof["mesh_tests"]


def test_bunny_native_winding(oracle_mod):
    check(scenes.bunny(flip_winding=False, stride=4), api.Options(160, 90), oracle_mod)


def test_reflection_grid_aa_both_depth_modes(oracle_mod):
    sc = scenes.bunny_spheres(stride=8)
    check(sc, api.Options(120, 68, antialias=api.Antialias(api.akGrid, 2), depthMode=api.NRT_DEPTH_INTENDED, maxRayDepth=8), oracle_mod)
    check(sc, api.Options(96, 54), oracle_mod)
    check(sc, api.Options(64, 36, depthMode=api.NRT_DEPTH_INTENDED, maxRayDepth=0), oracle_mod)
    check(sc, api.Options(64, 36, bounceCap=2), oracle_mod)        # safety cap reached: counted, same colours


@pytest.mark.parametrize("kind", [api.akJittered, api.akMultiJittered, api.akCorrelatedMultiJittered])
def test_jittered(oracle_mod, kind):
    check(scenes.spheres_reflection(), api.Options(48, 36, antialias=api.Antialias(kind, 3), seed=5), oracle_mod)


def test_progressive_and_line_ranges(oracle_mod):
    sc, o = scenes.bunny(stride=32), api.Options(64, 36)
    fb, rfb = api.newFramebuf(64, 36), api.newFramebuf(64, 36)
    step = 8
    while step >= 1:
        emu.render(sc, o, fb=fb, step=step, maxStep=8)
        oracle_mod.render(sc, o, fb=rfb, step=step, maxStep=8)
        assert (fb.data == rfb.data).all()
        step //= 2
    fb2 = api.newFramebuf(64, 36)
    for y in range(0, 36, 5):                                        # ragged line ranges
        emu.render(sc, o, fb=fb2, y0=y, y1=min(y + 5, 36))
    assert (fb2.data == fb.data).all()


def test_chunking_overflow_retry_and_exact_path(oracle_mod, monkeypatch):
    sc, o = scenes.bunny_spheres(stride=16), api.Options(100, 60, antialias=api.Antialias(api.akGrid, 2))
    monkeypatch.setenv("NRT_CHUNK_SAMPLES", "1000")
    check(sc, o, oracle_mod)
    monkeypatch.setenv("NRT_CAND_CAP", "8")
    check(sc, o, oracle_mod)
    monkeypatch.delenv("NRT_CAND_CAP")
    monkeypatch.setenv("NRT_PAIR_CAP", "4")      # (run, chunk) work list overflows -> re-render with 8x
    check(sc, o, oracle_mod)
    monkeypatch.delenv("NRT_PAIR_CAP")
    monkeypatch.setenv("NRT_FORCE_EXACT", "1")
    prof = check(sc, o, oracle_mod)
    assert prof["mesh_tests"] == 0 and prof["exact_rays"] == prof["mesh_rays"]


def test_filter_is_conservative_on_random_soup(oracle_mod):
    # random triangle soup (template test/test.nim:29-40), rays from many origins incl. reflections
    rs = np.random.RandomState(3)
    tri = (rs.uniform(-1, 1, (4000, 1, 3)) * 3.0) + rs.uniform(-0.4, 0.4, (4000, 3, 3))
    from nim_raytracer_b200 import loaders, linalg as L
    mesh = loaders.trianglesToMesh(tri)
    sc = scenes.mesh_scene(mesh, reflection=0.5)
    sc.objects[0].geometry.objectToWorld = L.translate(L.mat4(1.0), api.vec3(0.0, 3.5, -12.0))
    sc.objects[0].geometry.worldToObject = L.inverse(sc.objects[0].geometry.objectToWorld)
    check(sc, api.Options(120, 68), oracle_mod)


def test_general_transforms_point_light_mirror(oracle_mod):
    # rotation + non-uniform scale on every geometry kind: the literal mat*vec path, non-unit object-space
    # directions, GENERAL-mode shadow rays (point light) and reflection rays hitting a mesh
    prof = check(scenes.transformed_objects(), api.Options(120, 68), oracle_mod)
    assert prof["mesh_rays"] > 500 and prof["exact_rays"] == 0
    check(scenes.transformed_objects(stride=64), api.Options(64, 36, antialias=api.Antialias(api.akGrid, 2)), oracle_mod)


def test_stress_scene_small(oracle_mod):
    # BASELINE config 5 at reduced size (template test/test.nim:29-40)
    check(scenes.stress(ntri=800, nspheres=30), api.Options(64, 36, antialias=api.Antialias(api.akGrid, 2)), oracle_mod)


def test_sphere_clusters(oracle_mod):
    # >= 64 spheres: the object scan goes through the two-level sphere clusters (Morton-ordered groups of
    # 16 x 16 with bounding spheres) and evaluates the survivors in list order; mirrors among them
    sc = scenes.stress(ntri=300, nspheres=700)
    check(sc, api.Options(80, 45, antialias=api.Antialias(api.akGrid, 2), depthMode=api.NRT_DEPTH_INTENDED, maxRayDepth=3), oracle_mod)
    sc = scenes.stress(ntri=50, nspheres=70, seed=7)       # one partly filled level-1 cluster
    check(sc, api.Options(64, 36), oracle_mod)


@pytest.mark.parametrize("k", [1e-18, 1e-6, 1.0, 1e6, 1e12, 1e18, 1e25])
def test_extreme_scales(oracle_mod, k):
    # every length of the scene times k: float32 first looks (sphere / plane / box / chunk bounds) run out of
    # range and the float64 path takes over; bias 0 puts shadow and reflection origins exactly on the surfaces
    sc = scenes.scaled_scene(k)
    check(sc, api.Options(64, 40, antialias=api.Antialias(api.akGrid, 2), bias=1e-8 * k), oracle_mod)
    check(sc, api.Options(48, 30, bias=0.0), oracle_mod)


def test_separate_shadow_and_resolve_launches(oracle_mod, monkeypatch):
    # ShadowResolve serves up to 32 lights (occlusion bits in one register); beyond that, and with
    # NRT_FUSE_RESOLVE=0, ShadowTrace + Resolve run as separate launches through cs.occ
    sc = scenes.bunny_spheres(stride=16)
    o = api.Options(72, 40, antialias=api.Antialias(api.akGrid, 2), depthMode=api.NRT_DEPTH_INTENDED, maxRayDepth=3)
    monkeypatch.setenv("NRT_FUSE_RESOLVE", "0")
    check(sc, o, oracle_mod)
    check(scenes.spheres_reflection(), api.Options(96, 72), oracle_mod)        # point light: Resolve reads the hit point
    monkeypatch.delenv("NRT_FUSE_RESOLVE")
    check(scenes.with_many_lights(scenes.bunny_spheres(stride=32), 32), api.Options(48, 28), oracle_mod)   # fused, all 32 bits used
    check(scenes.with_many_lights(scenes.bunny_spheres(stride=32), 33), api.Options(48, 28), oracle_mod)   # one too many: separate launches


def test_degenerate_meshes(oracle_mod):
    from nim_raytracer_b200 import loaders, linalg as L
    # zero-area and needle triangles, duplicated coplanar faces (first index must win)
    tri = np.array([
        [[0, 0, 0], [0, 0, 0], [0, 0, 0]],
        [[-1, 0.5, 0], [1, 0.5, 0], [0, 2.0, 0]],
        [[-1, 0.5, 0], [1, 0.5, 0], [0, 2.0, 0]],
        [[-1, 0.5, -1e-9], [1, 0.5, -1e-9], [0, 2.0, -1e-9]],
        [[0, 0.5, 0], [1e-7, 0.5, 0], [0, 3, 0]],
    ], dtype=np.float64)
    mesh = loaders.trianglesToMesh(tri)
    sc = scenes.mesh_scene(mesh)
    a1 = api.Aov(96, 54)
    oracle_mod.render(sc, api.Options(96, 54), aov=a1)
    assert set(np.unique(a1.tri_id)) >= {-1, 1}          # duplicate #2 never wins over #1
    check(sc, api.Options(96, 54), oracle_mod)


def test_many_mesh_objects_and_lights(oracle_mod):
    # several mesh objects (kMaxWalkMO = 2 result slots of the path kernels: the third is walked by the thread) x many lights
    check(scenes.many_meshes_many_lights(3, 5), api.Options(64, 36, antialias=api.Antialias(api.akGrid, 2)), oracle_mod)
    check(scenes.many_meshes_many_lights(12, 33), api.Options(32, 18), oracle_mod)


def test_short_wavefront_list_goes_to_path_tail(oracle_mod, monkeypatch):
    # the library's default: a bounce's wavefront list below NRT_HARD_TAIL_BELOW (16384) samples is finished by ONE
    # PathTail launch — from bounce 0 too, where only the primary direction is stored (PathWarpT, bounce0 == 0)
    monkeypatch.delenv("NRT_HARD_TAIL_BELOW")
    prof = check(scenes.mesh_cube(), api.Options(64, 64), oracle_mod)
    assert prof["mesh_rays"] == 0            # no wavefront queue saw a ray
    sc = scenes.bunny_spheres(stride=16)
    check(sc, api.Options(96, 54, antialias=api.Antialias(api.akGrid, 2), depthMode=api.NRT_DEPTH_INTENDED, maxRayDepth=6), oracle_mod)
    check(sc, api.Options(80, 45, antialias=api.Antialias(api.akJittered, 2), seed=5), oracle_mod)
    monkeypatch.setenv("NRT_HARD_TAIL_BELOW", "300")   # bounce 0 through the wavefront, later (shorter) lists through PathTail
    check(sc, api.Options(96, 54, antialias=api.Antialias(api.akGrid, 2), depthMode=api.NRT_DEPTH_INTENDED, maxRayDepth=6), oracle_mod)


def test_fork_moves_the_continuing_pool_to_a_helper_pipeline(oracle_mod, monkeypatch):
    # After bounce 0 the samples FusedBounce finished with a reflection ray stored are moved into a helper pipeline's
    # sample space (GatherPool), taken to the end of their paths there and scattered back (nrt_renderer.h: the fork).
    sc = scenes.bunny_spheres(stride=16)
    o = api.Options(96, 54, antialias=api.Antialias(api.akGrid, 2), depthMode=api.NRT_DEPTH_INTENDED, maxRayDepth=6)
    monkeypatch.setenv("NRT_FORK_MIN", "0")
    off = check(sc, o, oracle_mod)
    monkeypatch.setenv("NRT_FORK_MIN", "16")
    on = check(sc, o, oracle_mod)
    assert on["launches"] != off["launches"]          # the helper's gather / scatter and its own chain of launches
    check(sc, api.Options(80, 45), oracle_mod)        # reference depth bug: paths up to the bounce cap
    check(scenes.transformed_objects(), api.Options(120, 68, antialias=api.Antialias(api.akGrid, 2)), oracle_mod)
    monkeypatch.delenv("NRT_HARD_TAIL_BELOW")          # with the short-list shortcut on both sides of the fork
    monkeypatch.setenv("NRT_HARD_TAIL_BELOW", "64")
    check(sc, o, oracle_mod)


def _grid_scene(big=False, far=False):
    """>= 64 spheres (clusters + the light-space shadow grids), a distant and a point light; optionally a few huge
    spheres (cells with long lists: the grid declines them) or everything moved far from the origin (the grid's
    float32 margin rejects the rays and the cluster traversal takes them)."""
    from nim_raytracer_b200.api import DistantLight, PointLight, point, vec, vec3
    from nim_raytracer_b200 import linalg as L
    sc = scenes.stress(ntri=200, nspheres=300, seed=3)
    sc.lights = [DistantLight(color=vec3(1.0), intensity=3.0, dir=L.normalize(vec(-1.0, -1.5, -0.4))),
                 PointLight(color=vec3(1.0, 0.8, 0.5), intensity=3000.0, pos=point(2.0, 9.0, -25.0)),
                 DistantLight(color=vec3(0.2, 0.3, 0.8), intensity=1.0, dir=L.normalize(vec(0.0, -1.0, 0.0)))]
    return sc


def test_light_space_shadow_grid(oracle_mod):
    # shadow rays of the DistantLights read one cell of the light-space grid of the clustered spheres (ShadowGridF);
    # the PointLight's rays and the path rays traverse the clusters
    o = api.Options(96, 54, antialias=api.Antialias(api.akGrid, 2), depthMode=api.NRT_DEPTH_INTENDED, maxRayDepth=3)
    check(_grid_scene(), o, oracle_mod)
    check(scenes.stress(ntri=100, nspheres=2000, seed=11), api.Options(120, 68), oracle_mod)   # dense: long cell lists


@pytest.mark.parametrize("hot", ["1", "0"])
def test_scene_header_by_value_with_and_without_hot_copies(oracle_mod, monkeypatch, hot):
    """The per-sample functors carry the scene header by value (nrt_pipeline.h: SceneArg); with DScene::hotOk the
    lights, grid headers and mesh gate records are read from the header's own copies, otherwise (NRT_HOT_HEADER=0, or a
    scene with more than two lights / more than one mesh object) from the device tables.  Same bits either way, on the
    fused path and on the wavefront-only path."""
    monkeypatch.setenv("NRT_HOT_HEADER", hot)
    sc = scenes.bunny_spheres(stride=8)
    o = api.Options(96, 54, antialias=api.Antialias(api.akGrid, 2), depthMode=api.NRT_DEPTH_INTENDED, maxRayDepth=4)
    for path in ("1", "0"):
        monkeypatch.setenv("NRT_PATH", path)
        check(sc, o, oracle_mod)
    monkeypatch.delenv("NRT_PATH")
    check(scenes.transformed_objects(), api.Options(80, 60), oracle_mod)   # point light + rotated boxes: hot copy of a PointLight


def test_overflowed_lists_grow_to_what_the_attempt_asked_for(oracle_mod, monkeypatch):
    """A one-entry candidate list / pair list (the smallest a caller can ask for) must still end in the right frame: the
    re-render sizes the lists from the overflowed attempt's counters (pairs -> pre-candidates -> candidates, one resize
    each at worst), not only by a fixed factor (nrt_renderer.h: renderRows, need_cand / need_pairs)."""
    sc, o = scenes.bunny_spheres(stride=16), api.Options(100, 60, antialias=api.Antialias(api.akGrid, 2))
    for path in ("0", "1"):
        monkeypatch.setenv("NRT_PATH", path)
        monkeypatch.setenv("NRT_CAND_CAP", "1")
        check(sc, o, oracle_mod)
        monkeypatch.setenv("NRT_PAIR_CAP", "1")
        check(sc, o, oracle_mod)
        monkeypatch.delenv("NRT_CAND_CAP")
        check(sc, o, oracle_mod)
        monkeypatch.delenv("NRT_PAIR_CAP")


def test_random_scenes_replay(oracle_mod):
    """A fixed slice of tools/fuzz_emu.py's seeds (random objects under random affine transforms, lights, cameras, render
    options and path knobs; every third seed also as progressive passes and ragged line ranges) and of its
    degenerate-input seeds (zero / negative radii, inverted boxes, zero-area / huge / tiny / non-finite vertices, singular
    transforms, zero light directions, extreme fov): ids, tHit, framebuffer bits and Stats equal the oracle's.  The tool
    has been run over seeds 0..40,000 and 12,000 degenerate ones without a difference."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import fuzz_emu
    seeds = [s for s in range(120) if s != 87] + list(range(fuzz_emu.DEGENERATE, fuzz_emu.DEGENERATE + 60))   # (87: 43 s)
    bad = {s: m for s in seeds if (m := fuzz_emu.run(s))}
    assert not bad, bad
