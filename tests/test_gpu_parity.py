"""-m gpu: parity of the CUDA path (through the C ABI, include/nrt.h) against the
CPU oracle on the same inputs, plus full-size golden fixtures and size-independent
properties.  Integer/index outputs must be bit-exact; the float32 framebuffer is
compared bit-exactly too (all shading is float64 in the reference's operation
order), with the north-star tolerance (|dRGB| <= 1/255 on >= 99.9 % of pixels)
asserted as the contractual bound."""
import hashlib
import os

import numpy as np
import pytest

from nim_raytracer_b200 import api, scenes

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module", autouse=True)
def _init():
    api.initRenderer(1)
    yield


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def gpu_render(scene, opts, **kw):
    fb, aov = api.newFramebuf(opts.width, opts.height), api.Aov(opts.width, opts.height)
    st = api.renderFrame(scene, opts, fb, aov=aov, **kw)
    return fb, st, aov


def assert_parity(scene, opts, oracle_mod, **kw):
    fb, st, aov = gpu_render(scene, opts, **kw)
    rfb, rst, raov = oracle_mod.render(scene, opts, aov=api.Aov(opts.width, opts.height), **kw)
    assert (aov.obj_id == raov.obj_id).all(), f"{(aov.obj_id != raov.obj_id).sum()} object ids differ"
    assert (aov.tri_id == raov.tri_id).all(), f"{(aov.tri_id != raov.tri_id).sum()} primitive ids differ"
    assert (aov.t_hit == raov.t_hit).all()
    d = np.abs(fb.image().astype(np.float64) - rfb.image().astype(np.float64)).max(axis=2)
    assert (d <= 1.0 / 255.0).mean() >= 0.999           # north-star tolerance
    assert (fb.data == rfb.data).all(), f"max |dRGB| = {d.max()}"  # and in fact bit-exact
    assert st == rst
    return fb, st, aov


def test_config1_spheres_reflection(oracle_mod):
    # BASELINE config 1: spheres-reflection.nim, 640x480, akNone, bias 1e-8, maxRayDepth 5
    assert_parity(scenes.spheres_reflection(), api.Options(640, 480), oracle_mod)


def test_boxtest_and_one_triangle_mesh(oracle_mod):
    assert_parity(scenes.boxtest(), api.Options(300, 200), oracle_mod)       # data/scenes/boxtest.nim
    assert_parity(scenes.mesh_cube(), api.Options(128, 128), oracle_mod)     # data/scenes/mesh-cube.nim


def test_meshperftest_triangle():
    # test/meshperftest.nim:22-44: 1-triangle mesh from the origin along -z => t = 5
    from nim_raytracer_b200 import linalg as L
    v = np.array([api.point(0, 1, -5), api.point(-2, -1, -5), api.point(2, -1, -5)])
    mesh = api.initTriangleMesh(v, np.array([api.vec(0, 0, 1)]), [[0, 1, 2]], [[0, 0, 0]], L.mat4(1.0))
    sc = api.Scene([api.Object("m", mesh, api.Material(api.vec3(1.0)))], [], 90.0, L.mat4(1.0), api.vec3(0.0))
    fb, st, aov = gpu_render(sc, api.Options(2, 2))
    assert aov.obj_id[3] == 0 and aov.tri_id[3] == 0 and aov.t_hit[3] == 5.0


def test_bunny_full_mesh_small_frame(oracle_mod):
    # all 69,451 triangles, 320x180 (the oracle finishes in seconds)
    _, _, aov = assert_parity(scenes.bunny(), api.Options(320, 180), oracle_mod)
    assert (aov.tri_id >= 0).sum() > 1000


def test_bunny_native_winding(oracle_mod):
    # secondary parity case of SURVEY §8d: the unflipped mesh (what the reference would literally do)
    assert_parity(scenes.bunny(flip_winding=False, stride=2), api.Options(240, 135), oracle_mod)


def test_config3_reflection_grid_aa(oracle_mod):
    sc = scenes.bunny_spheres(stride=4)
    o = api.Options(240, 135, antialias=api.Antialias(api.akGrid, 2), depthMode=api.NRT_DEPTH_INTENDED, maxRayDepth=8)
    assert_parity(sc, o, oracle_mod)
    o = api.Options(160, 90, antialias=api.Antialias(api.akGrid, 4))   # literal reference depth semantics (REFBUG)
    assert_parity(sc, o, oracle_mod)


@pytest.mark.parametrize("kind", [api.akJittered, api.akMultiJittered, api.akCorrelatedMultiJittered])
def test_jittered_kinds(oracle_mod, kind):
    o = api.Options(160, 120, antialias=api.Antialias(kind, 3), seed=20161018)
    assert_parity(scenes.spheres_reflection(), o, oracle_mod)


def test_general_transforms_point_light_mirror(oracle_mod):
    # every geometry kind under rotation + non-uniform scale, point light, mirror (GENERAL bundles)
    assert_parity(scenes.transformed_objects(), api.Options(240, 135), oracle_mod)
    assert_parity(scenes.transformed_objects(stride=4), api.Options(160, 90, antialias=api.Antialias(api.akGrid, 2)), oracle_mod)


def test_progressive_refinement(oracle_mod):
    # gui.nim:113-122,254-257: step halves from maxStep to 1; the result is the 1-step frame
    sc, o = scenes.bunny(stride=16), api.Options(128, 72)
    full, _, _ = gpu_render(sc, o)
    fb = api.newFramebuf(128, 72)
    rfb = api.newFramebuf(128, 72)
    step = 8
    while step >= 1:
        api.renderFrame(sc, o, fb, step=step, maxStep=8)
        oracle_mod.render(sc, o, fb=rfb, step=step, maxStep=8)
        assert (fb.data == rfb.data).all(), f"step {step}"
        step //= 2
    assert (fb.data == full.data).all()


def test_render_line_matches_frame(oracle_mod):
    # renderer.nim:162: the per-scanline entry point used by the worker-pool caller
    sc, o = scenes.boxtest(), api.Options(96, 64)
    ds = api.DeviceScene(sc)
    fb = api.newFramebuf(96, 64)
    total = api.Stats()
    for y in range(64):
        total += api.renderLine(ds, o, fb, y)
    rfb, rst, _ = oracle_mod.render(sc, o)
    assert (fb.data == rfb.data).all() and total == rst
    with pytest.raises(AssertionError):
        api.renderLine(ds, o, fb, 0, step=3)             # renderer.nim:166 assert isPowerOfTwo(step)


def test_chunked_and_exact_paths_agree(oracle_mod, monkeypatch):
    sc, o = scenes.bunny_spheres(stride=8), api.Options(200, 112, antialias=api.Antialias(api.akGrid, 2))
    ref, rst, raov = gpu_render(sc, o)
    monkeypatch.setenv("NRT_CHUNK_SAMPLES", "4096")      # many chunks
    fb, st, aov = gpu_render(sc, o)
    assert (fb.data == ref.data).all() and st == rst and (aov.tri_id == raov.tri_id).all()
    monkeypatch.setenv("NRT_FORCE_EXACT", "1")           # float64 brute force instead of filter + verify
    fb, st, aov = gpu_render(sc, o)
    assert (fb.data == ref.data).all() and st == rst and (aov.tri_id == raov.tri_id).all()
    monkeypatch.delenv("NRT_FORCE_EXACT")
    monkeypatch.setenv("NRT_CAND_CAP", "64")             # forces the overflow -> retry path
    fb, st, aov = gpu_render(sc, o)
    assert (fb.data == ref.data).all() and st == rst


def test_two_level_traversal_equals_brute_force(monkeypatch):
    # chunk bounds + (run, chunk) work list vs every chunk of the record set: same ids, image, stats and
    # candidates; far fewer executed tests.  Then the work list overflows on purpose: re-render path.
    sc, o = scenes.bunny_spheres(stride=2), api.Options(480, 270, antialias=api.Antialias(api.akGrid, 2))
    ds = api.DeviceScene(sc)

    def run():
        fb, aov = api.newFramebuf(o.width, o.height), api.Aov(o.width, o.height)
        st = api.renderFrame(ds, o, fb, aov=aov)
        return fb, st, aov, ds.profile()

    monkeypatch.setenv("NRT_PREFILTER_CULL", "0")
    b_fb, b_st, b_aov, b_p = run()
    monkeypatch.setenv("NRT_PREFILTER_CULL", "1")
    c_fb, c_st, c_aov, c_p = run()
    assert (b_fb.data == c_fb.data).all() and b_st == c_st
    assert (b_aov.tri_id == c_aov.tri_id).all() and (b_aov.obj_id == c_aov.obj_id).all() and (b_aov.t_hit == c_aov.t_hit).all()
    assert b_p.candidates == c_p.candidates and b_p.pre_candidates == c_p.pre_candidates
    assert c_p.mesh_tests < 0.5 * b_p.mesh_tests          # (a 256-ray run spans 8 pixels x 4 spp here; ~1/16 at 4K x 16 spp)
    monkeypatch.setenv("NRT_PAIR_CAP", "64")
    d_fb, d_st, d_aov, _ = run()
    assert (d_fb.data == c_fb.data).all() and d_st == c_st and (d_aov.tri_id == c_aov.tri_id).all()


def test_kernel_timing_hook():
    sc, o = scenes.bunny(stride=8), api.Options(320, 180)
    ds = api.DeviceScene(sc)
    fb = api.newFramebuf(o.width, o.height)
    api.setKernelTiming(True)
    try:
        api.renderFrame(ds, o, fb)
        kt = ds.kernelTimes()
    finally:
        api.setKernelTiming(False)
    assert "k_mesh_prefilter" in kt and "Shade" in kt and any(k.startswith("ShadowResolve") for k in kt)
    assert all(ms >= 0 and n > 0 for ms, n in kt.values())
    assert abs(sum(ms for ms, _ in kt.values()) - ds.profile().total_ms) < max(1.0, ds.profile().total_ms)


def _check_golden(name, scene, opts):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    fb, st, aov = gpu_render(scene, opts)
    h, w = opts.height, opts.width
    rows = g["rows"]
    obj_rows = aov.obj_id.reshape(h, w)[rows]
    tri_rows = aov.tri_id.reshape(h, w)[rows]
    assert (obj_rows == g["obj_rows"]).all(), f"{(obj_rows != g['obj_rows']).sum()} object ids differ on the sampled rows"
    assert (tri_rows == g["tri_rows"]).all(), f"{(tri_rows != g['tri_rows']).sum()} primitive ids differ on the sampled rows"
    d = np.abs(fb.image()[rows].astype(np.float64) - g["fb_rows"].astype(np.float64)).max(axis=2)
    assert (d <= 1.0 / 255.0).mean() >= 0.999
    assert (fb.image()[rows] == g["fb_rows"]).all(), f"max |dRGB| on sampled rows = {d.max()}"
    assert digest(aov.obj_id) == str(g["obj_sha256"]) and digest(aov.tri_id) == str(g["tri_sha256"])
    assert digest(aov.t_hit) == str(g["t_sha256"])
    assert digest(fb.data) == str(g["fb_sha256"])
    assert [st.numPrimaryRays, st.numIntersectionTests, st.numIntersectionHits, st.numRays,
            st.numCappedSamples] == g["stats"].tolist()
    return fb, st, aov


def test_golden_config1_full_size():
    _check_golden("config1_spheres_640x480", scenes.spheres_reflection(), api.Options(640, 480))


def test_golden_config2_full_size():
    # BASELINE config 2 at its full size: 1920x1080, 69,451 triangles, primary + shadow rays
    fb, st, aov = _check_golden("config2_bunny_1920x1080", scenes.bunny(), api.Options(1920, 1080))
    # size-independent properties: Stats identities of renderer.nim:58,138
    assert st.numPrimaryRays == 1920 * 1080 and st.numIntersectionTests == st.numRays * 2


def test_golden_config3_reduced():
    o = api.Options(480, 270, antialias=api.Antialias(api.akGrid, 2), depthMode=api.NRT_DEPTH_INTENDED, maxRayDepth=8)
    _check_golden("config3_bunny_spheres_480x270_g2", scenes.bunny_spheres(), o)


def _check_row_crcs(name, fb, aov):
    """Per-scanline CRC32 of the framebuffer and ids against the fixture: names the first rows that differ."""
    import zlib
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    if "fb_row_crc" not in g:
        return
    h, w = int(g["height"]), int(g["width"])
    crc = np.array([zlib.crc32(r.tobytes()) for r in fb.image()], dtype=np.uint32)
    bad = np.nonzero(crc != g["fb_row_crc"])[0]
    assert bad.size == 0, f"{bad.size} framebuffer rows differ from the oracle, first {bad[:8].tolist()}"
    icrc = np.array([zlib.crc32(a.tobytes() + b.tobytes()) for a, b in
                     zip(aov.obj_id.reshape(h, w), aov.tri_id.reshape(h, w))], dtype=np.uint32)
    bad = np.nonzero(icrc != g["id_row_crc"])[0]
    assert bad.size == 0, f"{bad.size} id rows differ from the oracle, first {bad[:8].tolist()}"


def test_golden_config3_full_size():
    # BASELINE config 3 at its stated size: bunny + 7 spheres (reflection 0 / 0.5 / 1), 1920x1080, 16 spp grid,
    # intended depth 8 - every pixel's ids, t, float32 RGB and the Stats equal the CPU oracle's (renderer.nim:144-159)
    o = api.Options(1920, 1080, antialias=api.Antialias(api.akGrid, 4), depthMode=api.NRT_DEPTH_INTENDED, maxRayDepth=8)
    fb, st, aov = _check_golden("config3_bunny_spheres_1920x1080_g4", scenes.bunny_spheres(), o)
    _check_row_crcs("config3_bunny_spheres_1920x1080_g4", fb, aov)


def test_golden_config4_full_size():
    # BASELINE config 4 — the workload bench.py times — at its stated size: 3840x2160, 16 spp grid, depth 8.
    # 132.7 M samples / 361.9 M rays: every pixel's ids, t, float32 RGB and the Stats equal the CPU oracle's frame
    # (tests/golden/make_goldens.py config4: 38 minutes of brute force on 8 cores; renderer.nim:144-159)
    o = api.Options(3840, 2160, antialias=api.Antialias(api.akGrid, 4), depthMode=api.NRT_DEPTH_INTENDED, maxRayDepth=8)
    fb, st, aov = _check_golden("config4_bunny_spheres_3840x2160_g4", scenes.bunny_spheres(), o)
    _check_row_crcs("config4_bunny_spheres_3840x2160_g4", fb, aov)


def test_golden_config5_class():
    # BASELINE config 5's class: 100,000 random triangles as ONE mesh + 10,000 spheres (sphere clusters in the object scan,
    # 10 % mirrors) + plane, 200x112, 4 spp, intended depth 4 - against the brute-force oracle's frame (27 s on 8 cores).
    # Rendered twice: the first frame takes the fused path, sees that most samples have a ray entering the mesh box and
    # the second takes the wavefront for every bounce (nrt_renderer.h: automatic NRT_PATH) - same frame both times.
    sc = scenes.stress(ntri=100_000, nspheres=10_000)
    o = api.Options(200, 112, antialias=api.Antialias(api.akGrid, 2), depthMode=api.NRT_DEPTH_INTENDED, maxRayDepth=4)
    ds = api.DeviceScene(sc)
    for _ in range(2):
        _check_golden("config5_class_100k_tris_10k_spheres_200x112_g2", ds, o)
    ds.close()


@pytest.mark.parametrize("path", ["0", "2"])
def test_path_modes_agree(oracle_mod, monkeypatch, path):
    # NRT_PATH: 0 = the wavefront for every bounce, 1 (default) = FusedBounce + wavefront + PathTail, 2 = PathMega
    # (one thread per sample, meshes walked by the thread): three formulations, one result
    monkeypatch.setenv("NRT_PATH", path)
    sc = scenes.bunny_spheres(stride=4)
    o = api.Options(320, 180, antialias=api.Antialias(api.akGrid, 2), depthMode=api.NRT_DEPTH_INTENDED, maxRayDepth=8)
    assert_parity(sc, o, oracle_mod)
    assert_parity(sc, api.Options(200, 112), oracle_mod)                 # reference depth bug: cap 64
    assert_parity(scenes.transformed_objects(), api.Options(240, 135, antialias=api.Antialias(api.akGrid, 2)), oracle_mod)
    assert_parity(scenes.bunny(flip_winding=False, stride=2), api.Options(240, 135), oracle_mod)
    assert_parity(scenes.spheres_reflection(), api.Options(160, 120, antialias=api.Antialias(api.akJittered, 3), seed=11), oracle_mod)


def test_scene_header_without_hot_copies_and_cub_select(oracle_mod, monkeypatch):
    # NRT_HOT_HEADER=0: the by-value scene header reads lights / grid headers / gate records from the device tables
    # (what scenes with > 2 lights or > 1 mesh object do); NRT_OWN_SELECT=0: cub::DeviceSelect instead of the
    # 16-flags-per-thread ordered select.  Same bits as the defaults (every other test) and as the oracle.
    sc = scenes.bunny_spheres(stride=4)
    o = api.Options(640, 360, antialias=api.Antialias(api.akGrid, 4), depthMode=api.NRT_DEPTH_INTENDED, maxRayDepth=8)   # 3.7 M samples: the own select runs
    a_fb, a_st, a_aov = gpu_render(sc, o)
    monkeypatch.setenv("NRT_HOT_HEADER", "0")
    monkeypatch.setenv("NRT_OWN_SELECT", "0")
    b_fb, b_st, b_aov = gpu_render(sc, o)
    assert (a_fb.data == b_fb.data).all() and a_st == b_st and (a_aov.tri_id == b_aov.tri_id).all() and (a_aov.t_hit == b_aov.t_hit).all()
    assert_parity(sc, api.Options(200, 112, antialias=api.Antialias(api.akGrid, 2), depthMode=api.NRT_DEPTH_INTENDED, maxRayDepth=8), oracle_mod)


def test_full_size_filter_vs_exact(monkeypatch):
    # BASELINE config 2 scene at 960x540: float32 filter + float64 verify == float64 brute force, all ids
    sc, o = scenes.bunny(), api.Options(960, 540)
    a_fb, a_st, a_aov = gpu_render(sc, o)
    monkeypatch.setenv("NRT_FORCE_EXACT", "1")
    b_fb, b_st, b_aov = gpu_render(sc, o)
    assert (a_aov.tri_id == b_aov.tri_id).all() and (a_aov.obj_id == b_aov.obj_id).all()
    assert (a_fb.data == b_fb.data).all() and a_st == b_st


def test_stress_scene_small(oracle_mod):
    # BASELINE config 5 at reduced size: random triangle soup as ONE mesh + many spheres (10 % mirrors)
    sc = scenes.stress(ntri=3000, nspheres=60)
    assert_parity(sc, api.Options(160, 90, antialias=api.Antialias(api.akGrid, 2)), oracle_mod)


@pytest.mark.parametrize("k", [1e-18, 1e-6, 1e6, 1e18, 1e25])
def test_extreme_scales(oracle_mod, k):
    sc = scenes.scaled_scene(k)
    assert_parity(sc, api.Options(96, 60, antialias=api.Antialias(api.akGrid, 2), bias=1e-8 * k), oracle_mod)
    assert_parity(sc, api.Options(64, 40, bias=0.0), oracle_mod)


def test_separate_shadow_and_resolve_launches(oracle_mod, monkeypatch):
    # NRT_FUSE_RESOLVE=0 and > 32 lights: ShadowTrace + Resolve as separate launches (occlusion flags in memory)
    sc = scenes.bunny_spheres(stride=8)
    o = api.Options(200, 112, antialias=api.Antialias(api.akGrid, 2), depthMode=api.NRT_DEPTH_INTENDED, maxRayDepth=3)
    monkeypatch.setenv("NRT_FUSE_RESOLVE", "0")
    assert_parity(sc, o, oracle_mod)
    assert_parity(scenes.spheres_reflection(), api.Options(160, 120), oracle_mod)
    monkeypatch.delenv("NRT_FUSE_RESOLVE")
    assert_parity(scenes.with_many_lights(scenes.bunny_spheres(stride=32), 32), api.Options(96, 54), oracle_mod)
    assert_parity(scenes.with_many_lights(scenes.bunny_spheres(stride=32), 33), api.Options(96, 54), oracle_mod)


def test_many_mesh_objects_and_lights(oracle_mod):
    # 12 mesh objects x 32 lights: the fused shadow-gate kernel would need 52 KiB of shared-memory counters (> the 48 KiB a
    # launch gets without opting in): the separate flags kernel takes over; 33 lights: separate ShadowTrace + Resolve
    for nl in (32, 33):
        assert_parity(scenes.many_meshes_many_lights(12, nl), api.Options(96, 54), oracle_mod)
    assert_parity(scenes.many_meshes_many_lights(3, 5), api.Options(120, 68, antialias=api.Antialias(api.akGrid, 2)), oracle_mod)


def test_stress_scene_sphere_clusters(oracle_mod):
    # BASELINE config 5's object count class: thousands of spheres -> clustered object scan
    sc = scenes.stress(ntri=20000, nspheres=3000)
    assert_parity(sc, api.Options(240, 135, antialias=api.Antialias(api.akGrid, 2), depthMode=api.NRT_DEPTH_INTENDED, maxRayDepth=4), oracle_mod)


def test_config4_full_size_properties(monkeypatch):
    # BASELINE config 4 at its full size (3840x2160, 16 spp, depth 8): size-independent properties —
    # Stats identities of renderer.nim:58,138,155 and invariance under the chunking of the frame
    sc = scenes.bunny_spheres()
    o = api.Options(3840, 2160, antialias=api.Antialias(api.akGrid, 4), depthMode=api.NRT_DEPTH_INTENDED, maxRayDepth=8)
    ds = api.DeviceScene(sc)
    fb1 = api.newFramebuf(o.width, o.height)
    st1 = api.renderFrame(ds, o, fb1)
    assert st1.numPrimaryRays == 3840 * 2160 * 16
    assert st1.numIntersectionTests == st1.numRays * len(sc.objects)
    assert st1.numCappedSamples == 0 and np.isfinite(fb1.data).all()
    monkeypatch.setenv("NRT_CHUNK_SAMPLES", str(5_000_000))
    fb2 = api.newFramebuf(o.width, o.height)
    st2 = api.renderFrame(ds, o, fb2)
    assert st1 == st2 and digest(fb1.data) == digest(fb2.data)
    # the 1/4-size frame of the same scene is checked against the oracle in test_golden_config3_reduced


def test_error_behaviour():
    import ctypes as C
    L = api.lib()
    ds = api.DeviceScene(scenes.boxtest())
    o = api.Options(16, 16).to_c()
    fb = api.newFramebuf(16, 16)
    p = fb.data.ctypes.data_as(C.c_void_p)
    assert L.nrt_render(ds.handle, C.byref(o), 0, 16, 3, 4, p, None, None) == -6      # NRT_ERR_UNSUPPORTED
    assert L.nrt_render(ds.handle, C.byref(o), 0, 16, 4, 2, p, None, None) == -6      # maxStep < step
    assert L.nrt_render(None, C.byref(o), 0, 16, 1, 1, p, None, None) == -1           # NRT_ERR_INVALID
    assert L.nrt_render(ds.handle, C.byref(o), 0, 16, 1, 1, None, None, None) == -1
    assert b"null" in L.nrt_last_error()
    assert L.nrt_render(ds.handle, C.byref(o), 5, 5, 1, 1, p, None, None) == 0        # empty line range: no-op
    bad = api.Options(16, 16, antialias=api.Antialias(api.akGrid, 0)).to_c()
    assert L.nrt_render(ds.handle, C.byref(bad), 0, 16, 1, 1, p, None, None) == -1


def test_empty_scene_and_no_lights(oracle_mod):
    from nim_raytracer_b200 import linalg as L
    sc = api.Scene([], [], 50.0, L.mat4(1.0), api.vec3(0.1, 0.2, 0.3))
    fb, st, aov = assert_parity(sc, api.Options(32, 16), oracle_mod)
    assert (aov.obj_id == -1).all() and np.allclose(fb.image()[0, 0], [0.1, 0.2, 0.3])
    sc = scenes.bunny(stride=32)
    sc.lights = []
    assert_parity(sc, api.Options(64, 36), oracle_mod)


def test_scene_update_and_srgb_output(oracle_mod):
    sc = scenes.bunny(stride=16)
    ds = api.DeviceScene(sc)
    o = api.Options(160, 90)
    fb = api.newFramebuf(160, 90)
    api.renderFrame(ds, o, fb)
    sc.objects[0].material.albedo = api.vec3(0.9, 0.2, 0.1)
    sc.cameraToWorld = scenes._camera(0.5, 5.0, 2.0)
    ds.update(sc)
    api.renderFrame(ds, o, fb)
    rfb, _, _ = oracle_mod.render(sc, o)
    assert (fb.data == rfb.data).all()
    # output stage (utils/framebuf.nim:74-78, utils/color.nim:17-22): equal to the oracle's samples of its own frame
    img = api.framebufToSrgb8(fb)
    assert (img.reshape(-1) == oracle_mod.outvalues(rfb.data, 8, True)).all()


def test_scene_update_identical_changed_and_rejected(oracle_mod):
    # nrt_scene_update: (1) a bit-identical description is copied to the device and compared there - the records are
    # kept; (2) one changed vertex is detected and everything is rebuilt; (3) a description whose shape differs
    # (object kind, mesh index, light kind) is rejected BEFORE anything is touched: the scene still renders
    sc = scenes.bunny_spheres(stride=16)
    o = api.Options(160, 90)
    ds = api.DeviceScene(sc)
    fb = api.newFramebuf(160, 90)
    api.renderFrame(ds, o, fb)
    ref, rst, _ = oracle_mod.render(sc, o)
    assert (fb.data == ref.data).all()
    ds.update()                                                  # (1)
    fb1 = api.newFramebuf(160, 90)
    st1 = api.renderFrame(ds, o, fb1)
    assert (fb1.data == ref.data).all() and st1 == rst
    sc.objects[0].geometry.vertices[7, 1] += 0.25                # (2) the mesh arrays are shared with the description
    sc.objects[0].geometry.vertices[8, 0] -= 0.125
    ds.update(sc)
    fb2 = api.newFramebuf(160, 90)
    st2 = api.renderFrame(ds, o, fb2)
    ref2, rst2, _ = oracle_mod.render(sc, o)
    assert (fb2.data == ref2.data).all() and st2 == rst2
    bad = scenes.bunny_spheres(stride=16)                        # (3)
    bad.objects[0].geometry.vertices[:] = sc.objects[0].geometry.vertices
    bad.objects[2], bad.objects[0] = bad.objects[0], bad.objects[2]   # a sphere where the mesh was and vice versa
    with pytest.raises(api.NrtError):
        ds.update(bad)
    bad2 = scenes.bunny_spheres(stride=16)
    bad2.lights[1] = api.PointLight(color=api.vec3(1.0), intensity=100.0, pos=api.point(0.0, 5.0, -5.0))
    with pytest.raises(api.NrtError):
        ds.update(bad2)
    fb3 = api.newFramebuf(160, 90)
    st3 = api.renderFrame(ds, o, fb3)
    assert (fb3.data == ref2.data).all() and st3 == rst2


def test_page_locked_caller_memory(oracle_mod):
    # nrt_host_register: the caller's framebuffer and mesh arrays page-locked in place (what bench.py's e2e
    # leg and a shared multi-process host framebuffer do); results and error behaviour unchanged
    import ctypes as C
    L = api.lib()
    sc, o = scenes.bunny(stride=16), api.Options(200, 120)
    pinned = api.pinSceneArrays(sc)
    assert 1 <= len(pinned) <= 4
    ds = api.DeviceScene(sc)
    fb = api.newFramebuf(o.width, o.height)
    p = fb.data.ctypes.data_as(C.c_void_p)
    assert L.nrt_host_register(p, fb.data.nbytes) == 0
    assert L.nrt_host_register(p, fb.data.nbytes) != 0          # already registered: an error code, not a crash
    assert b"cudaHostRegister" in L.nrt_last_error()
    try:
        st = api.renderFrame(ds, o, fb)
        ds.update()                                               # uploads from the page-locked arrays
        st2 = api.renderFrame(ds, o, fb)
    finally:
        assert L.nrt_host_unregister(p) == 0
        api.unpinSceneArrays(pinned)
    rfb, rst, _ = oracle_mod.render(sc, o)
    assert (fb.data == rfb.data).all() and st == rst and st2 == rst
    assert L.nrt_host_register(None, 16) != 0 and L.nrt_host_unregister(None) == 0


def test_output_stage_16_bit_and_rgba(tmp_path, oracle_mod):
    # utils/framebuf.nim:55-93 (any maxval, big-endian 16-bit samples) and utils/image.nim:45-54 on the GPU: byte work,
    # equal to the oracle's restatement of outvalue (tests/test_output_stage.py covers the cut points exhaustively)
    rs = np.random.RandomState(11)
    fb = api.newFramebuf(37, 23)
    fb.data[:] = rs.uniform(-0.2, 1.3, fb.data.shape).astype(np.float32)
    for bits, srgb in ((16, True), (16, False), (10, True), (5, False), (8, True)):
        img = api.framebufQuantize(fb, bits, srgb)
        assert img.shape == (23, 37, 3) and (img.reshape(-1) == oracle_mod.outvalues(fb.data, bits, srgb)).all()
    assert (api.framebufQuantize(fb, 8, True) == api.framebufToSrgb8(fb)).all()
    rgba = api.framebufToRgba8(fb, alpha=0x7F)
    assert (rgba.reshape(-1) == oracle_mod.rgba8(fb.data, 0x7F)).all()
    path, opath = tmp_path / "out16.ppm", tmp_path / "oracle16.ppm"
    assert api.writePpm(fb, str(path), bits=16) and oracle_mod.write_ppm(fb, str(opath), bits=16)
    raw = path.read_bytes()
    assert raw.startswith(b"P6 37 23 65535 ") and len(raw) == len(b"P6 37 23 65535 ") + 37 * 23 * 6
    assert raw == opath.read_bytes()


def test_two_gpus_in_process_match_one(oracle_mod):
    import ctypes as C
    api.shutdown()
    api.initRenderer(0)                                   # all visible devices
    try:
        n = api.lib().nrt_device_count()
        if n < 2:
            pytest.skip("single GPU box")
        assert_parity(scenes.bunny(stride=8), api.Options(256, 144), oracle_mod)
    finally:
        api.shutdown()
        api.initRenderer(1)


@pytest.mark.gpu
def test_short_wavefront_list_goes_to_path_tail(oracle_mod, monkeypatch):
    # library default (conftest pins it off for the other tests): wavefront lists below 16384 samples are finished by
    # one PathTail launch, from bounce 0 as well
    monkeypatch.delenv("NRT_HARD_TAIL_BELOW")
    sc = scenes.bunny_spheres(stride=4)
    o = api.Options(320, 180, antialias=api.Antialias(api.akGrid, 2), depthMode=api.NRT_DEPTH_INTENDED, maxRayDepth=8)
    assert_parity(sc, o, oracle_mod)
    assert_parity(sc, api.Options(200, 112), oracle_mod)
    assert_parity(scenes.bunny(flip_winding=False, stride=2), api.Options(240, 135), oracle_mod)
    monkeypatch.setenv("NRT_HARD_TAIL_BELOW", "2000")   # bounce 0 through the wavefront, the later lists through PathTail
    assert_parity(sc, o, oracle_mod)


@pytest.mark.gpu
def test_lane_plan_from_band_feedback_keeps_the_frame(oracle_mod, monkeypatch):
    # The second frame of a scene deals its bands out by the first frame's per-band wavefront counts (heavy bands to
    # the high-priority lanes, mesh-free bands to the low-priority ones, nrt.cu: BandFeedback): scheduling only — the
    # frame, the ids and the Stats stay the oracle's, frame after frame, whatever the number of lanes.
    monkeypatch.delenv("NRT_HARD_TAIL_BELOW")
    sc = scenes.bunny_spheres(stride=4)
    o = api.Options(640, 512, antialias=api.Antialias(api.akGrid, 4), depthMode=api.NRT_DEPTH_INTENDED, maxRayDepth=8)
    rfb, rst, raov = oracle_mod.render(sc, o, aov=api.Aov(o.width, o.height))
    monkeypatch.setenv("NRT_LANE_FEEDBACK", "1")
    for lanes, heavy in (("2", "1"), ("4", "2"), ("4", "1"), ("3", "2")):
        monkeypatch.setenv("NRT_LANES", lanes)
        monkeypatch.setenv("NRT_HEAVY_LANES", heavy)
        ds = api.DeviceScene(sc)
        for frame in range(3):
            fb, aov = api.newFramebuf(o.width, o.height), api.Aov(o.width, o.height)
            st = api.renderFrame(ds, o, fb, aov=aov)
            assert st == rst, (lanes, heavy, frame)
            assert (aov.obj_id == raov.obj_id).all() and (aov.tri_id == raov.tri_id).all()
            assert (fb.data == rfb.data).all(), (lanes, heavy, frame)
        ds.close()


@pytest.mark.gpu
def test_fork_helper_pipeline_keeps_the_frame(oracle_mod, monkeypatch):
    # the fork (nrt_renderer.h): the continuing pool of bounce 0 runs in a helper pipeline (own buffers, stream and
    # host thread) next to the wavefront chain of the samples with mesh rays; off (NRT_FORK_MIN=0) and on, same frame
    sc = scenes.bunny_spheres(stride=4)
    o = api.Options(320, 180, antialias=api.Antialias(api.akGrid, 2), depthMode=api.NRT_DEPTH_INTENDED, maxRayDepth=8)
    for fork_min in ("0", "64"):
        monkeypatch.setenv("NRT_FORK_MIN", fork_min)
        assert_parity(sc, o, oracle_mod)
        assert_parity(sc, api.Options(200, 112), oracle_mod)
        assert_parity(scenes.transformed_objects(), api.Options(240, 135, antialias=api.Antialias(api.akGrid, 2)), oracle_mod)


@pytest.mark.gpu
def test_light_space_shadow_grid_and_ray_bundles(oracle_mod):
    # >= 64 spheres: warp-bundle first look over the sphere clusters (RayBundle) for path rays, light-space grid
    # (ShadowGridF) for the DistantLights' shadow rays, cluster traversal for the PointLight's
    from test_emu_parity import _grid_scene
    o = api.Options(320, 180, antialias=api.Antialias(api.akGrid, 2), depthMode=api.NRT_DEPTH_INTENDED, maxRayDepth=4)
    assert_parity(_grid_scene(), o, oracle_mod)
    assert_parity(scenes.stress(ntri=100, nspheres=2000, seed=11), api.Options(240, 135), oracle_mod)
    assert_parity(scenes.stress(ntri=500, nspheres=40, seed=5), api.Options(240, 135, antialias=api.Antialias(api.akGrid, 2)), oracle_mod)   # flat scan


@pytest.mark.gpu
def test_partitions_render_the_rows_the_python_mirror_says(oracle_mod):
    # nrt_set_partition(k, n): the bands of partition k (nrt_unit_owner: serpentine deal) are rendered, every other row
    # is left untouched; distributed.owned_rows names the same rows and the n pieces assemble to the oracle's frame
    from nim_raytracer_b200 import distributed as D
    sc = scenes.bunny_spheres(stride=8)
    o = api.Options(256, 200, antialias=api.Antialias(api.akGrid, 4), depthMode=api.NRT_DEPTH_INTENDED, maxRayDepth=4)
    rfb, rst, _ = oracle_mod.render(sc, o)
    band = api.bandRows(o)
    ds = api.DeviceScene(sc)
    try:
        for n in (2, 3, 8):
            out = np.zeros_like(rfb.image())
            for k in range(n):
                api.setPartition(k, n)
                fb = api.newFramebuf(o.width, o.height)
                fb.data[:] = -1.0
                api.renderFrame(ds, o, fb)
                img = fb.image()
                rows = D.owned_rows(k, n, o.height, band)
                touched = np.where((img != -1.0).any(axis=(1, 2)))[0]
                assert touched.tolist() == rows.tolist(), (n, k)
                D.merge_rows(out, img, k, n, band=band)
            assert (out == rfb.image()).all(), n
    finally:
        api.setPartition(0, 1)
        ds.close()
