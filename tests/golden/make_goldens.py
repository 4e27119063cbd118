"""Generates the committed golden fixtures from the CPU oracle (oracle/ref_cpu.cpp).

Run here (no GPU needed):  python tests/golden/make_goldens.py
The oracle itself is pinned to the reference by tests/test_oracle_golden.py; these
fixtures let the -m gpu tests check full-size frames (BASELINE configs) without
re-running the brute-force CPU renderer on the GPU box.

Each fixture stores sha256 digests of the whole float32 framebuffer / id AOVs, the
Stats, and every 8th row verbatim (ids + float32 RGB) so a mismatch can be located.
"""
import hashlib
import os
import zlib
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from nim_raytracer_b200 import api, scenes  # noqa: E402


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def make(name, scene, opts, row_stride=8):
    t = time.time()
    aov = api.Aov(opts.width, opts.height)
    fb, st, _ = oracle.render(scene, opts, aov=aov)
    h, w = opts.height, opts.width
    rows = np.arange(0, h, row_stride)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        width=w, height=h,
        fb_sha256=digest(fb.data), obj_sha256=digest(aov.obj_id), tri_sha256=digest(aov.tri_id),
        t_sha256=digest(aov.t_hit),
        stats=np.array([st.numPrimaryRays, st.numIntersectionTests, st.numIntersectionHits, st.numRays,
                        st.numCappedSamples], dtype=np.int64),
        rows=rows,
        fb_rows=fb.image()[rows],
        obj_rows=aov.obj_id.reshape(h, w)[rows].astype(np.int8 if len(scene.objects) < 128 else np.int32),
        tri_rows=aov.tri_id.reshape(h, w)[rows],
        # one CRC32 per scanline of the float32 framebuffer / the ids: locates a mismatch in a full-size frame
        fb_row_crc=np.array([zlib.crc32(r.tobytes()) for r in fb.image()], dtype=np.uint32),
        id_row_crc=np.array([zlib.crc32(a.tobytes() + b.tobytes()) for a, b in
                             zip(aov.obj_id.reshape(h, w), aov.tri_id.reshape(h, w))], dtype=np.uint32),
    )
    print(f"{name}: {time.time() - t:.1f}s  {st}")


if __name__ == "__main__":
    which = sys.argv[1:] or ["config1", "config2", "config3_small"]
    if "config1" in which:   # BASELINE config 1: spheres-reflection 640x480, akNone
        make("config1_spheres_640x480", scenes.spheres_reflection(), api.Options(640, 480))
    if "config2" in which:   # BASELINE config 2: canonical bunny 1920x1080, akNone
        make("config2_bunny_1920x1080", scenes.bunny(), api.Options(1920, 1080), row_stride=8)
    if "config3_small" in which:  # config 3 scene, reduced size: 480x270, 4 spp grid, intended depth 8
        make("config3_bunny_spheres_480x270_g2", scenes.bunny_spheres(),
             api.Options(480, 270, antialias=api.Antialias(api.akGrid, 2), depthMode=api.NRT_DEPTH_INTENDED,
                         maxRayDepth=8), row_stride=4)
    if "config3" in which:   # BASELINE config 3 at its stated size: 1920x1080, 16 spp grid, intended depth 8 (~11 min on 8 cores)
        make("config3_bunny_spheres_1920x1080_g4", scenes.bunny_spheres(),
             api.Options(1920, 1080, antialias=api.Antialias(api.akGrid, 4), depthMode=api.NRT_DEPTH_INTENDED,
                         maxRayDepth=8), row_stride=16)
    if "config4" in which:   # BASELINE config 4 (the bench workload) at its stated size: 3840x2160, 16 spp (~45 min on 8 cores)
        make("config4_bunny_spheres_3840x2160_g4", scenes.bunny_spheres(),
             api.Options(3840, 2160, antialias=api.Antialias(api.akGrid, 4), depthMode=api.NRT_DEPTH_INTENDED,
                         maxRayDepth=8), row_stride=48)
    if "reference_scenes" in which:
        # every scene file of the reference (src/data/scenes/*.nim, restated in nim_raytracer_b200/scenes.py) with the
        # options of its default front-end (src/raytracer.nim:43-54: 300x200, akNone, bias 1e-8, maxRayDepth 5);
        # "mesh-bunny" is the scene that front-end includes: the teapot (src/data/meshes/teapot.obj)
        for name, build in scenes.REFERENCE_SCENES.items():
            make("ref_" + name.replace("-", "_") + "_300x200", build(), api.Options(300, 200, bias=0.00000001, maxRayDepth=5), row_stride=25)
    if "config5_class" in which:
        # BASELINE config 5's class at a size the brute-force oracle finishes in minutes: 100,000 random triangles as ONE
        # mesh + 10,000 spheres (10 % mirrors) + plane, 200x112, 4 spp grid, intended depth 4 (scenes.stress, the frozen
        # generator of BASELINE.md section 3)
        make("config5_class_100k_tris_10k_spheres_200x112_g2", scenes.stress(ntri=100_000, nspheres=10_000),
             api.Options(200, 112, antialias=api.Antialias(api.akGrid, 2), depthMode=api.NRT_DEPTH_INTENDED, maxRayDepth=4), row_stride=8)
