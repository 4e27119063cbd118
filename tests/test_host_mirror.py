"""The C++ host mirror (nim_raytracer_b200/host/nrt_host.hpp) of the reference's renderer interface:
compiles everywhere; on a GPU it reproduces the reference front-end's flow (src/raytracer.nim:42-124)
and must match the oracle bit for bit."""
import os
import subprocess

import numpy as np
import pytest

from nim_raytracer_b200 import api, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp")


def build(tmp):
    import __graft_entry__ as g
    g.build_cuda()
    exe = os.path.join(tmp, "host_mirror_test")
    libdir = os.path.dirname(api.LIB_PATH)
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", SRC, "-o", exe, "-L" + libdir, "-lnrt",
                    "-Wl,-rpath," + libdir], check=True)
    return exe


def test_host_mirror_compiles_and_refuses_without_gpu(tmp_path):
    exe = build(str(tmp_path))
    r = subprocess.run([exe, "--expect-no-device"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_host_mirror_matches_oracle(tmp_path, oracle_mod):
    exe = build(str(tmp_path))
    raw, ppm = str(tmp_path / "render.raw"), str(tmp_path / "render.ppm")
    r = subprocess.run([exe, scenes.BUNNY_GEOM, raw, ppm], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "meshperftest: t = 5" in r.stdout
    fb = np.fromfile(raw, dtype=np.float32)
    sc, o = scenes.bunny(), api.Options(300, 200)      # the same scene built by the Python fixtures
    rfb, rst, _ = oracle_mod.render(sc, o)
    assert (fb == rfb.data).all()
    assert f"numPrimaryRays {rst.numPrimaryRays} numIntersectionTests {rst.numIntersectionTests} numIntersectionHits {rst.numIntersectionHits}" in r.stdout
    head = open(ppm, "rb").read(16)
    assert head.startswith(b"P6 300 200 255 ")
