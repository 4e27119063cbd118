"""Loaders: loadObj (src/loaders/obj.nim:8-126), the .geom writer of
src/loaders/objconv.nim:139-153 and the finished .geom reader (geomloader.nim:30-49 is a stub in the reference),
in Python (nim_raytracer_b200/loaders.py) and in the C++ host mirror (host/nrt_host.hpp).

The reference holds one byte-exact vector for this path: test/bunny.geom is objconv's output for
src/data/meshes/bunny.obj.  tests/golden/bunny.obj.xz is that .obj (xz-compressed input data)."""
import lzma
import os
import subprocess

import numpy as np

from nim_raytracer_b200 import loaders, scenes

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _bunny_obj(tmp_path):
    p = str(tmp_path / "bunny.obj")
    with lzma.open(os.path.join(HERE, "golden", "bunny.obj.xz")) as f, open(p, "wb") as out:
        out.write(f.read())
    return p


def test_obj_to_geom_equals_the_reference_vector(tmp_path):
    # writeGeom(loadObj(bunny.obj)) == test/bunny.geom, byte for byte (objconv.nim:139-153)
    mesh = loaders.loadObj(_bunny_obj(tmp_path))
    assert mesh.vertices.shape == (35947, 4) and mesh.vertexIdx.shape == (69451, 3)
    out = str(tmp_path / "bunny.geom")
    loaders.writeGeom(out, mesh)
    assert open(out, "rb").read() == open(scenes.BUNNY_GEOM, "rb").read()


def test_geom_round_trip_and_flat_normals(tmp_path):
    mesh = loaders.loadObj(scenes.TEAPOT_OBJ)
    assert mesh.vertexIdx.shape == (6320, 3)
    # one flat normal per face (obj.nim:65-84): normalIdx = [k, k, k], unit length, orthogonal to both edges
    k = np.arange(6320)
    assert (mesh.normalIdx == np.stack([k, k, k], axis=1)).all()
    p = mesh.vertices[mesh.vertexIdx.reshape(-1), :3].reshape(-1, 3, 3)
    n = mesh.normals[:, :3]
    ok = np.isfinite(n).all(axis=1)                     # (degenerate faces normalise to NaN, as in the reference)
    assert ok.mean() > 0.99
    assert np.abs(np.linalg.norm(n[ok], axis=1) - 1.0).max() < 1e-12
    assert np.abs(np.einsum("ij,ij->i", n[ok], (p[:, 1] - p[:, 0])[ok])).max() < 1e-9
    out = str(tmp_path / "teapot.geom")
    loaders.writeGeom(out, mesh)
    tri = loaders.readGeom(out)
    assert tri.dtype == np.float32 and (tri == p.astype(np.float32)).all()
    back = loaders.loadGeom(out)                        # index-free mesh: vertex i of triangle k at 3k + i
    assert (back.vertices[:, :3] == tri.reshape(-1, 3).astype(np.float64)).all()
    assert (back.vertexIdx == np.arange(6320 * 3).reshape(-1, 3)).all()
    # truncated file: an error, not a short mesh
    open(str(tmp_path / "short.geom"), "wb").write(open(out, "rb").read()[:-40])
    try:
        loaders.readGeom(str(tmp_path / "short.geom"))
        raise AssertionError("truncated .geom accepted")
    except ValueError:
        pass


def test_cpp_host_mirror_loaders_equal_python(tmp_path):
    # nrt_host.hpp: loadObj / writeGeom / loadGeom produce the same arrays as loaders.py, bit for bit
    exe = str(tmp_path / "loader_tool")
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", os.path.join(HERE, "cpp", "loader_tool.cpp"), "-o", exe,
                    "-I" + os.path.join(ROOT, "include")], check=True)
    geom, dump = str(tmp_path / "t.geom"), str(tmp_path / "t.bin")
    subprocess.run([exe, scenes.TEAPOT_OBJ, geom, dump], check=True)
    mesh = loaders.loadObj(scenes.TEAPOT_OBJ)
    ref = str(tmp_path / "ref.geom")
    loaders.writeGeom(ref, mesh)
    assert open(geom, "rb").read() == open(ref, "rb").read()
    raw = np.fromfile(dump, dtype=np.uint8)
    nv, nf = raw[:16].view("<i8")
    off = 16
    v = raw[off:off + nv * 32].view("<f8").reshape(nv, 4); off += nv * 32
    n = raw[off:off + nf * 32].view("<f8").reshape(nf, 4); off += nf * 32
    vi = raw[off:off + nf * 24].view("<i8").reshape(nf, 3)
    assert (v == mesh.vertices).all() and (vi == mesh.vertexIdx).all()
    assert (n.view(np.uint64) == mesh.normals.view(np.uint64)).all()   # bit patterns: NaNs of degenerate faces included
    # and the C++ .geom reader on the reference's own file
    geom2, dump2 = str(tmp_path / "b.geom"), str(tmp_path / "b.bin")
    subprocess.run([exe, "--geom", scenes.BUNNY_GEOM, geom2, dump2], check=True)
    assert open(geom2, "rb").read() == open(scenes.BUNNY_GEOM, "rb").read()
    b = loaders.loadGeom(scenes.BUNNY_GEOM)
    raw = np.fromfile(dump2, dtype=np.uint8)
    nv, nf = raw[:16].view("<i8")
    assert nv == b.vertices.shape[0] and nf == 69451
    assert (raw[16:16 + nv * 32].view("<f8").reshape(nv, 4) == b.vertices).all()


def test_malformed_obj_tokens_follow_the_reference(tmp_path):
    # obj.nim:25-63: `except ValueError: discard` leaves a coordinate at 0.0 and a vertex index at 0; comments, other
    # records and empty lines are skipped; only the first three tokens of an `f` record are read (obj.nim:116-118).
    src = tmp_path / "odd.obj"
    src.write_text("# comment\n\nv 1 2 3\nv 4.5 oops 6\nv -1e0 0.25 7 1.0\nvn 0 1 0\nf 1 2 3\nf 3 x 1 2\nf 1/9/9 2//7 3\n")
    m = loaders.loadObj(str(src))
    assert m.vertices.tolist() == [[1, 2, 3, 1], [4.5, 0.0, 6, 1], [-1.0, 0.25, 7, 1]]
    assert m.vertexIdx.tolist() == [[0, 1, 2], [2, 0, 0], [0, 1, 2]]
    # the C++ host mirror reads the same arrays
    exe = str(tmp_path / "loader_tool")
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", os.path.join(HERE, "cpp", "loader_tool.cpp"), "-o", exe,
                    "-I" + os.path.join(ROOT, "include")], check=True)
    geom, dump = str(tmp_path / "o.geom"), str(tmp_path / "o.bin")
    subprocess.run([exe, str(src), geom, dump], check=True)
    raw = np.fromfile(dump, dtype=np.uint8)
    nv, nf = raw[:16].view("<i8")
    assert (nv, nf) == (3, 3)
    v = raw[16:16 + nv * 32].view("<f8").reshape(nv, 4)
    vi = raw[16 + nv * 32 + nf * 32:16 + nv * 32 + nf * 32 + nf * 24].view("<i8").reshape(nf, 3)
    assert (v == m.vertices).all() and (vi == m.vertexIdx).all()
