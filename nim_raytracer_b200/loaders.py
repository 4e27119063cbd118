"""Mesh loaders producing the TriangleMesh data the render path consumes.

  loadObj    — src/loaders/obj.nim:8-126 (v / f records only, 1-based indices,
               triangles, ONE flat normal per face: obj.nim:65-84)
  loadGeom   — reader for the `.geom` format written by src/loaders/objconv.nim:125-153
               (int32 triangle count, then 9 float32 per triangle).  The reference's
               own reader (src/loaders/geomloader.nim:30-49) is unfinished; this is
               the finished equivalent.
  writeGeom  — src/loaders/objconv.nim:139-153
"""
from __future__ import annotations

import numpy as np

from . import linalg
from .api import Geometry, initTriangleMesh


def calcNormals(vertices: np.ndarray, vertexIdx: np.ndarray):
    """obj.nim:65-84: n = normalize(cross(p1-p0, p2-p0)); normalIdx = [k,k,k]."""
    p0 = vertices[vertexIdx[:, 0], :3]
    p1 = vertices[vertexIdx[:, 1], :3]
    p2 = vertices[vertexIdx[:, 2], :3]
    a, b = p1 - p0, p2 - p0
    # glm cross: (a.y*b.z - b.y*a.z, a.z*b.x - b.z*a.x, a.x*b.y - b.x*a.y)
    n = np.stack([a[:, 1] * b[:, 2] - b[:, 1] * a[:, 2],
                  a[:, 2] * b[:, 0] - b[:, 2] * a[:, 0],
                  a[:, 0] * b[:, 1] - b[:, 0] * a[:, 1]], axis=1)
    d = (n[:, 0] * n[:, 0] + n[:, 1] * n[:, 1]) + n[:, 2] * n[:, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        n = n * (1.0 / np.sqrt(d))[:, None]
    normals = np.zeros((vertexIdx.shape[0], 4), dtype=np.float64)
    normals[:, :3] = n
    k = np.arange(vertexIdx.shape[0], dtype=np.int64)
    return normals, np.stack([k, k, k], axis=1)


def _objFloat(tok: str) -> float:
    """obj.nim:25-43: a coordinate that does not parse is left at its default 0.0 (`except ValueError: discard`)."""
    try:
        return float(tok)
    except ValueError:
        return 0.0


def _objIndex(tok: str) -> int:
    """obj.nim:46-63: parseInt(s) - 1; a token that does not parse leaves the index at its default 0.
    Deliberate superset: the vertex index of a `v/vt/vn` token is read (the reference's parseInt fails on the
    slash and yields index 0 — a degenerate face; it ships no such file)."""
    try:
        return int(tok.split("/")[0]) - 1
    except ValueError:
        return 0


def loadObj(fname: str, objectToWorld=None) -> Geometry:
    verts, faces = [], []
    with open(fname) as f:
        for line in f:
            c = line.split()
            if not c:
                continue
            if c[0] == "v":
                verts.append([_objFloat(c[1]), _objFloat(c[2]), _objFloat(c[3]), 1.0])
            elif c[0] == "f":
                faces.append([_objIndex(c[1]), _objIndex(c[2]), _objIndex(c[3])])
    vertices = np.asarray(verts, dtype=np.float64).reshape(-1, 4)
    vertexIdx = np.asarray(faces, dtype=np.int64).reshape(-1, 3)
    normals, normalIdx = calcNormals(vertices, vertexIdx)
    return initTriangleMesh(vertices, normals, vertexIdx, normalIdx,
                            linalg.mat4(1.0) if objectToWorld is None else objectToWorld)


def readGeom(fname: str) -> np.ndarray:
    """Raw triangle soup: (ntri, 3, 3) float32."""
    raw = np.fromfile(fname, dtype=np.uint8)
    n = int(raw[:4].view("<i4")[0])
    tri = raw[4:4 + n * 36].view("<f4").reshape(n, 3, 3)
    if tri.shape[0] != n:
        raise ValueError(f"{fname}: truncated .geom ({tri.shape[0]} of {n} triangles)")
    return tri


def trianglesToMesh(tri: np.ndarray, objectToWorld=None) -> Geometry:
    """Index-free mesh: vertex i of triangle k at 3k+i, one flat normal per face."""
    tri = np.asarray(tri, dtype=np.float64)
    n = tri.shape[0]
    vertices = np.ones((n * 3, 4), dtype=np.float64)
    vertices[:, :3] = tri.reshape(n * 3, 3)
    vertexIdx = np.arange(n * 3, dtype=np.int64).reshape(n, 3)
    normals, normalIdx = calcNormals(vertices, vertexIdx)
    return initTriangleMesh(vertices, normals, vertexIdx, normalIdx,
                            linalg.mat4(1.0) if objectToWorld is None else objectToWorld)


def loadGeom(fname: str, objectToWorld=None) -> Geometry:
    return trianglesToMesh(readGeom(fname), objectToWorld)


def writeGeom(fname: str, mesh: Geometry) -> None:
    tri = mesh.vertices[mesh.vertexIdx.reshape(-1), :3].astype("<f4")
    with open(fname, "wb") as f:
        f.write(np.int32(mesh.vertexIdx.shape[0]).astype("<i4").tobytes())
        f.write(tri.tobytes())
