// nrt_filter.h — float32 ray x triangle FILTER of the mesh path (the hot loop's arithmetic).
//
// TriangleMesh.intersect (geom.nim:339-358) tests every ray that passes the AABB gate
// against ALL faces with rayTriangleIntersectFast (geom.nim:283-336, float64).  Here
// a float32 evaluation with a proven error margin decides "possibly hit"; every such
// (ray, triangle) pair is then re-evaluated in float64 in the reference's exact
// operation order (nrt_core.h: rayTriangleExact), so results equal the reference's.
//
// Möller–Trumbore is rewritten with scalar triple products so that each test is a
// handful of dot products between per-triangle and per-ray vectors:
//   det = e1.(d x e2),  u' = (o - v0).(d x e2),  v' = d.((o - v0) x e1)   (u' = u det, v' = v det)
// The `t` tests of the reference (t >= 0, t < tMin) are left to the float64 pass:
// the mesh is only tested when the ray origin is outside its AABB with the box in
// front (geom.nim:340), so every line/triangle crossing already has t >= 0.
//
// Three formulations ("modes"), chosen per ray bundle:
//   GENERAL  arbitrary rays.  Frame centred at the mesh AABB centre C, ray = (d, m = (o-C) x d):
//              det = N.d,  u' = E2.m + A.d,  v' = E1n.m + B.d
//              N = -(e1 x e2), A = (v0-C) x e2, B = e1 x (v0-C), E1n = -e1     15 FFMA + 2 FADD
//   ORIGIN   rays sharing one origin O (primary rays).  Frame centred at O => m = 0:
//              det = N.d,  u' = A.d,  v' = B.d,  A = (v0-O) x e2, B = e1 x (v0-O)   9 FFMA + 2 FADD
//            Triangles whose plane faces away from O (t' = (O-v0).(e1 x e2) < 0) can never be
//            hit from O with det >= 1e-6 and t >= 0: they are dropped when the records are built.
//   DIR      rays sharing one direction D (shadow rays of a DistantLight).  det is a per-triangle
//            constant, so triangles with det < 1e-6 (geom.nim:306) are dropped at build time and
//            the barycentrics are pre-divided:  u = P.o' + pu0,  v = Q.o' + qv0,  w = 1 - u - v
//              P = (D x e2)/det, Q = (e1 x D)/det, pu0 = -(v0-C).P, qv0 = -(v0-C).Q,
//              o' = (o-C) minus its component along D (P.D = Q.D = 0)             6 FFMA + 2 FADD
//
// Records are stored pair-interleaved (two triangles per 64-bit lane pair) so the CUDA
// kernel evaluates two triangles per FFMA2 (fma.rn.f32x2, sm_100):
//   float index of coefficient k of record r:  ((r >> 1) * NC + k) * 2 + (r & 1)
//
// Error margins (u = 2^-24, round to nearest, no overflow; analysis in DESIGN.md §4):
// with S a per-triangle magnitude bound and Rr a per-ray one, Eb = S * Rr bounds the
// float32 error of u', v' (and det / 16), and the candidate test is
//   u' + Eb >= 0,  v' + Eb >= 0,  (det + 16 Eb) - (u' + Eb) - (v' + Eb) >= 0
// (DIR: det := 1).  Only sign bits are inspected: candidate <=> sign(u|v|w) clear.
#pragma once

#include "nrt_core.h"

namespace nrt {

enum FilterMode { FM_GENERAL = 0, FM_ORIGIN = 1, FM_DIR = 2 };

// floats per record / index of S / index of the original face id (bit pattern) in a record
NRT_HD constexpr int recFloats(int mode) { return mode == FM_GENERAL ? 16 : (mode == FM_ORIGIN ? 12 : 10); }
NRT_HD constexpr int recSlotS(int mode) { return mode == FM_GENERAL ? 3 : (mode == FM_ORIGIN ? 3 : 8); }
NRT_HD constexpr int recSlotId(int mode) { return mode == FM_GENERAL ? -1 : (mode == FM_ORIGIN ? 7 : 9); }
// floats per queued ray (always stored as float4 planes): GENERAL 8, ORIGIN/DIR 4
NRT_HD constexpr int rayPlanes(int mode) { return mode == FM_GENERAL ? 2 : 1; }

// Record lists are padded to a multiple of kRecPad with never-hit records (S = -1 => Eb < 0).
static constexpr int64_t kRecPad = 256;
NRT_HD int64_t paddedFaces(int64_t nfaces) { return (nfaces + kRecPad - 1) / kRecPad * kRecPad; }
NRT_HD int64_t recIndex(int64_t r, int k, int nc) { return ((r >> 1) * nc + k) * 2 + (r & 1); }

static constexpr double kFilterU = 5.9604644775390625e-8;    // 2^-24
static constexpr double kEps64 = 2.220446049250313e-16;       // 2^-52
static constexpr float kFilterKd = 16.0f;

NRT_HD float roundUpF(double v) {  // smallest-ish float >= v (v >= 0)
  float f = (float)v;
  if ((double)f < v) f = f * 1.0000002f + 1e-45f;
  return f;
}
NRT_HD float bitsToFloat(uint32_t u) { union { float f; uint32_t u; } c; c.u = u; return c.f; }

NRT_HD void neverHitRecord(int mode, float* c) {
  for (int k = 0; k < recFloats(mode); ++k) c[k] = 0.f;
  c[recSlotS(mode)] = -1.0f;
  if (recSlotId(mode) >= 0) c[recSlotId(mode)] = bitsToFloat(kNoTri);
}

// ----------------------------------------------------------------- records ----
// GENERAL: c = N.xyz, S | A.xyz, E2.x | E2.yz, B.xy | B.z, E1n.xyz
NRT_HD void makeRecGeneral(const DMesh& m, const double* p0, const double* p1, const double* p2, float* c) {
  const double e1x = p1[0] - p0[0], e1y = p1[1] - p0[1], e1z = p1[2] - p0[2];
  const double e2x = p2[0] - p0[0], e2y = p2[1] - p0[1], e2z = p2[2] - p0[2];
  const double cx = p0[0] - m.center[0], cy = p0[1] - m.center[1], cz = p0[2] - m.center[2];
  const double nx = -(e1y * e2z - e1z * e2y), ny = -(e1z * e2x - e1x * e2z), nz = -(e1x * e2y - e1y * e2x);
  const double ax = cy * e2z - cz * e2y, ay = cz * e2x - cx * e2z, az = cx * e2y - cy * e2x;  // c x e2
  const double bx = e1y * cz - e1z * cy, by = e1z * cx - e1x * cz, bz = e1x * cy - e1y * cx;  // e1 x c
  const double s1 = fmax(fabs(e1x) + fabs(e1y) + fabs(e1z), fabs(e2x) + fabs(e2y) + fabs(e2z));
  const double s2 = fmax(fabs(ax) + fabs(ay) + fabs(az), fabs(bx) + fabs(by) + fabs(bz));
  const double S = fmax(s1, s2 / m.L);
  c[0] = (float)nx; c[1] = (float)ny; c[2] = (float)nz; c[3] = roundUpF(S * 1.0000005);
  c[4] = (float)ax; c[5] = (float)ay; c[6] = (float)az; c[7] = (float)e2x;
  c[8] = (float)e2y; c[9] = (float)e2z; c[10] = (float)bx; c[11] = (float)by;
  c[12] = (float)bz; c[13] = (float)(-e1x); c[14] = (float)(-e1y); c[15] = (float)(-e1z);
}

// ORIGIN: c = N.xyz, S | A.xyz, id | B.xyz, 0      (frame centred at the shared origin O)
NRT_HD bool makeRecOrigin(const double* O, const double* p0, const double* p1, const double* p2, uint32_t id, float* c) {
  const double e1x = p1[0] - p0[0], e1y = p1[1] - p0[1], e1z = p1[2] - p0[2];
  const double e2x = p2[0] - p0[0], e2y = p2[1] - p0[1], e2z = p2[2] - p0[2];
  const double cx = p0[0] - O[0], cy = p0[1] - O[1], cz = p0[2] - O[2];          // v0 - O
  const double gx = e1y * e2z - e1z * e2y, gy = e1z * e2x - e1x * e2z, gz = e1x * e2y - e1y * e2x;  // e1 x e2
  // t' = (O - v0).(e1 x e2); with det > 0 the reference needs t = t'/det >= 0 (geom.nim:354).
  const double tp = -(cx * gx + cy * gy + cz * gz);
  const double s1 = fmax(fabs(e1x) + fabs(e1y) + fabs(e1z), fabs(e2x) + fabs(e2y) + fabs(e2z));
  const double scale = (fabs(cx) + fabs(cy) + fabs(cz)) * s1 * s1;
  if (tp < -1e-9 * scale) return false;  // faces away from O by far more than any rounding: never hit
  const double ax = cy * e2z - cz * e2y, ay = cz * e2x - cx * e2z, az = cx * e2y - cy * e2x;  // (v0-O) x e2
  const double bx = e1y * cz - e1z * cy, by = e1z * cx - e1x * cz, bz = e1x * cy - e1y * cx;  // e1 x (v0-O)
  const double S = fmax(fabs(gx) + fabs(gy) + fabs(gz), fmax(fabs(ax) + fabs(ay) + fabs(az), fabs(bx) + fabs(by) + fabs(bz)));
  c[0] = (float)(-gx); c[1] = (float)(-gy); c[2] = (float)(-gz); c[3] = roundUpF(S * 1.0000005);
  c[4] = (float)ax; c[5] = (float)ay; c[6] = (float)az; c[7] = bitsToFloat(id);
  c[8] = (float)bx; c[9] = (float)by; c[10] = (float)bz; c[11] = 0.f;
  return true;
}

// DIR: c = P.xyz, pu0 | Q.xyz, qv0 | S, id        (frame centred at the mesh AABB centre C)
// `D` is the shared object-space direction; det is computed exactly like rayTriangleExact.
NRT_HD bool makeRecDir(const DMesh& m, const double* D, const double* p0, const double* p1, const double* p2, uint32_t id,
                       float* c) {
  const double v0v1x = p1[0] - p0[0], v0v1y = p1[1] - p0[1], v0v1z = p1[2] - p0[2];
  const double v0v2x = p2[0] - p0[0], v0v2y = p2[1] - p0[1], v0v2z = p2[2] - p0[2];
  const double pvecx = D[1] * v0v2z - D[2] * v0v2y;
  const double pvecy = D[2] * v0v2x - D[0] * v0v2z;
  const double pvecz = D[0] * v0v2y - D[1] * v0v2x;
  const double det = v0v1x * pvecx + v0v1y * pvecy + v0v1z * pvecz;  // same operations as geom.nim:298-303
  if (det < 0.000001) return false;                                   // geom.nim:306, decided once per triangle
  if (!(det < 1e300)) return false;                                   // NaN / Inf: the reference cannot hit it either
  const double inv = 1.0 / det;
  const double px = pvecx * inv, py = pvecy * inv, pz = pvecz * inv;
  // v' = d.((o - v0) x e1) = (o - v0).(e1 x d)
  const double qx = (v0v1y * D[2] - v0v1z * D[1]) * inv, qy = (v0v1z * D[0] - v0v1x * D[2]) * inv,
               qz = (v0v1x * D[1] - v0v1y * D[0]) * inv;
  const double cx = p0[0] - m.center[0], cy = p0[1] - m.center[1], cz = p0[2] - m.center[2];
  const double pu0 = -(cx * px + cy * py + cz * pz), qv0 = -(cx * qx + cy * qy + cz * qz);
  const double S = fmax(fmax(fabs(px) + fabs(py) + fabs(pz), fabs(qx) + fabs(qy) + fabs(qz)), fmax(fabs(pu0), fabs(qv0)) / m.L);
  if (!(S < 1e30)) {  // float32 cannot represent it safely: keep it as an always-candidate record
    for (int k = 0; k < 8; ++k) c[k] = 0.f;
    c[8] = 1e30f; c[9] = bitsToFloat(id);
    return true;
  }
  c[0] = (float)px; c[1] = (float)py; c[2] = (float)pz; c[3] = (float)pu0;
  c[4] = (float)qx; c[5] = (float)qy; c[6] = (float)qz; c[7] = (float)qv0;
  c[8] = roundUpF(S * 1.0000005); c[9] = bitsToFloat(id);
  return true;
}

// -------------------------------------------------------------------- rays ----
struct FilterRay {    // plane 0: (a.xyz, rr); plane 1 (GENERAL only): (m.xyz, 0)
  float ax, ay, az, rr;
  float mx, my, mz, pad;
};

NRT_HD bool finiteMag(double v, double lo, double hi) { return (v > lo) && (v < hi); }  // false for NaN

// Returns false when the ray cannot be represented safely in float32 (it then takes the
// float64 brute-force path instead).  `r` is the object-space ray, `D` the bundle's shared
// direction (DIR), `O` its shared origin (ORIGIN).
NRT_HD bool makeFilterRay(int mode, const DMesh& m, const Ray& r, FilterRay& f) {
  const double big = 1e15, tiny = 1e-15;
  const double dx = r.dir.x, dy = r.dir.y, dz = r.dir.z;
  const double di = fmax(fabs(dx), fmax(fabs(dy), fabs(dz)));
  f.mx = f.my = f.mz = f.pad = 0.f;
  if (!finiteMag(di, tiny, big) || !finiteMag(m.L, tiny, big)) return false;
  if (mode == FM_ORIGIN) {
    f.ax = (float)dx; f.ay = (float)dy; f.az = (float)dz;
    f.rr = roundUpF(16.0 * kFilterU * di);
    return true;
  }
  const double ox = r.orig.x - m.center[0], oy = r.orig.y - m.center[1], oz = r.orig.z - m.center[2];
  const double oi = fmax(fabs(ox), fmax(fabs(oy), fabs(oz)));
  if (!(oi < big)) return false;
  if (mode == FM_DIR) {
    // slide the origin along D to the point closest to C: u, v do not depend on it (P.D = Q.D = 0)
    const double dd = dx * dx + dy * dy + dz * dz;
    const double s = (ox * dx + oy * dy + oz * dz) / dd;
    const double qx = ox - s * dx, qy = oy - s * dy, qz = oz - s * dz;
    const double qi = fmax(fabs(qx), fmax(fabs(qy), fabs(qz)));
    f.ax = (float)qx; f.ay = (float)qy; f.az = (float)qz;
    f.rr = roundUpF(16.0 * kFilterU * (qi + m.L) + 32.0 * kEps64 * oi);
    return true;
  }
  const double mx = oy * dz - oz * dy, my = oz * dx - ox * dz, mz = ox * dy - oy * dx;
  const double mi = fmax(fabs(mx), fmax(fabs(my), fabs(mz)));
  f.ax = (float)dx; f.ay = (float)dy; f.az = (float)dz;
  f.mx = (float)mx; f.my = (float)my; f.mz = (float)mz;
  f.rr = roundUpF(16.0 * kFilterU * (mi + m.L * di) + 8.0 * kEps64 * oi * di);
  return true;
}

// -------------------------------------------------------------------- test ----
// Scalar statement of one filter test (the CUDA kernel evaluates two records per FFMA2 with
// exactly these operations per component).  q = the record's NC floats, `a` = plane 0 xyz,
// `mm` = plane 1 xyz (GENERAL), rr = the thread's Rr.  Sign bit of the result clear <=> candidate.
NRT_HD uint32_t filterTest(int mode, const float* q, const float* a, const float* mm, float rr) {
  if (mode == FM_GENERAL) {
    const float eb = q[3] * rr, kd = eb * kFilterKd;
    const float u = fmaf(q[7], mm[0], fmaf(q[8], mm[1], fmaf(q[9], mm[2], fmaf(q[4], a[0], fmaf(q[5], a[1], fmaf(q[6], a[2], eb))))));
    const float v = fmaf(q[13], mm[0], fmaf(q[14], mm[1], fmaf(q[15], mm[2], fmaf(q[10], a[0], fmaf(q[11], a[1], fmaf(q[12], a[2], eb))))));
    const float det = fmaf(q[0], a[0], fmaf(q[1], a[1], fmaf(q[2], a[2], kd)));
    const float w = (det - u) - v;
    return fbits(u) | fbits(v) | fbits(w);
  }
  if (mode == FM_ORIGIN) {
    const float eb = q[3] * rr, kd = eb * kFilterKd;
    const float u = fmaf(q[4], a[0], fmaf(q[5], a[1], fmaf(q[6], a[2], eb)));
    const float v = fmaf(q[8], a[0], fmaf(q[9], a[1], fmaf(q[10], a[2], eb)));
    const float det = fmaf(q[0], a[0], fmaf(q[1], a[1], fmaf(q[2], a[2], kd)));
    const float w = (det - u) - v;
    return fbits(u) | fbits(v) | fbits(w);
  }
  const float eb = q[8] * rr, k1 = fmaf(eb, kFilterKd, 1.0f);
  const float u = fmaf(q[0], a[0], fmaf(q[1], a[1], fmaf(q[2], a[2], q[3] + eb)));
  const float v = fmaf(q[4], a[0], fmaf(q[5], a[1], fmaf(q[6], a[2], q[7] + eb)));
  const float w = (k1 - u) - v;
  return fbits(u) | fbits(v) | fbits(w);
}

// executed float32 flops per test (FFMA = 2): reported next to the roofline
NRT_HD constexpr int filterFlops(int mode) { return mode == FM_GENERAL ? 32 : (mode == FM_ORIGIN ? 20 : 14); }

}  // namespace nrt
