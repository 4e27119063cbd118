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

// Record lists are padded to a multiple of kRecPad (the prefilter's shared-memory chunk) with
// never-hit records (S = -1 => Eb < 0).
static constexpr int64_t kRecPad = 256;
NRT_HD int64_t paddedFaces(int64_t nfaces) { return (nfaces + kRecPad - 1) / kRecPad * kRecPad; }
NRT_HD int64_t recIndex(int64_t r, int k, int nc) { return ((r >> 1) * nc + k) * 2 + (r & 1); }   // hot records (FFMA2 pairs)
// Full records (read one at a time by the refine stage) are plain 64-byte rows: four 16-byte loads.
static constexpr int kFullStride = 16;
NRT_HD int64_t fullIndex(int64_t r, int k) { return r * kFullStride + k; }

static constexpr double kFilterU = 5.9604644775390625e-8;    // 2^-24
static constexpr double kEps64 = 2.220446049250313e-16;       // 2^-52
static constexpr float kFilterKd = 16.0f;

NRT_HD float roundUpF(double v) {  // smallest-ish float >= v (v >= 0)
  float f = (float)v;
  if ((double)f < v) f = f * 1.0000002f + 1e-45f;
  return f;
}
NRT_HD float bitsToFloat(uint32_t u) { union { float f; uint32_t u; } c; c.u = u; return c.f; }
NRT_HD float roundUpSigned(double v) {  // a float >= v, either sign
  float f = (float)v;
  if ((double)f < v) f = (v >= 0) ? f * 1.0000002f + 1e-45f : f * 0.9999998f;
  return f;
}

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

// =====================================================================================
// PREFILTER (the hot loop): bounding circle / sphere of every triangle.
//
// A ray can only hit a triangle if its line passes through the triangle's minimal enclosing
// sphere (centre c, radius r).  For ray bundles this collapses to 2-D:
//   DIR     orthographic projection along the shared direction D onto the plane (b1, b2):
//           the ray is a point p = (x, y); the triangle projects to a 2-D triangle with
//           enclosing circle (c, r);  hit  =>  |p - c|^2 <= r^2
//   ORIGIN  perspective projection from the shared origin O onto the plane z = 1 of the frame
//           (b1, b2, f):  p = (d.b1, d.b2) / d.f,  vertices (v-O).b / (v-O).f  — lines map to
//           lines, so the projected triangle (all three vertices in front: z > 0) bounds the
//           rays that can hit it; triangles touching z <= 0 become always-candidates.
//   GENERAL distance from the sphere centre to the ray line, with the ray given by its unit
//           direction dh and the point p0 of the line closest to the mesh centre (p0 . dh = 0):
//           dist^2 = |c - p0|^2 - (c . dh)^2 <= r^2
// Expanded so that every term is a product of one per-triangle and one per-ray quantity:
//   2-D:     (r^2 - |c|^2) + 2c.p                 >=  |p|^2           2 FFMA + 1 compare
//   GENERAL: (r^2 - |c|^2) + 2c.p0 + (c.dh)^2     >=  |p0|^2          1 FMUL + 6 FFMA + 1 compare
// (left side per (triangle, ray), right side a per-ray threshold).  Margins that bound the float32
// evaluation error are folded into the two additive constants at build time:
//   per triangle  mt = 8u (2|c|^2 + |r^2 - |c|^2|)   (GENERAL: 32u|c|^2 + 8u|K0|)
//   per ray       mr = 16u |p|^2 (+ float64 slack)
// Survivors ("pre-candidates") go through the float32 sign test above (refine stage) and
// only then to the float64 evaluation, so the hot loop touches 12-16 bytes per triangle.
// =====================================================================================

// floats per hot record (pair-interleaved like the full records)
NRT_HD constexpr int hotFloats(int mode) { return mode == FM_GENERAL ? 4 : 3; }
NRT_HD constexpr int prefilterFlops(int mode) { return mode == FM_GENERAL ? 13 : 4; }
// Rays of one prefilter "run": the consecutive queue entries one warp keeps in registers (32 lanes x
// 8 or 4 rays) and decides chunk culling for (nrt.cu: k_mesh_prefilter; mirrored by the emulation).
NRT_HD constexpr int prefilterRunRays(int mode) { return mode == FM_GENERAL ? 128 : 256; }

// Shared frame of a ray bundle (object space), built on the host in float64.
struct BundleFrame {
  double org[3];   // ORIGIN: the shared origin O;  DIR / GENERAL: the mesh AABB centre C
  double b1[3], b2[3], f[3];   // orthonormal; f = projection axis (DIR: the shared direction, unit)
  double valid;    // 0: no usable frame (degenerate direction) => the bundle's rays are treated as GENERAL
};
// One filter record set (device pointers), indexed like the frames: what a thread needs to walk the
// flattened hierarchy of a mesh object by itself (nrt_pipeline.h: meshIntersectWalk).
struct RecSet {
  const float* hot;      // hot records, pair-interleaved, padded to kRecPad
  const float* bounds;   // one hot-format bound per chunk of kRecPad records
  const float* sub;      // one hot-format bound per sub-chunk of kSubRecs records
  const float* recs;     // full records (64-byte rows)
  const uint32_t* ids;   // GENERAL: record position -> face (Morton order); else null (the id is in the record)
  uint32_t nrec;         // records kept (before padding)
  uint32_t usable;       // the set was built (valid frame)
};
NRT_HD int frameIndex(int nlights, int mo, int mode, int l) { return mo * (2 + nlights) + (mode == FM_GENERAL ? 0 : (mode == FM_ORIGIN ? 1 : 2 + l)); }

NRT_HD void neverHitHot(int mode, float* h) {
  for (int k = 0; k < hotFloats(mode); ++k) h[k] = 0.f;
  h[mode == FM_GENERAL ? 3 : 2] = -1e30f;
}
NRT_HD void alwaysHot(int mode, float* h) {
  for (int k = 0; k < hotFloats(mode); ++k) h[k] = 0.f;
  h[mode == FM_GENERAL ? 3 : 2] = 1e30f;
}

// Minimal enclosing circle of a 2-D triangle (float64).  Returns false if degenerate beyond use.
NRT_HD void enclosingCircle2(const double* a, const double* b, const double* c, double& cx, double& cy, double& r2) {
  // longest edge first
  const double ab = (b[0] - a[0]) * (b[0] - a[0]) + (b[1] - a[1]) * (b[1] - a[1]);
  const double bc = (c[0] - b[0]) * (c[0] - b[0]) + (c[1] - b[1]) * (c[1] - b[1]);
  const double ca = (a[0] - c[0]) * (a[0] - c[0]) + (a[1] - c[1]) * (a[1] - c[1]);
  const double *p = a, *q = b, *o = c; double l2 = ab;
  if (bc >= l2) { p = b; q = c; o = a; l2 = bc; }
  if (ca >= l2) { p = c; q = a; o = b; l2 = ca; }
  // angle at o >= 90 deg  <=>  (p-o).(q-o) <= 0: the circle on pq as diameter encloses o
  const double dotv = (p[0] - o[0]) * (q[0] - o[0]) + (p[1] - o[1]) * (q[1] - o[1]);
  cx = 0.5 * (p[0] + q[0]); cy = 0.5 * (p[1] + q[1]); r2 = 0.25 * l2;
  if (dotv > 0) {  // acute: circumcircle
    const double bx = q[0] - p[0], by = q[1] - p[1], ex = o[0] - p[0], ey = o[1] - p[1];
    const double d = 2 * (bx * ey - by * ex);
    const double b2 = bx * bx + by * by, e2 = ex * ex + ey * ey;
    if (d != 0) {
      const double ux = (ey * b2 - by * e2) / d, uy = (bx * e2 - ex * b2) / d;
      const double rr = ux * ux + uy * uy;
      if (rr < 4 * l2) { cx = p[0] + ux; cy = p[1] + uy; r2 = rr; }   // acute => R <= longest edge; else keep diameter circle
      else {  // numerically degenerate: fall back to a circle around the longest edge's midpoint
        const double mx = o[0] - cx, my = o[1] - cy;
        r2 = fmax(r2, mx * mx + my * my);
      }
    }
  }
  // make it enclose all three vertices despite rounding
  double m = 0;
  const double* v[3] = {a, b, c};
  for (int i = 0; i < 3; ++i) { const double x = v[i][0] - cx, y = v[i][1] - cy; m = fmax(m, x * x + y * y); }
  r2 = fmax(r2, m) * (1.0 + 1e-9);
}

// Minimal enclosing sphere of a 3-D triangle (centre in its plane), float64.
NRT_HD void enclosingSphere3(const double* a, const double* b, const double* c, double* ctr, double& r2) {
  double ab = 0, bc = 0, ca = 0;
  for (int k = 0; k < 3; ++k) { ab += (b[k] - a[k]) * (b[k] - a[k]); bc += (c[k] - b[k]) * (c[k] - b[k]); ca += (a[k] - c[k]) * (a[k] - c[k]); }
  const double *p = a, *q = b, *o = c; double l2 = ab;
  if (bc >= l2) { p = b; q = c; o = a; l2 = bc; }
  if (ca >= l2) { p = c; q = a; o = b; l2 = ca; }
  double dotv = 0;
  for (int k = 0; k < 3; ++k) { dotv += (p[k] - o[k]) * (q[k] - o[k]); ctr[k] = 0.5 * (p[k] + q[k]); }
  r2 = 0.25 * l2;
  if (dotv > 0) {
    double u[3], w[3], n[3];
    for (int k = 0; k < 3; ++k) { u[k] = q[k] - p[k]; w[k] = o[k] - p[k]; }
    n[0] = u[1] * w[2] - u[2] * w[1]; n[1] = u[2] * w[0] - u[0] * w[2]; n[2] = u[0] * w[1] - u[1] * w[0];
    const double n2 = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
    const double u2 = u[0] * u[0] + u[1] * u[1] + u[2] * u[2], w2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    if (n2 > 0) {
      // circumcentre - p = (|w|^2 (n x u)... ) : ( w2 * (u x n)... ) standard form: (u2 (w x n) ... )
      // c - p = ( |u|^2 (w x n) ... ) is error-prone; use: ((n x u) w2 + (w x n) u2) / (2 n2)
      double nu[3] = {n[1] * u[2] - n[2] * u[1], n[2] * u[0] - n[0] * u[2], n[0] * u[1] - n[1] * u[0]};
      double wn[3] = {w[1] * n[2] - w[2] * n[1], w[2] * n[0] - w[0] * n[2], w[0] * n[1] - w[1] * n[0]};
      double t[3], rr = 0;
      for (int k = 0; k < 3; ++k) { t[k] = (nu[k] * w2 + wn[k] * u2) / (2 * n2); rr += t[k] * t[k]; }
      if (rr < 4 * l2) { for (int k = 0; k < 3; ++k) ctr[k] = p[k] + t[k]; r2 = rr; }
      else { double m = 0; for (int k = 0; k < 3; ++k) m += (o[k] - ctr[k]) * (o[k] - ctr[k]); r2 = fmax(r2, m); }
    }
  }
  double m = 0;
  const double* v[3] = {a, b, c};
  for (int i = 0; i < 3; ++i) { double s = 0; for (int k = 0; k < 3; ++k) s += (v[i][k] - ctr[k]) * (v[i][k] - ctr[k]); m = fmax(m, s); }
  r2 = fmax(r2, m) * (1.0 + 1e-9);
}

NRT_HD double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// Hot record of triangle (p0, p1, p2) for a bundle frame.  h has hotFloats(mode) entries.
NRT_HD void makeHotRec(int mode, const BundleFrame& fr, const double* p0, const double* p1, const double* p2, float* h) {
  const double* v[3] = {p0, p1, p2};
  if (mode == FM_GENERAL) {
    double a[3], b[3], c[3], ctr[3], r2;
    for (int k = 0; k < 3; ++k) { a[k] = p0[k] - fr.org[k]; b[k] = p1[k] - fr.org[k]; c[k] = p2[k] - fr.org[k]; }
    enclosingSphere3(a, b, c, ctr, r2);
    const double c2 = dot3(ctr, ctr), K0 = r2 - c2;
    const double mt = 32.0 * kFilterU * c2 + 8.0 * kFilterU * fabs(K0) + 8.0 * kFilterU * r2;
    if (!(c2 < 1e30) || !(r2 < 1e30)) { alwaysHot(mode, h); return; }
    h[0] = (float)ctr[0]; h[1] = (float)ctr[1]; h[2] = (float)ctr[2]; h[3] = roundUpSigned(K0 + mt);
    return;
  }
  double q[3][2];
  for (int i = 0; i < 3; ++i) {
    double w[3] = {v[i][0] - fr.org[0], v[i][1] - fr.org[1], v[i][2] - fr.org[2]};
    double x = dot3(w, fr.b1), y = dot3(w, fr.b2);
    if (mode == FM_ORIGIN) {
      const double z = dot3(w, fr.f);
      const double wl = fabs(w[0]) + fabs(w[1]) + fabs(w[2]);
      if (!(z > 1e-6 * wl) || !(z > 1e-290)) { alwaysHot(mode, h); return; }   // vertex at / behind the projection plane
      x /= z; y /= z;
    }
    q[i][0] = x; q[i][1] = y;
  }
  double cx, cy, r2;
  enclosingCircle2(q[0], q[1], q[2], cx, cy, r2);
  const double c2 = cx * cx + cy * cy, c0 = r2 - c2;
  const double mt = 8.0 * kFilterU * (2.0 * c2 + fabs(c0)) + 8.0 * kFilterU * r2;
  if (!(c2 < 1e30) || !(r2 < 1e30)) { alwaysHot(mode, h); return; }
  h[0] = (float)(2.0 * cx); h[1] = (float)(2.0 * cy);
  h[2] = roundUpSigned(c0 + mt);
}

// Hot representation of a ray: plane H0 = (x, y, q, 0) [2-D] or (dh.xyz, q) [GENERAL]; plane H1 = (2 p0, 0).
struct HotRay { float a0, a1, a2, a3; float b0, b1, b2, b3; };

NRT_HD bool makeHotRay(int mode, const BundleFrame& fr, const Ray& r, HotRay& h) {
  const double big = 1e15;
  h.a0 = h.a1 = h.a2 = h.a3 = h.b0 = h.b1 = h.b2 = h.b3 = 0.f;
  const double d[3] = {r.dir.x, r.dir.y, r.dir.z};
  const double o[3] = {r.orig.x - fr.org[0], r.orig.y - fr.org[1], r.orig.z - fr.org[2]};
  const double oi = fmax(fabs(o[0]), fmax(fabs(o[1]), fabs(o[2])));
  if (mode == FM_ORIGIN) {
    const double z = dot3(d, fr.f), dl = fabs(d[0]) + fabs(d[1]) + fabs(d[2]);
    if (!(z > 1e-6 * dl) || !(z > 1e-290)) return false;   // not in front of the projection plane
    const double x = dot3(d, fr.b1) / z, y = dot3(d, fr.b2) / z;
    const double p2 = x * x + y * y;
    if (!(p2 < big)) return false;
    const double q = -p2 + 16.0 * kFilterU * p2 + 1e-14 * (1.0 + p2);
    h.a0 = (float)x; h.a1 = (float)y; h.a2 = -roundUpSigned(q);   // threshold T = -q, rounded down
    return true;
  }
  if (!(oi < big)) return false;
  if (mode == FM_DIR) {
    const double x = dot3(o, fr.b1), y = dot3(o, fr.b2);
    const double p2 = x * x + y * y;
    if (!(p2 < 1e29)) return false;
    const double q = -p2 + 16.0 * kFilterU * p2 + 64.0 * kEps64 * oi * oi;
    h.a0 = (float)x; h.a1 = (float)y; h.a2 = -roundUpSigned(q);   // threshold T = -q, rounded down
    return true;
  }
  const double dl2 = dot3(d, d);
  if (!(dl2 > 1e-290) || !(dl2 < 1e290)) return false;
  const double il = 1.0 / sqrt(dl2);
  const double dh[3] = {d[0] * il, d[1] * il, d[2] * il};
  const double s = dot3(o, dh);
  const double p0[3] = {o[0] - s * dh[0], o[1] - s * dh[1], o[2] - s * dh[2]};
  const double p2 = dot3(p0, p0);
  if (!(p2 < 1e29)) return false;
  const double q = -p2 + 16.0 * kFilterU * p2 + 64.0 * kEps64 * oi * oi;
  h.a0 = (float)dh[0]; h.a1 = (float)dh[1]; h.a2 = (float)dh[2]; h.a3 = -roundUpSigned(q);   // threshold T = -q
  h.b0 = (float)(2.0 * p0[0]); h.b1 = (float)(2.0 * p0[1]); h.b2 = (float)(2.0 * p0[2]);
  return true;
}

// ---- spatial order + chunk bounds (two-level flattened traversal of a record set) ----
// Records are stored in Morton order of the face centroids, so the kRecPad (256) consecutive
// records of a shared-memory chunk are spatially compact; one extra hot-format record per chunk
// encloses all of the chunk's circles / spheres.  A warp tests its rays against the chunk bound
// first and skips the chunk when none of them can reach it.
NRT_HD uint32_t expandBits10(uint32_t v) {
  v = (v * 0x00010001u) & 0xFF0000FFu;
  v = (v * 0x00000101u) & 0x0F00F00Fu;
  v = (v * 0x00000011u) & 0xC30C30C3u;
  v = (v * 0x00000005u) & 0x49249249u;
  return v;
}
NRT_HD uint32_t mortonKey(const DMesh& m, const double* p0, const double* p1, const double* p2) {
  uint32_t q[3];
  for (int k = 0; k < 3; ++k) {
    const double c = (p0[k] + p1[k] + p2[k]) * (1.0 / 3.0), ext = m.bmax[k] - m.bmin[k];
    double t = (ext > 0) ? (c - m.bmin[k]) / ext : 0.0;
    if (!(t > 0)) t = 0;          // also NaN
    if (t > 0.999999) t = 0.999999;
    q[k] = uint32_t(t * 1024.0);
  }
  return (expandBits10(q[0]) << 2) | (expandBits10(q[1]) << 1) | expandBits10(q[2]);
}

// Bound of chunk `ch` (records [ch*kRecPad, (ch+1)*kRecPad)) of a hot record array, in hot-record
// format: a circle / sphere (C, R) around the records' own circles {centre c, rho^2 = c0m + |c|^2}
// (2-D: c = (h0, h1)/2; GENERAL: c = h[0..2], rho^2 = K0m + |c|^2; rho >= the face's true radius
// because c0m carries the face margin).  A ray that truly hits a face of the chunk lies within
// d + rho <= R / (1 + 1e-6) of C, and the float32 evaluation of the bound test carries the same
// margins as a record's, so it passes: culling by the bound never loses a hit.
NRT_HD void rangeBound(int mode, const float* hot, int64_t first, int64_t count, float* b) {
  const int nh = hotFloats(mode), dim = (mode == FM_GENERAL) ? 3 : 2, last = nh - 1;
  const double half = (mode == FM_GENERAL) ? 1.0 : 0.5;
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  int64_t n = 0;
  bool always = false;
  for (int64_t r = first; r < first + count; ++r) {
    const float k0 = hot[recIndex(r, last, nh)];
    if (k0 <= -1e29f) continue;             // never-hit padding
    if (k0 >= 1e29f) { always = true; break; }
    double c[3] = {0, 0, 0}, c2 = 0;
    for (int k = 0; k < dim; ++k) { c[k] = half * double(hot[recIndex(r, k, nh)]); c2 += c[k] * c[k]; }
    const double rho2 = double(k0) + c2, rho = rho2 > 0 ? sqrt(rho2) : 0.0;
    for (int k = 0; k < dim; ++k) { lo[k] = fmin(lo[k], c[k] - rho); hi[k] = fmax(hi[k], c[k] + rho); }
    ++n;
  }
  if (always) { alwaysHot(mode, b); return; }
  if (n == 0) { neverHitHot(mode, b); return; }
  // centre of the box around the circles (tighter than their mean for elongated chunks)
  double C[3] = {0.5 * (lo[0] + hi[0]), 0.5 * (lo[1] + hi[1]), dim == 3 ? 0.5 * (lo[2] + hi[2]) : 0.0};
  double R = 0;
  for (int64_t r = first; r < first + count; ++r) {
    const float k0 = hot[recIndex(r, last, nh)];
    if (k0 <= -1e29f) continue;
    double c2 = 0, d2 = 0;
    for (int k = 0; k < dim; ++k) {
      const double c = half * double(hot[recIndex(r, k, nh)]);
      c2 += c * c; d2 += (c - C[k]) * (c - C[k]);
    }
    const double rho2 = double(k0) + c2;
    const double rho = rho2 > 0 ? sqrt(rho2) : 0.0;
    R = fmax(R, sqrt(d2) + rho);
  }
  R *= 1.0 + 1e-6;
  const double C2 = C[0] * C[0] + C[1] * C[1] + (dim == 3 ? C[2] * C[2] : 0.0), r2 = R * R, k0 = r2 - C2;
  if (!(C2 < 1e29) || !(r2 < 1e29)) { alwaysHot(mode, b); return; }
  if (mode == FM_GENERAL) {
    const double mt = 32.0 * kFilterU * C2 + 8.0 * kFilterU * fabs(k0) + 8.0 * kFilterU * r2;
    b[0] = (float)C[0]; b[1] = (float)C[1]; b[2] = (float)C[2]; b[3] = roundUpSigned(k0 + mt);
  } else {
    const double mt = 8.0 * kFilterU * (2.0 * C2 + fabs(k0)) + 8.0 * kFilterU * r2;
    b[0] = (float)(2.0 * C[0]); b[1] = (float)(2.0 * C[1]); b[2] = roundUpSigned(k0 + mt);
  }
}

NRT_HD void chunkBound(int mode, const float* hot, int64_t ch, float* b) { rangeBound(mode, hot, ch * kRecPad, kRecPad, b); }
// Third level: every chunk is cut into kSubPerChunk runs of kSubRecs consecutive records (sub-patches of
// the chunk's patch, still in Morton order) with their own bounds; a (ray run, chunk) pair evaluates only
// the sub-chunks some ray of the run can reach.
static constexpr int kSubRecs = 16, kSubPerChunk = int(kRecPad) / kSubRecs;
NRT_HD void subChunkBound(int mode, const float* hot, int64_t sub, float* b) { rangeBound(mode, hot, sub * kSubRecs, kSubRecs, b); }

// Scalar statement of one prefilter test (the CUDA kernel evaluates two records per FFMA2 with
// exactly these operations per component).  The per-ray constant is kept on the other side of
// the comparison (threshold T = -q, stored in the ray) instead of being added: one float32
// operation less per test and no extra rounding.  true <=> pre-candidate.
NRT_HD bool prefilterTest(int mode, const float* h, const HotRay& r) {
  if (mode == FM_GENERAL) {
    const float s = fmaf(h[0], r.a0, fmaf(h[1], r.a1, h[2] * r.a2));
    const float t = fmaf(h[0], r.b0, fmaf(h[1], r.b1, fmaf(h[2], r.b2, h[3])));
    return fmaf(s, s, t) >= r.a3;
  }
  return fmaf(h[0], r.a0, fmaf(h[1], r.a1, h[2])) >= r.a2;
}

// executed float32 flops per test (FFMA = 2): reported next to the roofline
NRT_HD constexpr int filterFlops(int mode) { return mode == FM_GENERAL ? 32 : (mode == FM_ORIGIN ? 20 : 14); }

}  // namespace nrt
