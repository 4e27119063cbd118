// nrt_core.h — per-ray / per-sample bodies of the render path.
//
// Everything here is `NRT_HD` (__host__ __device__): the CUDA build (libnrt.so)
// instantiates these inside its kernels; the test-only host-emulation build
// (tests/emu, never loadable through the product API) runs the same bodies in
// plain loops so the wavefront logic can be unit-tested without a GPU.
//
// Float64 code follows the reference operation by operation (citations are
// file:line in the nim-raytracer tree) and is compiled with -fmad=false so that
// no multiply-add is contracted: results are bit-identical to the IEEE oracle.
// Float32 code (the mesh filter) is explicitly fused with fmaf and is only ever
// used conservatively (see FilterRec below).
#pragma once

#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define NRT_HD __host__ __device__ __forceinline__
#else
#define NRT_HD inline
#endif

namespace nrt {

// ---------------------------------------------------------------- constants --
#define NRT_INF (__builtin_huge_val())
#define NRT_NEG_INF (-__builtin_huge_val())
static constexpr double kPi = 3.14159265358979323846;  // Nim math.PI
static constexpr uint32_t kNoTri = 0xFFFFFFFFu;
static constexpr uint32_t kInvalidRef = 0xFFFFFFFFu;

enum GeomKind { GEOM_SPHERE = 0, GEOM_PLANE = 1, GEOM_BOX = 2, GEOM_MESH = 3 };
enum LightKind { LIGHT_DISTANT = 0, LIGHT_POINT = 1 };
enum AaKind { AA_NONE = 0, AA_GRID = 1, AA_JITTERED = 2, AA_MULTI_JITTERED = 3, AA_CMJ = 4 };
enum DepthMode { DEPTH_REFBUG = 0, DEPTH_INTENDED = 1 };

// ------------------------------------------------------------- float64 math --
struct V3 { double x, y, z; };
struct V4 { double x, y, z, w; };

NRT_HD V4 v4(double x, double y, double z, double w) { V4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
NRT_HD V3 v3(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
NRT_HD V4 add(V4 a, V4 b) { return v4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
NRT_HD V4 sub(V4 a, V4 b) { return v4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
NRT_HD V4 scale(V4 a, double s) { return v4(a.x * s, a.y * s, a.z * s, a.w * s); }
NRT_HD V3 add(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
NRT_HD V3 scale(V3 a, double s) { return v3(a.x * s, a.y * s, a.z * s); }
NRT_HD V3 mul(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
NRT_HD V3 divs(V3 a, double s) { return v3(a.x / s, a.y / s, a.z / s); }
// glm dot / normalize / mat*vec: left-to-right sums (same choices as the oracle).
NRT_HD double dot(V4 a, V4 b) { return ((a.x * b.x + a.y * b.y) + a.z * b.z) + a.w * b.w; }
NRT_HD V4 normalize(V4 v) { const double s = 1.0 / sqrt(dot(v, v)); return scale(v, s); }
NRT_HD V4 mulm(const double* m, V4 v) {  // m[col*4+row]
  return v4(((m[0] * v.x + m[4] * v.y) + m[8] * v.z) + m[12] * v.w,
            ((m[1] * v.x + m[5] * v.y) + m[9] * v.z) + m[13] * v.w,
            ((m[2] * v.x + m[6] * v.y) + m[10] * v.z) + m[14] * v.w,
            ((m[3] * v.x + m[7] * v.y) + m[11] * v.z) + m[15] * v.w);
}
// Nim min/max operand order (NaN behaviour; geom.nim:88-89 relies on it)
NRT_HD double nim_min(double a, double b) { return (a <= b) ? a : b; }
NRT_HD double nim_max(double a, double b) { return (a >= b) ? a : b; }
NRT_HD double signd(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : 0.0); }  // utils/mathutils.nim:12-18

// ------------------------------------------------------------ scene (device) --
struct DObject {
  int32_t kind;
  int32_t mesh;        // index into DScene.meshes, or -1
  int32_t mesh_obj;    // index among the scene's MESH objects, or -1
  int32_t _pad;
  double o2w[16];
  double w2o[16];
  double radius;
  double bmin[4];      // Box.aabb.vmin, or the mesh AABB for MESH objects
  double bmax[4];
  double albedo[3];
  double reflection;
};

struct DLight {
  int32_t kind;
  int32_t _pad;
  double color[3];
  double intensity;
  double dir[4];
  double pos[4];
};

struct DMesh {
  const double* verts;      // nverts*4
  const double* normals;    // nnormals*4
  const int64_t* vidx;      // nfaces*3
  const int64_t* nidx;      // nfaces*3
  int64_t nverts, nnormals, nfaces;
  double bmin[4], bmax[4];  // calcAABB (geom.nim:175-188)
  double center[3];         // filter frame origin (AABB centre)
  double L;                 // filter length scale (max AABB half extent)
  float* recs;              // nfaces * 16 floats: general-mode filter records
};

struct DScene {
  int32_t nobjects, nlights, nmeshes, nmesh_objs;
  const DObject* objects;
  const DLight* lights;
  const DMesh* meshes;
  const int32_t* mesh_obj_index;  // mesh object k -> object index
  double c2w[16];
  double tan_half_fov;            // f of renderer.nim:38 (host libm, shared with nothing else)
  double bg[3];
};

// ----------------------------------------------------------------------- Ray --
struct Ray {            // geom.nim:32-39 (depth/x/y omitted: never read)
  V4 orig, dir;
  double ix, iy, iz;    // invDir
  int sx, sy, sz;       // sign
};

NRT_HD Ray initRay(V4 orig, V4 dir) {  // geom.nim:41-48
  Ray r;
  r.orig = orig; r.dir = dir;
  r.ix = 1 / dir.x; r.iy = 1 / dir.y; r.iz = 1 / dir.z;
  r.sx = r.ix < 0; r.sy = r.iy < 0; r.sz = r.iz < 0;
  return r;
}

// geom.nim:76-96
NRT_HD double aabbIntersect(const double* bmin, const double* bmax, const Ray& r) {
  double tmin = NRT_NEG_INF, tmax = NRT_INF;
  const double txmin = ((r.sx ? bmax[0] : bmin[0]) - r.orig.x) * r.ix;
  const double txmax = ((r.sx ? bmin[0] : bmax[0]) - r.orig.x) * r.ix;
  const double tymin = ((r.sy ? bmax[1] : bmin[1]) - r.orig.y) * r.iy;
  const double tymax = ((r.sy ? bmin[1] : bmax[1]) - r.orig.y) * r.iy;
  const double tzmin = ((r.sz ? bmax[2] : bmin[2]) - r.orig.z) * r.iz;
  const double tzmax = ((r.sz ? bmin[2] : bmax[2]) - r.orig.z) * r.iz;
  tmin = nim_max(tzmin, nim_max(tymin, nim_max(txmin, tmin)));
  tmax = nim_min(tzmax, nim_min(tymax, nim_min(txmax, tmax)));
  tmax *= 1.0000000000000004;
  return (tmin <= tmax) ? tmin : NRT_NEG_INF;
}

// geom.nim:215-237 ((x / 2) * a, min(t1, t2) even if negative)
NRT_HD double sphereIntersect(double radius, const Ray& r) {
  const double a = r.dir.x * r.dir.x + r.dir.y * r.dir.y + r.dir.z * r.dir.z;
  const double b = 2 * (r.dir.x * r.orig.x + r.dir.y * r.orig.y + r.dir.z * r.orig.z);
  const double c = r.orig.x * r.orig.x + r.orig.y * r.orig.y + r.orig.z * r.orig.z - radius * radius;
  const double delta = b * b - 4 * a * c;
  if (delta >= 0.0) {
    const double t1 = (-b - signd(b) * sqrt(delta)) / 2 * a;
    const double t2 = c / (a * t1);
    return nim_min(t1, t2);
  }
  return NRT_NEG_INF;
}

// geom.nim:240-248
NRT_HD double planeIntersect(const Ray& r) {
  const V4 n = v4(0.0, 1.0, 0.0, 0.0);
  const double denom = dot(n, r.dir);
  if (fabs(denom) > 1e-6) return -dot(r.orig, n) / denom;
  return NRT_NEG_INF;
}

// geom.nim:283-336 rayTriangleIntersectFast, float64, exact operation order
NRT_HD double rayTriangleExact(const Ray& r, const double* v0, const double* v1, const double* v2) {
  const double v0v1x = v1[0] - v0[0], v0v1y = v1[1] - v0[1], v0v1z = v1[2] - v0[2];
  const double v0v2x = v2[0] - v0[0], v0v2y = v2[1] - v0[1], v0v2z = v2[2] - v0[2];
  const double pvecx = r.dir.y * v0v2z - r.dir.z * v0v2y;
  const double pvecy = r.dir.z * v0v2x - r.dir.x * v0v2z;
  const double pvecz = r.dir.x * v0v2y - r.dir.y * v0v2x;
  const double det = v0v1x * pvecx + v0v1y * pvecy + v0v1z * pvecz;
  if (det < 0.000001) return NRT_NEG_INF;
  const double invDet = 1 / det;
  const double tvecx = r.orig.x - v0[0], tvecy = r.orig.y - v0[1], tvecz = r.orig.z - v0[2];
  const double u = (tvecx * pvecx + tvecy * pvecy + tvecz * pvecz) * invDet;
  if (u < 0 || u > 1) return NRT_NEG_INF;
  const double qvecx = tvecy * v0v1z - tvecz * v0v1y;
  const double qvecy = tvecz * v0v1x - tvecx * v0v1z;
  const double qvecz = tvecx * v0v1y - tvecy * v0v1x;
  const double v = (r.dir.x * qvecx + r.dir.y * qvecy + r.dir.z * qvecz) * invDet;
  if (v < 0 || u + v > 1) return NRT_NEG_INF;
  return (v0v2x * qvecx + v0v2y * qvecy + v0v2z * qvecz) * invDet;
}

// geom.nim:361-379 (non-mesh kinds)
NRT_HD V4 geomNormal(const DObject& o, V4 hit) {
  if (o.kind == GEOM_SPHERE) return normalize(v4(hit.x, hit.y, hit.z, 0.0));
  if (o.kind == GEOM_PLANE) return v4(0.0, 1.0, 0.0, 0.0);
  if (o.kind == GEOM_BOX) {
    const V4 bmin = v4(o.bmin[0], o.bmin[1], o.bmin[2], o.bmin[3]);
    const V4 bmax = v4(o.bmax[0], o.bmax[1], o.bmax[2], o.bmax[3]);
    const V4 c = scale(add(bmin, bmax), 0.5);
    const V4 p = sub(hit, c);
    const V4 d = scale(sub(bmin, bmax), 0.5);
    const double bias = 1.000001;
    return normalize(v4(trunc(p.x / fabs(d.x) * bias), trunc(p.y / fabs(d.y) * bias),
                        trunc(p.z / fabs(d.z) * bias), 0.0));
  }
  return v4(0, 0, 0, 0);
}

struct ShadingInfo { V4 lightDir; V3 lightIntensity; double lightDistance; };

// light.nim:46-62
NRT_HD ShadingInfo getShadingInfo(const DLight& l, V4 p) {
  ShadingInfo si;
  const V3 color = v3(l.color[0], l.color[1], l.color[2]);
  if (l.kind == LIGHT_DISTANT) {
    si.lightDir = v4(l.dir[0], l.dir[1], l.dir[2], l.dir[3]);
    si.lightIntensity = scale(color, l.intensity);
    si.lightDistance = NRT_INF;
    return si;
  }
  V4 lightDir = sub(p, v4(l.pos[0], l.pos[1], l.pos[2], l.pos[3]));
  const double r2 = dot(lightDir, lightDir);
  si.lightDir = normalize(lightDir);
  si.lightIntensity = divs(scale(color, l.intensity), (4 * kPi * r2));
  si.lightDistance = sqrt(r2);
  return si;
}

// shader.nim:12-17
NRT_HD V3 shadeDiffuse(const DObject& o, const ShadingInfo& si, V4 hitNormal) {
  const V3 albedo = v3(o.albedo[0], o.albedo[1], o.albedo[2]);
  const double c = nim_max(0.0, dot(hitNormal, scale(si.lightDir, -1.0)));
  return scale(mul(divs(albedo, kPi), si.lightIntensity), c);
}

// renderer.nim:31-44 (orig/dir only; initRay is applied per object in trace)
NRT_HD void castPrimaryRay(const DScene& sc, int w, int h, double x, double y, V4& orig, V4& dir) {
  const double r = double(w) / double(h);
  const double f = sc.tan_half_fov;
  const double cx = ((2 * x * r) / double(w) - r) * f;
  const double cy = (1 - 2 * y / double(h)) * f;
  orig = mulm(sc.c2w, v4(0.0, 0.0, 0.0, 1.0));
  dir = mulm(sc.c2w, normalize(v4(cx, cy, -1, 0.0)));
}

// counter-based RNG for the jittered AA kinds (same spec as the oracle)
NRT_HD uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
struct PixelRng {
  uint64_t key, ctr;
  NRT_HD double next() {
    const uint64_t v = splitmix64(key + 0xD1342543DE82EF95ull * (ctr++));
    return double(v >> 11) * (1.0 / 9007199254740992.0);
  }
  NRT_HD double random(double mx) { return next() * mx; }
};
NRT_HD PixelRng pixelRng(uint64_t seed, int width, int x, int y) {
  PixelRng r;
  r.key = splitmix64(seed ^ (0x632BE59BD9B4E019ull * (uint64_t(y) * uint64_t(width) + uint64_t(x) + 1)));
  r.ctr = 0;
  return r;
}

// ordered-bits helpers: non-negative doubles compare like their uint64 patterns
NRT_HD uint64_t dbits(double d) { union { double d; uint64_t u; } c; c.d = d; return c.u; }
NRT_HD double bitsd(uint64_t u) { union { double d; uint64_t u; } c; c.u = u; return c.d; }
NRT_HD uint32_t fbits(float f) { union { float f; uint32_t u; } c; c.f = f; return c.u; }

// ---------------------------------------------------- float32 mesh filter -----
// Möller–Trumbore (geom.nim:283-336) rewritten with scalar triple products in a
// frame centred at the mesh AABB centre C, ray = (d, m = (o - C) x d):
//   det = e1.(d x e2)            = N . d            N  = -(e1 x e2)
//   u'  = (o - v0).(d x e2)      = E2 . m + A . d   A  = (v0 - C) x e2
//   v'  = d.((o - v0) x e1)      = E1n . m + B . d  E1n = -e1, B = e1 x (v0 - C)
// (u' = u*det, v' = v*det.)  A float32 evaluation with margin Eb decides
// "possibly hit"; every such (ray, triangle) pair is re-evaluated in float64 with
// the reference's exact operation order (rayTriangleExact).  The `t` tests of the
// reference (t >= 0, t < tMin) are not needed in the filter: the mesh is only
// tested when the ray origin is outside its AABB and the box is in front
// (geom.nim:340), so every line/triangle crossing has t >= 0 up to rounding —
// and the float64 pass applies them exactly anyway.
//
// Record layout (16 floats = 4 x 16-byte vectors, loaded with LDS.128):
//   q0 = (N.x,  N.y,  N.z,  S)      S = per-triangle magnitude scale (see below)
//   q1 = (A.x,  A.y,  A.z,  E2.x)
//   q2 = (E2.y, E2.z, B.x,  B.y)
//   q3 = (B.z,  E1n.x,E1n.y,E1n.z)
//
// Error bound (unit roundoff u = 2^-24, round-to-nearest, no overflow): each of
// u', v' is a 6-term fmaf chain of float32-rounded inputs, so
//   |fl(u') - u'| <= 9u (|E2|.|m| + |A|.|d|) <= 9u S (|m|_inf + L |d|_inf)
// with S = max(|e1|_1, |e2|_1, |A|_1 / L, |B|_1 / L) and L = max AABB half extent;
// det (3 terms) errs by <= 5u |N|_1 |d|_inf <= 30u S L |d|_inf.  With
//   Rr = 16u (|m|_inf + L |d|_inf)  (per ray),  Eb = S * Rr,  Kd = 16 Eb
// the tests  u'+Eb >= 0,  v'+Eb >= 0,  (det+Kd) - (u'+Eb) - (v'+Eb) >= 0  hold for
// every pair the float64 reference accepts (margin analysis in DESIGN.md §4).
struct FilterRay {   // 8 floats = 2 x 16-byte vectors
  float dx, dy, dz, rr;
  float mx, my, mz, pad;
};

static constexpr double kFilterU = 5.9604644775390625e-8;  // 2^-24

NRT_HD float roundUpF(double v) {  // float >= v (v >= 0)
  float f = (float)v;
  if ((double)f < v) f = f * 1.0000002f + 1e-45f;
  return f;
}

// Builds the filter-side representation of an object-space ray (float64 in).
NRT_HD FilterRay makeFilterRay(const DMesh& m, const Ray& r) {
  const double ox = r.orig.x - m.center[0], oy = r.orig.y - m.center[1], oz = r.orig.z - m.center[2];
  const double dx = r.dir.x, dy = r.dir.y, dz = r.dir.z;
  const double mx = oy * dz - oz * dy, my = oz * dx - ox * dz, mz = ox * dy - oy * dx;
  const double mi = fmax(fabs(mx), fmax(fabs(my), fabs(mz)));
  const double di = fmax(fabs(dx), fmax(fabs(dy), fabs(dz)));
  // + the float64 cancellation error of m itself (2^-52 |o||d|, negligible but counted)
  const double oi = fmax(fabs(ox), fmax(fabs(oy), fabs(oz)));
  const double rr = 16.0 * kFilterU * (mi + m.L * di) + 4.0 * 2.220446049250313e-16 * oi * di;
  FilterRay f;
  f.dx = (float)dx; f.dy = (float)dy; f.dz = (float)dz;
  f.mx = (float)mx; f.my = (float)my; f.mz = (float)mz;
  f.rr = roundUpF(rr);
  f.pad = 0.f;
  return f;
}

// Builds the 16-float record of triangle (v0, v1, v2) (object space, float64).
NRT_HD void makeFilterRec(const DMesh& m, const double* p0, const double* p1, const double* p2, float* q) {
  const double e1x = p1[0] - p0[0], e1y = p1[1] - p0[1], e1z = p1[2] - p0[2];
  const double e2x = p2[0] - p0[0], e2y = p2[1] - p0[1], e2z = p2[2] - p0[2];
  const double cx = p0[0] - m.center[0], cy = p0[1] - m.center[1], cz = p0[2] - m.center[2];
  const double nx = -(e1y * e2z - e1z * e2y), ny = -(e1z * e2x - e1x * e2z), nz = -(e1x * e2y - e1y * e2x);
  const double ax = cy * e2z - cz * e2y, ay = cz * e2x - cx * e2z, az = cx * e2y - cy * e2x;   // c x e2
  const double bx = e1y * cz - e1z * cy, by = e1z * cx - e1x * cz, bz = e1x * cy - e1y * cx;   // e1 x c
  const double s1 = fmax(fabs(e1x) + fabs(e1y) + fabs(e1z), fabs(e2x) + fabs(e2y) + fabs(e2z));
  const double s2 = fmax(fabs(ax) + fabs(ay) + fabs(az), fabs(bx) + fabs(by) + fabs(bz));
  const double S = fmax(s1, s2 / m.L);
  q[0] = (float)nx; q[1] = (float)ny; q[2] = (float)nz; q[3] = roundUpF(S * 1.0000005);
  q[4] = (float)ax; q[5] = (float)ay; q[6] = (float)az; q[7] = (float)e2x;
  q[8] = (float)e2y; q[9] = (float)e2z; q[10] = (float)bx; q[11] = (float)by;
  q[12] = (float)bz; q[13] = (float)(-e1x); q[14] = (float)(-e1y); q[15] = (float)(-e1z);
}

// One filter test.  Returns the OR of the three sign words: sign bit clear <=> candidate.
NRT_HD uint32_t filterTest(const float* q, float dx, float dy, float dz, float mx, float my, float mz,
                           float eb, float kd) {
  const float u = fmaf(q[7], mx, fmaf(q[8], my, fmaf(q[9], mz, fmaf(q[4], dx, fmaf(q[5], dy, fmaf(q[6], dz, eb))))));
  const float v = fmaf(q[13], mx, fmaf(q[14], my, fmaf(q[15], mz, fmaf(q[10], dx, fmaf(q[11], dy, fmaf(q[12], dz, eb))))));
  const float det = fmaf(q[0], dx, fmaf(q[1], dy, fmaf(q[2], dz, kd)));
  const float w = (det - u) - v;
  return fbits(u) | fbits(v) | fbits(w);
}

static constexpr float kFilterKd = 16.0f;

// Records are padded to a multiple of kRecPad faces with never-hit records
// (all zero, S = -1 => Eb < 0 => u' + Eb < 0) so the hot loop has no tail.
static constexpr int64_t kRecPad = 256;
NRT_HD int64_t paddedFaces(int64_t nfaces) { return (nfaces + kRecPad - 1) / kRecPad * kRecPad; }

}  // namespace nrt
