// ref_cpu.cpp — CPU ORACLE for the nim-raytracer per-pixel render hot path.
//
// TEST INFRASTRUCTURE ONLY.  This file is a float64 CPU restatement of the
// reference renderer (johnnovak/nim-raytracer, Nim) used as the parity checker
// and as the CPU baseline ("port") of bench.py.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load it.  The product
// (nim_raytracer_b200/csrc, libnrt.so) never links, imports or calls anything here.
//
// Why a restatement: the reference cannot be built in this image (no nim, no
// clang, and its `glm` dependency ../nim-glm-fork is not in the tree; nim.cfg:1).
// Parity pins available from the reference itself (see tests/test_oracle_golden.py):
//   * utils/mathutils.nim:34-45   asserted quadratic golden (sphere solver formula)
//   * test/boxtest.nim:32-33      camera origin/direction golden (rotate/translate
//                                 composition, handedness, pixel mapping) — printed
//                                 with 16 digits, matches to ~1e-15 relative
//   * test/meshperftest.nim:8-44  triangle at z=-5 => t = 5
//   * test/geomtest.cpp:51-78     the reference's own C++ AABB slab test, compiled
//                                 from where it lies into oracle/_ref (Makefile)
// Everything else (glm's inverse/normalize/summation order, -ffast-math
// reassociation) is PARITY UNPINNED at the ulp level; choices are documented at
// each function.  Arithmetic here is IEEE float64, no contraction
// (-ffp-contract=off), explicit left-to-right summation.
//
// All citations are file:line in the nim-raytracer tree.

#include "../include/nrt.h"

#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

constexpr double kInf = std::numeric_limits<double>::infinity();
constexpr double kNegInf = -std::numeric_limits<double>::infinity();
constexpr double kPi = 3.14159265358979323846;  // Nim math.PI

// Nim `min`/`max` on float64 compile to ($1 <= $2 ? $1 : $2) / ($1 >= $2 ? $1 : $2)
// (SURVEY.md §3.4); operand order matters for NaN (geom.nim:88-89 relies on it).
inline double nim_min(double a, double b) { return (a <= b) ? a : b; }
inline double nim_max(double a, double b) { return (a >= b) ? a : b; }

struct Vec3 { double x, y, z; };
struct Vec4 { double x, y, z, w; };
struct Mat4 { double m[16]; };  // m[col*4+row], GLM column vectors

inline Vec4 vec(double x, double y, double z) { return {x, y, z, 0.0}; }     // geom.nim:11
inline Vec4 point(double x, double y, double z) { return {x, y, z, 1.0}; }   // geom.nim:14

inline Vec4 add(Vec4 a, Vec4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline Vec4 sub(Vec4 a, Vec4 b) { return {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
inline Vec4 scale(Vec4 a, double s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
inline Vec3 add(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 scale(Vec3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline Vec3 mul(Vec3 a, Vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline Vec3 divs(Vec3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }

// glm dot: left-to-right sum over all components (unpinned: summation order).
inline double dot(Vec4 a, Vec4 b) { return ((a.x * b.x + a.y * b.y) + a.z * b.z) + a.w * b.w; }
inline double dot(Vec3 a, Vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }

// glm normalize = v * inversesqrt(dot(v,v)), inversesqrt(x) = 1/sqrt(x)
// (GLM func_geometric; nim-glm-fork source absent => unpinned).
inline Vec4 normalize(Vec4 v) { double s = 1.0 / std::sqrt(dot(v, v)); return scale(v, s); }
inline Vec3 normalize(Vec3 v) { double s = 1.0 / std::sqrt(dot(v, v)); return scale(v, s); }

// glm mat * vec: column combination, left-to-right (unpinned: summation order).
inline Vec4 mul(const Mat4& M, Vec4 v) {
  const double* m = M.m;
  return {
      ((m[0] * v.x + m[4] * v.y) + m[8] * v.z) + m[12] * v.w,
      ((m[1] * v.x + m[5] * v.y) + m[9] * v.z) + m[13] * v.w,
      ((m[2] * v.x + m[6] * v.y) + m[10] * v.z) + m[14] * v.w,
      ((m[3] * v.x + m[7] * v.y) + m[11] * v.z) + m[15] * v.w,
  };
}

// utils/mathutils.nim:12-18
inline double sign(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : 0.0); }

// ------------------------------------------------------------------ Ray ----
// geom.nim:32-48.  `depth` is accepted by initRay but NOT stored (the depth
// bug): Ray.depth stays 0.  We keep a separate `bounce`
// counter only to implement the INTENDED mode and the safety cap.
struct Ray {
  Vec4 orig, dir;
  Vec3 invDir;
  int sign[3];
  int64_t triangleHit;  // face index, -1 == nil
};

inline Ray initRay(Vec4 orig, Vec4 dir) {  // geom.nim:41-48
  Ray r;
  r.orig = orig;
  r.dir = dir;
  r.invDir = {1 / dir.x, 1 / dir.y, 1 / dir.z};
  r.sign[0] = r.invDir.x < 0;
  r.sign[1] = r.invDir.y < 0;
  r.sign[2] = r.invDir.z < 0;
  r.triangleHit = -1;
  return r;
}

// ----------------------------------------------------------------- AABB ----
struct AABB { Vec4 bounds[2]; };  // geom.nim:63-66

// geom.nim:76-96 (Ize, "Robust BVH Ray Traversal")
inline double aabb_intersect(const AABB& b, const Ray& r) {
  double tmin = kNegInf, tmax = kInf;
  const double txmin = (b.bounds[r.sign[0]].x - r.orig.x) * r.invDir.x;
  const double txmax = (b.bounds[1 - r.sign[0]].x - r.orig.x) * r.invDir.x;
  const double tymin = (b.bounds[r.sign[1]].y - r.orig.y) * r.invDir.y;
  const double tymax = (b.bounds[1 - r.sign[1]].y - r.orig.y) * r.invDir.y;
  const double tzmin = (b.bounds[r.sign[2]].z - r.orig.z) * r.invDir.z;
  const double tzmax = (b.bounds[1 - r.sign[2]].z - r.orig.z) * r.invDir.z;
  tmin = nim_max(tzmin, nim_max(tymin, nim_max(txmin, tmin)));
  tmax = nim_min(tzmax, nim_min(tymax, nim_min(txmax, tmax)));
  tmax *= 1.0000000000000004;
  return (tmin <= tmax) ? tmin : kNegInf;
}

// ---------------------------------------------------------- intersections --
// geom.nim:215-237.  NOTE `/ 2*a` parses as (x/2)*a and the result is
// min(t1,t2) even when negative.
inline void solve_quadratic(double a, double b, double c, double delta, double& t1, double& t2) {
  t1 = (-b - sign(b) * std::sqrt(delta)) / 2 * a;  // utils/mathutils.nim:23-28
  t2 = c / (a * t1);
}

inline double sphere_intersect(double radius, const Ray& r) {
  const double a = r.dir.x * r.dir.x + r.dir.y * r.dir.y + r.dir.z * r.dir.z;
  const double b = 2 * (r.dir.x * r.orig.x + r.dir.y * r.orig.y + r.dir.z * r.orig.z);
  const double c = r.orig.x * r.orig.x + r.orig.y * r.orig.y + r.orig.z * r.orig.z - radius * radius;
  const double delta = b * b - 4 * a * c;
  if (delta >= 0.0) {
    double t1, t2;
    solve_quadratic(a, b, c, delta, t1, t2);
    return nim_min(t1, t2);
  }
  return kNegInf;
}

// geom.nim:240-248
inline double plane_intersect(const Ray& r) {
  const Vec4 n = vec(0.0, 1.0, 0.0);
  const double denom = dot(n, r.dir);
  if (std::fabs(denom) > 1e-6) return -dot(r.orig, n) / denom;
  return kNegInf;
}

// geom.nim:283-336 (rayTriangleIntersectFast — the scalar Möller–Trumbore the mesh uses)
inline double ray_triangle_fast(const Ray& r, const double* v0, const double* v1, const double* v2) {
  const double v0v1x = v1[0] - v0[0], v0v1y = v1[1] - v0[1], v0v1z = v1[2] - v0[2];
  const double v0v2x = v2[0] - v0[0], v0v2y = v2[1] - v0[1], v0v2z = v2[2] - v0[2];
  const double pvecx = r.dir.y * v0v2z - r.dir.z * v0v2y;
  const double pvecy = r.dir.z * v0v2x - r.dir.x * v0v2z;
  const double pvecz = r.dir.x * v0v2y - r.dir.y * v0v2x;
  const double det = v0v1x * pvecx + v0v1y * pvecy + v0v1z * pvecz;
  if (det < 0.000001) return kNegInf;
  const double invDet = 1 / det;
  const double tvecx = r.orig.x - v0[0], tvecy = r.orig.y - v0[1], tvecz = r.orig.z - v0[2];
  const double u = (tvecx * pvecx + tvecy * pvecy + tvecz * pvecz) * invDet;
  if (u < 0 || u > 1) return kNegInf;
  const double qvecx = tvecy * v0v1z - tvecz * v0v1y;
  const double qvecy = tvecz * v0v1x - tvecx * v0v1z;
  const double qvecz = tvecx * v0v1y - tvecy * v0v1x;
  const double v = (r.dir.x * qvecx + r.dir.y * qvecy + r.dir.z * qvecz) * invDet;
  if (v < 0 || u + v > 1) return kNegInf;
  return (v0v2x * qvecx + v0v2y * qvecy + v0v2z * qvecz) * invDet;
}

// ---------------------------------------------------------------- scene ----
struct Mesh {
  const double* vertices;   // nverts*4
  const double* normals;    // nnormals*4
  const int64_t* vertexIdx; // nfaces*3
  const int64_t* normalIdx; // nfaces*3
  int64_t nverts, nfaces;
  AABB aabb;
};

// geom.nim:175-188
AABB calcAABB(const double* verts, int64_t n) {
  Vec4 vmin = point(kInf, kInf, kInf), vmax = point(kNegInf, kNegInf, kNegInf);
  for (int64_t i = 0; i < n; ++i) {
    const double* v = verts + 4 * i;
    if (v[0] < vmin.x) vmin.x = v[0];
    if (v[1] < vmin.y) vmin.y = v[1];
    if (v[2] < vmin.z) vmin.z = v[2];
    if (v[0] > vmax.x) vmax.x = v[0];
    if (v[1] > vmax.y) vmax.y = v[1];
    if (v[2] > vmax.z) vmax.z = v[2];
  }
  return AABB{{vmin, vmax}};
}

// geom.nim:339-358
double mesh_intersect(const Mesh& m, Ray& r) {
  if (aabb_intersect(m.aabb, r) < 0) return kNegInf;
  double tMin = kInf;
  for (int64_t f = 0; f < m.nfaces; ++f) {
    const double* v0 = m.vertices + 4 * m.vertexIdx[3 * f + 0];
    const double* v1 = m.vertices + 4 * m.vertexIdx[3 * f + 1];
    const double* v2 = m.vertices + 4 * m.vertexIdx[3 * f + 2];
    const double tHit = ray_triangle_fast(r, v0, v1, v2);
    if (tHit >= 0 && tHit < tMin) {
      tMin = tHit;
      r.triangleHit = f;
    }
  }
  return tMin;
}

struct Scene {
  const nrt_scene_desc* d;
  std::vector<Mesh> meshes;
  std::vector<Mat4> o2w, w2o;
  std::vector<AABB> boxes;
  Mat4 c2w;
  double tanHalfFov;  // f of renderer.nim:38
};

struct Stats { int64_t primary = 0, tests = 0, hits = 0, rays = 0, capped = 0; };

// geom.nim:215-252,339 dispatched on kind (Nim `method`)
inline double intersect(const Scene& sc, int i, Ray& r) {
  const nrt_object& o = sc.d->objects[i];
  switch (o.kind) {
    case NRT_GEOM_SPHERE: return sphere_intersect(o.radius, r);
    case NRT_GEOM_PLANE: return plane_intersect(r);
    case NRT_GEOM_BOX: return aabb_intersect(sc.boxes[i], r);
    case NRT_GEOM_MESH: return mesh_intersect(sc.meshes[o.mesh], r);
    default: return kNegInf;  // geom.nim:213 base method
  }
}

// geom.nim:361-379
inline Vec4 geom_normal(const Scene& sc, int i, Vec4 hit) {
  const nrt_object& o = sc.d->objects[i];
  switch (o.kind) {
    case NRT_GEOM_SPHERE: return normalize(vec(hit.x, hit.y, hit.z));
    case NRT_GEOM_PLANE: return vec(0.0, 1.0, 0.0);
    case NRT_GEOM_BOX: {
      const AABB& b = sc.boxes[i];
      const Vec4 c = scale(add(b.bounds[0], b.bounds[1]), 0.5);
      const Vec4 p = sub(hit, c);
      const Vec4 d = scale(sub(b.bounds[0], b.bounds[1]), 0.5);
      const double bias = 1.000001;
      // `.int` truncates toward zero (Nim float->int conversion)
      return normalize(vec(std::trunc(p.x / std::fabs(d.x) * bias),
                           std::trunc(p.y / std::fabs(d.y) * bias),
                           std::trunc(p.z / std::fabs(d.z) * bias)));
    }
    default: return {0, 0, 0, 0};  // geom.nim:361-362
  }
}

// renderer.nim:47-67
struct Hit { int obj; double t; };
Hit trace(const Scene& sc, Ray& ray, double tNear, Stats& st) {
  double tmin = tNear;
  int objmin = -1;
  st.rays++;
  for (int i = 0; i < sc.d->nobjects; ++i) {
    Ray rayO = initRay(mul(sc.w2o[i], ray.orig), mul(sc.w2o[i], ray.dir));
    const double tHit = intersect(sc, i, rayO);
    st.tests++;
    if (tHit >= 0 && tHit < tmin) {
      tmin = tHit;
      objmin = i;
      ray.triangleHit = rayO.triangleHit;
      st.hits++;
    }
  }
  return {objmin, tmin};
}

struct ShadingInfo { Vec4 lightDir; Vec3 lightIntensity; double lightDistance; };

// light.nim:46-62
inline ShadingInfo getShadingInfo(const nrt_light& l, Vec4 p) {
  const Vec3 color{l.color[0], l.color[1], l.color[2]};
  if (l.kind == NRT_LIGHT_DISTANT) {
    return {Vec4{l.dir[0], l.dir[1], l.dir[2], l.dir[3]}, scale(color, l.intensity), kInf};
  }
  Vec4 lightDir = sub(p, Vec4{l.pos[0], l.pos[1], l.pos[2], l.pos[3]});
  const double r2 = dot(lightDir, lightDir);  // length2
  lightDir = normalize(lightDir);
  return {lightDir, divs(scale(color, l.intensity), (4 * kPi * r2)), std::sqrt(r2)};
}

// shader.nim:12-17
inline Vec3 shadeDiffuse(const nrt_object& o, const ShadingInfo& si, Vec4 hitNormal) {
  const Vec3 albedo{o.albedo[0], o.albedo[1], o.albedo[2]};
  const double c = nim_max(0.0, dot(hitNormal, scale(si.lightDir, -1.0)));
  return scale(mul(divs(albedo, kPi), si.lightIntensity), c);
}

struct Opts { const nrt_options* o; int cap; };

// renderer.nim:71-127.  `bounce` = number of reflections on the path so far.
Vec3 shade(const Scene& sc, const Opts& op, Ray& ray, int objHit, double tHit, int bounce, Stats& st) {
  const Vec3 bg{sc.d->bg_color[0], sc.d->bg_color[1], sc.d->bg_color[2]};
  if (objHit < 0) return bg;
  const nrt_object& obj = sc.d->objects[objHit];
  const Vec4 hitW = add(ray.orig, scale(ray.dir, tHit));
  const Vec4 hitO = mul(sc.w2o[objHit], hitW);
  Vec4 hitNormal;
  if (ray.triangleHit < 0) {
    hitNormal = mul(sc.o2w[objHit], geom_normal(sc, objHit, hitO));
  } else {
    const Mesh& m = sc.meshes[obj.mesh];
    const double* n = m.normals + 4 * m.normalIdx[3 * ray.triangleHit + 0];
    hitNormal = mul(sc.o2w[objHit], Vec4{n[0], n[1], n[2], n[3]});
  }
  Vec3 result{0.0, 0.0, 0.0};
  for (int l = 0; l < sc.d->nlights; ++l) {
    const ShadingInfo si = getShadingInfo(sc.d->lights[l], hitW);
    const Vec4 lightDir = scale(si.lightDir, -1.0);
    Ray shadowRay = initRay(add(hitW, scale(hitNormal, op.o->bias)), lightDir);
    const Hit sh = trace(sc, shadowRay, si.lightDistance, st);
    if (sh.obj < 0) result = add(result, shadeDiffuse(obj, si, hitNormal));
  }
  const double reflection = obj.reflection;
  // renderer.nim:108: `ray.depth <= opts.maxRayDepth`; REFBUG: ray.depth == 0 always.
  const int depth = (op.o->depth_mode == NRT_DEPTH_INTENDED) ? (1 + bounce) : 0;
  if (reflection > 0.0 && depth <= op.o->max_ray_depth) {
    if (bounce >= op.cap) {  // safety cap (not in the reference): stop here, count it
      st.capped++;
      return result;
    }
    const Vec4 i = ray.dir, n = hitNormal;
    const Vec4 r = sub(i, scale(n, 2 * dot(n, i)));
    Ray rayR = initRay(add(hitW, scale(r, op.o->bias)), r);
    const Hit hr = trace(sc, rayR, kInf, st);
    Vec3 reflColor = (hr.obj >= 0) ? shade(sc, op, rayR, hr.obj, hr.t, bounce + 1, st) : bg;
    result = add(scale(result, 1.0 - reflection), scale(reflColor, reflection));
  }
  return result;
}

// renderer.nim:31-44
inline Ray castPrimaryRay(const Scene& sc, int w, int h, double x, double y) {
  const double r = double(w) / double(h);
  const double f = sc.tanHalfFov;
  const double cx = ((2 * x * r) / double(w) - r) * f;
  const double cy = (1 - 2 * y / double(h)) * f;
  return initRay(mul(sc.c2w, point(0.0, 0.0, 0.0)), mul(sc.c2w, normalize(vec(cx, cy, -1))));
}

// counter-based RNG shared with the GPU path for the jittered AA kinds
// (the reference uses a time-seeded global RNG: no reproducible seed exists).
inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
struct PixelRng {
  uint64_t key, ctr;
  double next() {  // uniform [0,1)
    const uint64_t v = splitmix64(key + 0xD1342543DE82EF95ull * (ctr++));
    return double(v >> 11) * (1.0 / 9007199254740992.0);
  }
  double random(double mx) { return next() * mx; }
};

// sampling.nim:5-113.  out has m*n entries (x,y pairs).
void make_samples(int kind, int m, int n, PixelRng& rng, std::vector<double>& p) {
  p.assign(size_t(2) * m * n, 0.0);
  auto X = [&](int i) -> double& { return p[2 * i]; };
  auto Y = [&](int i) -> double& { return p[2 * i + 1]; };
  if (kind == NRT_AA_GRID) {  // sampling.nim:5-18 (yoffs uses xs, sic)
    const double xs = 1.0 / double(n), ys = 1.0 / double(m);
    const double xoffs = xs * 0.5, yoffs = xs * 0.5;
    for (int j = 0; j < m; ++j)
      for (int i = 0; i < n; ++i) { X(j * m + i) = double(i) * xs + xoffs; Y(j * m + i) = double(j) * ys + yoffs; }
  } else if (kind == NRT_AA_JITTERED) {  // sampling.nim:21-32
    const double xs = 1.0 / double(n), ys = 1.0 / double(m);
    for (int j = 0; j < m; ++j)
      for (int i = 0; i < n; ++i) {
        const double rx = rng.random(xs), ry = rng.random(ys);
        X(j * m + i) = double(i) * xs + rx;
        Y(j * m + i) = double(j) * ys + ry;
      }
  } else {  // multiJittered / correlatedMultiJittered(n, m): sampling.nim:38-113
    const int nn = m, mm = n;  // called as (m, m): parameter names are (n, m)
    const double xs = 1.0 / double(nn), ys = 1.0 / double(mm);
    for (int j = 0; j < nn; ++j)
      for (int i = 0; i < mm; ++i) {
        const double jj = j, ii = i;
        const double r1 = rng.random(1.0), r2 = rng.random(1.0);
        X(j * mm + i) = (ii + (jj + r1) * xs) * ys;
        Y(j * mm + i) = (jj + (ii + r2) * ys) * xs;
      }
    if (kind == NRT_AA_MULTI_JITTERED) {
      for (int j = 0; j < nn; ++j)
        for (int i = 0; i < mm; ++i) {
          const int k = j + int(rng.random(1.0) * double(nn - j));
          std::swap(X(j * mm + i), X(k * mm + i));
        }
      for (int i = 0; i < mm; ++i)
        for (int j = 0; j < nn; ++j) {
          const int k = i + int(rng.random(1.0) * double(mm - i));
          std::swap(Y(j * mm + i), Y(j * mm + k));
        }
    } else {
      for (int j = 0; j < nn; ++j) {
        const int k = j + int(rng.random(1.0) * double(nn - j));
        for (int i = 0; i < mm; ++i) std::swap(X(j * mm + i), X(k * mm + i));
      }
      for (int i = 0; i < mm; ++i) {
        const int k = i + int(rng.random(1.0) * double(mm - i));
        for (int j = 0; j < nn; ++j) std::swap(Y(j * mm + i), Y(j * mm + k));
      }
    }
  }
}

struct Target { float* fb; const nrt_aov* aov; };

inline void store_pixel(const Target& tg, int w, int x, int y, Vec3 c) {  // framebuf.nim:22-28
  float* p = tg.fb + (size_t(y) * w + x) * 3;
  p[0] = float(c.x); p[1] = float(c.y); p[2] = float(c.z);
}

// renderer.nim:162-211
void renderLine(const Scene& sc, const Opts& op, const Target& tg, int y, int step, int maxStep, Stats& stats) {
  const nrt_options& o = *op.o;
  const int w = o.width, h = o.height;
  std::vector<double> samples;
  for (int x = 0; x < w; x += step) {
    if (step < maxStep) {
      const int mask = step * 2 - 1;
      if (((x & mask) == 0) && ((y & mask) == 0)) continue;
    }
    Vec3 color{0, 0, 0};
    int aovObj = -1; int64_t aovTri = -1; double aovT = kInf;
    if (o.aa_kind == NRT_AA_NONE) {  // calcPixelNoSampling, renderer.nim:132-141
      Ray ray = castPrimaryRay(sc, w, h, double(x), double(y));
      stats.primary++;
      const Hit hit = trace(sc, ray, kInf, stats);
      aovObj = hit.obj; aovTri = hit.obj >= 0 ? ray.triangleHit : -1; aovT = hit.t;
      color = shade(sc, op, ray, hit.obj, hit.t, 0, stats);
    } else {  // calcPixel, renderer.nim:144-159
      PixelRng rng{splitmix64(o.seed ^ (0x632BE59BD9B4E019ull * (uint64_t(y) * uint64_t(w) + uint64_t(x) + 1))), 0};
      make_samples(o.aa_kind, o.grid_size, o.grid_size, rng, samples);
      const int ns = int(samples.size() / 2);
      for (int i = 0; i < ns; ++i) {
        Ray ray = castPrimaryRay(sc, w, h, double(x) + samples[2 * i], double(y) + samples[2 * i + 1]);
        stats.primary++;
        const Hit hit = trace(sc, ray, kInf, stats);
        if (i == 0) { aovObj = hit.obj; aovTri = hit.obj >= 0 ? ray.triangleHit : -1; aovT = hit.t; }
        color = add(color, shade(sc, op, ray, hit.obj, hit.t, 0, stats));
      }
      color = scale(color, 1 / double(ns));
    }
    if (tg.aov) {
      const size_t pi = size_t(y) * w + x;
      if (tg.aov->obj_id) tg.aov->obj_id[pi] = aovObj;
      if (tg.aov->tri_id) tg.aov->tri_id[pi] = int32_t(aovTri);
      if (tg.aov->t_hit) tg.aov->t_hit[pi] = aovT;
    }
    if (step > 1) {
      for (int i = x; i < std::min(x + step, w); ++i)
        for (int j = y; j < std::min(y + step, h); ++j) store_pixel(tg, w, i, j, color);
    } else {
      store_pixel(tg, w, x, y, color);
    }
  }
}

bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

Scene build_scene(const nrt_scene_desc* d) {
  Scene sc;
  sc.d = d;
  sc.meshes.resize(d->nmeshes);
  for (int i = 0; i < d->nmeshes; ++i) {
    const nrt_mesh& m = d->meshes[i];
    sc.meshes[i] = Mesh{m.vertices, m.normals, m.vertex_idx, m.normal_idx, m.nverts, m.nfaces,
                        calcAABB(m.vertices, m.nverts)};
  }
  sc.o2w.resize(d->nobjects); sc.w2o.resize(d->nobjects); sc.boxes.resize(d->nobjects);
  for (int i = 0; i < d->nobjects; ++i) {
    std::memcpy(sc.o2w[i].m, d->objects[i].object_to_world, sizeof(double) * 16);
    std::memcpy(sc.w2o[i].m, d->objects[i].world_to_object, sizeof(double) * 16);
    const double* a = d->objects[i].vmin; const double* b = d->objects[i].vmax;
    sc.boxes[i] = AABB{{Vec4{a[0], a[1], a[2], a[3]}, Vec4{b[0], b[1], b[2], b[3]}}};
  }
  std::memcpy(sc.c2w.m, d->camera_to_world, sizeof(double) * 16);
  // f = tan(degToRad(fov)/2), renderer.nim:38; Nim degToRad(d) = d * (PI/180)
  sc.tanHalfFov = std::tan((d->fov * (kPi / 180.0)) / 2);
  return sc;
}

}  // namespace

extern "C" {

// Whole-frame driver mirroring src/raytracer.nim:61-109: one scanline per work
// item, pulled from a shared counter by `nthreads` workers
// (0 => std::thread::hardware_concurrency(), like countProcessors(),
// concurrency/workerpool.nim:172).
int oracle_render(const nrt_scene_desc* desc, const nrt_options* opts, int y0, int y1, int step, int max_step,
                  float* fb, nrt_stats* stats, const nrt_aov* aov, int nthreads) {
  if (!desc || !opts || !fb) return NRT_ERR_INVALID;
  if (!is_pow2(step) || !is_pow2(max_step) || max_step < step) return NRT_ERR_UNSUPPORTED;
  Scene sc = build_scene(desc);
  Opts op{opts, opts->bounce_cap > 0 ? opts->bounce_cap : 64};
  Target tg{fb, aov};
  if (nthreads <= 0) nthreads = int(std::thread::hardware_concurrency());
  if (nthreads <= 0) nthreads = 1;
  std::vector<int> lines;
  for (int y = std::max(0, y0); y < std::min(y1, opts->height); ++y)
    if ((y - y0) % step == 0) lines.push_back(y);
  std::atomic<size_t> next{0};
  std::vector<Stats> per(nthreads);
  auto worker = [&](int tid) {
    for (;;) {
      const size_t i = next.fetch_add(1);
      if (i >= lines.size()) break;
      renderLine(sc, op, tg, lines[i], step, max_step, per[tid]);
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nthreads; ++t) th.emplace_back(worker, t);
  worker(0);
  for (auto& t : th) t.join();
  if (stats) {
    Stats s;
    for (auto& p : per) { s.primary += p.primary; s.tests += p.tests; s.hits += p.hits; s.rays += p.rays; s.capped += p.capped; }
    stats->num_primary_rays = s.primary;
    stats->num_intersection_tests = s.tests;
    stats->num_intersection_hits = s.hits;
    stats->num_rays = s.rays;
    stats->num_capped_samples = s.capped;
  }
  return NRT_OK;
}

// Same worker-pool loop over an explicit list of scanlines (step = maxStep = 1):
// used by bench.py to time a bounded, uniformly spread sample of a frame.
int oracle_render_rows(const nrt_scene_desc* desc, const nrt_options* opts, const int* rows, int nrows, float* fb,
                       nrt_stats* stats, int nthreads) {
  if (!desc || !opts || !fb || (nrows > 0 && !rows)) return NRT_ERR_INVALID;
  Scene sc = build_scene(desc);
  Opts op{opts, opts->bounce_cap > 0 ? opts->bounce_cap : 64};
  Target tg{fb, nullptr};
  if (nthreads <= 0) nthreads = int(std::thread::hardware_concurrency());
  if (nthreads <= 0) nthreads = 1;
  std::atomic<int> next{0};
  std::vector<Stats> per(nthreads);
  auto worker = [&](int tid) {
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= nrows) break;
      if (rows[i] >= 0 && rows[i] < opts->height) renderLine(sc, op, tg, rows[i], 1, 1, per[tid]);
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nthreads; ++t) th.emplace_back(worker, t);
  worker(0);
  for (auto& t : th) t.join();
  if (stats) {
    Stats s;
    for (auto& p : per) { s.primary += p.primary; s.tests += p.tests; s.hits += p.hits; s.rays += p.rays; s.capped += p.capped; }
    stats->num_primary_rays = s.primary;
    stats->num_intersection_tests = s.tests;
    stats->num_intersection_hits = s.hits;
    stats->num_rays = s.rays;
    stats->num_capped_samples = s.capped;
  }
  return NRT_OK;
}

int oracle_hardware_threads(void) { return int(std::thread::hardware_concurrency()); }

// ---- unit entry points used by the golden-vector tests ---------------------

// utils/mathutils.nim:20-28
void oracle_solve_quadratic(double a, double b, double c, double* t1, double* t2) {
  const double delta = b * b - 4 * a * c;
  solve_quadratic(a, b, c, delta, *t1, *t2);
}

// geom.nim:41-48 + 76-96; bounds = vmin xyz, vmax xyz
double oracle_aabb_intersect(const double* vmin, const double* vmax, const double* orig, const double* dir) {
  AABB b{{point(vmin[0], vmin[1], vmin[2]), point(vmax[0], vmax[1], vmax[2])}};
  Ray r = initRay(point(orig[0], orig[1], orig[2]), vec(dir[0], dir[1], dir[2]));
  return aabb_intersect(b, r);
}

double oracle_sphere_intersect(double radius, const double* orig, const double* dir) {
  Ray r = initRay(point(orig[0], orig[1], orig[2]), vec(dir[0], dir[1], dir[2]));
  return sphere_intersect(radius, r);
}

double oracle_plane_intersect(const double* orig, const double* dir) {
  Ray r = initRay(point(orig[0], orig[1], orig[2]), vec(dir[0], dir[1], dir[2]));
  return plane_intersect(r);
}

// geom.nim:283-336; v0,v1,v2 are xyz triples
double oracle_ray_triangle(const double* orig, const double* dir, const double* v0, const double* v1, const double* v2) {
  Ray r = initRay(point(orig[0], orig[1], orig[2]), vec(dir[0], dir[1], dir[2]));
  return ray_triangle_fast(r, v0, v1, v2);
}

// renderer.nim:31-44; c2w as m[col*4+row]; out = orig xyzw, dir xyzw
void oracle_cast_primary_ray(int w, int h, double x, double y, double fov, const double* c2w, double* out) {
  Scene sc;
  std::memcpy(sc.c2w.m, c2w, sizeof(double) * 16);
  sc.tanHalfFov = std::tan((fov * (kPi / 180.0)) / 2);
  Ray r = castPrimaryRay(sc, w, h, x, y);
  out[0] = r.orig.x; out[1] = r.orig.y; out[2] = r.orig.z; out[3] = r.orig.w;
  out[4] = r.dir.x; out[5] = r.dir.y; out[6] = r.dir.z; out[7] = r.dir.w;
}

// sampling.nim patterns for (m, m); out has 2*m*m doubles
void oracle_samples(int kind, int m, uint64_t seed, int width, int x, int y, double* out) {
  PixelRng rng{splitmix64(seed ^ (0x632BE59BD9B4E019ull * (uint64_t(y) * uint64_t(width) + uint64_t(x) + 1))), 0};
  std::vector<double> p;
  make_samples(kind, m, m, rng, p);
  std::memcpy(out, p.data(), p.size() * sizeof(double));
}

// ---- output stage --------------------------------------------------------------
// utils/color.nim:17-22 linearToSRGB(v: float32): `let a = 0.055` is a float64, the literals 12.92 and 1/2.4 take
// the float32 type of `v` (pow -> powf), the affine part is evaluated in float64 and narrowed to the float32
// result.  (Nim's float32 / float64 conversions are implicit both ways; which overload a mixed expression picks
// cannot be checked without a Nim compiler: this reading is the frozen one, "parity unpinned" like glm.)
static float linearToSRGB(float v) {
  const double a = 0.055;
  if (v <= 0.0031308f) return 12.92f * v;
  return float((1 + a) * double(powf(v, float(1 / 2.4))) - a);
}
// utils/framebuf.nim:74-78 outvalue: clamp(v, 0.0, 1.0) -> linearToSRGB -> Natural(round(c * maxval)), maxval =
// float32(2^bits - 1).  Nim's clamp lets NaN through and Natural(NaN) is undefined: defined as 0 here (and in the product).
static uint32_t outvalue(float v, int bits, int srgb) {
  const float maxval = float((1u << bits) - 1u);
  if (v != v) return 0;
  float c = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
  if (srgb) c = linearToSRGB(c);
  return uint32_t(roundf(c * maxval));
}
// writePpm's sample stream (framebuf.nim:80-89): n float32 components -> n uint8 (bits <= 8) or n big-endian
// uint16 (framebuf.nim:67-71)
void oracle_outvalues(const float* fb, int64_t n, int bits, int srgb, unsigned char* out) {
  for (int64_t i = 0; i < n; ++i) {
    const uint32_t v = outvalue(fb[i], bits, srgb);
    if (bits <= 8) out[i] = (unsigned char)v;
    else { out[2 * i] = (unsigned char)(v >> 8); out[2 * i + 1] = (unsigned char)(v & 0xFFu); }
  }
}
// utils/image.nim:45-54 ImageRGBA.copyFrom: round(v * 0xff).uint8 per channel + alpha.  The float -> uint8
// conversion is undefined outside 0..255 in the reference: clamped here (and in the product); NaN -> 0.
void oracle_rgba8(const float* fb, int64_t npix, unsigned char alpha, unsigned char* out) {
  for (int64_t p = 0; p < npix; ++p) {
    for (int k = 0; k < 3; ++k) {
      const float r = roundf(fb[3 * p + k] * 255.0f);
      out[4 * p + k] = (unsigned char)(r != r ? 0.0f : (r < 0.0f ? 0.0f : (r > 255.0f ? 255.0f : r)));
    }
    out[4 * p + 3] = alpha;
  }
}

}  // extern "C"
