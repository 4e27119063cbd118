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

// Keeps a loaded value materialised where it is written (the compiler would otherwise sink the load
// below the branch that decides whether it is needed, serialising two memory round trips).
#if defined(__CUDA_ARCH__)
#define NRT_KEEP_D(x) asm volatile("" : "+d"(x))
#else
#define NRT_KEEP_D(x) ((void)0)
#endif
// Compiler-level fence (no instruction): loads written after it are issued after it — neither hoisted
// above the code before it nor served from a value loaded earlier and kept alive in registers.
#if defined(__CUDA_ARCH__) && !defined(NRT_NO_FENCE)
#define NRT_COMPILER_FENCE() asm volatile("" ::: "memory")
#else
#define NRT_COMPILER_FENCE() ((void)0)
#endif

// Requests the line of `p` into L2 without a register or a scoreboard entry (no-op on the host).
#if defined(__CUDA_ARCH__)
#define NRT_PREFETCH_L2(p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p))
#else
#define NRT_PREFETCH_L2(p) ((void)0)
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
  int32_t xlate_only;  // worldToObject is exactly [I | t] (all reference scenes): see toObject()
  double o2w[16];
  double w2o[16];
  double radius;
  double bmin[4];      // Box.aabb.vmin, or the mesh AABB for MESH objects
  double bmax[4];
  double albedo[3];
  double reflection;
  double albedo_pi[3]; // albedo / PI (shader.nim:15), the same IEEE division done once at scene build
  double plane_nw[4];  // objectToWorld * (0,1,0,0): the world-space normal of a Plane (geom.nim:367-368, renderer.nim:85), done once
};

// Compact per-object record read by the in-order object scan of trace(): 48 bytes, three
// 16-byte vector loads (the full DObject, 448 B over four cache lines, is only needed for
// non-translation transforms, boxes and shading).
struct alignas(16) CObj {
  int32_t kind, mesh_obj, xlate_only, _pad;
  double t[3];     // translation column of worldToObject
  double radius;
};
NRT_HD CObj loadCObj(const CObj* p) {
#if defined(__CUDA_ARCH__)
  CObj c;
  const int4 h = __ldg(reinterpret_cast<const int4*>(p));
  const double2 a = __ldg(reinterpret_cast<const double2*>(p) + 1);
  const double2 b = __ldg(reinterpret_cast<const double2*>(p) + 2);
  c.kind = h.x; c.mesh_obj = h.y; c.xlate_only = h.z; c._pad = h.w;
  c.t[0] = a.x; c.t[1] = a.y; c.t[2] = b.x; c.radius = b.y;
  return c;
#else
  return *p;
#endif
}

// float32 mirror for the object scan's first look at an object (16 bytes, one vector load, the
// same address for every lane of a warp): see certainMissF() / planeMissF().
//   sphere the float32 test applies to:  (tx, ty, tz) = translation column of worldToObject rounded to
//                                        float32, r2m = radius^2 + object part of the margin (rounded up)
//   anything else: r2m = +Inf and tx carries a tag (bit pattern of an int):
//     COF_PLANE  plane with an exact [I | t] worldToObject: ty = t.y, tz = object part of the margin
//     COF_MESH   triangle mesh: ty = bit pattern of its mesh-object index (its gate code decides)
//     COF_SLOW   no float32 shortcut (boxes, general matrices, non-finite values)
enum { COF_SLOW = 0, COF_PLANE = 1, COF_MESH = 2 };
struct alignas(16) CObjF {
  float tx, ty, tz;
  float r2m;
};
NRT_HD CObjF loadCObjF(const CObjF* p) {
#if defined(__CUDA_ARCH__)
  CObjF c;
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  c.tx = a.x; c.ty = a.y; c.tz = a.z; c.r2m = a.w;
  return c;
#else
  return *p;
#endif
}

static constexpr int kClusterSize = 16;    // members per level-2 cluster, level-2 clusters per level-1 cluster
static constexpr int kClusterMin = 64;     // fewer clusterable spheres than this: flat scan
static constexpr int kSurvivorCap = 24;    // per-ray list of objects that need the float64 evaluation

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
  float rb2f;               // squared half diagonal of the AABB, rounded up (boxCertainMiss); Inf disables the pre-test
  float _padf;
  float* recs;              // GENERAL-mode full filter records (nrt_filter.h), pair-interleaved
  float* hot;               // GENERAL-mode hot (bounding sphere) records, pair-interleaved
  float* bounds;            // GENERAL-mode chunk bounds (one hot-format record per 256 records)
  uint32_t* order;          // faces in Morton order of their centroids: record r <-> face order[r]
};

// float32 first look at the AABB gate of a mesh object (geom.nim:340) from the WORLD-space ray, for objects whose
// worldToObject is exactly [I | t]: the bounding sphere of the mesh box moved to world space, c = centre - t
// (object-space ray origin - centre = world origin - c), in the format and with the margins of a sphere's CObjF:
//   r2m = rb^2 (1 + 2e-6) + 2e-7 |c|_inf^2 (rounded up),   mm = 2e-7 |c|_inf^2 (rounded up)
// valid == 0: no float32 shortcut for this object (general matrix, non-finite values).
struct alignas(16) MeshGateF { float cx, cy, cz, r2m; float mm, valid, pad0, pad1; };
// Light-space grid of the spheres, one per DistantLight (scenes with sphere clusters).  Every shadow ray of a distant
// light has the same direction u (renderer.nim:99, light.nim), so a sphere can only be hit by the rays whose origin
// projects — along u, onto the plane (e1, e2) — into the sphere's projected circle.  The plane is cut into G x G cells;
// a cell lists, ASCENDING (the in-order scan of trace() decides ties and Stats), the spheres whose circle, inflated by
// every error a ray's float32 projection can carry, touches it.  A shadow ray reads ONE cell.
//   circle of sphere i (centre c, radius r):  radius r (1 + 1e-6) + 4e-7 |c|_1 + margin + 1e-4 h   (h = cell edge)
//     1e-6      e1, e2 are float32 vectors: unit and perpendicular to u only to ~1e-7
//     4e-7|c|_1 the same tilt seen from the sphere's side (|e.(c - o)| is off by <= 1.7e-7 (|c| + |o|))
//     margin    what the ray may carry: the grid takes a ray only if 1e-6 |o|_1 <= margin (conversion of o to float32,
//               three float32 products and sums, the tilt seen from the ray's side); other rays use the cluster traversal
//     1e-4 h    rounding of (p - lo) / h
//
// The same structure serves the PRIMARY rays (entry `nlights` of DScene.sgrid, persp = 1): they share the camera origin
// (renderer.nim:42), so a sphere can only be hit by the directions inside its silhouette cone.  With (e1, e2, e3) the
// rows of the world-to-camera rotation, a direction d maps to (X, Y) = (e1.d, e2.d) / e3.d on the plane one unit in
// front of the camera, and a sphere to the bounding rectangle of its silhouette there (tangent angles in the (x, z)
// and (y, z) planes).  Built for rigid cameras only; rays outside the grid or not facing forward take the clusters.
//
// Scenes of at most 32 objects (every scene the reference ships) use the same grids with a 32-bit OBJECT MASK per cell
// instead of a list (start == null, items[cell] = mask): one load replaces the first look at every object, and the
// bounding spheres of the mesh boxes are binned too, so the same mask answers "can this ray enter mesh object mo's box".
struct ShadowGridF {
  float e1[3], e2[3];
  float lo1, lo2, invh, margin;
  int32_t G, persp;
  const uint32_t* start;   // G * G + 1 offsets into items
  const uint32_t* items;   // object indices
  float e3[3], _padf;
};
struct BundleFrame;  // nrt_filter.h
struct RecSet;       // nrt_filter.h

static constexpr int kHotLights = 2, kHotMO = 1;
struct DScene {
  int32_t nobjects, nlights, nmeshes, nmesh_objs;
  const DObject* objects;
  const CObj* cobjs;              // compact mirror of objects[] for the object scan
  const CObjF* cobjf;             // float32 mirror (first look at spheres)
  // Sphere clusters (scenes with many spheres): the spheres the float32 test applies to, in Morton order
  // of their centres, as groups of kClusterSize (level 2) inside groups of kClusterSize^2 (level 1),
  // every group with a bounding sphere in CObjF form (tx,ty,tz = -centre).  ncl1 == 0: not built.
  const CObjF* cl1;               // ncl1 level-1 bounds
  const CObjF* cl2;               // ncl1 * kClusterSize level-2 bounds (padding: never-hit)
  const CObjF* clm;               // ncl1 * kClusterSize^2 member records (padding: never-hit)
  const uint32_t* clmIdx;         // object index of every member
  const uint32_t* slowIdx;        // the objects that are not cluster members, ascending
  int32_t ncl1, nslow;
  const DLight* lights;
  const DMesh* meshes;
  const int32_t* mesh_obj_index;  // mesh object k -> object index
  const BundleFrame* frames;      // [mo * (2 + nlights) + j]: j = 0 GENERAL, 1 ORIGIN, 2 + l DIR(l)
  const RecSet* recsets;          // same index: the filter record set of the bundle (per-thread mesh walk of the path kernels)
  const MeshGateF* mgate;         // per mesh object: float32 world-space bounding sphere of its box (meshGateMissF)
  uint32_t slowMask;              // scenes of <= 32 objects with mask grids: the objects the grids do not cover (always looked at)
  int32_t maskGrids;              // 1: sgrid's cells hold ONE 32-bit object mask each (items[cell], start == null)
  const ShadowGridF* sgrid;       // nlights + 1 entries (G == 0: none), or null: light-space grids of the clustered spheres per DistantLight, then the camera grid
  double c2w[16];
  double cam_orig[4];             // c2w * (0,0,0,1): castPrimaryRay's origin (renderer.nim:42), the same product done once on the host
  double tan_half_fov;            // f of renderer.nim:38 (host libm, shared with nothing else)
  double bg[3];
  // Hot copies of the small tables, for the kernels that carry this header BY VALUE (FusedBounceT: kernel-parameter
  // space, so a light, a grid header or a gate record is a uniform constant load instead of a dependent chain of
  // global loads).  hotOk != 0 only in the host's by-value copy (SceneData::h) and only when the tables fit
  // (nlights <= kHotLights, nmesh_objs <= kHotMO); the device-resident header keeps 0 and reads the tables.
  int32_t hotOk, moIndexv[kHotMO];
  DLight lightv[kHotLights];
  ShadowGridF sgridv[kHotLights + 1];
  MeshGateF mgatev[kHotMO];
};
// a light / a grid header / a float32 gate record of the scene, from the hot copies when the header has them
NRT_HD DLight lightOf(const DScene& sc, int l) { if (sc.hotOk) return sc.lightv[l]; return sc.lights[l]; }
NRT_HD ShadowGridF sgridOf(const DScene& sc, int i) { if (sc.hotOk) return sc.sgridv[i]; return sc.sgrid[i]; }

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
NRT_HD double sphereIntersect(double radius, V4 orig, V4 dir) {
  const double a = dir.x * dir.x + dir.y * dir.y + dir.z * dir.z;
  const double b = 2 * (dir.x * orig.x + dir.y * orig.y + dir.z * orig.z);
  const double c = orig.x * orig.x + orig.y * orig.y + orig.z * orig.z - radius * radius;
  const double delta = b * b - 4 * a * c;
  if (delta >= 0.0) {
    const double t1 = (-b - signd(b) * sqrt(delta)) / 2 * a;
    const double t2 = c / (a * t1);
    return nim_min(t1, t2);
  }
  return NRT_NEG_INF;
}

// Conservative float32 pre-test of the sphere discriminant (geom.nim:216-230): returns true only
// if delta = b^2 - 4ac is certainly negative, i.e. the reference returns NegInf; the float64
// evaluation (and its sqrt / divisions) is then skipped.  oc = object-space origin (sphere at 0).
// |float32 error of b^2 - a c| <= 16u a (|oc|^2 + r^2) with u = 2^-24 covers the roundings of the
// converted inputs and of the nine float32 operations.
NRT_HD bool sphereCertainMiss(double radius, V4 oc, V4 dir) {
  const float ox = (float)oc.x, oy = (float)oc.y, oz = (float)oc.z;
  const float dx = (float)dir.x, dy = (float)dir.y, dz = (float)dir.z;
  const float r = (float)radius;
  const float a = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
  const float b = fmaf(dx, ox, fmaf(dy, oy, dz * oz));           // half of the reference's b
  const float o2 = fmaf(ox, ox, fmaf(oy, oy, oz * oz));
  const float c = o2 - r * r;
  const float disc = fmaf(b, b, -(a * c));                         // delta / 4
  const float margin = 9.5367431640625e-7f * (a * (o2 + r * r));  // 16 * 2^-24
  // NaN / Inf (overflowing inputs) compare false => not a certain miss
  return disc < -margin;
}

// The same decision from float32 copies of the WORLD-space ray (converted once per ray, not per
// object) and of the object's translation: oc = o + t is formed in float32.  With u = 2^-24,
// M = |o|_inf + |t|_inf and delta = 2.01 u M bounding the error of each component of oc, the
// float32 value of delta/4 = b^2 - a c differs from the exact one by at most
//   16u a (|oc|^2 + r^2)                    conversions of d and r^2, the float32 operations
// + a (7 |oc| delta + 6 delta^2)             perturbation of oc
//   <= a (3.4e-6 |oc|^2 + 1.1e-7 (|o|_inf^2 + |t|_inf^2))   (7 |oc| delta <= 3.5 (2^-20 |oc|^2 + 2^20 delta^2)).
// The test is b^2 - a c < -a (5e-6 |oc|^2 + [2e-6 r^2 + 2e-7 |t|_inf^2] + [2e-7 |o|_inf^2]), evaluated as
//   b^2 < a (|oc|^2 (1 - 5e-6) - r2m - mray)     r2m = r^2 + object part (CObjF), mray = ray part
// (the rearrangement costs another ~3u a |oc|^2, inside the slack between 4.3e-6 and 5e-6).
// true => the reference's delta is negative (NegInf).  NaN / Inf compare false.
struct RayF { float ox, oy, oz, dx, dy, dz, a, mray; };
NRT_HD RayF makeRayF(V4 o, V4 d) {
  RayF r;
  r.ox = (float)o.x; r.oy = (float)o.y; r.oz = (float)o.z;
  r.dx = (float)d.x; r.dy = (float)d.y; r.dz = (float)d.z;
  r.a = fmaf(r.dx, r.dx, fmaf(r.dy, r.dy, r.dz * r.dz));
  const float mo = fmaxf(fabsf(r.ox), fmaxf(fabsf(r.oy), fabsf(r.oz)));
  r.mray = 2e-7f * (mo * mo);
  return r;
}
NRT_HD bool certainMissF(const CObjF& c, const RayF& r) {
  const float ox = r.ox + c.tx, oy = r.oy + c.ty, oz = r.oz + c.tz;
  const float b = fmaf(r.dx, ox, fmaf(r.dy, oy, r.dz * oz));
  const float o2 = fmaf(ox, ox, fmaf(oy, oy, oz * oz));
  return b * b < r.a * fmaf(o2, 0.999995f, -(c.r2m + r.mray));
}

// The cell of a shadow ray (origin rf.o; its direction is the grid's light direction by construction).
// false: the grid does not take this ray (far origin, non-finite values): use the cluster traversal.
// true: items [b, e) are the clustered spheres the ray can possibly hit (b == e outside the grid).
NRT_HD bool shadowGridCell(const ShadowGridF& g, const RayF& rf, uint32_t& b, uint32_t& e) {
  float p1, p2;
  if (g.persp) {   // a primary ray: its direction on the plane one unit in front of the camera
    const float dz = fmaf(g.e3[0], rf.dx, fmaf(g.e3[1], rf.dy, g.e3[2] * rf.dz));
    if (!(dz * dz > 0.09f * rf.a && dz > 0.f)) return false;    // not facing forward enough (or non-finite)
    p1 = fmaf(g.e1[0], rf.dx, fmaf(g.e1[1], rf.dy, g.e1[2] * rf.dz)) / dz;
    p2 = fmaf(g.e2[0], rf.dx, fmaf(g.e2[1], rf.dy, g.e2[2] * rf.dz)) / dz;
  } else {
    const float err = 1e-6f * (fabsf(rf.ox) + fabsf(rf.oy) + fabsf(rf.oz));
    if (!(err <= g.margin)) return false;
    p1 = fmaf(g.e1[0], rf.ox, fmaf(g.e1[1], rf.oy, g.e1[2] * rf.oz));
    p2 = fmaf(g.e2[0], rf.ox, fmaf(g.e2[1], rf.oy, g.e2[2] * rf.oz));
  }
  const float f1 = (p1 - g.lo1) * g.invh, f2 = (p2 - g.lo2) * g.invh;
  b = e = 0;
  if (!(f1 >= 0.f && f2 >= 0.f && f1 < float(g.G) && f2 < float(g.G))) return !g.persp && f1 == f1 && f2 == f2;   // outside: nothing to hit (light grid) / not covered (camera grid); NaN: not taken
  int c1 = int(f1), c2 = int(f2);
  if (c1 > g.G - 1) c1 = g.G - 1;
  if (c2 > g.G - 1) c2 = g.G - 1;
  const uint32_t cell = uint32_t(c2) * uint32_t(g.G) + uint32_t(c1);
  if (!g.start) { b = g.items[cell]; e = 0; return true; }   // mask grid: the mask is returned in b
  b = g.start[cell]; e = g.start[cell + 1];
  return true;
}

#if defined(__CUDA_ARCH__)
// ---- the warp's rays as ONE bundle ---------------------------------------------------------------------------
// The lanes of a warp that trace together hold neighbouring samples (one or two pixels' worth), so their rays form a
// thin bundle.  The bundle is described by its leader's ray (o_c, unit u_c) and two spreads: dO >= |o_i - o_c| and
// dD >= |u_i - u_c| for every participating lane i.  For a sphere (centre c, radius r) and any lane i, with D_i the
// distance of c from lane i's LINE and s_i the parameter of the closest point (|s_i| <= |o_i - c| <= |o_c - c| + dO):
//     D_i >= D_c - |o_i - o_c| - |s_i| |u_i - u_c| >= D_c - (dO + (|o_c - c| + dO) dD) = D_c - Delta.
// So if the LEADER's line passes the certain-miss test of certainMissF() against the radius r + Delta (plus a margin
// of 1e-6 |o_c - c| that keeps every lane's exact discriminant away from zero by far more than the float64
// evaluation's rounding), every lane's reference discriminant is negative: the sphere — or, for a cluster bound, every
// member — is skipped by the whole warp after ONE test in ONE lane.  The tests of up to 32 records run side by
// side in the warp's lanes (bundleScan, nrt_pipeline.h).  Conservative by construction: only "certainly missed by
// all" is ever concluded, everything else goes to the per-ray first look and the float64 evaluation as before.
struct RayBundle {
  unsigned wm;            // participating lanes
  int rank, nact;         // this lane's rank among them, their number
  bool on;                // the bundle test applies (all rays float32-safe, spreads finite and small)
  float ox, oy, oz, ux, uy, uz, a, mray, dO, dD;
};
__device__ __forceinline__ RayBundle makeRayBundle(const RayF& r, bool f32ok) {
  RayBundle B;
  B.wm = __activemask();
  const unsigned lane = threadIdx.x & 31u;
  B.rank = __popc(B.wm & ((1u << lane) - 1u));
  B.nact = __popc(B.wm);
  const int leader = __ffs(int(B.wm)) - 1;
  const float inv = rsqrtf(r.a);
  const float ux = r.dx * inv, uy = r.dy * inv, uz = r.dz * inv;
  B.ox = __shfl_sync(B.wm, r.ox, leader); B.oy = __shfl_sync(B.wm, r.oy, leader); B.oz = __shfl_sync(B.wm, r.oz, leader);
  B.ux = __shfl_sync(B.wm, ux, leader); B.uy = __shfl_sync(B.wm, uy, leader); B.uz = __shfl_sync(B.wm, uz, leader);
  const float ex = r.ox - B.ox, ey = r.oy - B.oy, ez = r.oz - B.oz;
  const float fx = ux - B.ux, fy = uy - B.uy, fz = uz - B.uz;
  const float eo2 = fmaf(ex, ex, fmaf(ey, ey, ez * ez)), ed2 = fmaf(fx, fx, fmaf(fy, fy, fz * fz));
  // maxima over the lanes: non-negative floats order like their bit patterns (a NaN is larger than everything and
  // switches the bundle test off below)
  const float mo2 = __uint_as_float(__reduce_max_sync(B.wm, __float_as_uint(eo2)));
  const float md2 = __uint_as_float(__reduce_max_sync(B.wm, __float_as_uint(ed2)));
  const float moc = fmaxf(fabsf(B.ox), fmaxf(fabsf(B.oy), fabsf(B.oz)));
  // spreads, rounded up: the float32 conversions of o (2^-24 |o| per component), of u and the approximate rsqrt
  B.dO = sqrtf(mo2) * 1.0001f + 4e-7f * moc;
  B.dD = sqrtf(md2) * 1.0001f + 2e-6f;
  const float mo = moc + B.dO;
  B.mray = 2e-7f * (mo * mo);
  B.a = fmaf(B.ux, B.ux, fmaf(B.uy, B.uy, B.uz * B.uz));
  B.on = __all_sync(B.wm, f32ok) && B.dD < 0.5f && B.dO < 1e15f && B.a > 0.99f && B.a < 1.01f;
  return B;
}
// true => EVERY ray of the bundle certainly misses the sphere (or cluster bound) c: certainMissF()'s test for the
// leader's line against the radius sqrt(r2m) + Delta
__device__ __forceinline__ bool bundleMiss(const CObjF& c, const RayBundle& B) {
  if (c.r2m < 0.f) return true;    // never-hit padding record of a cluster
  const float ox = B.ox + c.tx, oy = B.oy + c.ty, oz = B.oz + c.tz;
  const float b = fmaf(B.ux, ox, fmaf(B.uy, oy, B.uz * oz));
  const float o2 = fmaf(ox, ox, fmaf(oy, oy, oz * oz));
  const float doc = sqrtf(o2) * 1.000001f;
  const float delta = fmaf(doc + B.dO, B.dD, B.dO) * 1.001f + 1e-6f * doc;
  const float rr = sqrtf(c.r2m) + delta;
  const float R2 = rr * rr * 1.000004f;
  // (records without a float32 sphere: r2m = +Inf -> R2 = +Inf -> false; NaN anywhere -> false)
  return b * b < B.a * fmaf(o2, 0.999995f, -(R2 + B.mray));
}
// Calls f(k), warp-uniformly and in ascending k, for every record recs[k], k in [0, count), that some ray of the
// bundle may hit: the participating lanes test `nact` records per step side by side, one ballot per step.
template <class F>
__device__ __forceinline__ void bundleScan(const RayBundle& B, const CObjF* recs, int count, F&& f) {
  for (int base = 0; base < count; base += B.nact) {
    const int k = base + B.rank;
    bool cand = false;
    if (k < count) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(recs + k));
      CObjF c; c.tx = a.x; c.ty = a.y; c.tz = a.z; c.r2m = a.w;
      cand = !bundleMiss(c, B);
    }
    unsigned m = __ballot_sync(B.wm, cand);
    while (m) {
      const int src = __ffs(int(m)) - 1;
      m &= m - 1;
      f(base + __popc(B.wm & ((1u << src) - 1u)));
    }
  }
}
#endif

// true => the ray (t >= 0) certainly does not enter the mesh object's box, i.e. the reference's slab test
// (geom.nim:76-96) returns NegInf or a negative tmin and TriangleMesh.intersect returns at geom.nim:340:
//   (1) the ray's LINE misses the bounding sphere — certainMissF's test with r2m = the sphere's; or
//   (2) the origin is outside the sphere and the ray moves away from its centre: b = d.(o - c) > 0.
// With oc the float32 value of o - c (each component off by at most delta = 2.01u(|o|_inf + |c|_inf), see
// certainMissF), |b - b*| <= 3u |d||oc| + sqrt(3) |d| delta, so b > 0 and b^2 > a (1e-5 |oc|^2 + mray + mm) — the
// margins exceed 2 (3u)^2 |oc|^2 + 6 delta^2 by ten orders of magnitude — prove b* > 0; "outside" is the same
// expression that is negative in (1): |oc|^2 (1 - 5e-6) - r2m - mray > 0 proves |o - c|^2 > rb^2.
// NaN / Inf compare false: not certain.  Only valid for rays with w components exactly (1, 0) and finite xyz.
NRT_HD bool meshGateMissF(const MeshGateF& g, const RayF& r) {
  const float ox = r.ox - g.cx, oy = r.oy - g.cy, oz = r.oz - g.cz;
  const float b = fmaf(r.dx, ox, fmaf(r.dy, oy, r.dz * oz));
  const float o2 = fmaf(ox, ox, fmaf(oy, oy, oz * oz));
  const float e = fmaf(o2, 0.999995f, -(g.r2m + r.mray));
  if (b * b < r.a * e) return true;
  return (e > 0.f) && (b > 0.f) && (b * b > r.a * fmaf(1e-5f, o2, r.mray + g.mm));
}

// Plane (object space y = 0, geom.nim:240-248) with worldToObject = [I | t]: when the object-space
// origin height o.y + t.y and the direction's y have the same strict sign the reference returns either
// NegInf (|d.y| <= 1e-6) or t = -o.y / d.y < 0, both rejected by trace() (renderer.nim:60).  Decided in
// float32 from the y components alone: |oy - (o.y + t.y)| <= 2u (1 + u) (|o.y| + |t.y|) < 2.5e-7 (|o.y| + |t.y|)
// =: margin (ray part from |oy|, object part in c.tz, which also adds 1e-30 so that the true height is
// far from the range where -o.y / d.y could round to -0.0).  float(d.y) keeps the sign of d.y.  A
// shadow ray that leaves the plane itself (height = bias = 1e-8, towards the light) is decided here too.
NRT_HD bool planeMissF(const CObjF& c, const RayF& r) {
  const float oy = r.oy + c.ty, m = fmaf(2.5e-7f, fabsf(r.oy), c.tz);
  return (oy > m && r.dy > 0.f) || (oy < -m && r.dy < 0.f);
}

// geom.nim:240-248
NRT_HD double planeIntersect(V4 orig, V4 dir) {
  const V4 n = v4(0.0, 1.0, 0.0, 0.0);
  const double denom = dot(n, dir);
  if (fabs(denom) > 1e-6) return -dot(orig, n) / denom;
  return NRT_NEG_INF;
}

// worldToObject * (orig, dir) of trace() (renderer.nim:54-55).  For a matrix that is exactly
// [I | t] (entries == 1.0 / == 0.0, t finite) and finite x, y, z the glm product
//   ((1*x + 0*y) + 0*z) + t*w
// has zero products that vanish, so per component
//   points  (w == 1): x + t  (one rounding)  unless x == 0 and t == 0 (a zero whose sign depends on
//                     the products: evaluated literally),
//   vectors (w == 0): x                       unless x == 0 (same).
// Bit for bit the literal product at a fraction of its 56 flops and 16 loads; any other input
// (w not exactly 1 / 0, non-finite components) takes the literal product.
NRT_HD bool finite3(V4 v) {
  const double big = 1.7976931348623157e308;
  return (fabs(v.x) <= big) && (fabs(v.y) <= big) && (fabs(v.z) <= big);   // false for NaN
}
NRT_HD double mulmRow(const double* m, int r, V4 v) { return ((m[r] * v.x + m[4 + r] * v.y) + m[8 + r] * v.z) + m[12 + r] * v.w; }
NRT_HD void toObject(const DObject& ob, V4 o, V4 d, V4& oo, V4& dd) {
  if (ob.xlate_only && o.w == 1.0 && d.w == 0.0 && finite3(o) && finite3(d)) {
    const double* m = ob.w2o;
    oo = v4(o.x + m[12], o.y + m[13], o.z + m[14], 1.0);
    dd = v4(d.x, d.y, d.z, 0.0);
    // components whose literal value is a zero of product-dependent sign (rare: one real branch)
    const bool rare = (o.x == 0.0 && m[12] == 0.0) || (o.y == 0.0 && m[13] == 0.0) || (o.z == 0.0 && m[14] == 0.0) ||
                      d.x == 0.0 || d.y == 0.0 || d.z == 0.0;
    if (rare) {
      if (o.x == 0.0 && m[12] == 0.0) oo.x = mulmRow(m, 0, o);
      if (o.y == 0.0 && m[13] == 0.0) oo.y = mulmRow(m, 1, o);
      if (o.z == 0.0 && m[14] == 0.0) oo.z = mulmRow(m, 2, o);
      if (d.x == 0.0) dd.x = mulmRow(m, 0, d);
      if (d.y == 0.0) dd.y = mulmRow(m, 1, d);
      if (d.z == 0.0) dd.z = mulmRow(m, 2, d);
    }
  } else {
    oo = mulm(ob.w2o, o);
    dd = mulm(ob.w2o, d);
  }
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
  const V3 albedoPi = v3(o.albedo_pi[0], o.albedo_pi[1], o.albedo_pi[2]);   // divs(albedo, kPi)
  const double c = nim_max(0.0, dot(hitNormal, scale(si.lightDir, -1.0)));
  return scale(mul(albedoPi, si.lightIntensity), c);
}

// renderer.nim:31-44 (orig/dir only; initRay is applied per object in trace)
// (cx, cy) of renderer.nim:39-40
NRT_HD double primaryCx(const DScene& sc, double r, int w, double x) { return ((2 * x * r) / double(w) - r) * sc.tan_half_fov; }
NRT_HD double primaryCy(const DScene& sc, int h, double y) { return (1 - 2 * y / double(h)) * sc.tan_half_fov; }
NRT_HD void castPrimaryRayC(const DScene& sc, double cx, double cy, V4& orig, V4& dir) {
  orig = v4(sc.cam_orig[0], sc.cam_orig[1], sc.cam_orig[2], sc.cam_orig[3]);
  dir = mulm(sc.c2w, normalize(v4(cx, cy, -1, 0.0)));
}
NRT_HD void castPrimaryRay(const DScene& sc, double r /* = double(w) / double(h) */, int w, int h, double x, double y, V4& orig, V4& dir) {
  castPrimaryRayC(sc, primaryCx(sc, r, w, x), primaryCy(sc, h, y), orig, dir);
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

}  // namespace nrt

#include "nrt_filter.h"
