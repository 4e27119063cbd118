// nrt_host.hpp — C++ mirror of the reference's renderer interface over the C ABI (include/nrt.h).
//
// The reference is Nim; Nim is not available in this image, so the host side above the C ABI
// is written in C++ (the reference is compiled code) with the SAME names, argument meaning and
// error behaviour as the reference's exported surface, so a caller written against
//   src/renderer/renderer.nim:7-28,162-215   (Options, Antialias, renderLine*, initRenderer*)
//   src/renderer/geom.nim:137-198            (Geometry, Sphere, Plane, Box, TriangleMesh, init*)
//   src/renderer/scene.nim:7-18, material.nim:4-7, light.nim:8-17, stats.nim:4-13
//   src/utils/framebuf.nim:7-93              (Framebuf, newFramebuf, writePpm)
//   src/loaders/obj.nim:86-126, objconv.nim:125-153 (loadObj, .geom)
// ports line by line.  The Nim `importc` shim a maintainer would add is in INTEGRATION.md.
//
// Header-only; link with libnrt.so.  There is no CPU fallback: without a B200 every call that
// would render throws nrt::Error(NRT_ERR_NO_DEVICE).
#pragma once

#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/nrt.h"

namespace nimrt {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};
inline void check(int rc, const char* what) {
  if (rc != NRT_OK) throw Error(rc, std::string(what) + ": " + nrt_last_error());
}

// ---- the glm subset the scene files use (float64, column vectors, post-multiplying builders) ----
struct Vec3 { double x = 0, y = 0, z = 0; };
struct Vec4 { double x = 0, y = 0, z = 0, w = 0; };
struct Mat4 { double m[16]; };  // m[col*4+row]

inline Vec3 vec3(double x, double y, double z) { return {x, y, z}; }
inline Vec3 vec3(double v) { return {v, v, v}; }
inline Vec4 vec(double x, double y, double z) { return {x, y, z, 0.0}; }     // geom.nim:11
inline Vec4 point(double x, double y, double z) { return {x, y, z, 1.0}; }   // geom.nim:14
const Vec3 X_AXIS{1, 0, 0}, Y_AXIS{0, 1, 0}, Z_AXIS{0, 0, 1};                // geom.nim:7-9
constexpr double PI = 3.14159265358979323846;
inline double degToRad(double d) { return d * (PI / 180.0); }

inline Mat4 mat4(double d = 1.0) {
  Mat4 r{};
  for (int i = 0; i < 4; ++i) r.m[i * 4 + i] = d;
  return r;
}
inline Vec4 normalize(Vec4 v) {
  const double s = 1.0 / std::sqrt(((v.x * v.x + v.y * v.y) + v.z * v.z) + v.w * v.w);
  return {v.x * s, v.y * s, v.z * s, v.w * s};
}
inline Vec3 normalize(Vec3 v) {
  const double s = 1.0 / std::sqrt((v.x * v.x + v.y * v.y) + v.z * v.z);
  return {v.x * s, v.y * s, v.z * s};
}
// GLM translate: m * T(v)
inline Mat4 translate(const Mat4& a, Vec3 v) {
  Mat4 r = a;
  for (int i = 0; i < 4; ++i) r.m[12 + i] = a.m[i] * v.x + a.m[4 + i] * v.y + a.m[8 + i] * v.z + a.m[12 + i];
  return r;
}
// GLM rotate with the fork's argument order (m, axis, angle): m * R
inline Mat4 rotate(const Mat4& a, Vec3 axis, double angle) {
  const Vec3 n = normalize(axis);
  const double c = std::cos(angle), s = std::sin(angle);
  const Vec3 t{(1 - c) * n.x, (1 - c) * n.y, (1 - c) * n.z};
  double R[3][3];  // R[col][row]
  R[0][0] = c + t.x * n.x; R[0][1] = t.x * n.y + s * n.z; R[0][2] = t.x * n.z - s * n.y;
  R[1][0] = t.y * n.x - s * n.z; R[1][1] = c + t.y * n.y; R[1][2] = t.y * n.z + s * n.x;
  R[2][0] = t.z * n.x + s * n.y; R[2][1] = t.z * n.y - s * n.x; R[2][2] = c + t.z * n.z;
  Mat4 r = a;
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 4; ++i) r.m[j * 4 + i] = a.m[i] * R[j][0] + a.m[4 + i] * R[j][1] + a.m[8 + i] * R[j][2];
  return r;
}
// cofactor inverse (geom.nim:162 `objectToWorld.inverse`)
inline Mat4 inverse(const Mat4& a) {
  auto at = [&](int r, int c) { return a.m[c * 4 + r]; };
  double cof[4][4];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      double s[3][3];
      int ri = 0;
      for (int i = 0; i < 4; ++i) {
        if (i == r) continue;
        int ci = 0;
        for (int j = 0; j < 4; ++j) {
          if (j == c) continue;
          s[ri][ci++] = at(i, j);
        }
        ++ri;
      }
      const double d = s[0][0] * (s[1][1] * s[2][2] - s[1][2] * s[2][1]) - s[0][1] * (s[1][0] * s[2][2] - s[1][2] * s[2][0]) +
                       s[0][2] * (s[1][0] * s[2][1] - s[1][1] * s[2][0]);
      cof[r][c] = ((r + c) % 2 == 0) ? d : -d;
    }
  double det = 0;
  for (int c = 0; c < 4; ++c) det += at(0, c) * cof[0][c];
  Mat4 out{};
  const double inv = 1.0 / det;
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) out.m[c * 4 + r] = cof[c][r] * inv;
  return out;
}

// ---- geom.nim:21-24,137-198 ----
struct Triangle { int64_t vertexIdx[3]; int64_t normalIdx[3]; };

struct Geometry {
  virtual ~Geometry() = default;
  Mat4 objectToWorld = mat4(1.0), worldToObject = mat4(1.0);
  virtual nrt_geom_kind kind() const = 0;
};
struct Sphere : Geometry { double r = 0; nrt_geom_kind kind() const override { return NRT_GEOM_SPHERE; } };
struct Plane : Geometry { nrt_geom_kind kind() const override { return NRT_GEOM_PLANE; } };
struct Box : Geometry { Vec4 vmin, vmax; nrt_geom_kind kind() const override { return NRT_GEOM_BOX; } };
struct TriangleMesh : Geometry {
  std::vector<Vec4> vertices, normals;
  std::vector<Triangle> faces;
  nrt_geom_kind kind() const override { return NRT_GEOM_MESH; }
};
using GeometryRef = std::shared_ptr<Geometry>;

inline std::shared_ptr<Sphere> initSphere(double r, const Mat4& objectToWorld) {
  auto s = std::make_shared<Sphere>();
  s->r = r; s->objectToWorld = objectToWorld; s->worldToObject = inverse(objectToWorld);
  return s;
}
inline std::shared_ptr<Plane> initPlane(const Mat4& objectToWorld) {
  auto p = std::make_shared<Plane>();
  p->objectToWorld = objectToWorld; p->worldToObject = inverse(objectToWorld);
  return p;
}
inline std::shared_ptr<Box> initBox(Vec4 vmin, Vec4 vmax, const Mat4& objectToWorld) {
  auto b = std::make_shared<Box>();
  b->vmin = vmin; b->vmax = vmax; b->objectToWorld = objectToWorld; b->worldToObject = inverse(objectToWorld);
  return b;
}
inline std::shared_ptr<TriangleMesh> initTriangleMesh(std::vector<Vec4> vertices, std::vector<Vec4> normals,
                                                      std::vector<Triangle> faces, const Mat4& objectToWorld) {
  auto m = std::make_shared<TriangleMesh>();
  m->vertices = std::move(vertices); m->normals = std::move(normals); m->faces = std::move(faces);
  m->objectToWorld = objectToWorld; m->worldToObject = inverse(objectToWorld);
  return m;  // the AABB (geom.nim:175-188) is computed by the callee, exactly as calcAABB
}

// ---- material.nim, scene.nim, light.nim ----
struct Material { Vec3 albedo; double reflection = 0.0; };
struct Object { std::string name; GeometryRef geometry; Material material; };
struct Light { virtual ~Light() = default; Vec3 color; double intensity = 0; virtual nrt_light_kind kind() const = 0; };
struct DistantLight : Light { Vec4 dir; nrt_light_kind kind() const override { return NRT_LIGHT_DISTANT; } };
struct PointLight : Light { Vec4 pos; nrt_light_kind kind() const override { return NRT_LIGHT_POINT; } };
inline std::shared_ptr<DistantLight> distantLight(Vec3 color, double intensity, Vec4 dir) {
  auto l = std::make_shared<DistantLight>(); l->color = color; l->intensity = intensity; l->dir = dir; return l;
}
inline std::shared_ptr<PointLight> pointLight(Vec3 color, double intensity, Vec4 pos) {
  auto l = std::make_shared<PointLight>(); l->color = color; l->intensity = intensity; l->pos = pos; return l;
}
struct Scene {
  std::vector<Object> objects;
  std::vector<std::shared_ptr<Light>> lights;
  double fov = 50.0;
  Mat4 cameraToWorld = mat4(1.0);
  Vec3 bgColor;
};

// ---- renderer.nim:10-28 ----
enum AntialiasKind { akNone, akGrid, akJittered, akMultiJittered, akCorrelatedMultiJittered };
struct Antialias { AntialiasKind kind = akNone; int gridSize = 1; };
struct Options {
  int width = 0, height = 0;
  Antialias antialias;
  double bias = 0.00000001;   // raytracer.nim:50
  int maxRayDepth = 5;        // raytracer.nim:51
  // build-specific: how renderer.nim:108 `ray.depth <= maxRayDepth` is read (see nrt.h)
  nrt_depth_mode depthMode = NRT_DEPTH_REFBUG;
  int bounceCap = 64;
  uint64_t seed = 0;
};

// ---- stats.nim:4-13 ----
struct Stats {
  int64_t numPrimaryRays = 0, numIntersectionTests = 0, numIntersectionHits = 0;
  Stats& operator+=(const Stats& r) {
    numPrimaryRays += r.numPrimaryRays; numIntersectionTests += r.numIntersectionTests;
    numIntersectionHits += r.numIntersectionHits;
    return *this;
  }
};

// ---- utils/framebuf.nim:7-28 ----
struct Framebuf {
  int w = 0, h = 0;
  std::vector<float> data;
  float* at(int x, int y) { assert(x < w && y < h); return data.data() + (size_t(y) * w + x) * 3; }
};
inline Framebuf newFramebuf(int w, int h) { Framebuf f; f.w = w; f.h = h; f.data.assign(size_t(w) * h * 3, 0.f); return f; }

// Page-locks caller-owned memory for the lifetime of the object (nrt_host_register): a Framebuf's data or
// a mesh array, so that renderLine / renderFrame / DeviceScene::update copy at full PCIe speed.  Optional:
// when the driver refuses (no device, range already registered) the memory simply stays pageable.
class PageLock {
 public:
  PageLock(void* p, size_t bytes) : p_(nrt_host_register(p, int64_t(bytes)) == NRT_OK ? p : nullptr) {}
  explicit PageLock(Framebuf& fb) : PageLock(fb.data.data(), fb.data.size() * sizeof(float)) {}
  ~PageLock() { if (p_) nrt_host_unregister(p_); }
  PageLock(const PageLock&) = delete;
  PageLock& operator=(const PageLock&) = delete;
  bool locked() const { return p_ != nullptr; }
 private:
  void* p_;
};

// ---- Scene -> nrt_scene_desc (element-wise, no memcpy of glm types) -----------------------
class DeviceScene {
 public:
  explicit DeviceScene(const Scene& s) { build(s); check(nrt_scene_create(&desc_, &handle_), "nrt_scene_create"); }
  ~DeviceScene() { if (handle_) nrt_scene_destroy(handle_); }
  DeviceScene(const DeviceScene&) = delete;
  DeviceScene& operator=(const DeviceScene&) = delete;
  void update(const Scene& s) { build(s); check(nrt_scene_update(handle_, &desc_), "nrt_scene_update"); }
  nrt_scene* handle() const { return handle_; }

 private:
  void build(const Scene& s) {
    objs_.assign(s.objects.size(), nrt_object{});
    lights_.assign(s.lights.size(), nrt_light{});
    meshes_.clear(); store_.clear();
    std::map<const Geometry*, int> index;
    for (size_t i = 0; i < s.objects.size(); ++i) {
      const Object& o = s.objects[i];
      nrt_object& d = objs_[i];
      d.kind = o.geometry->kind(); d.mesh = -1;
      for (int k = 0; k < 16; ++k) { d.object_to_world[k] = o.geometry->objectToWorld.m[k]; d.world_to_object[k] = o.geometry->worldToObject.m[k]; }
      if (auto sp = dynamic_cast<const Sphere*>(o.geometry.get())) d.radius = sp->r;
      if (auto bx = dynamic_cast<const Box*>(o.geometry.get())) {
        const double a[4] = {bx->vmin.x, bx->vmin.y, bx->vmin.z, bx->vmin.w}, b[4] = {bx->vmax.x, bx->vmax.y, bx->vmax.z, bx->vmax.w};
        for (int k = 0; k < 4; ++k) { d.vmin[k] = a[k]; d.vmax[k] = b[k]; }
      }
      d.albedo[0] = o.material.albedo.x; d.albedo[1] = o.material.albedo.y; d.albedo[2] = o.material.albedo.z;
      d.reflection = o.material.reflection;
      if (auto tm = dynamic_cast<const TriangleMesh*>(o.geometry.get())) {
        auto it = index.find(tm);
        if (it == index.end()) {
          auto st = std::make_unique<MeshStore>();
          for (const Vec4& v : tm->vertices) { st->v.push_back(v.x); st->v.push_back(v.y); st->v.push_back(v.z); st->v.push_back(v.w); }
          for (const Vec4& v : tm->normals) { st->n.push_back(v.x); st->n.push_back(v.y); st->n.push_back(v.z); st->n.push_back(v.w); }
          for (const Triangle& t : tm->faces)
            for (int k = 0; k < 3; ++k) { st->vi.push_back(t.vertexIdx[k]); st->ni.push_back(t.normalIdx[k]); }
          nrt_mesh m{};
          m.nverts = int64_t(tm->vertices.size()); m.vertices = st->v.data();
          m.nnormals = int64_t(tm->normals.size()); m.normals = st->n.data();
          m.nfaces = int64_t(tm->faces.size()); m.vertex_idx = st->vi.data(); m.normal_idx = st->ni.data();
          it = index.emplace(tm, int(meshes_.size())).first;
          meshes_.push_back(m);
          store_.push_back(std::move(st));
        }
        d.mesh = it->second;
      }
    }
    for (size_t i = 0; i < s.lights.size(); ++i) {
      const Light& l = *s.lights[i];
      nrt_light& d = lights_[i];
      d.kind = l.kind();
      d.color[0] = l.color.x; d.color[1] = l.color.y; d.color[2] = l.color.z;
      d.intensity = l.intensity;
      if (auto dl = dynamic_cast<const DistantLight*>(&l)) { d.dir[0] = dl->dir.x; d.dir[1] = dl->dir.y; d.dir[2] = dl->dir.z; d.dir[3] = dl->dir.w; }
      if (auto pl = dynamic_cast<const PointLight*>(&l)) { d.pos[0] = pl->pos.x; d.pos[1] = pl->pos.y; d.pos[2] = pl->pos.z; d.pos[3] = pl->pos.w; }
    }
    desc_ = nrt_scene_desc{};
    desc_.nobjects = int(objs_.size()); desc_.nlights = int(lights_.size()); desc_.nmeshes = int(meshes_.size());
    desc_.objects = objs_.data(); desc_.lights = lights_.data(); desc_.meshes = meshes_.data();
    desc_.fov = s.fov;
    for (int k = 0; k < 16; ++k) desc_.camera_to_world[k] = s.cameraToWorld.m[k];
    desc_.bg_color[0] = s.bgColor.x; desc_.bg_color[1] = s.bgColor.y; desc_.bg_color[2] = s.bgColor.z;
  }
  struct MeshStore { std::vector<double> v, n; std::vector<int64_t> vi, ni; };
  std::vector<nrt_object> objs_;
  std::vector<nrt_light> lights_;
  std::vector<nrt_mesh> meshes_;
  std::vector<std::unique_ptr<MeshStore>> store_;
  nrt_scene_desc desc_{};
  nrt_scene* handle_ = nullptr;
};

inline nrt_options toC(const Options& o) {
  nrt_options c{};
  c.width = o.width; c.height = o.height; c.aa_kind = int(o.antialias.kind); c.grid_size = o.antialias.gridSize;
  c.bias = o.bias; c.max_ray_depth = o.maxRayDepth; c.depth_mode = o.depthMode; c.bounce_cap = o.bounceCap; c.seed = o.seed;
  return c;
}

// renderer.nim:214-215 (+ the worker-pool start of raytracer.nim:61-65): selects the GPUs.
inline void initRenderer(int ngpu = 1) { check(nrt_init(ngpu, nullptr), "nrt_init"); }

inline bool isPowerOfTwo(int v) { return v > 0 && (v & (v - 1)) == 0; }

// renderer.nim:162-211 — one scanline (the worker-pool caller of raytracer.nim:25-32).
inline Stats renderLine(const DeviceScene& scene, const Options& opts, Framebuf& fb, int y, int step = 1, int maxStep = 1) {
  assert(isPowerOfTwo(step));      // renderer.nim:166-168
  assert(isPowerOfTwo(maxStep));
  assert(maxStep >= step);
  const nrt_options c = toC(opts);
  nrt_stats st{};
  check(nrt_render(scene.handle(), &c, y, y + 1, step, maxStep, fb.data.data(), &st, nullptr), "nrt_render");
  Stats r; r.numPrimaryRays = st.num_primary_rays; r.numIntersectionTests = st.num_intersection_tests; r.numIntersectionHits = st.num_intersection_hits;
  return r;
}

// All the lines the caller would queue (raytracer.nim:67-70 / gui.nim:113-122) in one call.
inline Stats renderFrame(const DeviceScene& scene, const Options& opts, Framebuf& fb, int step = 1, int maxStep = 1) {
  const nrt_options c = toC(opts);
  nrt_stats st{};
  check(nrt_render(scene.handle(), &c, 0, opts.height, step, maxStep, fb.data.data(), &st, nullptr), "nrt_render");
  Stats r; r.numPrimaryRays = st.num_primary_rays; r.numIntersectionTests = st.num_intersection_tests; r.numIntersectionHits = st.num_intersection_hits;
  return r;
}

// utils/framebuf.nim:55-93: P6, maxval 2^bits - 1, 8-bit samples up to 8 bits, big-endian 16-bit samples above;
// clamp -> linearToSRGB -> round on the GPU output stage (bit-exact: DESIGN.md section 13)
inline bool writePpmSamples(const std::string& filename, int w, int h, int bits, const std::vector<unsigned char>& samples) {
  std::ofstream f(filename, std::ios::binary);
  if (!f) return false;
  f << "P6 " << w << " " << h << " " << ((1 << bits) - 1) << " ";   // framebuf.nim:62-65
  f.write(reinterpret_cast<const char*>(samples.data()), std::streamsize(samples.size()));
  return bool(f);
}
inline bool writePpm(const Framebuf& fb, const std::string& filename, int bits = 8, bool sRGB = true) {
  if (bits < 1 || bits > 16) return false;   // framebuf.nim:56
  std::vector<unsigned char> img(size_t(fb.w) * fb.h * (bits <= 8 ? 3 : 6));
  if (nrt_framebuf_quantize(fb.data.data(), fb.w, fb.h, bits, sRGB ? 1 : 0, img.data()) != NRT_OK) return false;
  return writePpmSamples(filename, fb.w, fb.h, bits, img);
}

// renderFrame followed by writePpm in ONE call (nrt_render_quantized): the conversion runs as the epilogue of the
// kernel that stores the pixels and only the integer samples leave the GPU (3 or 6 bytes per pixel instead of 12).
// `samples` receives the PPM sample stream (what writePpm would write after its header).
inline Stats renderFrameQuantized(const DeviceScene& scene, const Options& opts, std::vector<unsigned char>& samples,
                                  int bits = 8, bool sRGB = true) {
  const nrt_options c = toC(opts);
  nrt_stats st{};
  samples.assign(size_t(opts.width) * opts.height * (bits <= 8 ? 3 : 6), 0);
  check(nrt_render_quantized(scene.handle(), &c, 0, opts.height, 1, 1, NRT_OUT_RGB, bits, sRGB ? 1 : 0, 255, samples.data(), &st),
        "nrt_render_quantized");
  Stats r; r.numPrimaryRays = st.num_primary_rays; r.numIntersectionTests = st.num_intersection_tests; r.numIntersectionHits = st.num_intersection_hits;
  return r;
}

// utils/image.nim:45-54 (ImageRGBA.copyFrom, the GUI's display copy): round(v * 255) per channel + a constant alpha
inline bool toRGBA8(const Framebuf& fb, std::vector<unsigned char>& rgba, unsigned char alpha = 255) {
  rgba.assign(size_t(fb.w) * fb.h * 4, 0);
  return nrt_framebuf_to_rgba8(fb.data.data(), fb.w, fb.h, alpha, rgba.data()) == NRT_OK;
}

// Which partition of `count` renders unit (band / rendered scanline) `unit` of a pass: the serpentine deal of
// raytracer.nim:67-70's work items to GPUs / processes (nrt_set_partition selects this process's share).
inline int unitOwner(long long unit, int count) { return nrt_unit_owner(unit, count); }
// The first scanlines of the units partition `index` of `count` renders in [y0, y1) (nrt_partition_rows).
inline std::vector<int> partitionRows(int height, int index, int count, int y0, int y1, int step = 1, int band = 1) {
  const int n = nrt_partition_rows(height, y0, y1, step, band, index, count, nullptr, 0);
  if (n < 0) throw Error(NRT_ERR_INVALID, "nrt_partition_rows: bad arguments");
  std::vector<int> rows(static_cast<size_t>(n));
  if (n > 0) nrt_partition_rows(height, y0, y1, step, band, index, count, rows.data(), n);
  return rows;
}
inline void shutdown() { nrt_shutdown(); }

// ---- loaders: obj.nim:65-126 (v / f only, one flat normal per face) and the .geom format ----
inline void calcNormals(TriangleMesh& m) {  // obj.nim:65-84
  m.normals.assign(m.faces.size(), Vec4{});
  for (size_t k = 0; k < m.faces.size(); ++k) {
    Triangle& t = m.faces[k];
    const Vec4 &p0 = m.vertices[t.vertexIdx[0]], &p1 = m.vertices[t.vertexIdx[1]], &p2 = m.vertices[t.vertexIdx[2]];
    const Vec3 a{p1.x - p0.x, p1.y - p0.y, p1.z - p0.z}, b{p2.x - p0.x, p2.y - p0.y, p2.z - p0.z};
    const Vec3 n = normalize(Vec3{a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y});
    m.normals[k] = vec(n.x, n.y, n.z);
    t.normalIdx[0] = t.normalIdx[1] = t.normalIdx[2] = int64_t(k);
  }
}
inline std::shared_ptr<TriangleMesh> loadObj(const std::string& fname) {
  std::ifstream f(fname);
  if (!f) throw Error(NRT_ERR_INVALID, "cannot open " + fname);
  auto m = std::make_shared<TriangleMesh>();
  std::string line;
  while (std::getline(f, line)) {
    std::istringstream ss(line);
    std::string tag;
    if (!(ss >> tag)) continue;
    // obj.nim:25-63: a token that does not parse leaves the coordinate at 0.0 / the index at 0 (`except ValueError:
    // discard`); the vertex index of a v/vt/vn token is read (a deliberate superset, as in loaders.py)
    if (tag == "v") {
      double xyz[3] = {0, 0, 0};
      for (int k = 0; k < 3; ++k) {
        std::string tok; ss >> tok;
        char* end = nullptr;
        const double v = std::strtod(tok.c_str(), &end);
        xyz[k] = (!tok.empty() && end && *end == '\0') ? v : 0.0;
      }
      m->vertices.push_back(point(xyz[0], xyz[1], xyz[2]));
    }
    else if (tag == "f") {
      Triangle t{};
      for (int k = 0; k < 3; ++k) {
        std::string tok; ss >> tok;
        tok = tok.substr(0, tok.find('/'));
        char* end = nullptr;
        const long long v = std::strtoll(tok.c_str(), &end, 10);
        t.vertexIdx[k] = (!tok.empty() && end && *end == '\0') ? int64_t(v) - 1 : 0;
      }
      m->faces.push_back(t);
    }
  }
  calcNormals(*m);
  return m;
}
// .geom: int32 triangle count + 9 float32 per triangle (objconv.nim:139-153); the reference's own
// reader (geomloader.nim:30-49) is unfinished — this is the finished equivalent.
inline std::shared_ptr<TriangleMesh> loadGeom(const std::string& fname) {
  std::ifstream f(fname, std::ios::binary);
  if (!f) throw Error(NRT_ERR_INVALID, "cannot open " + fname);
  int32_t n = 0;
  f.read(reinterpret_cast<char*>(&n), 4);
  std::vector<float> buf(size_t(n) * 9);
  f.read(reinterpret_cast<char*>(buf.data()), std::streamsize(buf.size() * 4));
  if (!f) throw Error(NRT_ERR_INVALID, "truncated " + fname);
  auto m = std::make_shared<TriangleMesh>();
  for (int32_t k = 0; k < n; ++k) {
    Triangle t{};
    for (int i = 0; i < 3; ++i) {
      m->vertices.push_back(point(buf[size_t(k) * 9 + i * 3], buf[size_t(k) * 9 + i * 3 + 1], buf[size_t(k) * 9 + i * 3 + 2]));
      t.vertexIdx[i] = int64_t(k) * 3 + i;
    }
    m->faces.push_back(t);
  }
  calcNormals(*m);
  return m;
}

// objconv.nim:139-153: int32 triangle count, then the three vertices of every face as 9 float32
inline void writeGeom(const std::string& fname, const TriangleMesh& m) {
  std::ofstream f(fname, std::ios::binary);
  if (!f) throw Error(NRT_ERR_INVALID, "cannot create " + fname);
  const int32_t n = int32_t(m.faces.size());
  f.write(reinterpret_cast<const char*>(&n), 4);
  for (const Triangle& t : m.faces)
    for (int i = 0; i < 3; ++i) {
      const Vec4& v = m.vertices[size_t(t.vertexIdx[i])];
      const float p[3] = {float(v.x), float(v.y), float(v.z)};
      f.write(reinterpret_cast<const char*>(p), 12);
    }
  if (!f) throw Error(NRT_ERR_INVALID, "write failed: " + fname);
}

}  // namespace nimrt
