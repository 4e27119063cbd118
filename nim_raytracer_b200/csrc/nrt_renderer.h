// nrt_renderer.h — backend-independent host orchestration of the wavefront
// pipeline (nrt_pipeline.h): scene flattening/upload and the per-chunk wave loop.
// `BE` supplies memory, launches and atomics: CudaBackend in nrt.cu (the
// product) or the loop backend of the test-only emulation.
#pragma once

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/nrt.h"
#include "nrt_pipeline.h"

namespace nrt {

struct ProfileAcc {
  int64_t mesh_tests = 0, mesh_tests_ref = 0, mesh_rays = 0, candidates = 0, exact_rays = 0;
  int64_t tests_by_mode[3] = {0, 0, 0};
  int64_t pre_candidates = 0;
  int64_t active[8] = {0}, wavefront[8] = {0}, tail = 0;   // fused path: samples per bounce / through the wavefront / finished by PathTail
  int64_t need_cand = 0, need_pairs = 0;   // the largest list any (wave, mesh object) asked for (counters keep counting past a full list)
};

inline bool isPow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// log2 of the tile edge T of a pass (FrameParams::tshift): whole-resolution passes use the largest power of two
// with T * T * spp <= 256 (one prefilter run = one tile where the numbers allow: 16 spp -> 4 x 4 pixels), capped
// at 16; progressive passes (step > 1 or pixels to skip) keep scanline order.  NRT_TILE overrides (1 = scanlines).
inline int tileShiftFor(const nrt_options& o, int step, int max_step) {
  if (step != 1 || max_step != 1) return 0;
  const int spp = o.aa_kind == NRT_AA_NONE ? 1 : o.grid_size * o.grid_size;
  int t = 0;
  while (t < 4 && (int64_t(1) << (2 * (t + 1))) * spp <= 256) ++t;
  if (const char* e = std::getenv("NRT_TILE")) {
    const int v = std::atoi(e);
    if (v >= 1 && v <= 64 && isPow2(v)) { t = 0; while ((1 << t) < v) ++t; }
  }
  return t;
}

// ------------------------------------------------------------------ scene ----
template <class BE>
struct SceneData {
  BE* be = nullptr;
  DScene h{};                 // host copy of the header (device pointers inside)
  DScene* d = nullptr;        // device copy
  std::vector<DObject> objs;
  std::vector<DLight> lights;
  std::vector<DMesh> meshes;  // device pointers inside
  std::vector<int32_t> moIndex;
  std::vector<CObj> cobjs; CObj* dCObjs = nullptr;
  std::vector<CObjF> cobjf; CObjF* dCObjF = nullptr;
  // sphere clusters (see DScene): device copies are re-allocated when a scene update changes their sizes
  CObjF *dCl1 = nullptr, *dCl2 = nullptr, *dClm = nullptr; uint32_t *dClmIdx = nullptr, *dSlowIdx = nullptr;
  int64_t capCl1 = 0, capSlow = 0;
  DObject* dObjs = nullptr; DLight* dLights = nullptr; DMesh* dMeshes = nullptr; int32_t* dMo = nullptr;
  std::vector<void*> owned;
  bool anyReflective = false, anyPointLight = false;
  int64_t bytes_uploaded = 0;
  // nrt_scene_update of a description that is bit-identical to the resident one: the inputs are still copied to
  // the device (into staging arrays, compared there with the live ones), the device-side precompute is skipped
  struct MeshStage { double* verts = nullptr; double* normals = nullptr; int64_t* vidx = nullptr; int64_t* nidx = nullptr; };
  std::vector<MeshStage> stage;
  std::vector<nrt_object> lastObjs;
  std::vector<nrt_light> lastLights;
  double lastCam[16] = {0}, lastFov = 0, lastBg[3] = {0, 0, 0};
  bool haveLast = false;
  int64_t rebuilds = 0, rebuildsSkipped = 0;
  bool sameSmallParts(const nrt_scene_desc* desc) const {
    if (!haveLast || size_t(desc->nobjects) != lastObjs.size() || size_t(desc->nlights) != lastLights.size()) return false;
    if (desc->nobjects && std::memcmp(desc->objects, lastObjs.data(), sizeof(nrt_object) * lastObjs.size()) != 0) return false;
    if (desc->nlights && std::memcmp(desc->lights, lastLights.data(), sizeof(nrt_light) * lastLights.size()) != 0) return false;
    return std::memcmp(desc->camera_to_world, lastCam, sizeof(lastCam)) == 0 && std::memcmp(&desc->fov, &lastFov, sizeof(double)) == 0 &&
           std::memcmp(desc->bg_color, lastBg, sizeof(lastBg)) == 0;
  }
  void rememberSmallParts(const nrt_scene_desc* desc) {
    lastObjs.assign(desc->objects, desc->objects + desc->nobjects);
    lastLights.assign(desc->lights, desc->lights + desc->nlights);
    std::memcpy(lastCam, desc->camera_to_world, sizeof(lastCam));
    lastFov = desc->fov;
    std::memcpy(lastBg, desc->bg_color, sizeof(lastBg));
    haveLast = true;
  }
  // Filter record sets per mesh object (device): ORIGIN (camera) and DIR per DistantLight.
  struct MoRecs { float* origin = nullptr; float* originHot = nullptr; float* originBounds = nullptr; std::vector<float*> dir, dirHot, dirBounds; };
  std::vector<BundleFrame> frames;        // host copy, [mo * recStride() + j]
  BundleFrame* dFrames = nullptr;
  RecSet* dRecSets = nullptr;             // device table of the record sets, indexed like the frames (DScene.recsets)
  MeshGateF* dMGate = nullptr;            // float32 gate records per mesh object (DScene.mgate)
  std::vector<MeshGateF> hMGate;          // host copy (mask grids bin the boxes' bounding spheres)
  std::vector<ShadowGridF> hSGrid;        // host copy of the grid headers (DScene::sgridv)
  static bool envHot() { const char* e = std::getenv("NRT_HOT_HEADER"); return !(e && *e == '0'); }
  bool anyGeneralShadow = false;          // some shadow rays need the GENERAL bundle (point light / unusable frame)
  std::vector<MoRecs> moRecs;
  uint32_t* dRecCount = nullptr;          // [mo * recStride() + j]: j = 0 GENERAL, 1 ORIGIN, 2 + l DIR(l)
  std::vector<uint32_t> hRecCount;
  int recStride() const { return 2 + h.nlights; }
  const float* recsOf(int mo, int mode, int l) const {
    return mode == FM_GENERAL ? meshes[objs[moIndex[mo]].mesh].recs : (mode == FM_ORIGIN ? moRecs[mo].origin : moRecs[mo].dir[l]);
  }
  const float* hotOf(int mo, int mode, int l) const {
    return mode == FM_GENERAL ? meshes[objs[moIndex[mo]].mesh].hot : (mode == FM_ORIGIN ? moRecs[mo].originHot : moRecs[mo].dirHot[l]);
  }
  const float* boundsOf(int mo, int mode, int l) const {
    return mode == FM_GENERAL ? meshes[objs[moIndex[mo]].mesh].bounds : (mode == FM_ORIGIN ? moRecs[mo].originBounds : moRecs[mo].dirBounds[l]);
  }
  static int64_t numChunks(int64_t nfaces) { return paddedFaces(nfaces) / kRecPad; }
  bool frameValid(int mo, int mode, int l) const { return frames[frameIndex(h.nlights, mo, mode, l)].valid > 0; }

  static bool makeBasis(const double* axis, BundleFrame& fr) {
    const double n2 = axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2];
    if (!(n2 > 1e-280) || !(n2 < 1e280)) return false;
    const double il = 1.0 / std::sqrt(n2);
    for (int k = 0; k < 3; ++k) fr.f[k] = axis[k] * il;
    const double ax = std::fabs(fr.f[0]), ay = std::fabs(fr.f[1]), az = std::fabs(fr.f[2]);
    double a[3] = {0, 0, 0};
    a[(ax <= ay && ax <= az) ? 0 : (ay <= az ? 1 : 2)] = 1.0;
    double c[3] = {a[1] * fr.f[2] - a[2] * fr.f[1], a[2] * fr.f[0] - a[0] * fr.f[2], a[0] * fr.f[1] - a[1] * fr.f[0]};
    const double cl = 1.0 / std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
    for (int k = 0; k < 3; ++k) fr.b1[k] = c[k] * cl;
    fr.b2[0] = fr.f[1] * fr.b1[2] - fr.f[2] * fr.b1[1];
    fr.b2[1] = fr.f[2] * fr.b1[0] - fr.f[0] * fr.b1[2];
    fr.b2[2] = fr.f[0] * fr.b1[1] - fr.f[1] * fr.b1[0];
    return true;
  }
  const uint32_t* recCountOf(int mo, int mode, int l) const { return dRecCount + mo * recStride() + (mode == FM_GENERAL ? 0 : (mode == FM_ORIGIN ? 1 : 2 + l)); }
  uint32_t hostRecCount(int mo, int mode, int l) const { return hRecCount[mo * recStride() + (mode == FM_GENERAL ? 0 : (mode == FM_ORIGIN ? 1 : 2 + l))]; }

  // geom.nim:175-188
  static void calcAABB(const double* v, int64_t n, double* bmin, double* bmax) {
    bmin[0] = bmin[1] = bmin[2] = NRT_INF; bmin[3] = 1.0;
    bmax[0] = bmax[1] = bmax[2] = NRT_NEG_INF; bmax[3] = 1.0;
    for (int64_t i = 0; i < n; ++i) {
      const double* p = v + 4 * i;
      for (int k = 0; k < 3; ++k) {
        if (p[k] < bmin[k]) bmin[k] = p[k];
        if (p[k] > bmax[k]) bmax[k] = p[k];
      }
    }
  }

  template <class T>
  T* up(const T* src, int64_t n, T* reuse = nullptr) {
    if (n <= 0) n = 1;
    T* p = reuse;
    if (!p) { p = static_cast<T*>(be->dalloc(sizeof(T) * n)); owned.push_back(p); }
    if (src) { be->upload(p, src, sizeof(T) * n); bytes_uploaded += int64_t(sizeof(T)) * n; }
    return p;
  }

  // Sphere clusters for the object scan of scenes with many spheres (DScene.cl1 ...): the spheres the
  // float32 test applies to are sorted by the Morton key of their centres and grouped 16 x 16; every
  // group gets a bounding sphere in CObjF form, i.e. "a big sphere at the group's centre".  A ray that
  // certainly misses the big sphere is farther than R >= |c_i - C| + r_i from C, hence farther than r_i
  // from every member centre: every member's discriminant is negative with the same float32 slack.
  static CObjF boundRecord(const double* C, double R) {
    CObjF f;
    const double mt = std::max(std::fabs(C[0]), std::max(std::fabs(C[1]), std::fabs(C[2]))), r2 = R * R;
    f.tx = float(-C[0]); f.ty = float(-C[1]); f.tz = float(-C[2]);
    f.r2m = (std::isfinite(r2) && r2 < 1e30 && mt < 1e15) ? roundUpF(r2 + 2e-6 * r2 + 2e-7 * mt * mt) : float(NRT_INF);   // Inf: never skipped
    return f;
  }
  void buildClusters(bool reuse) {
    h.ncl1 = 0; h.nslow = 0;
    std::vector<uint32_t> fast, slow;
    for (size_t i = 0; i < cobjf.size(); ++i) (cobjf[i].r2m < 3.0e38f ? fast : slow).push_back(uint32_t(i));
    if (int64_t(fast.size()) < kClusterMin) return;
    // Morton order of the centres (centre = -t for worldToObject = [I | t])
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (uint32_t i : fast) for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], -cobjs[i].t[k]); hi[k] = std::max(hi[k], -cobjs[i].t[k]); }
    auto key = [&](uint32_t i) {
      uint32_t q[3];
      for (int k = 0; k < 3; ++k) {
        const double ext = hi[k] - lo[k];
        double t = ext > 0 ? (-cobjs[i].t[k] - lo[k]) / ext : 0.0;
        t = std::min(std::max(t, 0.0), 0.999999);
        q[k] = uint32_t(t * 1024.0);
      }
      return (uint64_t(expandBits10(q[0])) << 2) | (uint64_t(expandBits10(q[1])) << 1) | uint64_t(expandBits10(q[2]));
    };
    std::vector<std::pair<uint64_t, uint32_t>> order;
    for (uint32_t i : fast) order.push_back({key(i), i});
    std::sort(order.begin(), order.end());
    const int64_t g2 = kClusterSize, g1 = int64_t(kClusterSize) * kClusterSize;
    const int64_t n1 = (int64_t(order.size()) + g1 - 1) / g1;
    CObjF never; never.tx = never.ty = never.tz = 0.f; never.r2m = -float(NRT_INF);   // b^2 < a * (+Inf): always a certain miss
    std::vector<CObjF> clm(size_t(n1 * g1), never), cl2(size_t(n1 * g2), never), cl1(size_t(n1), never);
    std::vector<uint32_t> idx(size_t(n1 * g1), kInvalidRef);
    for (size_t k = 0; k < order.size(); ++k) { clm[k] = cobjf[order[k].second]; idx[k] = order[k].second; }
    // bounding sphere of a set of spheres (centre c, radius r): centre of their box, R = max(|c - C| + r)
    auto bound = [&](const std::vector<std::array<double, 4>>& sp) {
      double blo[3] = {1e300, 1e300, 1e300}, bhi[3] = {-1e300, -1e300, -1e300};
      for (auto& q : sp) for (int k = 0; k < 3; ++k) { blo[k] = std::min(blo[k], q[k] - q[3]); bhi[k] = std::max(bhi[k], q[k] + q[3]); }
      std::array<double, 4> b{0.5 * (blo[0] + bhi[0]), 0.5 * (blo[1] + bhi[1]), 0.5 * (blo[2] + bhi[2]), 0.0};
      for (auto& q : sp) {
        const double dx = q[0] - b[0], dy = q[1] - b[1], dz = q[2] - b[2];
        b[3] = std::max(b[3], std::sqrt(dx * dx + dy * dy + dz * dz) + q[3]);
      }
      b[3] *= 1.0 + 1e-9;
      return b;
    };
    for (int64_t a1 = 0; a1 < n1; ++a1) {
      std::vector<std::array<double, 4>> l2s;
      for (int64_t a2 = a1 * g2; a2 < (a1 + 1) * g2; ++a2) {
        std::vector<std::array<double, 4>> ms;
        for (int64_t m = a2 * g2; m < (a2 + 1) * g2 && m < int64_t(order.size()); ++m) {
          const CObj& c = cobjs[order[size_t(m)].second];
          ms.push_back({-c.t[0], -c.t[1], -c.t[2], std::fabs(c.radius)});
        }
        if (ms.empty()) continue;
        const auto b = bound(ms);
        cl2[size_t(a2)] = boundRecord(b.data(), b[3]);
        l2s.push_back(b);
      }
      const auto b = bound(l2s);
      cl1[size_t(a1)] = boundRecord(b.data(), b[3]);
    }
    const bool fit = reuse && n1 <= capCl1 && int64_t(slow.size()) <= capSlow;
    if (!fit) { capCl1 = n1; capSlow = std::max<int64_t>(1, int64_t(slow.size())); }
    dCl1 = up(cl1.data(), n1, fit ? dCl1 : nullptr);
    dCl2 = up(cl2.data(), n1 * g2, fit ? dCl2 : nullptr);
    dClm = up(clm.data(), n1 * g1, fit ? dClm : nullptr);
    dClmIdx = up(idx.data(), n1 * g1, fit ? dClmIdx : nullptr);
    if (slow.empty()) slow.push_back(0u);
    dSlowIdx = up(slow.data(), int64_t(slow.size()), fit ? dSlowIdx : nullptr);
    h.ncl1 = int32_t(n1);
    h.nslow = int32_t(std::count_if(cobjf.begin(), cobjf.end(), [](const CObjF& f) { return !(f.r2m < 3.0e38f); }));
  }

  // Light-space grids of the clustered spheres, one per DistantLight, and the camera grid (nrt_core.h: ShadowGridF).
  // Host, float64.
  std::vector<void*> gridOwned;
  ShadowGridF* dSGrid = nullptr;
  struct GridRect { double a1, b1, a2, b2; uint32_t obj; };   // conservative rectangle of an object on the grid's plane
  // bins the rectangles (ascending obj) into G x G cells from (lo1, lo2), cell edge hcell; false: too many items
  bool fillGrid(ShadowGridF& g, const std::vector<GridRect>& rects, double lo1, double lo2, double hcell, int G, double margin, bool masks = false) {
    g.lo1 = float(lo1); g.lo2 = float(lo2); g.invh = float(1.0 / hcell); g.margin = float(margin * 0.999); g.G = G;
    if (masks) {   // one 32-bit object mask per cell (scenes of <= 32 objects)
      const double flo1 = double(g.lo1), flo2 = double(g.lo2), finv = double(g.invh), pad = 1e-4 * hcell;
      std::vector<uint32_t> cell(size_t(G) * G, 0u);
      for (const GridRect& q : rects) {
        int a1 = int(std::floor((q.a1 - pad - flo1) * finv)), b1 = int(std::floor((q.b1 + pad - flo1) * finv));
        int a2 = int(std::floor((q.a2 - pad - flo2) * finv)), b2 = int(std::floor((q.b2 + pad - flo2) * finv));
        a1 = std::max(a1, 0); a2 = std::max(a2, 0); b1 = std::min(b1, G - 1); b2 = std::min(b2, G - 1);
        for (int y = a2; y <= b2; ++y) for (int x = a1; x <= b1; ++x) cell[size_t(y) * G + x] |= 1u << (q.obj & 31u);
      }
      uint32_t* p = static_cast<uint32_t*>(be->dalloc(sizeof(uint32_t) * cell.size()));
      gridOwned.push_back(p);
      be->upload(p, cell.data(), sizeof(uint32_t) * cell.size());
      bytes_uploaded += int64_t(sizeof(uint32_t) * cell.size());
      be->sync();
      g.start = nullptr; g.items = p;
      return true;
    }
    // (cell coordinates are computed by the rays as (p - float(lo)) * float(1/h): the build uses the same two floats,
    // and every rectangle is widened by 1e-4 h for the rounding of that expression)
    const double flo1 = double(g.lo1), flo2 = double(g.lo2), finv = double(g.invh), pad = 1e-4 * hcell;
    std::vector<uint32_t> count(size_t(G) * G + 1, 0);
    auto range = [&](const GridRect& q, int& a1, int& b1, int& a2, int& b2) {
      a1 = int(std::floor((q.a1 - pad - flo1) * finv)); b1 = int(std::floor((q.b1 + pad - flo1) * finv));
      a2 = int(std::floor((q.a2 - pad - flo2) * finv)); b2 = int(std::floor((q.b2 + pad - flo2) * finv));
      a1 = std::max(a1, 0); a2 = std::max(a2, 0); b1 = std::min(b1, G - 1); b2 = std::min(b2, G - 1);
    };
    int64_t total = 0;
    for (const GridRect& q : rects) {
      int a1, b1, a2, b2; range(q, a1, b1, a2, b2);
      for (int y = a2; y <= b2; ++y) for (int x = a1; x <= b1; ++x) { ++count[size_t(y) * G + x + 1]; ++total; }
    }
    if (total > int64_t(64) * int64_t(rects.size()) + 65536) return false;   // huge footprints: no grid
    for (size_t k = 1; k < count.size(); ++k) count[k] += count[k - 1];
    std::vector<uint32_t> items(size_t(std::max<int64_t>(total, 1)), 0), fill(count.begin(), count.end() - 1);
    for (const GridRect& q : rects) {   // rects ascend by obj, so every cell's list does
      int a1, b1, a2, b2; range(q, a1, b1, a2, b2);
      for (int y = a2; y <= b2; ++y) for (int x = a1; x <= b1; ++x) items[fill[size_t(y) * G + x]++] = q.obj;
    }
    auto gup = [&](const uint32_t* src, size_t n) {
      uint32_t* p = static_cast<uint32_t*>(be->dalloc(sizeof(uint32_t) * std::max<size_t>(n, 1)));
      gridOwned.push_back(p);
      if (n) { be->upload(p, src, sizeof(uint32_t) * n); bytes_uploaded += int64_t(sizeof(uint32_t) * n); }
      return p;
    };
    g.start = gup(count.data(), count.size());
    g.items = gup(items.data(), size_t(total));
    be->sync();   // (the staging vectors die with this scope)
    return true;
  }
  void buildShadowGrids(const nrt_scene_desc* desc) {
    for (void* p : gridOwned) be->dfree(p);
    gridOwned.clear();
    dSGrid = nullptr;
    h.sgrid = nullptr;
    h.maskGrids = 0; h.slowMask = 0;
    hSGrid.clear();
    // list grids for clustered scenes, mask grids for scenes of at most 32 objects, nothing in between (flat scan)
    const bool masks = h.ncl1 <= 0 && cobjf.size() <= 32;
    if (h.ncl1 <= 0 && !masks) return;
    // the bounding spheres the grids bin: the spheres the float32 first look applies to and, in mask grids, the mesh
    // boxes with a float32 gate record (centre and radius with the record's margins)
    struct Ball { double c[3], r; uint32_t obj; };
    std::vector<Ball> balls;
    std::vector<uint32_t> fast;
    uint32_t covered = 0;
    for (size_t i = 0; i < cobjf.size(); ++i)
      if (cobjf[i].r2m < 3.0e38f) {
        fast.push_back(uint32_t(i));
        balls.push_back(Ball{{-cobjs[i].t[0], -cobjs[i].t[1], -cobjs[i].t[2]}, std::fabs(cobjs[i].radius), uint32_t(i)});
        covered |= 1u << (i & 31u);
      }
    if (masks) {
      for (size_t mo = 0; mo < moIndex.size() && mo < hMGate.size(); ++mo) {
        const MeshGateF& mg = hMGate[mo];
        if (!(mg.valid > 0.f) || !(mg.r2m < 3.0e38f)) continue;
        balls.push_back(Ball{{double(mg.cx), double(mg.cy), double(mg.cz)}, std::sqrt(double(mg.r2m)) * (1.0 + 1e-6), uint32_t(moIndex[mo])});
        covered |= 1u << (uint32_t(moIndex[mo]) & 31u);
      }
      std::sort(balls.begin(), balls.end(), [](const Ball& a, const Ball& b) { return a.obj < b.obj; });
      h.slowMask = ~covered;
    }
    if (balls.empty()) return;
    std::vector<ShadowGridF> grids(lights.size() + 1, ShadowGridF{});
    bool any = false;
    for (size_t l = 0; l < lights.size(); ++l) {
      if (lights[l].kind != NRT_LIGHT_DISTANT) continue;
      // ray direction u = -dir (renderer.nim:99); (e1, e2) spans the plane perpendicular to it
      double u[3] = {-lights[l].dir[0], -lights[l].dir[1], -lights[l].dir[2]};
      const double ul = std::sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
      if (!(ul > 1e-3 && ul < 1e3) || lights[l].dir[3] != 0.0) continue;
      for (double& v : u) v /= ul;
      int ax = 0;
      if (std::fabs(u[1]) < std::fabs(u[ax])) ax = 1;
      if (std::fabs(u[2]) < std::fabs(u[ax])) ax = 2;
      double a[3] = {0, 0, 0}; a[ax] = 1.0;
      double e1[3] = {u[1] * a[2] - u[2] * a[1], u[2] * a[0] - u[0] * a[2], u[0] * a[1] - u[1] * a[0]};
      const double e1l = std::sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
      for (double& v : e1) v /= e1l;
      double e2[3] = {u[1] * e1[2] - u[2] * e1[1], u[2] * e1[0] - u[0] * e1[2], u[0] * e1[1] - u[1] * e1[0]};
      ShadowGridF g{};
      for (int k = 0; k < 3; ++k) { g.e1[k] = float(e1[k]); g.e2[k] = float(e2[k]); }
      // (the projections below use the float32 vectors the rays will use)
      const double f1[3] = {g.e1[0], g.e1[1], g.e1[2]}, f2[3] = {g.e2[0], g.e2[1], g.e2[2]};
      struct Circ { double p1, p2, r, c1; };
      std::vector<Circ> cs;
      cs.reserve(fast.size());
      double lo1 = 1e300, lo2 = 1e300, hi1 = -1e300, hi2 = -1e300;
      bool ok = true;
      for (const Ball& bl : balls) {
        const double* c = bl.c;
        Circ q;
        q.p1 = f1[0] * c[0] + f1[1] * c[1] + f1[2] * c[2];
        q.p2 = f2[0] * c[0] + f2[1] * c[1] + f2[2] * c[2];
        q.r = bl.r;
        q.c1 = std::fabs(c[0]) + std::fabs(c[1]) + std::fabs(c[2]);
        if (!(std::isfinite(q.p1) && std::isfinite(q.p2) && std::isfinite(q.r) && q.c1 < 1e12)) { ok = false; break; }
        lo1 = std::min(lo1, q.p1 - q.r); hi1 = std::max(hi1, q.p1 + q.r);
        lo2 = std::min(lo2, q.p2 - q.r); hi2 = std::max(hi2, q.p2 + q.r);
        cs.push_back(q);
      }
      if (!ok) continue;
      const double ext = std::max(hi1 - lo1, hi2 - lo2);
      if (!(ext > 0) || !std::isfinite(ext)) continue;
      const int G = masks ? 256 : int(std::min(1024.0, std::max(16.0, 4.0 * std::sqrt(double(cs.size())))));
      double hcell = ext * 1.001 / G;
      const double margin = 0.02 * hcell;
      // the whole grid is moved out by the largest circle inflation, so that a ray outside it hits nothing
      double infl = 0;
      for (const Circ& q : cs) infl = std::max(infl, q.r * 1e-6 + 4e-7 * q.c1 + margin);
      lo1 -= 2 * infl; lo2 -= 2 * infl;
      hcell = (ext + 4 * infl) * 1.001 / G;
      std::vector<GridRect> rects;
      rects.reserve(cs.size());
      for (size_t n = 0; n < cs.size(); ++n) {
        const double R = cs[n].r * (1.0 + 1e-6) + 4e-7 * cs[n].c1 + margin;
        rects.push_back(GridRect{cs[n].p1 - R, cs[n].p1 + R, cs[n].p2 - R, cs[n].p2 + R, balls[n].obj});
      }
      if (!fillGrid(g, rects, lo1, lo2, hcell, G, margin, masks)) continue;
      grids[l] = g;
      any = true;
    }
    // ---- the camera grid: rigid cameras only (the rows of the inverse rotation are then orthonormal, and a sphere
    // stays a sphere of its own radius in camera space)
    {
      const double* M = desc->camera_to_world;   // m[col*4+row]
      double E[3][3];   // rows of the inverse rotation = columns of the rotation; camera looks down -z: e3 = -(third column)
      for (int k = 0; k < 3; ++k) { E[0][k] = M[0 * 4 + k]; E[1][k] = M[1 * 4 + k]; E[2][k] = -M[2 * 4 + k]; }
      bool rigid = M[3] == 0.0 && M[7] == 0.0 && M[11] == 0.0 && M[15] == 1.0;
      for (int i = 0; i < 3 && rigid; ++i)
        for (int j = 0; j < 3; ++j) {
          const double dp = E[i][0] * E[j][0] + E[i][1] * E[j][1] + E[i][2] * E[j][2];
          if (!(std::fabs(dp - (i == j ? 1.0 : 0.0)) < 1e-9)) rigid = false;
        }
      const double fext = 2.5 * h.tan_half_fov;   // covers aspect ratios up to 2.5; other rays take the clusters
      if (rigid && std::isfinite(fext) && fext > 1e-3 && fext <= 5.0) {   // (tan(1.45) = 8.2 lies outside the grid)
        ShadowGridF g{};
        g.persp = 1;
        for (int k = 0; k < 3; ++k) { g.e1[k] = float(E[0][k]); g.e2[k] = float(E[1][k]); g.e3[k] = float(E[2][k]); }
        const int G = masks ? 512 : int(std::min(512.0, std::max(32.0, 8.0 * std::sqrt(double(fast.size())))));
        const double hcell = 2.0 * fext / G, margin = 0.02 * hcell;
        // a ray's (X, Y) carries <= ~1.4e-6 (1 + |X|) of float32 error (three products per dot, the division, e3.d >= 0.3 |d|)
        bool ok = margin >= 4e-6 * (1.0 + fext);
        std::vector<GridRect> rects;
        const double co[3] = {h.cam_orig[0], h.cam_orig[1], h.cam_orig[2]};
        for (size_t n = 0; n < balls.size() && ok; ++n) {
          const uint32_t i = balls[n].obj;
          const double c[3] = {balls[n].c[0] - co[0], balls[n].c[1] - co[1], balls[n].c[2] - co[2]};
          // camera-space centre through the float32 rows the rays use
          double v[3];
          for (int k = 0; k < 3; ++k) {
            const float* e = k == 0 ? g.e1 : (k == 1 ? g.e2 : g.e3);
            v[k] = double(e[0]) * c[0] + double(e[1]) * c[1] + double(e[2]) * c[2];
          }
          const double cl = std::fabs(c[0]) + std::fabs(c[1]) + std::fabs(c[2]);
          const double R = balls[n].r * (1.0 + 1e-5) + 1e-6 * cl;   // float32 rows: rigid only to ~1e-7
          if (!(std::isfinite(R) && std::isfinite(cl) && cl < 1e12)) { ok = false; break; }
          if (v[2] + R < 0.0) continue;   // entirely behind the camera plane: no forward ray reaches it
          // extent of the silhouette in X: tangent angles of the circle (v[0], v[2]; R) seen from the origin
          auto extent = [&](double x, double z, double& lo, double& hi) {
            const double rho = std::sqrt(x * x + z * z);
            lo = -fext; hi = fext;
            if (!(rho > R * 1.001)) return;   // the origin is inside (or on) the circle: every direction
            const double th = std::atan2(x, z), al = std::asin(R / rho);
            if (th - al > -1.45) lo = std::tan(th - al);
            if (th + al < 1.45) hi = std::tan(th + al);
            if (th - al >= 1.45) lo = fext * 2;    // entirely beyond the grid on the + side
            if (th + al <= -1.45) hi = -fext * 2;
          };
          GridRect q; q.obj = i;
          extent(v[0], v[2], q.a1, q.b1);
          extent(v[1], v[2], q.a2, q.b2);
          q.a1 -= margin; q.b1 += margin; q.a2 -= margin; q.b2 += margin;
          if (q.b1 < -fext || q.a1 > fext || q.b2 < -fext || q.a2 > fext) continue;   // outside the grid
          rects.push_back(q);
        }
        if (ok && fillGrid(g, rects, -fext, -fext, hcell, G, margin, masks)) { grids[lights.size()] = g; any = true; }
      }
    }
    if (!any) return;
    dSGrid = static_cast<ShadowGridF*>(be->dalloc(sizeof(ShadowGridF) * grids.size()));
    gridOwned.push_back(dSGrid);
    be->upload(dSGrid, grids.data(), sizeof(ShadowGridF) * grids.size());
    be->sync();
    h.sgrid = dSGrid;
    h.maskGrids = masks ? 1 : 0;
    hSGrid = grids;
  }

  // Validates and flattens `desc`; with `reuse` the existing device buffers are refilled.
  int build(BE* backend, const nrt_scene_desc* desc, bool reuse, std::string& err) {
    be = backend;
    if (!desc || desc->nobjects < 0 || desc->nlights < 0 || desc->nmeshes < 0 ||
        (desc->nobjects > 0 && !desc->objects) || (desc->nlights > 0 && !desc->lights) ||
        (desc->nmeshes > 0 && !desc->meshes)) { err = "null or negative-sized scene description"; return NRT_ERR_INVALID; }
    if (reuse && (desc->nobjects != h.nobjects || desc->nlights != h.nlights || desc->nmeshes != h.nmeshes)) {
      err = "nrt_scene_update: scene shape differs from the created scene"; return NRT_ERR_INVALID;
    }
    // Validate the WHOLE description before anything of the live scene is touched: a failed
    // nrt_scene_update leaves the scene exactly as it was (and renderable).
    for (int i = 0; i < desc->nmeshes; ++i) {
      const nrt_mesh& m = desc->meshes[i];
      if (m.nverts < 0 || m.nfaces < 0 || m.nnormals < 0 || (m.nfaces > 0 && (!m.vertices || !m.vertex_idx || !m.normal_idx)) ||
          (m.nnormals > 0 && !m.normals) || m.nfaces > 0xFFFFFFF0ll) { err = "bad mesh description"; return NRT_ERR_INVALID; }
      if (reuse && (m.nverts != meshes[i].nverts || m.nfaces != meshes[i].nfaces || m.nnormals != meshes[i].nnormals)) {
        err = "nrt_scene_update: mesh shape differs"; return NRT_ERR_INVALID;
      }
      for (int64_t f = 0; f < m.nfaces * 3; ++f) {
        if (m.vertex_idx[f] < 0 || m.vertex_idx[f] >= m.nverts) { err = "vertex index out of range"; return NRT_ERR_INVALID; }
        // only normalIdx[0] of a face is ever read (renderer.nim:87)
        if (f % 3 == 0 && (m.normal_idx[f] < 0 || m.normal_idx[f] >= m.nnormals)) { err = "normal index out of range"; return NRT_ERR_INVALID; }
      }
    }
    for (int i = 0; i < desc->nobjects; ++i) {
      const nrt_object& o = desc->objects[i];
      if (o.kind < 0 || o.kind > 3) { err = "bad geometry kind"; return NRT_ERR_INVALID; }
      if (o.kind == NRT_GEOM_MESH && (o.mesh < 0 || o.mesh >= desc->nmeshes)) { err = "mesh index out of range"; return NRT_ERR_INVALID; }
      // the reused buffers (record sets per mesh object, frames, counters) are sized by which objects are
      // meshes and by the mesh each one refers to: an update must keep both
      if (reuse && (o.kind != objs[i].kind || (o.kind == NRT_GEOM_MESH && o.mesh != objs[i].mesh))) {
        err = "nrt_scene_update: object kind or mesh index differs from the created scene"; return NRT_ERR_INVALID;
      }
    }
    for (int i = 0; i < desc->nlights; ++i) {
      const int k = desc->lights[i].kind;
      if (k != NRT_LIGHT_DISTANT && k != NRT_LIGHT_POINT) { err = "bad light kind"; return NRT_ERR_INVALID; }
      // DIR record sets exist per DistantLight: the kind of a light is part of the scene's shape too
      if (reuse && k != lights[i].kind) { err = "nrt_scene_update: light kind differs from the created scene"; return NRT_ERR_INVALID; }
    }
    bytes_uploaded = 0;
    if (reuse && sameSmallParts(desc) && std::getenv("NRT_UPDATE_ALWAYS_REBUILD") == nullptr) {
      stage.resize(size_t(desc->nmeshes));
      be->diffBegin();
      for (int i = 0; i < desc->nmeshes; ++i) {
        const nrt_mesh& m = desc->meshes[i];
        MeshStage& st = stage[size_t(i)];
        st.verts = up(m.vertices, m.nverts * 4, st.verts);     be->diffAdd(st.verts, meshes[i].verts, sizeof(double) * size_t(m.nverts) * 4);
        st.normals = up(m.normals, m.nnormals * 4, st.normals); be->diffAdd(st.normals, meshes[i].normals, sizeof(double) * size_t(m.nnormals) * 4);
        st.vidx = up(m.vertex_idx, m.nfaces * 3, st.vidx);      be->diffAdd(st.vidx, meshes[i].vidx, sizeof(int64_t) * size_t(m.nfaces) * 3);
        st.nidx = up(m.normal_idx, m.nfaces * 3, st.nidx);      be->diffAdd(st.nidx, meshes[i].nidx, sizeof(int64_t) * size_t(m.nfaces) * 3);
      }
      if (!be->diffEnd()) { ++rebuildsSkipped; return NRT_OK; }   // identical: the resident records are this description's
    }
    ++rebuilds;
    std::vector<DMesh> old = meshes;
    std::vector<MoRecs> oldRecs = moRecs;
    objs.assign(desc->nobjects, DObject{});
    lights.assign(desc->nlights, DLight{});
    meshes.assign(desc->nmeshes, DMesh{});
    moIndex.clear();
    for (int i = 0; i < desc->nmeshes; ++i) {
      const nrt_mesh& m = desc->meshes[i];
      DMesh& dm = meshes[i];
      dm.nverts = m.nverts; dm.nnormals = m.nnormals; dm.nfaces = m.nfaces;
      dm.verts = up(m.vertices, m.nverts * 4, reuse ? const_cast<double*>(old[i].verts) : nullptr);
      dm.normals = up(m.normals, m.nnormals * 4, reuse ? const_cast<double*>(old[i].normals) : nullptr);
      dm.vidx = up(m.vertex_idx, m.nfaces * 3, reuse ? const_cast<int64_t*>(old[i].vidx) : nullptr);
      dm.nidx = up(m.normal_idx, m.nfaces * 3, reuse ? const_cast<int64_t*>(old[i].nidx) : nullptr);
      dm.recs = up<float>(nullptr, paddedFaces(m.nfaces) * kFullStride, reuse ? old[i].recs : nullptr);
      dm.hot = up<float>(nullptr, paddedFaces(m.nfaces) * hotFloats(FM_GENERAL), reuse ? old[i].hot : nullptr);
      dm.bounds = up<float>(nullptr, std::max<int64_t>(1, numChunks(m.nfaces)) * 4 * (1 + kSubPerChunk), reuse ? old[i].bounds : nullptr);
      dm.order = up<uint32_t>(nullptr, std::max<int64_t>(1, m.nfaces), reuse ? old[i].order : nullptr);
      calcAABB(m.vertices, m.nverts, dm.bmin, dm.bmax);
      double L = 0;
      for (int k = 0; k < 3; ++k) {
        dm.center[k] = 0.5 * (dm.bmin[k] + dm.bmax[k]);
        L = std::max(L, 0.5 * (dm.bmax[k] - dm.bmin[k]));
      }
      if (!(L > 0) || !std::isfinite(L)) { L = 0; for (int k = 0; k < 3; ++k) dm.center[k] = 0; }  // => every ray takes the exact path
      dm.L = L;
      {
        double rb2 = 0;
        for (int k = 0; k < 3; ++k) { const double hk = 0.5 * (dm.bmax[k] - dm.bmin[k]); rb2 += hk * hk; }
        rb2 *= 1.0 + 1e-6;   // covers the rounding of center[] and of the sum
        dm.rb2f = (std::isfinite(rb2) && rb2 < 1e30 && L > 0) ? roundUpF(rb2) : float(NRT_INF);
        dm._padf = 0.f;
      }
    }
    anyReflective = false; anyPointLight = false;
    for (int i = 0; i < desc->nobjects; ++i) {
      const nrt_object& o = desc->objects[i];
      DObject& dob = objs[i];
      dob.kind = o.kind; dob.mesh = -1; dob.mesh_obj = -1;
      {  // exactly [I | t] with finite t?
        const double* w = o.world_to_object;
        bool x = true;
        for (int c = 0; c < 4 && x; ++c)
          for (int r = 0; r < 4 && x; ++r) {
            const double v = w[c * 4 + r];
            if (c < 3 || r == 3) x = (v == ((c == r) ? 1.0 : 0.0));
            else x = std::isfinite(v);
          }
        dob.xlate_only = x ? 1 : 0;
      }
      std::memcpy(dob.o2w, o.object_to_world, sizeof(dob.o2w));
      std::memcpy(dob.w2o, o.world_to_object, sizeof(dob.w2o));
      dob.radius = o.radius;
      std::memcpy(dob.bmin, o.vmin, sizeof(dob.bmin));
      std::memcpy(dob.bmax, o.vmax, sizeof(dob.bmax));
      std::memcpy(dob.albedo, o.albedo, sizeof(dob.albedo));
      for (int k = 0; k < 3; ++k) dob.albedo_pi[k] = o.albedo[k] / kPi;
      { const V4 nw = mulm(dob.o2w, v4(0.0, 1.0, 0.0, 0.0)); dob.plane_nw[0] = nw.x; dob.plane_nw[1] = nw.y; dob.plane_nw[2] = nw.z; dob.plane_nw[3] = nw.w; }
      dob.reflection = o.reflection;
      if (o.reflection > 0.0) anyReflective = true;
      if (o.kind == NRT_GEOM_MESH) {
        dob.mesh = o.mesh;
        dob.mesh_obj = int32_t(moIndex.size());
        moIndex.push_back(i);
        std::memcpy(dob.bmin, meshes[o.mesh].bmin, sizeof(dob.bmin));
        std::memcpy(dob.bmax, meshes[o.mesh].bmax, sizeof(dob.bmax));
      }
    }
    for (int i = 0; i < desc->nlights; ++i) {
      const nrt_light& l = desc->lights[i];
      DLight& dl = lights[i];
      dl.kind = l.kind;
      if (l.kind == NRT_LIGHT_POINT) anyPointLight = true;
      std::memcpy(dl.color, l.color, sizeof(dl.color));
      dl.intensity = l.intensity;
      std::memcpy(dl.dir, l.dir, sizeof(dl.dir));
      std::memcpy(dl.pos, l.pos, sizeof(dl.pos));
    }
    cobjs.assign(objs.size(), CObj{});
    for (size_t i = 0; i < objs.size(); ++i) {
      CObj& c = cobjs[i];
      c.kind = objs[i].kind; c.mesh_obj = objs[i].mesh_obj; c.xlate_only = objs[i].xlate_only; c._pad = 0;
      c.t[0] = objs[i].w2o[12]; c.t[1] = objs[i].w2o[13]; c.t[2] = objs[i].w2o[14];
      c.radius = objs[i].radius;
    }
    cobjf.assign(objs.size(), CObjF{});
    for (size_t i = 0; i < objs.size(); ++i) {
      CObjF& f = cobjf[i];
      const CObj& c = cobjs[i];
      const double mt = std::max(std::fabs(c.t[0]), std::max(std::fabs(c.t[1]), std::fabs(c.t[2])));
      const double r2 = c.radius * c.radius;
      const bool finiteT = std::isfinite(c.t[0]) && std::isfinite(c.t[1]) && std::isfinite(c.t[2]) && mt < 1e15;
      f.tx = bitsToFloat(uint32_t(COF_SLOW)); f.ty = 0.f; f.tz = 0.f; f.r2m = float(NRT_INF);
      if (c.kind == NRT_GEOM_SPHERE && c.xlate_only && finiteT && std::isfinite(c.radius) && r2 < 1e30 && r2 > 1e-30) {
        f.tx = float(c.t[0]); f.ty = float(c.t[1]); f.tz = float(c.t[2]);
        f.r2m = roundUpF(r2 + 2e-6 * r2 + 2e-7 * mt * mt);
      } else if (c.kind == NRT_GEOM_PLANE && c.xlate_only && finiteT) {
        f.tx = bitsToFloat(uint32_t(COF_PLANE)); f.ty = float(c.t[1]); f.tz = roundUpF(2.5e-7 * std::fabs(c.t[1]) + 1e-30);
      } else if (c.kind == NRT_GEOM_MESH) {
        f.tx = bitsToFloat(uint32_t(COF_MESH)); f.ty = bitsToFloat(uint32_t(c.mesh_obj));
      }
    }
    dCObjF = up(cobjf.data(), int64_t(cobjf.size()), reuse ? dCObjF : nullptr);
    buildClusters(reuse);
    dCObjs = up(cobjs.data(), int64_t(cobjs.size()), reuse ? dCObjs : nullptr);
    dObjs = up(objs.data(), int64_t(objs.size()), reuse ? dObjs : nullptr);
    dLights = up(lights.data(), int64_t(lights.size()), reuse ? dLights : nullptr);
    dMeshes = up(meshes.data(), int64_t(meshes.size()), reuse ? dMeshes : nullptr);
    dMo = up(moIndex.data(), int64_t(moIndex.size()), reuse ? dMo : nullptr);
    h.nobjects = desc->nobjects; h.nlights = desc->nlights; h.nmeshes = desc->nmeshes;
    h.nmesh_objs = int32_t(moIndex.size());
    // ---- bundle frames (float64, host): projection frames of the ORIGIN / DIR bundles ----
    {
      const int nMOf = int(moIndex.size()), rsf = 2 + desc->nlights;
      frames.assign(size_t(std::max(1, nMOf * rsf)), BundleFrame{});
      anyGeneralShadow = anyPointLight;
      for (int mo = 0; mo < nMOf; ++mo) {
        const DObject& ob = objs[moIndex[mo]];
        const DMesh& m = meshes[ob.mesh];
        BundleFrame& fg = frames[mo * rsf];
        for (int k = 0; k < 3; ++k) fg.org[k] = m.center[k];
        fg.valid = 1.0;
        // ORIGIN: shared origin = trace()'s object-space image of castPrimaryRay's origin; axis = camera forward
        BundleFrame& fo = frames[mo * rsf + 1];
        const V4 ow = mulm(desc->camera_to_world, v4(0.0, 0.0, 0.0, 1.0));
        const V4 oo = mulm(ob.w2o, ow);
        const V4 fw = mulm(ob.w2o, mulm(desc->camera_to_world, v4(0.0, 0.0, -1.0, 0.0)));
        fo.org[0] = oo.x; fo.org[1] = oo.y; fo.org[2] = oo.z;
        const double fa[3] = {fw.x, fw.y, fw.z};
        fo.valid = (makeBasis(fa, fo) && std::isfinite(oo.x) && std::isfinite(oo.y) && std::isfinite(oo.z)) ? 1.0 : 0.0;
        for (int l = 0; l < desc->nlights; ++l) {
          BundleFrame& fd = frames[mo * rsf + 2 + l];
          for (int k = 0; k < 3; ++k) fd.org[k] = m.center[k];
          fd.valid = 0.0;
          if (lights[l].kind != NRT_LIGHT_DISTANT) continue;
          const V4 dw = scale(v4(lights[l].dir[0], lights[l].dir[1], lights[l].dir[2], lights[l].dir[3]), -1.0);
          const V4 dobj = mulm(ob.w2o, dw);
          const double da[3] = {dobj.x, dobj.y, dobj.z};
          fd.valid = makeBasis(da, fd) ? 1.0 : 0.0;
          if (!(fd.valid > 0)) anyGeneralShadow = true;
        }
      }
      {   // float32 gate records (nrt_core.h: MeshGateF)
        std::vector<MeshGateF> mg(size_t(std::max(1, nMOf)), MeshGateF{});
        for (int mo = 0; mo < nMOf; ++mo) {
          const DObject& ob = objs[moIndex[mo]];
          const DMesh& m = meshes[ob.mesh];
          MeshGateF& g = mg[size_t(mo)];
          const double c[3] = {m.center[0] - ob.w2o[12], m.center[1] - ob.w2o[13], m.center[2] - ob.w2o[14]};
          const double mc = std::max(std::fabs(c[0]), std::max(std::fabs(c[1]), std::fabs(c[2])));
          const double rb2 = double(m.rb2f);
          if (ob.xlate_only && std::isfinite(mc) && mc < 1e15 && rb2 < 1e30 && m.L > 0) {
            g.cx = float(c[0]); g.cy = float(c[1]); g.cz = float(c[2]);
            g.r2m = roundUpF(rb2 * (1.0 + 2e-6) + 2e-7 * mc * mc);
            g.mm = roundUpF(2e-7 * mc * mc);
            g.valid = 1.f;
          }
        }
        dMGate = up(mg.data(), int64_t(mg.size()), reuse ? dMGate : nullptr);
        hMGate = mg;
      }
      dFrames = up(frames.data(), int64_t(frames.size()), reuse ? dFrames : nullptr);
      dRecSets = up<RecSet>(nullptr, int64_t(frames.size()), reuse ? dRecSets : nullptr);   // filled once the records exist
    }
    h.cl1 = dCl1; h.cl2 = dCl2; h.clm = dClm; h.clmIdx = dClmIdx; h.slowIdx = dSlowIdx;
    h.objects = dObjs; h.cobjs = dCObjs; h.cobjf = dCObjF; h.lights = dLights; h.meshes = dMeshes; h.mesh_obj_index = dMo; h.frames = dFrames; h.recsets = dRecSets; h.mgate = dMGate;
    std::memcpy(h.c2w, desc->camera_to_world, sizeof(h.c2w));
    { const V4 co = mulm(h.c2w, v4(0.0, 0.0, 0.0, 1.0)); h.cam_orig[0] = co.x; h.cam_orig[1] = co.y; h.cam_orig[2] = co.z; h.cam_orig[3] = co.w; }
    h.tan_half_fov = std::tan((desc->fov * (kPi / 180.0)) / 2);  // renderer.nim:38; Nim degToRad = d * (PI/180)
    std::memcpy(h.bg, desc->bg_color, sizeof(h.bg));
    buildShadowGrids(desc);   // (sets h.sgrid; needs h.cam_orig / h.tan_half_fov)
    h.hotOk = 0;
    d = up(&h, 1, reuse ? d : nullptr);
    // the by-value copy of the header (FusedBounceT) carries the small tables itself (nrt_core.h: DScene::hotOk)
    if (int(lights.size()) <= kHotLights && int(moIndex.size()) <= kHotMO && hMGate.size() >= moIndex.size() &&
        (h.sgrid == nullptr || hSGrid.size() == lights.size() + 1) && envHot()) {
      for (size_t l = 0; l < lights.size(); ++l) h.lightv[l] = lights[l];
      for (size_t g = 0; g < lights.size() + 1; ++g) h.sgridv[g] = (h.sgrid != nullptr) ? hSGrid[g] : ShadowGridF{};
      for (size_t mo = 0; mo < moIndex.size(); ++mo) { h.mgatev[mo] = hMGate[mo]; h.moIndexv[mo] = moIndex[mo]; }
      h.hotOk = 1;
    }
    // ---- filter records: GENERAL per mesh; ORIGIN / DIR per mesh object (device-side build) ----
    for (auto& m : meshes)
      if (m.nfaces > 0) be->sortFaces(m);   // fills m.order (Morton order of the face centroids)
    dMeshes = up(meshes.data(), int64_t(meshes.size()), dMeshes);   // (the device copy was uploaded above with the same pointers)
    for (auto& m : meshes)
      if (m.nfaces > 0) be->forEach(paddedFaces(m.nfaces), BuildRecsGeneral{m});
    const int nMO = int(moIndex.size()), rs = recStride();
    moRecs.assign(nMO, MoRecs{});
    hRecCount.assign(size_t(std::max(1, nMO * rs)), 0u);
    for (int mo = 0; mo < nMO; ++mo) hRecCount[mo * rs] = uint32_t(meshes[objs[moIndex[mo]].mesh].nfaces);
    dRecCount = up(hRecCount.data(), int64_t(hRecCount.size()), reuse ? dRecCount : nullptr);
    for (int mo = 0; mo < nMO; ++mo) {
      const int64_t nf = meshes[objs[moIndex[mo]].mesh].nfaces;
      MoRecs& r = moRecs[mo];
      r.origin = up<float>(nullptr, paddedFaces(nf) * kFullStride, reuse ? oldRecs[mo].origin : nullptr);
      r.originHot = up<float>(nullptr, paddedFaces(nf) * hotFloats(FM_ORIGIN), reuse ? oldRecs[mo].originHot : nullptr);
      r.originBounds = up<float>(nullptr, std::max<int64_t>(1, numChunks(nf)) * 4 * (1 + kSubPerChunk), reuse ? oldRecs[mo].originBounds : nullptr);
      r.dirBounds.assign(size_t(desc->nlights), nullptr);
      r.dir.assign(size_t(desc->nlights), nullptr);
      r.dirHot.assign(size_t(desc->nlights), nullptr);
      if (nf > 0 && frameValid(mo, FM_ORIGIN, 0))
        be->compactRecs(nf, BuildRecsOrigin{d, mo}, r.origin, r.originHot, FM_ORIGIN, dRecCount + mo * rs + 1);
      for (int l = 0; l < desc->nlights; ++l) {
        if (lights[l].kind != NRT_LIGHT_DISTANT || !frameValid(mo, FM_DIR, l)) continue;
        r.dir[l] = up<float>(nullptr, paddedFaces(nf) * kFullStride, reuse ? oldRecs[mo].dir[l] : nullptr);
        r.dirHot[l] = up<float>(nullptr, paddedFaces(nf) * hotFloats(FM_DIR), reuse ? oldRecs[mo].dirHot[l] : nullptr);
        r.dirBounds[l] = up<float>(nullptr, std::max<int64_t>(1, numChunks(nf)) * 4 * (1 + kSubPerChunk), reuse ? oldRecs[mo].dirBounds[l] : nullptr);
        if (nf > 0) be->compactRecs(nf, BuildRecsDir{d, mo, l}, r.dir[l], r.dirHot[l], FM_DIR, dRecCount + mo * rs + 2 + l);
      }
    }
    // chunk bounds of every record set (after the records exist)
    for (int mo = 0; mo < nMO; ++mo) {
      const DMesh& m = meshes[objs[moIndex[mo]].mesh];
      const int64_t nch = numChunks(m.nfaces);
      if (nch == 0) continue;
      be->forEach(nch * (1 + kSubPerChunk), BuildBounds{FM_GENERAL, m.hot, dRecCount + mo * rs, m.bounds, nch});
      if (frameValid(mo, FM_ORIGIN, 0)) be->forEach(nch * (1 + kSubPerChunk), BuildBounds{FM_ORIGIN, moRecs[mo].originHot, dRecCount + mo * rs + 1, moRecs[mo].originBounds, nch});
      for (int l = 0; l < desc->nlights; ++l)
        if (moRecs[mo].dirHot[l]) be->forEach(nch * (1 + kSubPerChunk), BuildBounds{FM_DIR, moRecs[mo].dirHot[l], dRecCount + mo * rs + 2 + l, moRecs[mo].dirBounds[l], nch});
    }
    be->download(hRecCount.data(), dRecCount, sizeof(uint32_t) * hRecCount.size());
    {   // the record-set table of the per-thread mesh walk (meshIntersectWalk)
      std::vector<RecSet> sets(frames.size(), RecSet{});
      for (int mo = 0; mo < nMO; ++mo) {
        const DMesh& m = meshes[objs[moIndex[mo]].mesh];
        for (int j = 0; j < rs; ++j) {
          const int mode = j == 0 ? FM_GENERAL : (j == 1 ? FM_ORIGIN : FM_DIR), l = j >= 2 ? j - 2 : 0;
          RecSet& r = sets[size_t(mo * rs + j)];
          const float* hot = hotOf(mo, mode, l);
          r.usable = (m.nfaces > 0 && hot && (mode == FM_GENERAL || frameValid(mo, mode, l))) ? 1u : 0u;
          if (!r.usable) continue;
          r.hot = hot; r.bounds = boundsOf(mo, mode, l); r.sub = r.bounds + 4 * numChunks(m.nfaces);
          r.recs = recsOf(mo, mode, l); r.ids = mode == FM_GENERAL ? m.order : nullptr;
          r.nrec = hRecCount[size_t(mo * rs + j)];
        }
      }
      be->upload(dRecSets, sets.data(), sizeof(RecSet) * sets.size());
      bytes_uploaded += int64_t(sizeof(RecSet) * sets.size());
    }
    rememberSmallParts(desc);
    return NRT_OK;
  }

  void destroy() {
    for (void* p : owned) be->dfree(p);
    owned.clear();
    for (void* p : gridOwned) be->dfree(p);
    gridOwned.clear();
  }
};

// --------------------------------------------------------------- renderer ----
template <class BE>
struct Renderer {
  BE* be = nullptr;
  ChunkState cs{};
  int64_t capS = 0, capNR = 0, capCand = 0, capPairs = 0;
  int64_t wantedS = 0;   // the chunk-size request the current buffers were sized for (before the memory cap)
  int capMO = -1, capNL = -1, capWaves = 0, capRows = 0;
  std::vector<void*> owned;
  int32_t* dRows = nullptr;
  int64_t tailBelow = 32768;   // NRT_TAIL_BELOW: active lists shorter than this are finished by one PathTail launch
  // NRT_HARD_TAIL_BELOW: a bounce's wavefront list shorter than this is finished by one PathTail launch as well (the
  // ~22 dependent launches of a wavefront bounce cost ~0.35 ms however few samples they carry: measured on a 1/8 frame)
  int64_t hardTailBelow = 16384;
  // ---- the fork.  After bounce 0 a frame's samples are two disjoint pools: H0, the samples with a mesh ray (one
  // long chain of wavefront kernels), and C0, the samples FusedBounce finished with a reflection ray stored (bounce 1:
  // FusedBounce, a short wavefront chain, ... PathTail).  The two do not depend on each other, and the later bounces'
  // chains are latency-bound (a 1/8 frame: ~1.1 ms for 13 % of the samples), so C0 is moved (GatherPool) into the sample
  // space of a HELPER pipeline — its own buffers, stream and host thread — which takes it to the end of its paths
  // while this pipeline runs H0's chain; the accumulators come back (ScatterAccum) before Finalize.
  Renderer* sub = nullptr;          // the helper (set by the owner; null: no fork)
  // NRT_FORK_MIN: pools smaller than this stay here; 0 (the default) = never fork.  MEASURED AND LEFT OFF: the chains'
  // kernels are persistent grids that occupy every SM while they wait on memory, so two chains in flight mostly
  // take turns (1/8 frame: 3.56 -> 3.39 ms with one lane, nothing on top of three lanes; whole frame: 19.8 -> 20.4 ms)
  int64_t forkMin = 0;
  struct SubResult { int rc = 0; bool overflow = false; std::string err; unsigned long long stats[ST_COUNT] = {0}; ProfileAcc pacc; int64_t n = 0; };
  bool wantBandCounts = false;      // set by the owner when the next frame's lane plan will read bandHard
  std::vector<uint32_t> bandHard;   // per row unit of the last frame: samples on the bounce-0 wavefront list (fused path; else empty)
  // NRT_PATH: 0 = the wavefront for every bounce (round-1 pipeline), 1 = FusedBounce + wavefront for the samples
  // with mesh rays at bounce 0 + PathTail (default), 2 = PathMega (one thread per sample start to end)
  int pathMode = 1;
  double meshShare = 0.0;   // see render(): picks the path of the next frame when NRT_PATH is not set
  int autoMode = 1;
  ProfileAcc prof;
  int64_t launches_hint = 0;
  bool shadowGatePerSample = envInt("NRT_SHADOW_GATE_PER_SAMPLE", 1) != 0;
  bool shadowTracePerSample = envInt("NRT_SHADOW_TRACE_PER_SAMPLE", 1) != 0;
  bool fuseResolve = envInt("NRT_FUSE_RESOLVE", 1) != 0;   // ShadowTrace + Resolve in one launch (per-sample shadow kernels, <= 32 lights)
  // one entry per prefilter launch of the last frame, in launch order (NRT_TRACE_PREFILTER)
  struct PreLaunch { int wave, mo, b, mode; int64_t rays, work, pre, nch; };
  std::vector<PreLaunch> preLog;

  void freeAll() {
    for (void* p : owned) be->dfree(p);
    owned.clear();
    capS = capNR = capCand = capPairs = 0; capMO = -1; capNL = -1; capWaves = 0; capRows = 0; dRows = nullptr;
    wantedS = 0;
  }
  template <class T> T* al(int64_t n) { T* p = static_cast<T*>(be->dalloc(sizeof(T) * std::max<int64_t>(n, 1))); owned.push_back(p); return p; }

  static int64_t envInt(const char* name, int64_t dflt) {
    const char* v = std::getenv(name);
    return (v && *v) ? std::atoll(v) : dflt;
  }

  void ensure(int64_t S, int nL, int nMO, int waves, int64_t cand, int nrows, int64_t pairsReq) {
    const int64_t NR = S * std::max(1, nL);
    const int64_t pairsWant = pairsReq > 0 ? pairsReq : std::max<int64_t>(int64_t(1) << 20, cand / 8);   // ~0.03 pairs per queued ray on the bunny scenes
    if (S > capS || NR > capNR || nMO > capMO || nL != capNL || waves > capWaves || cand > capCand || nrows > capRows || pairsWant > capPairs ||
        (pairsReq > 0 && pairsReq != capPairs)) {
      freeAll();
      capS = S; capNR = NR; capMO = nMO; capNL = nL; capWaves = waves; capCand = cand; capRows = nrows;
      cs.rayO = al<double>(4 * S); cs.rayD = al<double>(4 * S); cs.hitW = al<double>(4 * S); cs.nrm = al<double>(4 * S);
      cs.accum = al<double>(3 * S); cs.weight = al<double>(S);
      cs.hitObj = al<int32_t>(S); cs.active = al<uint8_t>(S);
      const int64_t m = int64_t(std::max(nMO, 1)) * NR;
      const int64_t qcap = NR + int64_t(nL) * S, mq = int64_t(std::max(nMO, 1)) * qcap;
      cs.tBest = al<uint64_t>(m); cs.triBest = al<uint32_t>(m);
      cs.qref = al<uint32_t>(mq); cs.qray0 = al<float>(mq * 4); cs.qray1 = al<float>(m * 4); cs.xref = al<uint32_t>(m);
      cs.qhot0 = al<float>(mq * 4); cs.qhot1 = al<float>(m * 4);
      cs.preRay = al<uint32_t>(4 * cand); cs.preRec = al<uint32_t>(4 * cand);
      capPairs = pairsWant;
      cs.pairs = al<uint32_t>(2 * capPairs);
      cs.runc = al<float>(4 * (qcap / 128 + 8));   // (a run is >= 128 queue entries)
      cs.candRef = al<uint32_t>(cand); cs.candTri = al<uint32_t>(cand); cs.candT = al<double>(cand);
      cs.counters = al<uint32_t>(int64_t(waves) * std::max(nMO, 1) * cntStride(nL) + nrows);   // + the per-band counts (BandCount) behind the wave counters
      cs.alist = al<uint32_t>(2 * S); cs.hlist = al<uint32_t>(S); cs.flist = al<uint32_t>(S); cs.acount = al<uint32_t>(waves + 4);
      cs.stats = al<unsigned long long>(ST_COUNT);
      cs.gvb = (NR + 255) / 256 + 1;
      cs.gsn = (cs.gvb + 255) / 256;
      cs.gflag = al<uint8_t>(m);
      cs.occ = al<uint8_t>(NR);
      const int64_t grows = int64_t(std::max(nMO, 1)) * (2 + nL);
      cs.gcnt = al<uint32_t>(grows * cs.gvb);
      cs.gseg = al<uint32_t>(grows * cs.gsn);
      cs.gsegBase = al<uint32_t>(grows * cs.gsn);
      cs.gne = al<uint32_t>(cs.gvb);
      be->zero(cs.gseg, sizeof(uint32_t) * grows * cs.gsn);
      dRows = al<int32_t>(nrows);
    }
    cs.pairCap = capPairs;
    cs.S = capS; cs.NR = capNR; cs.QCAP = capNR + int64_t(nL) * capS; cs.nMO = nMO; cs.nL = nL; cs.candCap = capCand; cs.preCap = 4 * capCand; cs.rows = dRows;
  }

  Gate makeGate(const SceneData<BE>& sd, const FrameParams& fp, int kind, const ActiveSet& act, int bounce, int force_exact) const {
    return Gate{NRT_SCENE_ARG(sd), fp, cs, kind, act, force_exact, (bounce == 0 && kind == WAVE_PATH) ? FM_ORIGIN : FM_GENERAL, bounce};
  }
  uint32_t* waveCounters(int wave) const { return cs.counters + int64_t(wave) * cs.nMO * cntStride(cs.nL); }

  // One mesh wave: gate + per mesh object { filter per ray bundle, exact, verify }.  `gated`: the
  // kernel that produced the wave's rays already evaluated and counted the gate codes (produceGate).
  void meshWave(const SceneData<BE>& sd, const FrameParams& fp, int kind, const ActiveSet& act, int wave, int bounce, int force_exact, bool gated) {
    const bool primary = bounce == 0 && kind == WAVE_PATH;
    const int nMO = cs.nMO, nL = cs.nL, cst = cntStride(nL);
    if (nMO == 0 || (!act.list && act.n == 0)) return;
    uint32_t* cnt = waveCounters(wave);
    const int mult = (kind == WAVE_SHADOW) ? nL : 1;
    const Gate g = makeGate(sd, fp, kind, act, bounce, force_exact);
    if (gated) be->gateFinish(g, act.n * mult, nMO, cnt);
    else if (kind == WAVE_SHADOW && shadowGatePerSample && be->produceGateFits(nL, nMO, nL)) {
      be->produceGate(act.n, nL, ShadowGate{g, nMO}, cs, nMO, cnt, nullptr);
      be->gateFinish(g, act.n * mult, nMO, cnt);
    }
    else be->gate(g, nullptr, act.n * mult, mult, nMO, cnt);
    for (int mo = 0; mo < nMO; ++mo) {
      uint32_t* c = cnt + mo * cst;
      const DMesh& m = sd.meshes[sd.objs[sd.moIndex[mo]].mesh];
      if (m.nfaces == 0) continue;
      if (!force_exact) {
        // prefilter (hot) -> refine (float32 sign test) per ray bundle; both feed the candidate list of (wave, mo)
        auto run = [&](int mode, int l, int b) {
          preLog.push_back(PreLaunch{wave, mo, b, mode, 0, 0, 0, 0});
          be->filter(mode, sd.hotOf(mo, mode, l), sd.boundsOf(mo, mode, l), sd.boundsOf(mo, mode, l) + 4 * SceneData<BE>::numChunks(m.nfaces), sd.recCountOf(mo, mode, l), cs, mo, b, c);
          be->forEachCounted(c + cntPre(b), cs.preCap,
                             Refine<typename BE::Atom>{cs, mode, sd.recsOf(mo, mode, l), mode == FM_GENERAL ? m.order : nullptr,
                                                       mo, b, c + CNT_CAND});
        };
        if (kind == WAVE_PATH) {
          const int mode = (primary && sd.frameValid(mo, FM_ORIGIN, 0)) ? FM_ORIGIN : FM_GENERAL;
          run(mode, 0, 0);
        } else {
          if (sd.anyGeneralShadow) run(FM_GENERAL, 0, 0);
          for (int l = 0; l < nL; ++l)
            if (sd.lights[l].kind == NRT_LIGHT_DISTANT && sd.frameValid(mo, FM_DIR, l)) run(FM_DIR, l, 1 + l);
        }
      }
      // (one launch: the float64 brute-force queue is empty on every scene seen so far)
      be->forEachCounted2(c + CNT_EXACT, cs.NR, ExactMesh{NRT_SCENE_ARG(sd), fp, cs, kind, mo, c + CNT_EXACT, bounce},
                          c + CNT_CAND, cs.candCap, Verify1<typename BE::Atom>{NRT_SCENE_ARG(sd), fp, cs, kind, mo, bounce});
      be->forEachCounted(c + CNT_CAND, cs.candCap, Verify2<typename BE::Atom>{cs, mo});
    }
  }

  // trace() of the shadow rays of the active samples + shadeDiffuse / reflection set-up (renderer.nim:90-127)
  void shadowAndResolve(const SceneData<BE>& sd, const FrameParams& fp, const ActiveSet& act, int bounce) {
    const int nL = cs.nL, pl = sd.anyPointLight ? 1 : 0;
    if (nL > 0 && nL <= 32 && fuseResolve && (sd.h.ncl1 > 0 || shadowTracePerSample)) {
      if (sd.h.ncl1 > 0) be->forEachStats(nullptr, act.n, ShadowResolveClustered{NRT_SCENE_ARG(sd), fp, cs, act, bounce, pl}, cs.stats);
      else be->forEachStats(nullptr, act.n, ShadowResolve{NRT_SCENE_ARG(sd), fp, cs, act, bounce, pl}, cs.stats);
    } else {
      if (nL > 0) {
        if (sd.h.ncl1 > 0) be->forEachStats(nullptr, act.n, ShadowTraceSampleClustered{NRT_SCENE_ARG(sd), fp, cs, act}, cs.stats);
        else if (shadowTracePerSample) be->forEachStats(nullptr, act.n, ShadowTraceSample{NRT_SCENE_ARG(sd), fp, cs, act}, cs.stats);
        else be->forEachStats(nullptr, act.n * nL, ShadowTrace{NRT_SCENE_ARG(sd), fp, cs, act}, cs.stats);
      }
      be->forEachStats(nullptr, act.n, Resolve{NRT_SCENE_ARG(sd), fp, cs, act, bounce, pl}, cs.stats);
    }
  }

  // ---- the frame being rendered, for the member functions below
  struct FrameCtx { const SceneData<BE>* sd; FrameParams fp; int maxBounces; int force_exact; bool jitter; int waves; };
  struct ForkState { bool forked = false; int attempt = 0; SubResult res; };

  static void addProfile(ProfileAcc& a, const ProfileAcc& b) {
    a.mesh_tests += b.mesh_tests; a.mesh_tests_ref += b.mesh_tests_ref; a.mesh_rays += b.mesh_rays; a.candidates += b.candidates;
    a.exact_rays += b.exact_rays; a.pre_candidates += b.pre_candidates; a.tail += b.tail;
    for (int k = 0; k < 3; ++k) a.tests_by_mode[k] += b.tests_by_mode[k];
    for (int k = 0; k < 8; ++k) { a.active[k] += b.active[k]; a.wavefront[k] += b.wavefront[k]; }
    a.need_cand = std::max(a.need_cand, b.need_cand); a.need_pairs = std::max(a.need_pairs, b.need_pairs);
  }

  // one bounce of the samples `set` through the wavefront: mesh wave of the path rays, Shade, mesh wave of the shadow
  // rays, ShadowResolve
  void wavefrontBounce(const FrameCtx& fc, const ActiveSet& set, int bounce, bool gated, int& wave) {
    const SceneData<BE>& sd = *fc.sd;
    meshWave(sd, fc.fp, WAVE_PATH, set, 2 * bounce, bounce, fc.force_exact, gated);
    if (sd.h.ncl1 > 0) be->forEachStats(nullptr, set.n, ShadeClustered{NRT_SCENE_ARG(sd), fc.fp, cs, set, bounce}, cs.stats);
    else be->forEachStats(nullptr, set.n, Shade{NRT_SCENE_ARG(sd), fc.fp, cs, set, bounce}, cs.stats);
    if (cs.nL > 0) meshWave(sd, fc.fp, WAVE_SHADOW, set, 2 * bounce + 1, bounce, fc.force_exact, false);
    shadowAndResolve(sd, fc.fp, set, bounce);
    wave = std::max(wave, 2 * bounce + 2);
  }
  void pathTail(const FrameCtx& fc, const ActiveSet& set, int bounce) {
    const SceneData<BE>& sd = *fc.sd;
    if (sd.h.ncl1 > 0) be->pathWarp(set.count, set.n, PathTailClustered{NRT_SCENE_ARG(sd), fc.fp, cs, fc.force_exact, 0, set, bounce}, cs.stats);
    else be->pathWarp(set.count, set.n, PathTail{NRT_SCENE_ARG(sd), fc.fp, cs, fc.force_exact, 0, set, bounce}, cs.stats);
  }

  // The fused path from bounce `bounce0` to the end of every path of the pool `act`.  `fk` != null (the main
  // pipeline at bounce 0): the fork — see `sub`.
  void fusedLoop(const FrameCtx& fc, ActiveSet act, int bounce0, ProfileAcc& pacc, int& wave, ForkState* fk,
                 int64_t nband, uint32_t* dBand, int64_t p0, int64_t nS) {
    const SceneData<BE>& sd = *fc.sd;
    const FrameParams& fp = fc.fp;
    const int nMO = cs.nMO;
    int par = (bounce0 + 1) & 1;   // half of cs.alist the next list goes to
    for (int bounce = bounce0;; ++bounce) {
      if (bounce > bounce0 && act.n < tailBelow) { pathTail(fc, act, bounce); pacc.tail += act.n; break; }   // a small wave: one launch to the end of its paths
      if (bounce < 8) pacc.active[bounce] += act.n;
      const int gfs = (bounce == 0 && fc.jitter) ? 1 : 0;
      if (sd.h.ncl1 > 0) be->forEachStats(nullptr, act.n, FusedBounceClustered{NRT_SCENE_ARG(sd), fp, cs, fc.force_exact, gfs, act, bounce}, cs.stats);
      else be->forEachStats(nullptr, act.n, FusedBounce{NRT_SCENE_ARG(sd), fp, cs, fc.force_exact, gfs, act, bounce}, cs.stats);
      bool forkedHere = false;
      uint32_t nh = 0;
      uint32_t* hardCount = cs.acount + 2 * bounce;
      if (nMO > 0) {
        be->compactActive(cs, act, cs.hlist, hardCount, kFlagWavefront);
        if (bounce == 0 && nband > 0) be->forEach(nband, BandCount{cs.hlist, hardCount, p0, fp.band_pix, nS, fp.spp, dBand});
        be->download(&nh, hardCount, sizeof(nh));
        if (fk && bounce == 0 && sub && forkMin > 0 && bounce < fc.maxBounces && int64_t(nh) >= std::max(hardTailBelow, forkMin)) {
          // the samples that continue WITHOUT the wavefront (flag kFlagContinues) leave for the helper pipeline
          uint32_t* fcount = cs.acount + fc.waves;   // (slots behind the per-bounce counts)
          be->compactActive(cs, act, cs.flist, fcount, kFlagContinues);
          uint32_t nf = 0;
          be->download(&nf, fcount, sizeof(nf));
          if (int64_t(nf) >= forkMin && sub->poolFits(int64_t(nf), cs.nL, nMO)) {
            forkedHere = fk->forked = true;
            sub->tailBelow = tailBelow; sub->hardTailBelow = hardTailBelow; sub->pathMode = 1;
            sub->shadowGatePerSample = shadowGatePerSample; sub->shadowTracePerSample = shadowTracePerSample; sub->fuseResolve = fuseResolve;
            Renderer* const helper = sub;
            const ChunkState parent = cs;
            const FrameCtx fcc = fc;
            SubResult* const out = &fk->res;
            const int attempt = fk->attempt;
            const int64_t n = int64_t(nf);
            be->fork([helper, parent, fcc, out, attempt, n] { helper->renderPool(fcc, parent, parent.flist, n, 1, attempt, *out); });
          }
        }
        if (nh > 0 && int64_t(nh) < hardTailBelow) {
          // a short wavefront list: its samples leave the pool here (one launch to the end of their paths)
          pathTail(fc, ActiveSet{cs.hlist, hardCount, int64_t(nh)}, bounce); pacc.tail += nh;
        } else if (nh > 0) {
          wavefrontBounce(fc, ActiveSet{cs.hlist, hardCount, int64_t(nh)}, bounce, false, wave);
          if (bounce < 8) pacc.wavefront[bounce] += nh;
        }
      }
      if (bounce >= fc.maxBounces) break;
      // the next bounce's active set: every sample of this one whose flag is nonzero (FusedBounce: continues;
      // Resolve: continues), in sample order — after a fork only the wavefront's samples are still this pipeline's
      uint32_t* nextList = cs.alist + int64_t(par) * cs.S;
      uint32_t* nextCount = cs.acount + 2 * bounce + 1;
      par ^= 1;
      if (forkedHere) be->compactActive(cs, ActiveSet{cs.hlist, hardCount, int64_t(nh)}, nextList, nextCount, 0);
      else be->compactActive(cs, act, nextList, nextCount, 0);
      uint32_t cont = 0;
      be->download(&cont, nextCount, sizeof(cont));
      if (cont == 0) break;
      act = ActiveSet{nextList, nextCount, int64_t(cont)};
    }
  }

  // counters of the waves [0, wave) -> profile; true: a list overflowed (the frame is rendered again with larger lists)
  bool countersToProfile(const SceneData<BE>& sd, const uint32_t* hc, int wave, ProfileAcc& pacc) const {
    const int nMO = cs.nMO, nL = cs.nL, cst = cntStride(nL);
    bool overflow = false;
    for (int w = 0; w < wave && nMO > 0; ++w)
      for (int mo = 0; mo < nMO; ++mo) {
        const uint32_t* c = hc + (int64_t(w) * nMO + mo) * cst;
        const int64_t nf = sd.meshes[sd.objs[sd.moIndex[mo]].mesh].nfaces;
        if (c[CNT_CAND] > uint64_t(cs.candCap)) overflow = true;
        pacc.need_cand = std::max<int64_t>(pacc.need_cand, c[CNT_CAND]);
        int64_t queued = 0;
        for (int b = 0; b <= nL; ++b) {
          const int64_t q = c[cntQueue(b)];
          if (c[cntPre(b)] > uint64_t(cs.preCap) || c[cntWork(b)] > uint64_t(cs.pairCap)) overflow = true;
          pacc.need_cand = std::max<int64_t>(pacc.need_cand, (int64_t(c[cntPre(b)]) + 3) / 4);   // (preCap = 4 * candCap)
          pacc.need_pairs = std::max<int64_t>(pacc.need_pairs, c[cntWork(b)]);
          pacc.pre_candidates += c[cntPre(b)];
          if (!q) continue;
          // wave parity: even = path wave (primary for w == 0), odd = shadow wave
          const int mode = b > 0 ? FM_DIR : ((w & 1) ? FM_GENERAL : ((w == 0 && sd.frameValid(mo, FM_ORIGIN, 0)) ? FM_ORIGIN : FM_GENERAL));
          // executed prefilter tests: chunk bounds tested ray by ray (2-D bundles: only the chunks whose circle overlaps
          // the run's circle), sub-chunk bounds of the admitted pairs, and the records of the admitted sub-chunks; a
          // run is the prefilterRunRays(mode) consecutive queue entries of one warp
          const int64_t run = prefilterRunRays(mode);
          // (cntBnd: chunk bounds AND sub-chunk bounds that were tested ray by ray; the sub-chunk circles of the 2-D
          // bundles' first look at level 2 cost one test each)
          const int64_t t = int64_t(c[cntBnd(b)]) * run + int64_t(c[cntWork(b)]) * kSubPerChunk + int64_t(c[cntSub(b)]) * run * kSubRecs;
          pacc.mesh_tests += t; pacc.tests_by_mode[mode] += t;
          queued += q;
        }
        pacc.mesh_rays += queued + c[CNT_EXACT];
        pacc.exact_rays += c[CNT_EXACT];
        pacc.mesh_tests_ref += (queued + int64_t(c[CNT_EXACT])) * nf;
        pacc.candidates += c[CNT_CAND];
      }
    return overflow;
  }

  // ---- helper side of the fork
  // room for a pool of n samples without asking the device for more than it has?
  bool poolFits(int64_t n, int nL, int nMO) {
    if (n <= capS) return true;
    const int64_t perSample = 260 + int64_t(110) * std::max(1, nL) * std::max(1, nMO);
    return n * perSample < int64_t(0.5 * double(be->memAvailable(capS * perSample)));
  }
  // The pool `list[0..n)` of the pipeline `parent` (samples with rays stored for bounce `bounce0`), to the end of its
  // paths; its accumulators are written back into parent.accum.
  void renderPool(const FrameCtx& fc, const ChunkState& parent, const uint32_t* list, int64_t n, int bounce0, int attempt, SubResult& out) {
    out = SubResult{};
    out.n = n;
    try {
      const SceneData<BE>& sd = *fc.sd;
      const int nL = sd.h.nlights, nMO = sd.h.nmesh_objs;
      int64_t cand = std::max<int64_t>(int64_t(1) << 20, n * std::max(1, nL) / 2);
      for (int k = 0; k < attempt; ++k) cand *= 4;
      // (a pool a little larger than the last one does not reallocate: 1/8 of headroom)
      const int64_t S = (n <= capS) ? capS : n + n / 8;
      ensure(S, nL, nMO, fc.waves, std::max(cand, capCand), 1, 0);
      cs.fb = nullptr; cs.aovObj = nullptr; cs.aovTri = nullptr; cs.aovT = nullptr; cs.q = OutStage{};
      cs.p0 = 0; cs.npix = 0;
      preLog.clear();
      const int64_t ncnt = int64_t(fc.waves) * std::max(nMO, 1) * cntStride(nL);
      be->zero(cs.counters, sizeof(uint32_t) * ncnt);
      be->zero(cs.stats, sizeof(unsigned long long) * ST_COUNT);
      be->zero(cs.acount, sizeof(uint32_t) * (fc.waves + 4));
      be->forEach(n, GatherPool{parent, cs, list});
      int wave = 0;
      fusedLoop(fc, ActiveSet{nullptr, nullptr, n}, bounce0, out.pacc, wave, nullptr, 0, nullptr, 0, n);
      be->forEach(n, ScatterAccum{parent, cs, list});
      std::vector<uint32_t> hc(ncnt);
      be->download(hc.data(), cs.counters, sizeof(uint32_t) * ncnt);
      be->download(out.stats, cs.stats, sizeof(out.stats));   // (a stream sync: the accumulators are back)
      out.overflow = countersToProfile(sd, hc.data(), wave, out.pacc);
      const int cst = cntStride(nL);
      for (auto& e : preLog)
        if (e.nch == 0 && e.wave < wave) {
          const uint32_t* c = hc.data() + (int64_t(e.wave) * nMO + e.mo) * cst;
          e.rays = c[cntQueue(e.b)]; e.work = c[cntWork(e.b)]; e.pre = c[cntPre(e.b)];
          e.nch = std::max<int64_t>(1, SceneData<BE>::numChunks(int64_t(sd.hostRecCount(e.mo, e.mode, e.b > 0 ? e.b - 1 : 0))));
        }
    } catch (const std::exception& ex) {
      out.rc = NRT_ERR_CUDA; out.err = ex.what();
    }
  }

  // Renders the rows `rows` (already filtered by step) of one worker.
  int render(const SceneData<BE>& sd, const nrt_options& o, const std::vector<int32_t>& rows, int step, int max_step,
             float* fb, int32_t* aovObj, int32_t* aovTri, double* aovT, unsigned long long* statsOut, std::string& err,
             const OutStage* qout = nullptr, int tshift = 0, int yEnd = -1) {
    FrameParams fp{};
    fp.width = o.width; fp.height = o.height; fp.aa_kind = o.aa_kind;
    fp.grid = o.aa_kind == NRT_AA_NONE ? 1 : o.grid_size;
    fp.spp = fp.grid * fp.grid;
    fp.step = step; fp.max_step = max_step;
    fp.depth_mode = o.depth_mode; fp.max_ray_depth = o.max_ray_depth;
    fp.bounce_cap = o.bounce_cap > 0 ? o.bounce_cap : 64;
    fp.nx = (o.width + step - 1) / step;
    fp.tshift = tshift;                       // > 0: `rows` are the first rows of bands of T scanlines (tile order inside)
    fp.band_pix = tshift > 0 ? (((o.width + (1 << tshift) - 1) >> tshift) << (2 * tshift)) : fp.nx;
    fp.y_end = (yEnd < 0 || yEnd > o.height) ? o.height : yEnd;
    fp.bias = o.bias; fp.seed = o.seed;
    fp.aspect = double(o.width) / double(o.height);
    fp.inv_grid = 1.0 / double(fp.grid);
    for (int k = 0; k < ST_COUNT; ++k) statsOut[k] = 0;
    if (rows.empty() || fp.nx == 0) return NRT_OK;

    const int nL = sd.h.nlights, nMO = sd.h.nmesh_objs;
    const bool jitter = o.aa_kind >= NRT_AA_JITTERED;
    pathMode = int(envInt("NRT_PATH", -1));
    tailBelow = envInt("NRT_TAIL_BELOW", 32768);
    hardTailBelow = envInt("NRT_HARD_TAIL_BELOW", 16384);
    forkMin = envInt("NRT_FORK_MIN", 0);
    bandHard.clear();
    if (pathMode < 0 || pathMode > 2) {
      // automatic: the fused path, unless the last frame of this pipeline showed that most samples have a ray entering
      // a mesh box (a scene that is mostly mesh: FusedBounce would do a sample's bounce only to hand it over) — then the
      // wavefront for every bounce, until the share of rays entering a box says otherwise
      // (hysteresis: the two estimates below are not the same quantity)
      autoMode = (meshShare > (autoMode == 0 ? 0.15 : 0.3)) ? 0 : 1;
      pathMode = autoMode;
    }
    // An INTENDED-mode frame can reflect at most max_ray_depth times.
    int maxBounces = sd.anyReflective ? fp.bounce_cap : 0;
    if (sd.anyReflective && o.depth_mode == NRT_DEPTH_INTENDED) maxBounces = std::min(maxBounces, std::max(0, o.max_ray_depth));
    const int waves = 2 * (maxBounces + 1);
    const int64_t npixTotal = int64_t(rows.size()) * fp.band_pix;
    // Samples per chunk: as many as fit in 96 GB or 60 % of the free device memory (~470 B of state per
    // sample for one mesh object and two lights; a 3840x2160x16 frame is ONE chunk of 62 GB on a 180 GB
    // B200).  Every bounce costs ~25 launches and a host check whatever the chunk size, so large chunks pay.
    int64_t S = envInt("NRT_CHUNK_SAMPLES", int64_t(1) << 28);
    const int64_t perSample = 260 + int64_t(110) * std::max(1, nL) * std::max(1, nMO);
    S = std::min<int64_t>(S, (int64_t(1) << 31) / std::max(1, nL));   // wave-ray indices are 32-bit
    S = std::min<int64_t>(S, npixTotal * fp.spp);
    if (S <= wantedS && capS > 0) {
      S = std::min(S, capS);          // the existing buffers (sized for a request at least this large) serve
    } else {
      // (re)allocation ahead: ask the device (cudaMemGetInfo is far too slow to call every frame)
      wantedS = S;
      const int64_t memCap = std::min<int64_t>(int64_t(96) << 30, int64_t(0.6 * double(be->memAvailable(capS * perSample))));
      S = std::min<int64_t>(S, std::max<int64_t>(memCap, int64_t(64) << 20) / perSample);
    }
    S = std::max<int64_t>(S, fp.spp);
    S = std::min<int64_t>(S, npixTotal * fp.spp);
    int64_t chunkPix = std::max<int64_t>(1, S / fp.spp);
    S = chunkPix * fp.spp;
    int64_t cand = std::max<int64_t>(envInt("NRT_CAND_CAP", 0), 0);
    // observed: ~0.06 candidates and ~0.3 pre-candidates per sample on the bunny scenes; an overflow re-renders
    // the frame with 4x the capacity
    if (cand == 0) cand = std::max<int64_t>(int64_t(1) << 20, S * std::max(1, nL) / 2);
    const int force_exact = int(envInt("NRT_FORCE_EXACT", 0));
    int64_t pairsReq = std::max<int64_t>(envInt("NRT_PAIR_CAP", 0), 0);   // (run, chunk) work-list capacity; 0: sized from cand

    for (int attempt = 0;; ++attempt) {
      ensure(S, nL, nMO, waves, cand, int(rows.size()), pairsReq);
      cs.fb = fb; cs.aovObj = aovObj; cs.aovTri = aovTri; cs.aovT = aovT;
      cs.q = qout ? *qout : OutStage{};
      be->upload(dRows, rows.data(), sizeof(int32_t) * rows.size());
      preLog.clear();
      bool overflow = false;
      unsigned long long total[ST_COUNT] = {0};
      ProfileAcc pacc;
      // (an overflowed attempt still walks every chunk: its frame is discarded, but the re-render then knows what the
      // WHOLE frame's lists asked for instead of finding the next chunk's overflow one attempt at a time)
      for (int64_t p0 = 0; p0 < npixTotal; p0 += chunkPix) {
        const int64_t npix = std::min(chunkPix, npixTotal - p0);
        const int64_t nS = npix * fp.spp;
        cs.p0 = p0; cs.npix = npix;
        const int64_t ncnt = int64_t(waves) * std::max(nMO, 1) * cntStride(nL);
        const int64_t nband = (pathMode == 1 && nMO > 0 && wantBandCounts) ? int64_t(rows.size()) : 0;
        uint32_t* const dBand = cs.counters + ncnt;
        be->zero(cs.counters, sizeof(uint32_t) * (ncnt + (p0 == 0 ? nband : 0)));
        be->zero(cs.stats, sizeof(unsigned long long) * ST_COUNT);
        be->zero(cs.acount, sizeof(uint32_t) * (waves + 4));
        int wave = 0;
        ActiveSet act{nullptr, nullptr, nS};   // bounce 0: every sample of the chunk
        const FrameCtx fc{&sd, fp, maxBounces, force_exact, jitter, waves};
        ForkState fk;
        fk.attempt = attempt;
        struct Joiner { BE* be; ForkState* fk; ~Joiner() { if (fk->forked) be->join(); } } joiner{be, &fk};   // (an exception between fork and join must not leave the helper running on this frame's state)
        if (pathMode == 2) {
          // ---- PathMega: every sample start to end in one launch ----
          if (jitter) be->forEach(npix, GenJittered{sd.d, fp, cs});
          if (sd.h.ncl1 > 0) be->pathWarp(nullptr, nS, PathMegaClustered{NRT_SCENE_ARG(sd), fp, cs, force_exact, jitter ? 1 : 0, act, 0}, cs.stats);
          else be->pathWarp(nullptr, nS, PathMega{NRT_SCENE_ARG(sd), fp, cs, force_exact, jitter ? 1 : 0, act, 0}, cs.stats);
        } else if (pathMode == 1) {
          // ---- fused path (nrt_pipeline.h: FusedBounceT, PathWarpT): per bounce, ONE kernel takes every active sample
          // through the whole bounce in registers unless one of its rays enters a mesh box; those samples (flag
          // kFlagWavefront, listed in sample order) go through the wavefront for that bounce.
          if (jitter) be->forEach(npix, GenJittered{sd.d, fp, cs});
          fusedLoop(fc, act, 0, pacc, wave, &fk, nband, dBand, p0, nS);
        } else {
          // ---- the wavefront for every bounce (round-1 pipeline; NRT_PATH=0) ----
          // primary rays: generated and gated in one kernel (the jittered kinds generate per pixel: separate gate)
          const bool fuseGen = !jitter && nMO > 0;
          if (jitter) be->forEach(npix, GenJittered{sd.d, fp, cs});
          else if (fuseGen) be->produceGate(nS, 1, GenGate{GenSimple{sd.d, fp, cs}, makeGate(sd, fp, WAVE_PATH, act, 0, force_exact), nMO}, cs, nMO, waveCounters(0), nullptr);
          else be->forEach(nS, GenSimple{sd.d, fp, cs});
          for (int bounce = 0;; ++bounce) {
            uint32_t* nextList = cs.alist + int64_t((bounce + 1) & 1) * cs.S;
            uint32_t* nextCount = cs.acount + 2 * bounce + 1;
            // (act.n is exact on the host for every bounce: launches are sized to it)
            wavefrontBounce(fc, act, bounce, bounce == 0 && fuseGen, wave);
            if (bounce >= maxBounces) break;
            be->compactActive(cs, act, nextList, nextCount, 0);
            uint32_t cont = 0;
            be->download(&cont, nextCount, sizeof(cont));
            if (cont == 0) break;  // no sample continued
            act = ActiveSet{nextList, nextCount, int64_t(cont)};
          }
        }
        if (fk.forked) {   // the helper's pool is back in this pipeline's accumulators before Finalize reads them
          be->join();
          if (fk.res.rc != NRT_OK) { err = fk.res.err; return fk.res.rc; }
          if (fk.res.overflow) overflow = true;
          for (int k = 0; k < ST_COUNT; ++k) total[k] += fk.res.stats[k];
          addProfile(pacc, fk.res.pacc);
        }
        be->finalize(npix, Finalize{fp, cs});
        // chunk epilogue: counters (profile + overflow check) and stats
        const bool lastChunk = p0 + chunkPix >= npixTotal;
        std::vector<uint32_t> hc(ncnt + (lastChunk ? nband : 0));
        unsigned long long hs[ST_COUNT];
        be->download(hc.data(), cs.counters, sizeof(uint32_t) * hc.size());
        if (lastChunk && nband > 0) bandHard.assign(hc.begin() + ncnt, hc.end());
        be->download(hs, cs.stats, sizeof(hs));
        if (countersToProfile(sd, hc.data(), wave, pacc)) overflow = true;
        for (int k = 0; k < ST_COUNT; ++k) total[k] += hs[k];
        const int cst = cntStride(nL);
        for (auto& e : preLog)
          if (e.nch == 0 && e.wave < wave) {   // entries of this chunk (filled once)
            const uint32_t* c = hc.data() + (int64_t(e.wave) * nMO + e.mo) * cst;
            e.rays = c[cntQueue(e.b)]; e.work = c[cntWork(e.b)]; e.pre = c[cntPre(e.b)];
            e.nch = std::max<int64_t>(1, SceneData<BE>::numChunks(int64_t(sd.hostRecCount(e.mo, e.mode, e.b > 0 ? e.b - 1 : 0))));
          }
      }
      if (!overflow) {
        for (int k = 0; k < ST_COUNT; ++k) statsOut[k] = total[k];
        prof = pacc;
        // the share of mesh work seen by this frame: samples handed to the wavefront at bounce 0 (fused path), or rays
        // that entered a box per sample (wavefront path: <= 1 + nL rays per sample and bounce)
        if (pathMode == 1 && pacc.active[0] > 0) meshShare = double(pacc.wavefront[0]) / double(pacc.active[0]);
        else if (pathMode == 0 && total[ST_PRIMARY] > 0) meshShare = std::min(1.0, double(pacc.mesh_rays) / double(total[ST_PRIMARY]));
        return NRT_OK;
      }
      if (attempt >= 4) { err = "candidate buffer overflow"; return NRT_ERR_OVERFLOW; }
      // re-render the frame with larger candidate / pair buffers (outputs are simply overwritten): at least what the
      // overflowed attempt asked for (+ 25 %) — a full list hides what the lists behind it would have needed (pairs ->
      // pre-candidates -> candidates), so up to three resizes can follow one another — and never less than 4x / 8x
      cand = std::max(cand * 4, pacc.need_cand + pacc.need_cand / 4 + 1);
      const int64_t pairsNeed = pacc.need_pairs + pacc.need_pairs / 4 + 1;
      if (pairsReq > 0) pairsReq = std::max(pairsReq * 8, pairsNeed);
      else if (pairsNeed > std::max<int64_t>(int64_t(1) << 20, cand / 8)) pairsReq = pairsNeed;
    }
  }
};

}  // namespace nrt
