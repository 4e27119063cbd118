// nrt_emu.cpp — TEST-ONLY host emulation of the CUDA pipeline.
//
// Compiles the SAME per-element bodies (nrt_core.h / nrt_pipeline.h) and the SAME
// orchestration (nrt_renderer.h) as libnrt.so, with plain loops in place of kernel
// launches, so the wavefront logic, the float64 arithmetic and the float32 filter's
// conservativeness can be unit-tested on a machine without a GPU.
//
// This is NOT a CPU fallback of the product: it is built only by tests/ into
// tests/emu/libnrt_emu.so, exports `emu_*` symbols only, and nothing in
// nim_raytracer_b200/ or include/nrt.h can load or reach it.
#include <cstdio>
#include <cstring>

#include "../../nim_raytracer_b200/csrc/nrt_renderer.h"

namespace nrt {

struct LoopBackend {
  int64_t launches = 0;
  int64_t filter_tests = 0;
  struct Atom {
    static void min64(uint64_t* p, uint64_t v) { if (v < *p) *p = v; }
    static void min32(uint32_t* p, uint32_t v) { if (v < *p) *p = v; }
  };
  void* dalloc(size_t b) { return std::calloc(b ? b : 16, 1); }
  void dfree(void* p) { std::free(p); }
  void zero(void* p, size_t b) { std::memset(p, 0, b); }
  void upload(void* d, const void* s, size_t b) { std::memcpy(d, s, b); }
  void download(void* d, const void* s, size_t b) { std::memcpy(d, s, b); }
  void sync() {}
  template <class F> void forEach(int64_t n, const F& f) { for (int64_t i = 0; i < n; ++i) f(i); ++launches; }
  template <class F> void forEachStats(int64_t n, const F& f, unsigned long long* st) {
    for (int64_t i = 0; i < n; ++i) { const StatDelta d = f(i); for (int k = 0; k < ST_COUNT; ++k) st[k] += d.v[k]; }
    ++launches;
  }
  template <class F> void forEachCounted(const uint32_t* c, int64_t cap, const F& f) {
    const int64_t n = std::min<int64_t>(*c, cap);
    for (int64_t i = 0; i < n; ++i) f(i);
    ++launches;
  }
  void gate(const Gate& g, int64_t n, int nMO, uint32_t* cnt) {
    const ChunkState& cs = g.cs;
    for (int64_t i = 0; i < n; ++i)
      for (int mo = 0; mo < nMO; ++mo) {
        const GateOut o = g(i, mo);
        if (o.pass && o.safe) {
          const int64_t q = cnt[mo * CNT_STRIDE + CNT_QUEUE]++;
          cs.qref[int64_t(mo) * cs.NR + q] = uint32_t(i);
          float* p0 = cs.qray + int64_t(mo) * cs.NR * 8;
          float* p1 = p0 + cs.NR * 4;
          p0[4 * q + 0] = o.fr.dx; p0[4 * q + 1] = o.fr.dy; p0[4 * q + 2] = o.fr.dz; p0[4 * q + 3] = o.fr.rr;
          p1[4 * q + 0] = o.fr.mx; p1[4 * q + 1] = o.fr.my; p1[4 * q + 2] = o.fr.mz; p1[4 * q + 3] = 0.f;
        } else if (o.pass) {
          const int64_t q = cnt[mo * CNT_STRIDE + CNT_EXACT]++;
          cs.xref[int64_t(mo) * cs.NR + q] = uint32_t(i);
        }
      }
    ++launches;
  }
  // Same thread/ray assignment as k_mesh_filter (256 threads x 4 rays, rr = max over a thread's rays).
  void filter(const DMesh& m, const ChunkState& cs, int mo, uint32_t* cnt) {
    const int T = 256, R = 4;
    const uint32_t nq = cnt[CNT_QUEUE];
    const float* p0 = cs.qray + int64_t(mo) * cs.NR * 8;
    const float* p1 = p0 + cs.NR * 4;
    const uint32_t* qref = cs.qref + int64_t(mo) * cs.NR;
    const int64_t np = paddedFaces(m.nfaces);
    for (uint32_t rt = 0; rt * T * R < nq; ++rt)
      for (int tid = 0; tid < T; ++tid) {
        uint32_t ic[R], ref[R]; float rr = 0.f;
        for (int r = 0; r < R; ++r) {
          const uint32_t idx = rt * T * R + r * T + tid;
          ic[r] = idx < nq ? idx : nq - 1;
          ref[r] = idx < nq ? qref[ic[r]] : kInvalidRef;
          rr = std::max(rr, p0[4 * ic[r] + 3]);
        }
        bool any = false;
        for (int r = 0; r < R; ++r) any |= ref[r] != kInvalidRef;
        if (!any) continue;
        for (int64_t t = 0; t < np; ++t) {
          const float* q = m.recs + 16 * t;
          const float eb = q[3] * rr, kd = eb * kFilterKd;
          for (int r = 0; r < R; ++r) {
            if (ref[r] == kInvalidRef) continue;
            const uint32_t x = filterTest(q, p0[4 * ic[r]], p0[4 * ic[r] + 1], p0[4 * ic[r] + 2], p1[4 * ic[r]],
                                          p1[4 * ic[r] + 1], p1[4 * ic[r] + 2], eb, kd);
            ++filter_tests;
            if (int32_t(x) >= 0) {
              const uint32_t slot = cnt[CNT_CAND]++;
              if (slot < cs.candCap) { cs.candRef[slot] = ref[r]; cs.candTri[slot] = uint32_t(t); }
            }
          }
        }
      }
    ++launches;
  }
};

}  // namespace nrt

using namespace nrt;

extern "C" {

// Whole-frame render through the emulated pipeline (single worker).
int emu_render(const nrt_scene_desc* desc, const nrt_options* o, int y0, int y1, int step, int max_step, float* fb,
               nrt_stats* stats, const nrt_aov* aov, int64_t* prof /* 6 values, may be null */) {
  if (!desc || !o || !fb) return NRT_ERR_INVALID;
  if (!isPow2(step) || !isPow2(max_step) || max_step < step) return NRT_ERR_UNSUPPORTED;
  LoopBackend be;
  SceneData<LoopBackend> sd;
  std::string err;
  int rc = sd.build(&be, desc, false, err);
  if (rc != NRT_OK) { std::fprintf(stderr, "emu: %s\n", err.c_str()); return rc; }
  Renderer<LoopBackend> rn;
  rn.be = &be;
  std::vector<int32_t> rows;
  for (int y = std::max(0, y0); y < std::min(y1, o->height); ++y)
    if ((y - y0) % step == 0) rows.push_back(y);
  unsigned long long st[ST_COUNT];
  rc = rn.render(sd, *o, rows, step, max_step, fb, aov ? aov->obj_id : nullptr, aov ? aov->tri_id : nullptr,
                 aov ? aov->t_hit : nullptr, st, err);
  if (rc != NRT_OK) std::fprintf(stderr, "emu: %s\n", err.c_str());
  if (stats) {
    stats->num_primary_rays = int64_t(st[ST_PRIMARY]);
    stats->num_intersection_tests = int64_t(st[ST_TESTS]);
    stats->num_intersection_hits = int64_t(st[ST_HITS]);
    stats->num_rays = int64_t(st[ST_RAYS]);
    stats->num_capped_samples = int64_t(st[ST_CAPPED]);
  }
  if (prof) {
    prof[0] = rn.prof.mesh_rays; prof[1] = rn.prof.mesh_tests; prof[2] = rn.prof.candidates;
    prof[3] = rn.prof.exact_rays; prof[4] = be.launches; prof[5] = be.filter_tests;
  }
  rn.freeAll();
  sd.destroy();
  return rc;
}

}  // extern "C"
