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
  template <class F> void compactRecs(int64_t n, const F& f, float* recs, int mode, uint32_t* count) {
    const int nc = recFloats(mode);
    for (int64_t i = 0; i < n; ++i) {
      const RecOut o = f(i);
      if (!o.keep) continue;
      const int64_t r = (*count)++;
      for (int k = 0; k < nc; ++k) recs[recIndex(r, k, nc)] = o.c[k];
    }
    float c[16];
    neverHitRecord(mode, c);
    for (int64_t r = *count; r < paddedFaces(*count); ++r)
      for (int k = 0; k < nc; ++k) recs[recIndex(r, k, nc)] = c[k];
    ++launches;
  }
  void gate(const Gate& g, int64_t n, int nMO, uint32_t* cnt) {
    const ChunkState& cs = g.cs;
    const int cst = cntStride(cs.nL);
    for (int64_t i = 0; i < n; ++i)
      for (int mo = 0; mo < nMO; ++mo) {
        const GateOut o = g(i, mo);
        uint32_t* c = cnt + mo * cst;
        if (o.pass && o.safe) {
          const int64_t q = c[cntQueue(o.bundle)]++;
          const int64_t at = queueBase(cs, mo, o.bundle) + q;
          cs.qref[at] = uint32_t(i);
          float* p0 = cs.qray0 + 4 * at;
          p0[0] = o.fr.ax; p0[1] = o.fr.ay; p0[2] = o.fr.az; p0[3] = o.fr.rr;
          if (o.bundle == 0) {
            float* p1 = cs.qray1 + 4 * (int64_t(mo) * cs.NR + q);
            p1[0] = o.fr.mx; p1[1] = o.fr.my; p1[2] = o.fr.mz; p1[3] = 0.f;
          }
        } else if (o.pass) {
          const int64_t q = c[CNT_EXACT]++;
          cs.xref[int64_t(mo) * cs.NR + q] = uint32_t(i);
        }
      }
    ++launches;
  }
  // Same thread/ray assignment as k_mesh_filter (FT_THREADS threads x FT_R rays, Rr = max over a thread's rays).
  void filter(int mode, const float* recs, const uint32_t* count, const ChunkState& cs, int mo, int b, uint32_t* cnt) {
    const int T = 256, R = 8;
    const uint32_t nq = cnt[cntQueue(b)];
    const int64_t base = queueBase(cs, mo, b);
    const float* p0 = cs.qray0 + 4 * base;
    const float* p1 = cs.qray1 + 4 * (int64_t(mo) * cs.NR);
    const uint32_t* qref = cs.qref + base;
    const int64_t np = paddedFaces(*count);
    const int nc = recFloats(mode), idSlot = recSlotId(mode);
    for (uint32_t rt = 0; rt * T * R < nq; ++rt)
      for (int tid = 0; tid < T; ++tid) {
        uint32_t ic[R], ref[R]; float rr = 0.f;
        bool any = false;
        for (int r = 0; r < R; ++r) {
          const uint32_t idx = rt * T * R + r * T + tid;
          ic[r] = idx < nq ? idx : nq - 1;
          ref[r] = idx < nq ? qref[ic[r]] : kInvalidRef;
          rr = std::max(rr, p0[4 * ic[r] + 3]);
          any |= ref[r] != kInvalidRef;
        }
        if (!any) continue;
        for (int64_t t = 0; t < np; ++t) {
          float q[16];
          for (int k = 0; k < nc; ++k) q[k] = recs[recIndex(t, k, nc)];
          for (int r = 0; r < R; ++r) {
            if (ref[r] == kInvalidRef) continue;
            const uint32_t x = filterTest(mode, q, p0 + 4 * ic[r], p1 + 4 * ic[r], rr);
            ++filter_tests;
            if (int32_t(x) >= 0) {
              const uint32_t slot = cnt[CNT_CAND]++;
              uint32_t tri = uint32_t(t);
              if (idSlot >= 0) std::memcpy(&tri, &q[idSlot], 4);
              if (slot < cs.candCap) { cs.candRef[slot] = ref[r]; cs.candTri[slot] = tri; }
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
