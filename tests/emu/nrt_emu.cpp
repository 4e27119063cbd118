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
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../nim_raytracer_b200/csrc/nrt_renderer.h"

namespace nrt {

struct LoopBackend {
  int64_t launches = 0;
  int64_t filter_tests = 0;
  struct Atom {
    static void min64(uint64_t* p, uint64_t v) { if (v < *p) *p = v; }
    static void min32(uint32_t* p, uint32_t v) { if (v < *p) *p = v; }
    static uint32_t add32(uint32_t* p, uint32_t v) { const uint32_t o = *p; *p += v; return o; }
  };
  void* dalloc(size_t b) { return std::calloc(b ? b : 16, 1); }
  void dfree(void* p) { std::free(p); }
  void zero(void* p, size_t b) { std::memset(p, 0, b); }
  void upload(void* d, const void* s, size_t b) { std::memcpy(d, s, b); }
  void download(void* d, const void* s, size_t b) { std::memcpy(d, s, b); }
  void sync() {}
  template <class F> void fork(F f) { f(); }   // the helper pipeline of a fork runs inline
  void join() {}
  bool produceGateFits(int, int, int) const { return true; }
  bool diff_ = false;
  void diffBegin() { diff_ = false; }
  void diffAdd(const void* a, const void* b, size_t bytes) { if (bytes && std::memcmp(a, b, bytes) != 0) diff_ = true; }
  bool diffEnd() { return diff_; }
  int64_t memAvailable(int64_t) { return int64_t(1) << 40; }
  template <class F> void forEach(int64_t n, const F& f) { for (int64_t i = 0; i < n; ++i) f(i); ++launches; }
  template <class F> void forEachStats(const uint32_t* count, int64_t n, const F& f, unsigned long long* st) {
    if (count) n = *count;
    for (int64_t i = 0; i < n; ++i) { const StatDelta d = f(i); for (int k = 0; k < ST_COUNT; ++k) st[k] += d.v[k]; }
    ++launches;
  }
  // PathTail / PathMega: elements [0, *count) (or [0, n) when count is null), one "lane" at a time (ScalarCoop)
  template <class F> void pathWarp(const uint32_t* count, int64_t n, const F& f, unsigned long long* st) {
    forEachStats(nullptr, count ? std::min<int64_t>(*count, n) : n, f, st);
  }
  template <class F> void forEachCounted(const uint32_t* c, int64_t cap, const F& f) {
    const int64_t n = std::min<int64_t>(*c, cap);
    for (int64_t i = 0; i < n; ++i) f(i);
    ++launches;
  }
  template <class FA, class FB> void forEachCounted2(const uint32_t* ca, int64_t capA, const FA& fa, const uint32_t* cb, int64_t capB, const FB& fb) {
    forEachCounted(ca, capA, fa);
    forEachCounted(cb, capB, fb);
    --launches;
  }
  void finalize(int64_t npix, const Finalize& f) { forEach(npix, f); }
  void sortFaces(const DMesh& m) {
    std::vector<uint32_t> keys(m.nfaces), idx(m.nfaces);
    FaceKeys fk{m, keys.data(), idx.data()};
    for (int64_t f = 0; f < m.nfaces; ++f) fk(f);
    std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
    std::memcpy(m.order, idx.data(), sizeof(uint32_t) * m.nfaces);
    ++launches;
  }
  template <class F> void compactRecs(int64_t n, const F& f, float* recs, float* hot, int mode, uint32_t* count) {
    const int nc = recFloats(mode), nh = hotFloats(mode);
    for (int64_t i = 0; i < n; ++i) {
      const RecOut o = f(i);
      if (!o.keep) continue;
      const int64_t r = (*count)++;
      for (int k = 0; k < nc; ++k) recs[fullIndex(r, k)] = o.c[k];
      for (int k = 0; k < nh; ++k) hot[recIndex(r, k, nh)] = o.h[k];
    }
    float c[16], h[4];
    neverHitRecord(mode, c);
    neverHitHot(mode, h);
    for (int64_t r = *count; r < paddedFaces(*count); ++r) {
      for (int k = 0; k < nc; ++k) recs[fullIndex(r, k)] = c[k];
      for (int k = 0; k < nh; ++k) hot[recIndex(r, k, nh)] = h[k];
    }
    ++launches;
  }
  void gate(const Gate& g, const uint32_t* count, int64_t n, int mult, int nMO, uint32_t* cnt) {
    const ChunkState& cs = g.cs;
    const int cst = cntStride(cs.nL);
    if (count) n = int64_t(*count) * mult;
    for (int64_t i = 0; i < n; ++i)
      for (int mo = 0; mo < nMO; ++mo) {
        const GateOut o = g(i, mo);
        uint32_t* c = cnt + mo * cst;
        cs.gflag[int64_t(mo) * cs.NR + i] = o.pass ? (o.safe ? uint8_t(1 + o.bundle) : uint8_t(255)) : uint8_t(0);
        if (o.pass) {
          cs.tBest[int64_t(mo) * cs.NR + o.wi] = dbits(NRT_INF);
          cs.triBest[int64_t(mo) * cs.NR + o.wi] = kNoTri;
        }
        if (o.pass && o.safe) {
          const int64_t q = c[cntQueue(o.bundle)]++;
          const int64_t at = queueBase(cs, mo, o.bundle) + q;
          cs.qref[at] = o.wi;
          float* p0 = cs.qray0 + 4 * at;
          p0[0] = o.fr.ax; p0[1] = o.fr.ay; p0[2] = o.fr.az; p0[3] = o.fr.rr;
          float* h0 = cs.qhot0 + 4 * at;
          h0[0] = o.hr.a0; h0[1] = o.hr.a1; h0[2] = o.hr.a2; h0[3] = o.hr.a3;
          if (o.bundle == 0) {
            float* p1 = cs.qray1 + 4 * (int64_t(mo) * cs.NR + q);
            p1[0] = o.fr.mx; p1[1] = o.fr.my; p1[2] = o.fr.mz; p1[3] = 0.f;
            float* h1 = cs.qhot1 + 4 * (int64_t(mo) * cs.NR + q);
            h1[0] = o.hr.b0; h1[1] = o.hr.b1; h1[2] = o.hr.b2; h1[3] = 0.f;
          }
        } else if (o.pass) {
          const int64_t q = c[CNT_EXACT]++;
          cs.xref[int64_t(mo) * cs.NR + q] = o.wi;
        }
      }
    ++launches;
  }
  // fused producer + gate flags: here the producer runs first and the ordinary gate follows
  struct NoEmit { void operator()(int, int, uint8_t) const {} };
  template <class P> void produceGate(int64_t n, int, const P& p, const ChunkState&, int, uint32_t*, unsigned long long*) {
    NoEmit e;
    for (int64_t i = 0; i < n; ++i) p(i, e);
    ++launches;
  }
  void gateFinish(const Gate& g, int64_t n, int nMO, uint32_t* cnt) { gate(g, nullptr, n, 1, nMO, cnt); --launches; }
  // Prefilter, two-level like the CUDA kernel: the queue is cut into runs of prefilterRunRays(mode)
  // consecutive rays (one warp's registers); a run evaluates the 256 hot records of a chunk in full
  // iff at least one of its rays passes the chunk's bound test.
  void filter(int mode, const float* hot, const float* bounds, const float* sub, const uint32_t* count, const ChunkState& cs, int mo, int b, uint32_t* cnt) {
    const uint32_t nq = cnt[cntQueue(b)];
    const int64_t base = queueBase(cs, mo, b);
    const float* h0 = cs.qhot0 + 4 * base;
    const float* h1 = cs.qhot1 + 4 * (int64_t(mo) * cs.NR);
    const int64_t np = paddedFaces(*count), nch = np / kRecPad;
    const int nh = hotFloats(mode);
    const bool cull = std::getenv("NRT_PREFILTER_CULL") ? std::atoi(std::getenv("NRT_PREFILTER_CULL")) != 0 : true;
    const uint32_t run = uint32_t(prefilterRunRays(mode));
    auto rayAt = [&](uint32_t rq) {
      HotRay r;
      r.a0 = h0[4 * rq]; r.a1 = h0[4 * rq + 1]; r.a2 = h0[4 * rq + 2]; r.a3 = h0[4 * rq + 3];
      r.b0 = r.b1 = r.b2 = r.b3 = 0.f;
      if (mode == FM_GENERAL) { r.b0 = h1[4 * rq]; r.b1 = h1[4 * rq + 1]; r.b2 = h1[4 * rq + 2]; }
      return r;
    };
    for (uint32_t r0 = 0; r0 < nq; r0 += run) {
      const uint32_t r1 = std::min(nq, r0 + run);
      cnt[cntBnd(b)] += uint32_t(nch);   // (the CUDA kernel skips the chunks whose circle misses the run's circle)
      for (int64_t ch = 0; ch < nch; ++ch) {
        bool any = !cull;
        for (uint32_t rq = r0; rq < r1 && !any; ++rq) any = prefilterTest(mode, bounds + 4 * ch, rayAt(rq));
        if (!any) continue;
        ++cnt[cntWork(b)];
        cnt[cntBnd(b)] += uint32_t(kSubPerChunk);   // its sub-chunk bounds, tested ray by ray (nrt_renderer.h: countersToProfile)
        for (int64_t sb = ch * kSubPerChunk; sb < (ch + 1) * kSubPerChunk; ++sb) {
          bool anyS = !cull;
          for (uint32_t rq = r0; rq < r1 && !anyS; ++rq) anyS = prefilterTest(mode, sub + 4 * sb, rayAt(rq));
          if (!anyS) continue;
          ++cnt[cntSub(b)];
          for (uint32_t rq = r0; rq < r1; ++rq) {
            const HotRay r = rayAt(rq);
            for (int64_t t = sb * kSubRecs; t < (sb + 1) * kSubRecs; ++t) {
              float h[4];
              for (int k = 0; k < nh; ++k) h[k] = hot[recIndex(t, k, nh)];
              ++filter_tests;
              if (prefilterTest(mode, h, r)) {
                const uint32_t slot = cnt[cntPre(b)]++;
                if (slot < cs.preCap) { cs.preRay[slot] = rq; cs.preRec[slot] = uint32_t(t); }
              }
            }
          }
        }
      }
    }
    ++launches;
  }
  // next bounce's active list: the samples of the current set with active == 1, in sample order
  void compactActive(const ChunkState& cs, const ActiveSet& act, uint32_t* list, uint32_t* count, uint8_t match) {
    uint32_t n = 0;
    for (int64_t i = 0; i < act.n; ++i) {
      const int64_t s = act.list ? int64_t(act.list[i]) : i;
      if (match ? cs.active[s] == match : cs.active[s] != 0) list[n++] = uint32_t(s);
    }
    *count = n;
    ++launches;
  }
};

}  // namespace nrt

using namespace nrt;

extern "C" {

// Whole-frame render through the emulated pipeline (single worker).
int emu_render(const nrt_scene_desc* desc, const nrt_options* o, int y0, int y1, int step, int max_step, float* fb,
               nrt_stats* stats, const nrt_aov* aov, int64_t* prof /* 7 values, may be null */) {
  if (!desc || !o || !fb) return NRT_ERR_INVALID;
  if (!isPow2(step) || !isPow2(max_step) || max_step < step) return NRT_ERR_UNSUPPORTED;
  LoopBackend be;
  SceneData<LoopBackend> sd;
  std::string err;
  int rc = sd.build(&be, desc, false, err);
  if (rc != NRT_OK) { std::fprintf(stderr, "emu: %s\n", err.c_str()); return rc; }
  Renderer<LoopBackend> rn, helper;
  rn.be = &be;
  helper.be = &be;
  rn.sub = &helper;
  std::vector<int32_t> rows;
  const int tshift = tileShiftFor(*o, step, max_step), T = 1 << tshift;   // T > 1: bands of T rows, tile order inside
  for (int y = std::max(0, y0); y < std::min(y1, o->height); ++y)
    if ((y - y0) % (step * T) == 0) rows.push_back(y);
  unsigned long long st[ST_COUNT];
  rc = rn.render(sd, *o, rows, step, max_step, fb, aov ? aov->obj_id : nullptr, aov ? aov->tri_id : nullptr,
                 aov ? aov->t_hit : nullptr, st, err, nullptr, tshift, std::min(y1, o->height));
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
    prof[3] = rn.prof.exact_rays; prof[4] = be.launches; prof[5] = be.filter_tests; prof[6] = rn.prof.pre_candidates;
  }
  rn.freeAll();
  helper.freeAll();
  sd.destroy();
  return rc;
}

}  // extern "C"
