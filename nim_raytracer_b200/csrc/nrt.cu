// nrt.cu — CUDA backend (sm_100a) + the C ABI of include/nrt.h.
//
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo
//        -fmad=false --shared -Xcompiler -fPIC nrt.cu -o libnrt.so
// -fmad=false: float64 code must round exactly like the IEEE oracle; the float32
// hot loop fuses explicitly with fmaf (FFMA in SASS).
//
// There is no CPU implementation in this library: every entry point that
// computes anything needs a CUDA device and fails with NRT_ERR_NO_DEVICE otherwise.

#include <cuda_runtime.h>
#include <cuda_pipeline.h>
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/permutation_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include <atomic>
#include <type_traits>
#include <cstdio>
#include <stdexcept>
#include <mutex>
#include <thread>
#include <sched.h>
#include <pthread.h>
#include <cctype>
#include <condition_variable>
#include <deque>
#include <functional>

#include "nrt_renderer.h"

// Resident CTAs per SM the register allocation of the per-sample kernels aims at (0 = the compiler's
// choice).  Measured on B200, BASELINE config 4: k_gate_flags 5.87 ms at 56 registers -> 4.52 ms at 40
// (6 CTAs/SM, 116 bytes of spills); Resolve 3.16 -> 2.95 ms at 64; Shade and ShadowTrace are fastest at
// the compiler's 64 registers without spills (48 registers: +7 %, 40: +15 %).
#ifndef NRT_OCC_ST
#define NRT_OCC_ST 0
#endif
// ShadowResolve (ShadowTrace + Resolve in one launch): 80 registers with ~100 bytes of spills in the
// Resolve tail (7.1 ms); 116 registers without spills at 2 CTAs/SM measured 8.5 ms
#ifndef NRT_OCC_SR
#define NRT_OCC_SR NRT_OCC_ST
#endif
#ifndef NRT_OCC_SHADE
#define NRT_OCC_SHADE 0
#endif
// FusedBounce (a whole bounce of a sample in registers).  Measured on B200, config 4 (all launches of a frame): the
// compiler's choice, 126 registers without spills (2 CTAs/SM), 13.06 ms; 3 CTAs/SM = 80 registers with ~300 bytes of
// spills 11.77 ms; 4 CTAs/SM = 64 registers 12.25 ms.  The kernel waits on fixed-latency float64 chains: warps in
// flight are worth more than the spills cost.
#ifndef NRT_OCC_FUSED
#define NRT_OCC_FUSED 4   // r02 (mask grids, merged shadow loop): 3 CTAs/SM at 80 registers 10.28 ms, 4 at 64 registers 9.84 ms
#endif
#ifndef NRT_OCC_TAIL
#define NRT_OCC_TAIL 1
#endif
#ifndef NRT_OCC_DEFAULT
#define NRT_OCC_DEFAULT 4
#endif
// prefilter evaluation kernel: resident CTAs per SM = persistent grid size (2-D bundles: 3 at 80 registers; 4 at 64
// registers measured 8 % slower; GENERAL: 2 at 94 registers, 3 no faster)
#ifndef NRT_OCC_PRE_2D
#define NRT_OCC_PRE_2D 3
#endif
#ifndef NRT_OCC_PRE_GEN
#define NRT_OCC_PRE_GEN 2
#endif
#ifndef NRT_OCC_GATE_WRITE
#define NRT_OCC_GATE_WRITE 4   // k_gate_write: 1.63 ms at 80 registers -> 1.33 ms at 64
#endif
#ifndef NRT_OCC_GATE
#define NRT_OCC_GATE 6
#endif

namespace nrt {

// ------------------------------------------------------------------ errors ----
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define NRT_CUDA(call)                                                                          \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess) {                                                                    \
      char b_[512];                                                                             \
      snprintf(b_, sizeof(b_), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      throw std::runtime_error(b_);                                                             \
    }                                                                                           \
  } while (0)

// ----------------------------------------------------------------- kernels ----
static constexpr int kBlock = 256;

template <class F>
__global__ void __launch_bounds__(kBlock) k_for_each(F f, int64_t n) {
  const int64_t i = int64_t(blockIdx.x) * kBlock + threadIdx.x;
  if (i < n) f(i);
}

// Block-wide sum of the per-thread Stats deltas into the global counters: one redux.sync per counter
// and warp (per-thread values are below 2^27: one trace() call tests each object once), the warp sums
// go to shared memory with plain stores, and one thread per counter adds them up in 64 bits for a
// single global atomic per counter and block.
__device__ __forceinline__ void blockStatsAdd(const StatDelta& d, unsigned long long* stats) {
  __shared__ unsigned sh[kBlock / 32][ST_COUNT];
  const unsigned warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k <= ST_CONT; ++k) {
    const unsigned v = __reduce_add_sync(0xffffffffu, unsigned(d.v[k]));
    if ((threadIdx.x & 31) == 0) sh[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x <= ST_CONT) {
    unsigned long long v = 0;
#pragma unroll
    for (int w = 0; w < kBlock / 32; ++w) v += sh[w][threadIdx.x];
    if (v) atomicAdd(&stats[threadIdx.x], v);
  }
}

// resident CTAs per SM the register allocation aims at (measured on B200: see DESIGN.md)
template <class F> struct MinBlocks { static constexpr int v = NRT_OCC_DEFAULT; };
template <> struct MinBlocks<ShadowTrace> { static constexpr int v = NRT_OCC_ST; };
template <bool CL> struct MinBlocks<ShadowTraceSampleT<CL>> { static constexpr int v = NRT_OCC_ST; };
template <bool CL> struct MinBlocks<ShadowResolveT<CL>> { static constexpr int v = NRT_OCC_SR; };
template <bool CL> struct MinBlocks<ShadeT<CL>> { static constexpr int v = NRT_OCC_SHADE; };
template <bool CL> struct MinBlocks<FusedBounceT<CL>> { static constexpr int v = NRT_OCC_FUSED; };
template <class F>
__global__ void __launch_bounds__(kBlock, MinBlocks<F>::v) k_for_each_stats(F f, int64_t n, unsigned long long* stats, int64_t ahead) {
  const int64_t i = int64_t(blockIdx.x) * kBlock + threadIdx.x;
  // A thread lives for one element: its first instruction after the index arithmetic waits for DRAM.
  // Requesting the inputs of the element `ahead` positions later into L2 (about one wave of CTAs ahead)
  // turns that wait into an L2 hit for the thread that will own it.
  if (ahead > 0 && i + ahead < n) f.prefetch(i + ahead);
  StatDelta d = zeroStats();
  if (i < n) d = f(i);
  blockStatsAdd(d, stats);
}

// ---- warp-cooperative mesh walk (PathTail / PathMega, nrt_pipeline.h: PathWarpT) ---------------------
// TriangleMesh.intersect (geom.nim:339-358) of ONE ray by the 32 lanes of a warp, over the flattened
// hierarchy of a record set: 32 chunk bounds per step (lane j <-> chunk base + j, one ballot), the sub-chunk
// bounds of two admitted chunks per step (16 lanes each), the records of two admitted sub-chunks per step
// (16 lanes each: bounding circle / sphere, float32 sign test, the reference's float64 evaluation — side by
// side in different lanes), then the warp's minimum of (t, face index).
__device__ __forceinline__ MeshHit walkWarp(const DMesh& m, const RecSet& rs, int mode, const WalkRay& w, unsigned lane) {
  const Ray r = walkRayAsRay(w);
  const HotRay hr = walkRayHot(w);
  double best = NRT_INF; uint32_t bt = kNoTri;     // geom.nim:343
  const uint32_t nch = uint32_t(paddedFaces(int64_t(rs.nrec)) / kRecPad);
  const unsigned half = lane >> 4, sl = lane & 15u;
  for (uint32_t base = 0; base < nch; base += 32) {
    const uint32_t c = base + lane;
    unsigned m1 = __ballot_sync(0xffffffffu, c < nch && prefilterTest(mode, rs.bounds + 4 * size_t(c), hr));
    while (m1) {
      const int b0 = __ffs(m1) - 1; m1 &= m1 - 1;
      int b1 = -1;
      if (m1) { b1 = __ffs(m1) - 1; m1 &= m1 - 1; }
      const int myb = half ? b1 : b0;
      unsigned m2 = __ballot_sync(0xffffffffu, myb >= 0 && prefilterTest(mode, rs.sub + 4 * (size_t(base + myb) * kSubPerChunk + sl), hr));
      while (m2) {
        const int s0 = __ffs(m2) - 1; m2 &= m2 - 1;
        int s1 = -1;
        if (m2) { s1 = __ffs(m2) - 1; m2 &= m2 - 1; }
        const int mys = half ? s1 : s0;
        if (mys >= 0) {
          const int64_t sub = int64_t(base + (mys < 16 ? b0 : b1)) * kSubPerChunk + (mys & 15);
          walkRecord(m, rs, mode, sub * kSubRecs + sl, w, r, hr, best, bt);
        }
      }
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double ot = __shfl_xor_sync(0xffffffffu, best, off);
    const uint32_t otri = __shfl_xor_sync(0xffffffffu, bt, off);
    if (ot < best || (ot == best && otri < bt)) { best = ot; bt = otri; }
  }
  MeshHit h; h.t = (best == 0) ? 0.0 : best; h.tri = bt;
  return h;
}
struct WarpCoop {
  __device__ __forceinline__ bool any(bool x) const { return __any_sync(0xffffffffu, x); }
  // every lane arrives here (dead lanes with need == false): the rays that passed their AABB gate are taken one
  // at a time — ballot, broadcast of the ray from its lane, walk by the whole warp, result back to its lane
  __device__ __forceinline__ void meshAll(const DScene& sc, bool need, int mode, int l, V4 o, V4 d, int force_exact, MeshRes& mr) const {
    const unsigned lane = threadIdx.x & 31u;
    for (int mo = 0; mo < sc.nmesh_objs && mo < kMaxWalkMO; ++mo) {
      mr.t[mo] = NRT_NEG_INF; mr.tri[mo] = kNoTri;
      const int md = walkMode(sc, mo, mode, l);
      WalkRay w;
      w.ox = w.oy = w.oz = w.dx = w.dy = w.dz = 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k) { w.f[k] = 0.f; w.h[k] = 0.f; }
      bool safe = false;
      const bool pass = need && walkPrep(sc, mo, md, l, o, d, force_exact, w, safe);
      if (pass) { mr.t[mo] = NRT_INF; }
      if (pass && !safe) {   // float64 over all faces by the lane itself (rays the float32 filter cannot take: rare)
        const MeshHit h = walkScalar(sc, mo, md, l, w, false);
        mr.t[mo] = h.t; mr.tri[mo] = h.tri;
      }
      unsigned todo = __ballot_sync(0xffffffffu, pass && safe);
      if (!todo) continue;
      const DMesh& m = sc.meshes[sc.objects[sc.mesh_obj_index[mo]].mesh];
      const RecSet rs = sc.recsets[frameIndex(sc.nlights, mo, md, l)];
      while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        WalkRay b;
        b.ox = __shfl_sync(0xffffffffu, w.ox, src); b.oy = __shfl_sync(0xffffffffu, w.oy, src); b.oz = __shfl_sync(0xffffffffu, w.oz, src);
        b.dx = __shfl_sync(0xffffffffu, w.dx, src); b.dy = __shfl_sync(0xffffffffu, w.dy, src); b.dz = __shfl_sync(0xffffffffu, w.dz, src);
#pragma unroll
        for (int k = 0; k < 8; ++k) { b.f[k] = __shfl_sync(0xffffffffu, w.f[k], src); b.h[k] = __shfl_sync(0xffffffffu, w.h[k], src); }
        const MeshHit h = walkWarp(m, rs, md, b, lane);
        if (int(lane) == src) { mr.t[mo] = h.t; mr.tri[mo] = h.tri; }
      }
    }
  }
};
// lane <-> element i of [0, n) (n = min(*count, nHost) when the count lives on the device); whole warps stay in
// f.run() until their last lane's path has ended; CTA-uniform trip count (blockStatsAdd synchronises)
// `lpw` lanes of every warp own an element, the others only help with the walks: a short list is spread over
// many warps (a warp takes its rays' walks one after the other, so the length of a launch is the longest warp's).
template <class F>
__global__ void __launch_bounds__(kBlock, NRT_OCC_TAIL) k_path_warp(F f, const uint32_t* count, int64_t nHost, int lpw, unsigned long long* stats) {
  int64_t n = nHost;
  if (count) { n = *count; if (n > nHost) n = nHost; }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t perCta = int64_t(kBlock / 32) * lpw;
  for (int64_t base = int64_t(blockIdx.x) * perCta; base < n; base += int64_t(gridDim.x) * perCta) {
    const int64_t i = base + int64_t(warp) * lpw + lane;
    const StatDelta d = f.run(i, lane < lpw && i < n, WarpCoop{});
    blockStatsAdd(d, stats);
    __syncthreads();
  }
}

// Elements [0, min(*count, cap)) with a device-resident count (no host sync).
template <class F>
__global__ void __launch_bounds__(kBlock) k_for_each_counted(F f, const uint32_t* count, int64_t cap) {
  int64_t n = *count;
  if (n > cap) n = cap;
  if (int64_t(blockIdx.x) * kBlock >= n) return;
  for (int64_t i = int64_t(blockIdx.x) * kBlock + threadIdx.x; i < n; i += int64_t(gridDim.x) * kBlock) f(i);
}

// Two counted lists in one launch (fa over [0, min(*ca, capA)), then fb over [0, min(*cb, capB))).
template <class FA, class FB>
__global__ void __launch_bounds__(kBlock) k_for_each_counted2(FA fa, const uint32_t* ca, int64_t capA, FB fb, const uint32_t* cb, int64_t capB) {
  int64_t na = *ca, nb = *cb;
  if (na > capA) na = capA;
  if (nb > capB) nb = capB;
  for (int64_t i = int64_t(blockIdx.x) * kBlock + threadIdx.x; i < na; i += int64_t(gridDim.x) * kBlock) fa(i);
  for (int64_t i = int64_t(blockIdx.x) * kBlock + threadIdx.x; i < nb; i += int64_t(gridDim.x) * kBlock) fb(i);
}

// AABB gate of TriangleMesh.intersect (geom.nim:340) + ORDERED compaction of the rays that enter
// each mesh's box into one queue per ray bundle (0 = arbitrary / shared-origin rays, 1 + l = shadow
// rays of DistantLight l) and the float64 brute-force queue.  Three launches:
//   k_gate_flags  evaluates the gate per (wave position, mesh object) — a float32 certain-miss test
//                 first, the reference's float64 slab test only for rays near the box — stores a
//                 one-byte code, counts the codes of every 256-ray block with warp ballots, adds
//                 the counts to their segment (256 blocks) and lists the non-empty blocks;
//   k_gate_scan   exclusive scan of the (few) segment totals per (mesh object, queue) row ->
//                 segment offsets and the queue totals;
//   k_gate_write  only over the non-empty blocks: offset = segment offset + counts of the preceding
//                 blocks of the segment + ballot rank; re-evaluates the passing rays and stores them.
// Queue order therefore equals wave order (scanline order of the samples): the 256 consecutive
// queue entries a prefilter warp works on belong to neighbouring pixels, which is what makes the
// chunk bounds of the two-level traversal selective.
static constexpr int kSegBlocks = 256;   // 256-ray blocks per segment
__device__ __forceinline__ uint8_t gateWant(int r, int nB) { return r < nB ? uint8_t(1 + r) : uint8_t(255); }

__global__ void __launch_bounds__(kBlock, NRT_OCC_GATE) k_gate_flags(Gate g, const uint32_t* count, int64_t nHost, int mult, int nMO, uint32_t* neCount) {
  extern __shared__ uint32_t sh_cnt[];   // nMO * nRow
  const unsigned lane = threadIdx.x & 31u;
  const ChunkState& cs = g.cs;
  const int nB = 1 + cs.nL, nRow = nB + 1, rows = nMO * nRow;
  const int64_t n = count ? int64_t(*count) * mult : nHost;
  const int64_t nVB = (n + kBlock - 1) / kBlock;
  for (int64_t vb = blockIdx.x; vb < nVB; vb += gridDim.x) {
    for (int k = threadIdx.x; k < rows; k += kBlock) sh_cnt[k] = 0;
    __syncthreads();
    const int64_t i = vb * kBlock + threadIdx.x;
    for (int mo = 0; mo < nMO; ++mo) {
      uint8_t code = 0;
      if (i < n) {
        code = gateCode(g(i, mo));
        cs.gflag[int64_t(mo) * cs.NR + i] = code;
      }
      if (__ballot_sync(0xffffffffu, code != 0)) {
        for (int r = 0; r < nRow; ++r) {
          const unsigned m = __ballot_sync(0xffffffffu, code == gateWant(r, nB));
          if (m && lane == 0) atomicAdd(&sh_cnt[mo * nRow + r], uint32_t(__popc(m)));
        }
      }
    }
    __syncthreads();
    bool any = false;
    for (int k = threadIdx.x; k < rows; k += kBlock) {
      const uint32_t c = sh_cnt[k];
      cs.gcnt[int64_t(k) * cs.gvb + vb] = c;
      if (c) { atomicAdd(&cs.gseg[int64_t(k) * cs.gsn + vb / kSegBlocks], c); any = true; }
    }
    if (__syncthreads_or(any) && threadIdx.x == 0) cs.gne[atomicAdd(neCount, 1u)] = uint32_t(vb);
  }
}

// Fused producer + gate flags (nrt_pipeline.h: GenGate, ShadeGate): thread i creates `mult` wave
// positions i * mult + k, evaluates their gate codes while the rays are in registers and counts
// them for the 256-position blocks it overlaps (a CTA of 256 threads covers exactly `mult` of them).
template <class P>
struct ProduceEmit {
  const ChunkState& cs; uint8_t* flags; uint32_t* sh; int64_t i; int mult, rows, nRow, nB;
  __device__ __forceinline__ void operator()(int k, int mo, uint8_t code) const {
    const int64_t pos = i * mult + k;
    flags[int64_t(mo) * cs.NR + pos] = code;
    // warp-aggregated shared-memory count: lanes with the same (block, row) key elect one adder
    const unsigned act = __activemask();
    if (!__any_sync(act, code != 0)) return;
    const int vbl = int((int64_t(threadIdx.x) * mult + k) >> 8);
    const int r = (code == 255) ? nB : int(code) - 1;
    const int key = code ? vbl * rows + mo * nRow + r : -1;
    const unsigned peers = __match_any_sync(act, key);
    if (key >= 0 && (threadIdx.x & 31) == unsigned(__ffs(peers) - 1)) atomicAdd(&sh[key], uint32_t(__popc(peers)));
  }
};
__device__ __forceinline__ void produceCounts(const ChunkState& cs, const uint32_t* sh, int64_t n, int mult, int rows, uint32_t* neCount) {
  const int64_t nVB = (n * mult + kBlock - 1) / kBlock;
  for (int k = threadIdx.x; k < mult * rows; k += kBlock) {
    const int vbl = k / rows, row = k - vbl * rows;
    const int64_t vb = int64_t(blockIdx.x) * mult + vbl;
    if (vb >= nVB) continue;
    const uint32_t c = sh[k];
    cs.gcnt[int64_t(row) * cs.gvb + vb] = c;
    if (c) atomicAdd(&cs.gseg[int64_t(row) * cs.gsn + vb / kSegBlocks], c);
  }
  for (int vbl = threadIdx.x; vbl < mult; vbl += kBlock) {
    const int64_t vb = int64_t(blockIdx.x) * mult + vbl;
    if (vb >= nVB) continue;
    uint32_t any = 0;
    for (int row = 0; row < rows; ++row) any |= sh[vbl * rows + row];
    if (any) cs.gne[atomicAdd(neCount, 1u)] = uint32_t(vb);
  }
}
// resident CTAs per SM of the fused producer + gate kernels (0 = the compiler's choice: 64 registers, 4 CTAs).
// Measured on B200, config 4: ShadowGate 3.67 ms at 64 registers, 4.01 at 48 (108 bytes of spills), 4.45 at 40,
// 5.18 at 32; GenGate 2.66 / 2.71 / 2.61 / 3.19 ms.
#ifndef NRT_OCC_PG_SHADOW
#define NRT_OCC_PG_SHADOW 0
#endif
#ifndef NRT_OCC_PG_GEN
#define NRT_OCC_PG_GEN 0
#endif
template <class P> struct MinBlocksPG { static constexpr int v = 0; };
template <> struct MinBlocksPG<ShadowGate> { static constexpr int v = NRT_OCC_PG_SHADOW; };
template <> struct MinBlocksPG<GenGate> { static constexpr int v = NRT_OCC_PG_GEN; };
template <class P>
__global__ void __launch_bounds__(kBlock, MinBlocksPG<P>::v) k_produce_gate(P p, ChunkState cs, int64_t n, int mult, int nMO, uint32_t* neCount, int64_t ahead) {
  extern __shared__ uint32_t sh_pc[];   // mult * rows
  const int nB = 1 + cs.nL, nRow = nB + 1, rows = nMO * nRow;
  {   // (see k_for_each_stats)
    const int64_t j = int64_t(blockIdx.x) * kBlock + threadIdx.x + ahead;
    if (ahead > 0 && j < n) p.prefetch(j);
  }
  for (int k = threadIdx.x; k < mult * rows; k += kBlock) sh_pc[k] = 0;
  __syncthreads();
  const int64_t i = int64_t(blockIdx.x) * kBlock + threadIdx.x;
  ProduceEmit<P> emit{cs, cs.gflag, sh_pc, i, mult, rows, nRow, nB};
  if (i < n) p(i, emit);
  __syncthreads();
  produceCounts(cs, sh_pc, n, mult, rows, neCount);
}
// one CTA per (mesh object, queue) row: exclusive scan of the segment totals (and their reset for the
// next gate); queue total -> counter block
__global__ void __launch_bounds__(1024) k_gate_scan(ChunkState cs, const uint32_t* count, int64_t nHost, int mult, uint32_t* cnt) {
  __shared__ uint32_t sh[1024];
  const int nB = 1 + cs.nL, nRow = nB + 1, cst = cntStride(cs.nL);
  const int row = blockIdx.x, mo = row / nRow, r = row - mo * nRow;
  uint32_t* seg = cs.gseg + int64_t(row) * cs.gsn;
  uint32_t* base = cs.gsegBase + int64_t(row) * cs.gsn;
  const int64_t n = count ? int64_t(*count) * mult : nHost;
  const int64_t nSeg = ((n + kBlock - 1) / kBlock + kSegBlocks - 1) / kSegBlocks;
  uint32_t carry = 0;
  for (int64_t s0 = 0; s0 < nSeg; s0 += 1024) {   // one pass unless a wave has more than 67 M rays
    const int64_t k = s0 + threadIdx.x;
    const uint32_t v = (k < nSeg) ? seg[k] : 0u;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
      const uint32_t a = (int(threadIdx.x) >= off) ? sh[threadIdx.x - off] : 0u;
      __syncthreads();
      sh[threadIdx.x] += a;
      __syncthreads();
    }
    if (k < nSeg) { base[k] = carry + sh[threadIdx.x] - v; seg[k] = 0; }
    carry += sh[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) cnt[mo * cst + (r < nB ? cntQueue(r) : CNT_EXACT)] = carry;
}

__global__ void __launch_bounds__(kBlock, NRT_OCC_GATE_WRITE) k_gate_write(Gate g, const uint32_t* count, int64_t nHost, int mult, int nMO, const uint32_t* neCount) {
  __shared__ uint32_t wc[kBlock / 32];
  __shared__ uint32_t sh_pre;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const ChunkState& cs = g.cs;
  const int nB = 1 + cs.nL, nRow = nB + 1;
  const int64_t n = count ? int64_t(*count) * mult : nHost;
  const uint32_t nNE = *neCount;
  for (uint32_t e = blockIdx.x; e < nNE; e += gridDim.x) {
    const int64_t vb = cs.gne[e];
    const int64_t i = vb * kBlock + threadIdx.x;
    const int64_t sg = vb / kSegBlocks, inSeg = vb - sg * kSegBlocks;
    for (int mo = 0; mo < nMO; ++mo) {
      const uint8_t code = (i < n) ? cs.gflag[int64_t(mo) * cs.NR + i] : uint8_t(0);
      if (code != 0) {   // passing rays start from "box hit, no face yet" (geom.nim:343: tMin = Inf)
        const uint32_t wi = g.waveIndex(i);
        cs.tBest[int64_t(mo) * cs.NR + wi] = dbits(NRT_INF);
        cs.triBest[int64_t(mo) * cs.NR + wi] = kNoTri;
      }
      for (int r = 0; r < nRow; ++r) {
        const int64_t row = int64_t(mo) * nRow + r;
        const uint32_t* pc = cs.gcnt + row * cs.gvb;
        if (pc[vb] == 0) continue;           // nothing of this block goes to this queue (block-uniform)
        // queue offset of the block: segment offset + the counts of the segment's preceding blocks
        if (threadIdx.x == 0) sh_pre = 0;
        __syncthreads();
        uint32_t v = (int64_t(threadIdx.x) < inSeg) ? pc[sg * kSegBlocks + threadIdx.x] : 0u;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        if (lane == 0 && v) atomicAdd(&sh_pre, v);
        const bool mine = code == gateWant(r, nB);
        const unsigned m = __ballot_sync(0xffffffffu, mine);
        if (lane == 0) wc[warp] = uint32_t(__popc(m));
        __syncthreads();
        uint32_t pre = cs.gsegBase[row * cs.gsn + sg] + sh_pre;
        for (unsigned w = 0; w < warp; ++w) pre += wc[w];
        __syncthreads();
        if (!mine) continue;
        const int64_t q = int64_t(pre) + __popc(m & lt);
        const GateOut o = g(i, mo);
        if (r < nB) {
          const int64_t at = queueBase(cs, mo, r) + q;
          cs.qref[at] = o.wi;
          reinterpret_cast<float4*>(cs.qray0)[at] = make_float4(o.fr.ax, o.fr.ay, o.fr.az, o.fr.rr);
          reinterpret_cast<float4*>(cs.qhot0)[at] = make_float4(o.hr.a0, o.hr.a1, o.hr.a2, o.hr.a3);
          if (r == 0) {
            reinterpret_cast<float4*>(cs.qray1)[int64_t(mo) * cs.NR + q] = make_float4(o.fr.mx, o.fr.my, o.fr.mz, 0.f);
            reinterpret_cast<float4*>(cs.qhot1)[int64_t(mo) * cs.NR + q] = make_float4(o.hr.b0, o.hr.b1, o.hr.b2, 0.f);
          }
        } else {
          cs.xref[int64_t(mo) * cs.NR + q] = o.wi;
        }
      }
    }
  }
}

// Filter-record build with culling, order preserving (records stay in the Morton order of the
// faces): keep flags -> exclusive scan (cub) -> scatter; stored pair-interleaved (nrt_filter.h).
template <class F>
__global__ void __launch_bounds__(kBlock) k_rec_flags(F f, int64_t n, uint32_t* flag) {
  const int64_t i = int64_t(blockIdx.x) * kBlock + threadIdx.x;
  if (i < n) flag[i] = f(i).keep ? 1u : 0u;
}
template <class F, int MODE>
__global__ void __launch_bounds__(kBlock) k_rec_scatter(F f, int64_t n, const uint32_t* pos, float* recs, float* hot, uint32_t* count) {
  constexpr int NC = recFloats(MODE), NH = hotFloats(MODE);
  const int64_t i = int64_t(blockIdx.x) * kBlock + threadIdx.x;
  if (i >= n) return;
  const RecOut o = f(i);
  const int64_t r = pos[i];
  if (o.keep) {
#pragma unroll
    for (int k = 0; k < NC; ++k) recs[fullIndex(r, k)] = o.c[k];
#pragma unroll
    for (int k = 0; k < NH; ++k) hot[recIndex(r, k, NH)] = o.h[k];
  }
  if (i == n - 1) *count = uint32_t(r) + (o.keep ? 1u : 0u);
}
template <int MODE>
__global__ void __launch_bounds__(kBlock) k_pad_recs(float* recs, float* hot, const uint32_t* count) {
  constexpr int NC = recFloats(MODE), NH = hotFloats(MODE);
  const int64_t n = *count, np = paddedFaces(n);
  float c[16], h[4];
  neverHitRecord(MODE, c);
  neverHitHot(MODE, h);
  for (int64_t r = n + threadIdx.x; r < np; r += kBlock) {
#pragma unroll
    for (int k = 0; k < NC; ++k) recs[fullIndex(r, k)] = c[k];
#pragma unroll
    for (int k = 0; k < NH; ++k) hot[recIndex(r, k, NH)] = h[k];
  }
}

// ---------------------------------------------------------- mesh prefilter ----
// The hot path: the queued rays of a bundle x the hot records (bounding circle / sphere,
// nrt_filter.h) of one record set, float32 — a flattened two-level traversal of the mesh.
//
// Level 1 (k_prefilter_bounds)  Records are stored in Morton order of the face centroids, so the
//          256 records of a chunk are a compact patch of the surface with one bound (circle / sphere
//          around the chunk's circles, same record format and same test).  A warp keeps a RUN of
//          32 x R consecutive queue entries (neighbouring pixels) in registers and tests them against
//          the bounds of 32 chunks: R tests per lane per bound, the per-lane pass bits are OR-reduced
//          across the warp (redux.sync), and every admitted (run, chunk) pair is appended to a work list.
// Level 2 (k_mesh_prefilter)  Persistent warps pop pairs from the list — uniform work items, so a
//          launch ends within one item of its last warp (short lists are split into half / quarter
//          chunks).  The warp stages the chunk's hot records into its shared-memory slice (cp.async,
//          16-byte vectors) and evaluates run x chunk in full: two faces per FFMA2 (fma.rn.f32x2;
//          records are pair-interleaved, ray components are scalar operands broadcast to both
//          halves), records read back as warp-broadcast LDS.128.  Per (ray, face) test:
//            ORIGIN / DIR  2 FFMA (2-D point in circle),   GENERAL  1 FMUL + 6 FFMA
//          plus the shared compare: max of four left-hand sides against the ray's threshold.
//          Survivors go through a per-warp shared-memory buffer to the pre-candidate list
//          (one global atomic per item, coalesced stores).
static constexpr int FT_THREADS = 256;
static constexpr int FT_WARPS = FT_THREADS / 32;
static constexpr int FT_TC = 256;     // records per chunk
static constexpr int FT_WB = 256;     // per-warp survivor buffer entries (flushed after every item)
static_assert(kRecPad == FT_TC, "one bound per shared-memory chunk");

struct PreArgs {
  const float4* hot;        // pair-interleaved hot records, padded to a multiple of kRecPad
  const float4* bounds;     // one hot-format record per chunk
  const float4* sub;        // one hot-format record per sub-chunk (kSubRecs records)
  const uint32_t* nrec;     // number of records (device)
  const float4* h0;         // ray plane H0: (x, y, T, 0) | (dh, T)
  const float4* h1;         // ray plane H1: (2 p0, 0)   (GENERAL)
  const uint32_t* qcount;   // queued rays (device)
  uint32_t* itemctr;        // level 2 work-item counter
  uint32_t* prectr;         // pre-candidate counter
  uint32_t* workctr;        // (run, chunk) pairs admitted by level 1 (may exceed pairCap: the host re-renders)
  uint32_t* subctr;         // (run, sub-chunk) pairs evaluated in full
  uint32_t* bndctr;         // (run, chunk) pairs whose bound was tested ray by ray
  float4* runc;             // per run: (cx, cy, Rr, margin) of its points' bounding circle, Rr < 0: none (2-D bundles, level 1 -> level 2)
  uint2* pairs;             // the work list
  uint32_t pairCap;
  uint32_t* preRay;
  uint32_t* preRec;
  uint32_t preCap;
  uint32_t cull;            // 0: every chunk is evaluated (brute force over the record set)
  uint32_t splitBelow;      // lists shorter than splitBelow x (resident warps) are split into half / quarter chunk items
};

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 dup2(float a) { return make_float2(a, a); }

// left-hand side of one prefilter test pair (two records) for one ray: exactly the operations of prefilterTest()
template <int MODE>
__device__ __forceinline__ float2 prefilterPair(const float2* h, float a0, float a1, float a2, float b0, float b1, float b2) {
  if (MODE == FM_GENERAL) {
    const float2 s = ffma2(h[0], dup2(a0), ffma2(h[1], dup2(a1), fmul2(h[2], dup2(a2))));
    const float2 tt = ffma2(h[0], dup2(b0), ffma2(h[1], dup2(b1), ffma2(h[2], dup2(b2), h[3])));
    return ffma2(s, s, tt);
  }
  return ffma2(h[0], dup2(a0), ffma2(h[1], dup2(a1), h[2]));
}

// the R rays of this lane in run `run` (tail entries duplicate the last real ray and are never emitted)
template <int MODE, int R>
struct LaneRays {
  float a0[R], a1[R], a2[R], a3[R], b0[R], b1[R], b2[R];
  uint32_t idx[R];
  __device__ __forceinline__ void load(const PreArgs& a, uint32_t run, uint32_t nq, int lane) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const uint32_t i = run * (32u * R) + r * 32 + lane;
      const uint32_t ic = i < nq ? i : nq - 1;
      const float4 p0 = __ldg(a.h0 + ic);
      a0[r] = p0.x; a1[r] = p0.y; a2[r] = p0.z; a3[r] = p0.w;
      if (MODE == FM_GENERAL) {
        const float4 p1 = __ldg(a.h1 + ic);
        b0[r] = p1.x; b1[r] = p1.y; b2[r] = p1.z;
      } else { b0[r] = b1[r] = b2[r] = 0.f; }
      idx[r] = i < nq ? i : kInvalidRef;
    }
  }
};

template <int MODE, int R>
__global__ void __launch_bounds__(FT_THREADS) k_prefilter_bounds(PreArgs a) {
  constexpr uint32_t RUN = 32 * R;
  static_assert(RUN == uint32_t(prefilterRunRays(MODE)), "run size is part of the executed-test accounting");
  const uint32_t nq = *a.qcount;
  if (nq == 0) return;
  const uint32_t nChunks = uint32_t(paddedFaces(int64_t(*a.nrec))) / FT_TC;
  if (nChunks == 0) return;
  const int lane = threadIdx.x & 31;
  const uint32_t nRuns = (nq + RUN - 1) / RUN, nGroups = (nChunks + 31) / 32;
  const uint32_t totalWarps = gridDim.x * FT_WARPS, gw = blockIdx.x * FT_WARPS + (threadIdx.x >> 5);
  uint32_t bnd = 0;
  for (uint64_t item = gw; item < uint64_t(nRuns) * nGroups; item += totalWarps) {
    const uint32_t run = uint32_t(item / nGroups), gr = uint32_t(item - uint64_t(run) * nGroups);
    const uint32_t c0 = gr * 32, c1 = min(nChunks, c0 + 32);
    uint32_t mask;
    if (a.cull) {
      LaneRays<MODE, R> ry;
      ry.load(a, run, nq, lane);
      uint32_t cand = (c1 - c0 >= 32) ? 0xffffffffu : ((1u << (c1 - c0)) - 1u);
      if (MODE != FM_GENERAL) {
        // 2-D bundles: one circle around the run's points first; lane j tests it against the circle of
        // chunk c0 + j (32 chunks per instruction instead of 8 ray tests per lane per chunk).  A ray that
        // passes a chunk's bound test lies within sqrt(Rc^2 + its margin + evaluation error) of the chunk
        // centre and within Rr of the run centre, so the two circles overlap: no admitted pair is lost.
        float xlo = ry.a0[0], xhi = ry.a0[0], ylo = ry.a1[0], yhi = ry.a1[0], mr = 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          xlo = fminf(xlo, ry.a0[r]); xhi = fmaxf(xhi, ry.a0[r]); ylo = fminf(ylo, ry.a1[r]); yhi = fmaxf(yhi, ry.a1[r]);
          mr = fmaxf(mr, fmaf(ry.a0[r], ry.a0[r], ry.a1[r] * ry.a1[r]) - ry.a2[r]);   // |p|^2 - T = the ray's margin
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          xlo = fminf(xlo, __shfl_xor_sync(0xffffffffu, xlo, off)); xhi = fmaxf(xhi, __shfl_xor_sync(0xffffffffu, xhi, off));
          ylo = fminf(ylo, __shfl_xor_sync(0xffffffffu, ylo, off)); yhi = fmaxf(yhi, __shfl_xor_sync(0xffffffffu, yhi, off));
          mr = fmaxf(mr, __shfl_xor_sync(0xffffffffu, mr, off));
        }
        const float cx = 0.5f * (xlo + xhi), cy = 0.5f * (ylo + yhi);
        float r2 = 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) { const float dx = ry.a0[r] - cx, dy = ry.a1[r] - cy; r2 = fmaxf(r2, fmaf(dx, dx, dy * dy)); }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) r2 = fmaxf(r2, __shfl_xor_sync(0xffffffffu, r2, off));
        const float Rr = sqrtf(r2) * 1.00001f;
        const float p2max = fmaf(fmaxf(fabsf(xlo), fabsf(xhi)), fmaxf(fabsf(xlo), fabsf(xhi)), fmaxf(fabsf(ylo), fabsf(yhi)) * fmaxf(fabsf(ylo), fabsf(yhi)));
        const bool circOk = Rr < 1e15f && mr < 1e30f && p2max < 1e30f;
        // the run's circle, for level 2's look at the sub-chunk bounds (k_mesh_prefilter)
        if (gr == 0 && lane == 0) a.runc[run] = make_float4(cx, cy, circOk ? Rr : -1.f, mr);
        if (circOk) {
          bool pass = false;
          if (c0 + lane < c1) {
            const float4 bd = __ldg(a.bounds + c0 + lane);
            const float Cx = 0.5f * bd.x, Cy = 0.5f * bd.y, C2 = fmaf(Cx, Cx, Cy * Cy);
            const float Rc2 = bd.z + C2 + 2e-6f * (fabsf(bd.z) + C2 + p2max) + 1.01f * mr;
            const float Rs = Rr + sqrtf(fmaxf(Rc2, 0.f)) + 5e-7f * (fabsf(cx) + fabsf(cy) + fabsf(Cx) + fabsf(Cy));
            const float dx = cx - Cx, dy = cy - Cy;
            pass = (bd.z >= 1e29f) || (bd.z > -1e29f && fmaf(dx, dx, dy * dy) <= Rs * Rs * 1.00001f);
          }
          cand &= __ballot_sync(0xffffffffu, pass);
        }
      }
      mask = 0;
      bnd += uint32_t(__popc(cand));
      for (uint32_t m = cand; m; m &= m - 1) {     // the exact per-ray bound test for the remaining chunks
        const uint32_t bit = uint32_t(__ffs(m) - 1);
        const float4 bd = __ldg(a.bounds + c0 + bit);
        bool p = false;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (MODE == FM_GENERAL) {
            const float sdot = fmaf(bd.x, ry.a0[r], fmaf(bd.y, ry.a1[r], bd.z * ry.a2[r]));
            const float tt = fmaf(bd.x, ry.b0[r], fmaf(bd.y, ry.b1[r], fmaf(bd.z, ry.b2[r], bd.w)));
            p = p || (fmaf(sdot, sdot, tt) >= ry.a3[r]);
          } else {
            p = p || (fmaf(bd.x, ry.a0[r], fmaf(bd.y, ry.a1[r], bd.z)) >= ry.a2[r]);
          }
        }
        if (__any_sync(0xffffffffu, p)) mask |= 1u << bit;
      }
    } else {
      mask = (c1 - c0 >= 32) ? 0xffffffffu : ((1u << (c1 - c0)) - 1u);
    }
    if (mask) {
      const uint32_t n = uint32_t(__popc(mask));
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(a.workctr, n);
      base = __shfl_sync(0xffffffffu, base, 0);
      if (uint32_t(lane) < n && base + lane < a.pairCap) a.pairs[base + lane] = make_uint2(run, c0 + __fns(mask, 0, lane + 1));
    }
  }
  if (lane == 0 && bnd) atomicAdd(a.bndctr, bnd);
}

// writes a warp's buffered survivors to the global pre-candidate list: one atomic, coalesced stores (every lane calls)
__device__ __noinline__ void preFlush(const uint2* wbuf, uint32_t wn, uint32_t* prectr, uint32_t* preRay, uint32_t* preRec, uint32_t preCap) {
  const uint32_t lane = threadIdx.x & 31u;
  __syncwarp();
  if (wn) {
    uint32_t gb = 0;
    if (lane == 0) gb = atomicAdd(prectr, wn);
    gb = __shfl_sync(0xffffffffu, gb, 0);
    for (uint32_t k = lane; k < wn; k += 32) {
      const uint2 e = wbuf[k];
      if (gb + k < preCap) { preRay[gb + k] = e.x; preRec[gb + k] = e.y; }
    }
  }
  __syncwarp();
}

template <int MODE, int R, int U>
__global__ void __launch_bounds__(FT_THREADS, MODE == FM_GENERAL ? NRT_OCC_PRE_GEN : NRT_OCC_PRE_2D) k_mesh_prefilter(PreArgs a) {
  constexpr int NH = hotFloats(MODE);      // float2 per record pair == float4 per record quad
  constexpr int CH4 = (FT_TC / 4) * NH;    // float4 per chunk
  constexpr int SQ = kSubRecs / 4;         // record quads per sub-chunk
  extern __shared__ __align__(16) float4 smem_tiles[];   // [FT_WARPS][CH4 + kSubPerChunk]
  const uint32_t nq = *a.qcount;
  if (nq == 0) return;
  const uint32_t nPairs = min(*a.workctr, a.pairCap);
  if (nPairs == 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4* const tile = smem_tiles + size_t(warp) * (CH4 + kSubPerChunk);
  float4* const stile = tile + CH4;        // the chunk's sub-chunk bounds
  // per-warp survivor buffer behind the tiles: FT_WB (ray, record) pairs + a counter
  uint2* const wbuf = reinterpret_cast<uint2*>(smem_tiles + size_t(FT_WARPS) * (CH4 + kSubPerChunk)) + size_t(warp) * FT_WB;
  uint32_t* const wcnt = reinterpret_cast<uint32_t*>(reinterpret_cast<uint2*>(smem_tiles + size_t(FT_WARPS) * (CH4 + kSubPerChunk)) + size_t(FT_WARPS) * FT_WB) + warp;
  (void)wcnt;
  // Survivors are compacted with warp votes: every lane reaches emit() (p = this lane has a survivor), one ballot
  // gives the lanes their slots behind the warp's running count `wn` (a register, the same in every lane) — no
  // shared-memory atomic and no wait for one (r02 ncu: the warp-aggregated ATOMS of the per-lane emit and the SHFLs
  // that broadcast their results carried ~45 % of the kernel's stall samples).
  // The buffer is written out when it is full (and once at the end), not per item: one global atomic per FT_WB
  // survivors (the per-item flush's atomic and, in dense items, one global atomic per emit were ~55 % of the stalls).
  uint32_t wn = 0;
  const uint32_t ltmask = (1u << lane) - 1u;
  auto flush = [&]() {   // every lane (a call, not inline code: emit() is instantiated at 32 sites, and the first
                         // inlined version ran out of instruction cache — 18 cycles of no_instruction stall per issue)
    preFlush(wbuf, wn, a.prectr, a.preRay, a.preRec, a.preCap);
    wn = 0;
  };
  auto emit = [&](bool p, uint32_t ray, uint32_t rec) {   // every lane
    const uint32_t bal = __ballot_sync(0xffffffffu, p);
    if (!bal) return;
    const uint32_t n = uint32_t(__popc(bal));
    if (wn + n > uint32_t(FT_WB)) flush();
    if (p) wbuf[wn + uint32_t(__popc(bal & ltmask))] = make_uint2(ray, rec);
    wn += n;
  };
  // short lists: items of half / quarter chunks so that the launch still fills the GPU
  const uint32_t totalWarps = gridDim.x * FT_WARPS;
  uint32_t SP = 1;
  while (SP < 4 && uint64_t(nPairs) * SP < uint64_t(a.splitBelow) * totalWarps) SP <<= 1;
  const uint32_t SN = kSubPerChunk / SP;               // sub-chunks per item
  const uint64_t nItems = uint64_t(nPairs) * SP;
  uint32_t subWork = 0, sbTested = 0;   // sub-chunks evaluated in full / sub-chunk bounds tested ray by ray
  // The item counter and the work list are read TWO / ONE items ahead: the global atomic (~700 cycles) and the
  // dependent load of the pair are in flight while the current item is evaluated (they were 11 % + 5 % of the stalls).
  auto pop = [&]() { uint32_t v = 0; if (lane == 0) v = atomicAdd(a.itemctr, 1u); return v; };   // (lane 0's register; broadcast where it is used)
  uint32_t itemA = __shfl_sync(0xffffffffu, pop(), 0);
  uint32_t rawB = pop();
  uint2 pairA = make_uint2(0u, 0u);
  if (itemA < nItems) pairA = a.pairs[itemA / SP];
  for (;;) {
    const uint32_t item = itemA;
    if (item >= nItems) break;
    const uint32_t pi = item / SP, part = item - pi * SP;
    const uint2 pr = pairA;
    // next item: its index arrived during the previous item; its pair is requested now, the index after it as well
    itemA = __shfl_sync(0xffffffffu, rawB, 0);
    if (itemA < nItems) pairA = a.pairs[itemA / SP];
    rawB = pop();
    // stage this item's part of the chunk and its sub-chunk bounds (16 bytes per lane and step), rays meanwhile
    {
      const float4* src = a.hot + size_t(pr.y) * CH4 + size_t(part) * SN * SQ * NH;
      for (uint32_t k = lane; k < SN * SQ * NH; k += 32) __pipeline_memcpy_async(tile + k, src + k, 16);
      if (uint32_t(lane) < SN) __pipeline_memcpy_async(stile + lane, a.sub + size_t(pr.y) * kSubPerChunk + part * SN + lane, 16);
      __pipeline_commit();
    }
    LaneRays<MODE, R> ry;
    ry.load(a, pr.x, nq, lane);
    __pipeline_wait_prior(0);
    __syncwarp();                            // the records are in the slice for every lane
    const uint32_t base = pr.y * FT_TC + part * SN * kSubRecs;
    // ---- level 2, first look (2-D bundles): the run's circle against the circles of the item's sub-chunks, one
    // sub-chunk per lane (the test of k_prefilter_bounds' level 1, same margins); the exact ray-by-ray bound test
    // below then runs for the overlapping sub-chunks only (it was ~30 % of an item's instructions)
    uint32_t smask = (SN >= 32u) ? 0xffffffffu : ((1u << SN) - 1u);
    if (MODE != FM_GENERAL && a.cull) {
      const float4 rc = __ldg(a.runc + pr.x);
      if (rc.z >= 0.f) {
        bool pass = false;
        if (uint32_t(lane) < SN) {
          const float4 bd = stile[lane];
          const float Cx = 0.5f * bd.x, Cy = 0.5f * bd.y, C2 = fmaf(Cx, Cx, Cy * Cy);
          const float px = fabsf(rc.x) + rc.z, py = fabsf(rc.y) + rc.z, p2max = fmaf(px, px, py * py);
          const float Rc2 = bd.z + C2 + 2e-6f * (fabsf(bd.z) + C2 + p2max) + 1.01f * rc.w;
          const float Rs = rc.z + sqrtf(fmaxf(Rc2, 0.f)) + 5e-7f * (fabsf(rc.x) + fabsf(rc.y) + fabsf(Cx) + fabsf(Cy));
          const float dx = rc.x - Cx, dy = rc.y - Cy;
          pass = (bd.z >= 1e29f) || (bd.z > -1e29f && fmaf(dx, dx, dy * dy) <= Rs * Rs * 1.00001f) || !(p2max < 1e30f);
        }
        smask &= __ballot_sync(0xffffffffu, pass);
      }
    }
    sbTested += uint32_t(__popc(smask));
    for (uint32_t sm = smask; sm; sm &= sm - 1) {
      const uint32_t sb = uint32_t(__ffs(int(sm)) - 1);
      // ---- level 2: can any ray of the run reach this sub-chunk? ----
      if (a.cull) {
        const float4 bd = stile[sb];
        bool p = false;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (MODE == FM_GENERAL) {
            const float sdot = fmaf(bd.x, ry.a0[r], fmaf(bd.y, ry.a1[r], bd.z * ry.a2[r]));
            const float tt = fmaf(bd.x, ry.b0[r], fmaf(bd.y, ry.b1[r], fmaf(bd.z, ry.b2[r], bd.w)));
            p = p || (fmaf(sdot, sdot, tt) >= ry.a3[r]);
          } else {
            p = p || (fmaf(bd.x, ry.a0[r], fmaf(bd.y, ry.a1[r], bd.z)) >= ry.a2[r]);
          }
        }
        if (!__any_sync(0xffffffffu, p)) continue;
      }
      ++subWork;
      // ---- level 3: run x the sub-chunk's records ----
      // A ray keeps ONE running maximum of its left-hand sides over the sub-chunk's 16 records (FMNMX3: two new values
      // per instruction) and is compared with its threshold once per sub-chunk; per (ray, record quad) that is
      // 4 FFMA2 + 2 FMNMX3 instead of 4 FFMA2 + FMNMX + FMNMX3 + FSETP + SEL + IADD (the ALU pipe, not the FMA pipe,
      // was the busier one: r01q ncu).  The rare sub-chunk with a passing ray is evaluated again for those rays.
      float mx[R];
#pragma unroll
      for (int r = 0; r < R; ++r) mx[r] = -3.0e38f;
#pragma unroll
      for (uint32_t tq = 0; tq < uint32_t(SQ); ++tq) {
        const uint32_t t = sb * SQ + tq;
        float2 q[2 * NH];   // two record pairs: q[j * NH + k] = coefficient k of pair j
#pragma unroll
        for (int c = 0; c < NH; ++c) {
          const float4 v4 = tile[t * NH + c];
          q[2 * c] = make_float2(v4.x, v4.y);
          q[2 * c + 1] = make_float2(v4.z, v4.w);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float2 g0 = prefilterPair<MODE>(q, ry.a0[r], ry.a1[r], ry.a2[r], ry.b0[r], ry.b1[r], ry.b2[r]);
          const float2 g1 = prefilterPair<MODE>(q + NH, ry.a0[r], ry.a1[r], ry.a2[r], ry.b0[r], ry.b1[r], ry.b2[r]);
          mx[r] = fmaxf(fmaxf(mx[r], g0.x), g0.y);
          mx[r] = fmaxf(fmaxf(mx[r], g1.x), g1.y);
        }
      }
      uint32_t hm = 0;   // bit r: ray r passed one of the sub-chunk's tests
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float thr = (MODE == FM_GENERAL) ? ry.a3[r] : ry.a2[r];
        if (mx[r] >= thr) hm |= 1u << r;
      }
      if (__any_sync(0xffffffffu, hm != 0)) {  // (warp-uniform) re-evaluate the rays that passed and emit their survivors
#pragma unroll 1
        for (uint32_t tq = 0; tq < uint32_t(SQ); ++tq) {
          const uint32_t t = sb * SQ + tq;
          float2 q[2 * NH];
#pragma unroll
          for (int c = 0; c < NH; ++c) {
            const float4 v4 = tile[t * NH + c];
            q[2 * c] = make_float2(v4.x, v4.y);
            q[2 * c + 1] = make_float2(v4.z, v4.w);
          }
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const bool mine = (hm & (1u << r)) != 0 && ry.idx[r] != kInvalidRef;
            if (!__any_sync(0xffffffffu, mine)) continue;   // no lane's ray r passed: warp-uniform
            const float thr = (MODE == FM_GENERAL) ? ry.a3[r] : ry.a2[r];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const float2 g = prefilterPair<MODE>(q + j * NH, ry.a0[r], ry.a1[r], ry.a2[r], ry.b0[r], ry.b1[r], ry.b2[r]);
              emit(mine && g.x >= thr, ry.idx[r], base + 4 * t + 2 * j);
              emit(mine && g.y >= thr, ry.idx[r], base + 4 * t + 2 * j + 1);
            }
          }
        }
      }
    }
    __syncwarp();                            // every lane left the slice before it is staged again
  }
  flush();
  if (lane == 0 && subWork) atomicAdd(a.subctr, subWork);
  if (lane == 0 && sbTested) atomicAdd(a.bndctr, sbTested);
}

// ---- ordered stream compaction of the per-sample flags ----------------------------------------------------
// The next bounce's list / the wavefront's list of a whole chunk: the indices i in [0, n) whose one-byte flag matches,
// ascending.  cub::DeviceSelect reads the flags a byte at a time (0.26 ms for the 132.7 M flags of a config-4 frame,
// twice per bounce-0); here a thread takes 16 flags with one 16-byte load and compares them four at a time
// (__vcmpeq4 / __vcmpne4), a CTA owns a tile of kSelTile flags: count per tile -> exclusive scan over the tiles (one CTA)
// -> the same tiles again, every thread writing its matches behind its exclusive prefix.  match == 0: flag != 0.
static constexpr int kSelTile = kBlock * 16;
__device__ __forceinline__ uint32_t selMask16(const uint8_t* flags, const uint32_t* in, int64_t base, int64_t n, uint32_t match) {
  uint32_t m = 0;
  if (in) {   // the flags of the samples in[base .. base + 16): gathered (the list ascends, so neighbours share sectors)
    uint8_t f[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) f[k] = (base + k < n) ? flags[in[base + k]] : uint8_t(match ? match + 1 : 0);
#pragma unroll
    for (int k = 0; k < 16; ++k) if (match ? (f[k] == match) : (f[k] != 0)) m |= 1u << k;
    return m;
  }
  if (base + 16 <= n) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(flags + base));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    const uint32_t mm = match * 0x01010101u;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t c = match ? __vcmpeq4(w[k], mm) : __vcmpne4(w[k], 0u);   // 0xff per matching byte
      m |= ((c & 1u) | ((c >> 7) & 2u) | ((c >> 14) & 4u) | ((c >> 21) & 8u)) << (4 * k);
    }
  } else {
    for (int k = 0; k < 16 && base + k < n; ++k) {
      const uint8_t f = flags[base + k];
      if (match ? (f == match) : (f != 0)) m |= 1u << k;
    }
  }
  return m;
}
__global__ void __launch_bounds__(kBlock) k_sel_count(const uint8_t* flags, const uint32_t* in, int64_t n, uint32_t match, uint32_t* tileCount) {
  const int64_t base = int64_t(blockIdx.x) * kSelTile + int64_t(threadIdx.x) * 16;
  const uint32_t c = base < n ? uint32_t(__popc(selMask16(flags, in, base, n, match))) : 0u;
  const uint32_t ws = __reduce_add_sync(0xffffffffu, c);
  __shared__ uint32_t sh[kBlock / 32];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = ws;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
#pragma unroll
    for (int w = 0; w < kBlock / 32; ++w) t += sh[w];
    tileCount[blockIdx.x] = t;
  }
}
// exclusive scan of the tile counts in place, total -> *count (one CTA of 1024 threads)
__global__ void __launch_bounds__(1024) k_sel_scan(uint32_t* tileCount, int64_t ntiles, uint32_t* count) {
  __shared__ uint32_t sh[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t b = 0; b < ntiles; b += 1024) {
    const int64_t i = b + threadIdx.x;
    const uint32_t v = i < ntiles ? tileCount[i] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, off); if ((threadIdx.x & 31) >= unsigned(off)) x += y; }
    if ((threadIdx.x & 31) == 31) sh[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      uint32_t t = sh[threadIdx.x];
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, t, off); if (threadIdx.x >= unsigned(off)) t += y; }
      sh[threadIdx.x] = t;   // inclusive over the warps
    }
    __syncthreads();
    const uint32_t wbase = (threadIdx.x >> 5) ? sh[(threadIdx.x >> 5) - 1] : 0u;
    const uint32_t c0 = carry;
    if (i < ntiles) tileCount[i] = c0 + wbase + x - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = c0 + wbase + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = carry;
}
__global__ void __launch_bounds__(kBlock) k_sel_write(const uint8_t* flags, const uint32_t* in, int64_t n, uint32_t match, const uint32_t* tileOffset, uint32_t* list) {
  const int64_t base = int64_t(blockIdx.x) * kSelTile + int64_t(threadIdx.x) * 16;
  uint32_t m = base < n ? selMask16(flags, in, base, n, match) : 0u;
  const uint32_t c = uint32_t(__popc(m));
  uint32_t x = c;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, off); if ((threadIdx.x & 31) >= unsigned(off)) x += y; }
  __shared__ uint32_t sh[kBlock / 32];
  if ((threadIdx.x & 31) == 31) sh[threadIdx.x >> 5] = x;
  __syncthreads();
  uint32_t wbase = 0;
#pragma unroll
  for (int w = 0; w < kBlock / 32; ++w) if (w < int(threadIdx.x >> 5)) wbase += sh[w];
  uint32_t o = tileOffset[blockIdx.x] + wbase + x - c;
  for (; m; m &= m - 1) { const uint32_t i = uint32_t(base) + uint32_t(__ffs(int(m)) - 1); list[o++] = in ? in[i] : i; }
}

// Finalize with coalesced reads: a pixel's samples are contiguous in the accumulator planes, so one
// thread per pixel would read 8-byte words 8*spp bytes apart.  The CTA copies a tile of
// pixels x spp samples into shared memory with coalesced loads (rows padded by one double: no bank
// conflicts), then every thread adds its pixel's row in sample order (renderer.nim:147-159).
static constexpr int kFinSmemDoubles = 5120;   // 40 KiB
__global__ void __launch_bounds__(kBlock) k_finalize_tiled(Finalize f, int64_t npix, int tilePix) {
  extern __shared__ double sh_fin[];
  const int spp = f.fp.spp, row = spp + 1;
  const ChunkState& cs = f.cs;
  for (int64_t p0 = int64_t(blockIdx.x) * tilePix; p0 < npix; p0 += int64_t(gridDim.x) * tilePix) {
    const int np = (npix - p0 < int64_t(tilePix)) ? int(npix - p0) : tilePix;
    double sum[3] = {0.0, 0.0, 0.0};
    for (int ch = 0; ch < 3; ++ch) {
      const double* src = cs.accum + int64_t(ch) * cs.S + p0 * spp;
      __syncthreads();
      for (int k = threadIdx.x; k < np * spp; k += kBlock) { const int p = k / spp; sh_fin[p * row + (k - p * spp)] = src[k]; }
      __syncthreads();
      if (int(threadIdx.x) < np) {
        double a = 0.0;
        const double* r = sh_fin + threadIdx.x * row;
        for (int k = 0; k < spp; ++k) a = a + r[k];
        sum[ch] = a;
      }
    }
    if (int(threadIdx.x) < np) f.store(p0 + threadIdx.x, sum[0], sum[1], sum[2]);
  }
}

// Output stage on an existing float32 framebuffer (nrt_framebuf_*): the same conversions Finalize's fused epilogue
// makes (nrt_pipeline.h: OutStage / storeQ), one thread per pixel.
__global__ void __launch_bounds__(kBlock) k_out_stage(const float* fb, OutStage q, int64_t npix) {
  const int64_t p = int64_t(blockIdx.x) * kBlock + threadIdx.x;
  if (p >= npix) return;
  storeQ(q, p, fb[3 * p], fb[3 * p + 1], fb[3 * p + 2]);
}

// word-wise comparison of two device arrays (nrt_scene_update: is the uploaded description the resident one?)
__global__ void __launch_bounds__(kBlock) k_diff(const uint2* a, const uint2* b, int64_t n, uint32_t* flag) {
  bool d = false;
  for (int64_t i = int64_t(blockIdx.x) * kBlock + threadIdx.x; i < n; i += int64_t(gridDim.x) * kBlock) {
    const uint2 x = a[i], y = b[i];
    d = d || x.x != y.x || x.y != y.y;
  }
  if (__syncthreads_or(d) && threadIdx.x == 0) atomicOr(flag, 1u);
}

// register-resident FFMA loop: the float32 roofline denominator
__global__ void __launch_bounds__(256) k_ffma_peak(float* out, int iters, float a, float b) {
  float x[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) x[k] = float(threadIdx.x + k);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = fmaf(x[k], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += x[k];
  if (s == 12345.678f) out[0] = s;
}

// ------------------------------------------------------- kernel categories ----
// Per-category device time of a frame (nrt_set_kernel_timing): every launch is bracketed by two
// CUDA events on the render stream.  Off by default (the events cost a little); bench.py turns it on
// for an untimed frame to report each kernel family's share of the step.
enum KernelCat { KC_GEN = 0, KC_GATE_FLAGS, KC_GATE_SCAN, KC_GATE_WRITE, KC_PREFILTER, KC_REFINE, KC_EXACT, KC_VERIFY,
                 KC_SHADE, KC_SHADOW_TRACE, KC_RESOLVE, KC_COMPACT, KC_FINALIZE, KC_OTHER, KC_SHADOW_RESOLVE, KC_FUSED_PRIMARY, KC_PATH_TAIL, KC_COUNT };
static const char* const kKernelCatNames[KC_COUNT] = {
  "gen+gate (GenGate / GenSimple / GenJittered)", "k_gate_flags", "k_gate_scan", "k_gate_write", "k_mesh_prefilter", "Refine",
  "ExactMesh", "Verify1+Verify2", "Shade", "ShadowTrace", "Resolve", "compactActive (ordered select)", "Finalize", "other",
  "ShadowResolve (ShadowTrace + Resolve)", "FusedBounce (a whole bounce of a sample in registers)", "PathTail / PathMega (per-thread paths + mesh walk)"};
static_assert(KC_COUNT <= NRT_KERNEL_CATEGORIES, "nrt_kernel_times is too small");
template <class F> struct CatOf { static constexpr int v = KC_OTHER; };
template <> struct CatOf<GenSimple> { static constexpr int v = KC_GEN; };
template <> struct CatOf<GenJittered> { static constexpr int v = KC_GEN; };
template <> struct CatOf<GenGate> { static constexpr int v = KC_GEN; };
template <> struct CatOf<ShadowGate> { static constexpr int v = KC_GATE_FLAGS; };
template <bool CL> struct CatOf<ShadeT<CL>> { static constexpr int v = KC_SHADE; };
template <> struct CatOf<ShadowTrace> { static constexpr int v = KC_SHADOW_TRACE; };
template <bool CL> struct CatOf<ShadowTraceSampleT<CL>> { static constexpr int v = KC_SHADOW_TRACE; };
template <bool CL> struct CatOf<ShadowResolveT<CL>> { static constexpr int v = KC_SHADOW_RESOLVE; };
template <> struct CatOf<Resolve> { static constexpr int v = KC_RESOLVE; };
template <bool CL> struct CatOf<FusedBounceT<CL>> { static constexpr int v = KC_FUSED_PRIMARY; };
template <bool CL, int KIND> struct CatOf<PathWarpT<CL, KIND>> { static constexpr int v = KC_PATH_TAIL; };
template <> struct CatOf<Finalize> { static constexpr int v = KC_FINALIZE; };
template <> struct CatOf<ExactMesh> { static constexpr int v = KC_EXACT; };
template <class A> struct CatOf<Refine<A>> { static constexpr int v = KC_REFINE; };
template <class A> struct CatOf<Verify1<A>> { static constexpr int v = KC_VERIFY; };
template <class A> struct CatOf<Verify2<A>> { static constexpr int v = KC_VERIFY; };

// One persistent host thread that runs one job at a time (the helper pipeline of a fork, nrt_renderer.h: Renderer::sub)
class Helper {
 public:
  Helper() : th_([this] { loop(); }) {}
  ~Helper() {
    { std::unique_lock<std::mutex> lk(mu_); stop_ = true; }
    cv_.notify_all();
    th_.join();
  }
  void start(std::function<void()> f) {
    { std::unique_lock<std::mutex> lk(mu_); job_ = std::move(f); busy_ = true; }
    cv_.notify_all();
  }
  void wait() {
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [this] { return !busy_; });
  }
 private:
  void loop() {
    for (;;) {
      std::function<void()> job;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [this] { return stop_ || bool(job_); });
        if (stop_ && !job_) return;
        job.swap(job_);
      }
      job();
      { std::unique_lock<std::mutex> lk(mu_); busy_ = false; }
      done_.notify_all();
    }
  }
  std::mutex mu_;
  std::condition_variable cv_, done_;
  std::function<void()> job_;
  bool busy_ = false, stop_ = false;
  std::thread th_;
};

// ----------------------------------------------------------------- backend ----
struct CudaBackend {
  Helper* helper = nullptr;    // runs the fork's helper pipeline (null: the job runs inline)
  void fork(std::function<void()> f) { if (helper) helper->start(std::move(f)); else f(); }
  void join() { if (helper) helper->wait(); }
  int device = 0;
  cudaStream_t stream = nullptr;               // the stream this backend launches on (one of the two below)
  cudaStream_t sHi = nullptr, sLo = nullptr;   // highest / lowest priority: a lane with the frame's mesh work launches on sHi,
                                               // a lane of mesh-free bands on sLo (its FusedBounce fills the SMs the other's chain of small kernels leaves idle)
  void createStreams() {
    int least = 0, greatest = 0;
    NRT_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    NRT_CUDA(cudaStreamCreateWithPriority(&sHi, cudaStreamNonBlocking, greatest));
    NRT_CUDA(cudaStreamCreateWithPriority(&sLo, cudaStreamNonBlocking, least));
    stream = sHi;
  }
  int sms = 148;
  int64_t launches = 0;
  // profiling of the mesh filter
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> filterEvents;
  size_t filterUsed = 0;
  std::vector<int> filterModes;
  bool cull = true;    // NRT_PREFILTER_CULL=0: evaluate every chunk (brute force over the record set)
  int splitBelow = 8;  // NRT_PREFILTER_SPLIT: measured on the 1/8-frame partitions of an 8-GPU run (2: 1.08 ms of prefilter, 8: 1.04, 32: 1.01; full frame unchanged)
  int64_t prefetchAhead = 0;   // elements ahead the per-sample kernels prefetch into L2 (NRT_PREFETCH_AHEAD; 0 = off): set in init
  bool smemOptIn = false;
  static constexpr size_t kPinnedBytes = 1 << 16;
  void* pinned = nullptr;
  void* scratchPtr[2] = {nullptr, nullptr};
  size_t scratchBytes[2] = {0, 0};

  struct Atom {
    static __device__ __forceinline__ void min64(uint64_t* p, uint64_t v) {
      atomicMin(reinterpret_cast<unsigned long long*>(p), static_cast<unsigned long long>(v));
    }
    static __device__ __forceinline__ void min32(uint32_t* p, uint32_t v) { atomicMin(p, v); }
    static __device__ __forceinline__ uint32_t add32(uint32_t* p, uint32_t v) { return atomicAdd(p, v); }
  };

  // ---- optional per-category timing ----
  bool timing = false;
  std::vector<cudaEvent_t> tEvents;     // pairs (start, stop)
  std::vector<int> tCats;
  size_t tUsed = 0;
  struct Timed {
    CudaBackend* be; size_t slot;
    Timed(CudaBackend* b, int cat) : be(b), slot(size_t(-1)) {
      if (!be->timing) return;
      if (be->tUsed == be->tCats.size()) {
        cudaEvent_t a, c;
        if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&c) != cudaSuccess) return;
        be->tEvents.push_back(a); be->tEvents.push_back(c); be->tCats.push_back(0);
      }
      slot = be->tUsed++;
      be->tCats[slot] = cat;
      cudaEventRecord(be->tEvents[2 * slot], be->stream);
    }
    ~Timed() { if (slot != size_t(-1)) cudaEventRecord(be->tEvents[2 * slot + 1], be->stream); }
  };
  // NRT_TIMELINE=<file>: every timed launch of this backend as "dev lane category start_us end_us" relative to `base`
  // (the frame's start event on the device's first stream); call before collectTimes
  void dumpTimeline(FILE* f, cudaEvent_t base, int dev, int lane) {
    for (size_t i = 0; i < tUsed; ++i) {
      float a = 0, b = 0;
      if (cudaEventElapsedTime(&a, base, tEvents[2 * i]) == cudaSuccess && cudaEventElapsedTime(&b, base, tEvents[2 * i + 1]) == cudaSuccess)
        fprintf(f, "%d %d %-14.14s %10.1f %10.1f\n", dev, lane, kKernelCatNames[tCats[i]], a * 1e3, b * 1e3);
    }
  }
  // call after a stream sync; adds this frame's per-category times and launch counts
  void collectTimes(double* ms, int64_t* n, double* mx) {
    for (size_t i = 0; i < tUsed; ++i) {
      float t = 0;
      if (cudaEventElapsedTime(&t, tEvents[2 * i], tEvents[2 * i + 1]) == cudaSuccess) {
        ms[tCats[i]] += t; n[tCats[i]] += 1;
        if (double(t) > mx[tCats[i]]) mx[tCats[i]] = double(t);
      }
    }
    tUsed = 0;
  }

  void use() { NRT_CUDA(cudaSetDevice(device)); }
  void* dalloc(size_t bytes) { use(); void* p = nullptr; NRT_CUDA(cudaMalloc(&p, bytes ? bytes : 16)); return p; }
  void dfree(void* p) { if (p) { cudaSetDevice(device); cudaFree(p); } }
  void zero(void* p, size_t bytes) { use(); NRT_CUDA(cudaMemsetAsync(p, 0, bytes, stream)); }
  void upload(void* dst, const void* src, size_t bytes) {
    use();
    NRT_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream));
    // pageable sources are staged by the runtime before the call returns; callers may reuse src
  }
  // device -> host with a stream sync; small reads (the per-bounce continuation count, counters, stats) go
  // through a pinned staging buffer: a copy into pageable memory is staged by the driver and costs ~2x
  void download(void* dst, const void* src, size_t bytes) {
    use();
    if (bytes <= kPinnedBytes) {
      if (!pinned) NRT_CUDA(cudaHostAlloc(&pinned, kPinnedBytes, cudaHostAllocDefault));
      NRT_CUDA(cudaMemcpyAsync(pinned, src, bytes, cudaMemcpyDeviceToHost, stream));
      NRT_CUDA(cudaStreamSynchronize(stream));
      std::memcpy(dst, pinned, bytes);
      return;
    }
    NRT_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream));
    NRT_CUDA(cudaStreamSynchronize(stream));
  }
  void sync() { use(); NRT_CUDA(cudaStreamSynchronize(stream)); }
  uint32_t* dDiff = nullptr;
  void diffBegin() {
    use();
    if (!dDiff) NRT_CUDA(cudaMalloc(&dDiff, 16));
    NRT_CUDA(cudaMemsetAsync(dDiff, 0, 4, stream));
  }
  void diffAdd(const void* a, const void* b, size_t bytes) {   // (all scene arrays are float64 / int64: 8-byte words)
    const int64_t n = int64_t(bytes / 8);
    if (n <= 0) return;
    k_diff<<<unsigned(std::min<int64_t>(int64_t(sms) * 8, (n + kBlock - 1) / kBlock)), kBlock, 0, stream>>>(
        static_cast<const uint2*>(a), static_cast<const uint2*>(b), n, dDiff);
    NRT_CUDA(cudaGetLastError()); ++launches;
  }
  bool diffEnd() { uint32_t f = 0; download(&f, dDiff, sizeof(f)); return f != 0; }
  // free device memory + what the caller already holds (its buffers are reused or replaced)
  int64_t memAvailable(int64_t held) {
    use();
    size_t fr = 0, tot = 0;
    NRT_CUDA(cudaMemGetInfo(&fr, &tot));
    return int64_t(fr) + held;
  }
  static unsigned blocksFor(int64_t n) { return unsigned((n + kBlock - 1) / kBlock); }

  template <class F> void forEach(int64_t n, const F& f) {
    if (n <= 0) return;
    use();
    Timed tm(this, CatOf<F>::v);
    k_for_each<F><<<blocksFor(n), kBlock, 0, stream>>>(f, n);
    NRT_CUDA(cudaGetLastError()); ++launches;
  }
  // over [0, n) (every count on this path is known on the host: the first argument is kept for the backend interface)
  template <class F> void forEachStats(const uint32_t*, int64_t n, const F& f, unsigned long long* stats) {
    use();
    if (n <= 0) return;
    Timed tm(this, CatOf<F>::v);
    k_for_each_stats<F><<<blocksFor(n), kBlock, 0, stream>>>(f, n, stats, prefetchAhead);
    NRT_CUDA(cudaGetLastError()); ++launches;
  }
  // PathTail (count on the device, at most n) / PathMega (count == nullptr: n elements)
  template <class F> void pathWarp(const uint32_t* count, int64_t n, const F& f, unsigned long long* stats) {
    use();
    if (n <= 0) return;
    Timed tm(this, CatOf<F>::v);
    // elements per warp: all 32 lanes when the list fills the GPU's resident warps (8 per SM at this kernel's
    // register count), fewer for short lists
    int lpw = 32;
    static const int64_t fill = [] { const char* e = std::getenv("NRT_TAIL_FILL"); return e ? std::max<int64_t>(1, std::atoll(e)) : int64_t(4); }();   // measured on a 1/8 frame: 1 -> 0.29 ms, 4 -> 0.19 ms, 16 -> 0.21 ms
    while (lpw > 1 && n * 32 / lpw < int64_t(sms) * 8 * 32 * fill / 2) lpw >>= 1;   // i.e. while warps(n, lpw) < fill/2 x the resident warps
    const int64_t perCta = int64_t(kBlock / 32) * lpw;
    const int64_t blocks = (n + perCta - 1) / perCta;
    k_path_warp<F><<<unsigned(blocks), kBlock, 0, stream>>>(f, count, n, lpw, stats);
    NRT_CUDA(cudaGetLastError()); ++launches;
  }
  template <class F> void forEachCounted(const uint32_t* count, int64_t cap, const F& f) {
    use();
    Timed tm(this, CatOf<F>::v);
    k_for_each_counted<F><<<unsigned(sms * 8), kBlock, 0, stream>>>(f, count, cap);
    NRT_CUDA(cudaGetLastError()); ++launches;
  }
  void finalize(int64_t npix, const Finalize& f) {
    if (npix <= 0) return;
    use();
    const int spp = f.fp.spp;
    if (f.fp.aa_kind == AA_NONE || spp < 2 || spp + 1 > kFinSmemDoubles / 8) { forEach(npix, f); return; }
    Timed tm(this, KC_FINALIZE);
    const int tilePix = std::min<int>(kBlock, kFinSmemDoubles / (spp + 1));
    const int64_t tiles = (npix + tilePix - 1) / tilePix;
    k_finalize_tiled<<<unsigned(std::min<int64_t>(tiles, int64_t(sms) * 16)), kBlock, sizeof(double) * kFinSmemDoubles, stream>>>(f, npix, tilePix);
    NRT_CUDA(cudaGetLastError()); ++launches;
  }
  template <class FA, class FB> void forEachCounted2(const uint32_t* ca, int64_t capA, const FA& fa, const uint32_t* cb, int64_t capB, const FB& fb) {
    use();
    Timed tm(this, CatOf<FB>::v);
    k_for_each_counted2<FA, FB><<<unsigned(sms * 8), kBlock, 0, stream>>>(fa, ca, capA, fb, cb, capB);
    NRT_CUDA(cudaGetLastError()); ++launches;
  }
  // grow-only scratch buffers (temp storage of cub, keep flags, sort keys)
  void* scratch(int slot, size_t bytes) {
    if (scratchBytes[slot] < bytes) {
      use();
      NRT_CUDA(cudaStreamSynchronize(stream));
      if (scratchPtr[slot]) NRT_CUDA(cudaFree(scratchPtr[slot]));
      scratchPtr[slot] = nullptr; scratchBytes[slot] = 0;
      const size_t want = std::max<size_t>(bytes + bytes / 4, 1 << 16);
      NRT_CUDA(cudaMalloc(&scratchPtr[slot], want));
      scratchBytes[slot] = want;
    }
    return scratchPtr[slot];
  }
  void gate(const Gate& g, const uint32_t* count, int64_t n, int mult, int nMO, uint32_t* cnt) {
    use();
    if (!count && n <= 0) return;
    const ChunkState& cs = g.cs;
    const int rows = nMO * (2 + cs.nL);
    const unsigned grid = count ? unsigned(sms * 8) : blocksFor(n);
    uint32_t* ne = cnt + CNT_NE;   // zeroed with the counter blocks at the start of the chunk
    { Timed tm(this, KC_GATE_FLAGS); k_gate_flags<<<grid, kBlock, sizeof(uint32_t) * rows, stream>>>(g, count, n, mult, nMO, ne); }
    { Timed tm(this, KC_GATE_SCAN); k_gate_scan<<<unsigned(rows), 1024, 0, stream>>>(cs, count, n, mult, cnt); }
    { Timed tm(this, KC_GATE_WRITE); k_gate_write<<<unsigned(sms * 8), kBlock, 0, stream>>>(g, count, n, mult, nMO, ne); }
    NRT_CUDA(cudaGetLastError()); launches += 3;
  }
  // the fused producer + gate kernel counts in mult * nMO * (2 + nL) shared-memory words: beyond the 48 KiB a launch gets
  // without opting in (e.g. 32 lights x 12 mesh objects) the caller takes the separate flags kernel instead
  bool produceGateFits(int mult, int nMO, int nL) const { return sizeof(uint32_t) * size_t(mult) * size_t(nMO) * size_t(2 + nL) <= size_t(48) * 1024; }
  // fused producer + gate flags (GenGate: mult 1; ShadeGate: mult nL, with Stats); gateFinish() completes the gate
  template <class P> void produceGate(int64_t n, int mult, const P& p, const ChunkState& cs, int nMO, uint32_t* cnt, unsigned long long* stats) {
    use();
    if (n <= 0) return;
    const size_t sm = sizeof(uint32_t) * size_t(mult) * nMO * (2 + cs.nL);
    (void)stats;
    Timed tm(this, CatOf<P>::v);
    k_produce_gate<P><<<blocksFor(n), kBlock, sm, stream>>>(p, cs, n, mult, nMO, cnt + CNT_NE, prefetchAhead);
    NRT_CUDA(cudaGetLastError()); ++launches;
  }
  void gateFinish(const Gate& g, int64_t n, int nMO, uint32_t* cnt) {
    use();
    if (n <= 0) return;
    const int rows = nMO * (2 + g.cs.nL);
    { Timed tm(this, KC_GATE_SCAN); k_gate_scan<<<unsigned(rows), 1024, 0, stream>>>(g.cs, nullptr, n, 1, cnt); }
    { Timed tm(this, KC_GATE_WRITE); k_gate_write<<<unsigned(sms * 8), kBlock, 0, stream>>>(g, nullptr, n, 1, nMO, cnt + CNT_NE); }
    NRT_CUDA(cudaGetLastError()); launches += 2;
  }
  // Morton order of the face centroids -> m.order (stable: ties keep face order)
  void sortFaces(const DMesh& m) {
    use();
    const int64_t n = m.nfaces;
    uint32_t* keys = static_cast<uint32_t*>(scratch(1, sizeof(uint32_t) * 3 * n));
    uint32_t* keys2 = keys + n; uint32_t* idx = keys + 2 * n;
    k_for_each<FaceKeys><<<blocksFor(n), kBlock, 0, stream>>>(FaceKeys{m, keys, idx}, n);
    size_t tb = 0;
    NRT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, keys2, idx, m.order, int(n), 0, 30, stream));
    void* tmp = scratch(0, tb);
    NRT_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, keys, keys2, idx, m.order, int(n), 0, 30, stream));
    NRT_CUDA(cudaGetLastError()); launches += 2;
  }
  template <class F> void compactRecs(int64_t n, const F& f, float* recs, float* hot, int mode, uint32_t* count) {
    use();
    uint32_t* flag = static_cast<uint32_t*>(scratch(1, sizeof(uint32_t) * 2 * n));
    uint32_t* pos = flag + n;
    k_rec_flags<F><<<blocksFor(n), kBlock, 0, stream>>>(f, n, flag);
    size_t tb = 0;
    NRT_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, flag, pos, int(n), stream));
    void* tmp = scratch(0, tb);
    NRT_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, flag, pos, int(n), stream));
    if (mode == FM_ORIGIN) {
      k_rec_scatter<F, FM_ORIGIN><<<blocksFor(n), kBlock, 0, stream>>>(f, n, pos, recs, hot, count);
      k_pad_recs<FM_ORIGIN><<<1, kBlock, 0, stream>>>(recs, hot, count);
    } else {
      k_rec_scatter<F, FM_DIR><<<blocksFor(n), kBlock, 0, stream>>>(f, n, pos, recs, hot, count);
      k_pad_recs<FM_DIR><<<1, kBlock, 0, stream>>>(recs, hot, count);
    }
    NRT_CUDA(cudaGetLastError()); launches += 4;
  }
  // next bounce's active list: the samples of the current set with active == 1, in sample order
  // `match` != 0: only the samples whose flag equals it (FusedBounce's kFlagWavefront), identity set only
  struct FlagIs { uint8_t v; __host__ __device__ bool operator()(uint8_t f) const { return f == v; } };
  void compactActive(const ChunkState& cs, const ActiveSet& act, uint32_t* list, uint32_t* count, uint8_t match) {
    use();
    Timed tm(this, KC_COMPACT);
    size_t tb = 0;
    const char* const ose = std::getenv("NRT_OWN_SELECT");   // (read per call: the tests switch it inside one process)
    const bool ownSelect = !(ose && *ose == '0');
    if (!act.list && ownSelect && act.n >= (int64_t(1) << 20)) {
      // a whole chunk's flags: three small launches taking 16 flags per thread (see k_sel_count); the same list as cub's.
      // (List-based sets stay with cub: the gathered variant of the same kernels — `in` != null — measured slower,
      // compactActive 0.49 -> 0.60 ms per config-4 frame.)
      const int64_t ntiles = (act.n + kSelTile - 1) / kSelTile;
      uint32_t* tiles = static_cast<uint32_t*>(scratch(0, sizeof(uint32_t) * size_t(ntiles)));
      k_sel_count<<<unsigned(ntiles), kBlock, 0, stream>>>(cs.active, act.list, act.n, match, tiles);
      k_sel_scan<<<1, 1024, 0, stream>>>(tiles, ntiles, count);
      k_sel_write<<<unsigned(ntiles), kBlock, 0, stream>>>(cs.active, act.list, act.n, match, tiles, list);
      NRT_CUDA(cudaGetLastError());
      launches += 3;   // (three kernels of this library; a cub select counts as one launch of its sweep)
      return;
    }
    if (!act.list && match) {
      thrust::counting_iterator<uint32_t> ids(0u);
      auto flags = thrust::make_transform_iterator(cs.active, FlagIs{match});
      NRT_CUDA(cub::DeviceSelect::Flagged(nullptr, tb, ids, flags, list, count, int(act.n), stream));
      void* tmp = scratch(0, tb);
      NRT_CUDA(cub::DeviceSelect::Flagged(tmp, tb, ids, flags, list, count, int(act.n), stream));
    } else if (match) {
      auto flags = thrust::make_transform_iterator(thrust::make_permutation_iterator(cs.active, act.list), FlagIs{match});
      NRT_CUDA(cub::DeviceSelect::Flagged(nullptr, tb, act.list, flags, list, count, int(act.n), stream));
      void* tmp = scratch(0, tb);
      NRT_CUDA(cub::DeviceSelect::Flagged(tmp, tb, act.list, flags, list, count, int(act.n), stream));
    } else if (!act.list) {
      thrust::counting_iterator<uint32_t> ids(0u);
      NRT_CUDA(cub::DeviceSelect::Flagged(nullptr, tb, ids, cs.active, list, count, int(act.n), stream));
      void* tmp = scratch(0, tb);
      NRT_CUDA(cub::DeviceSelect::Flagged(tmp, tb, ids, cs.active, list, count, int(act.n), stream));
    } else {
      auto flags = thrust::make_permutation_iterator(cs.active, act.list);
      NRT_CUDA(cub::DeviceSelect::Flagged(nullptr, tb, act.list, flags, list, count, int(act.n), stream));
      void* tmp = scratch(0, tb);
      NRT_CUDA(cub::DeviceSelect::Flagged(tmp, tb, act.list, flags, list, count, int(act.n), stream));
    }
    ++launches;
  }
  // prefilter launch for one ray bundle of one mesh object
  void filter(int mode, const float* hot, const float* bounds, const float* sub, const uint32_t* nrec, const ChunkState& cs, int mo, int b, uint32_t* cnt) {
    use();
    PreArgs a;
    a.hot = reinterpret_cast<const float4*>(hot);
    a.bounds = reinterpret_cast<const float4*>(bounds);
    a.sub = reinterpret_cast<const float4*>(sub);
    a.nrec = nrec;
    const int64_t base = queueBase(cs, mo, b);
    a.h0 = reinterpret_cast<const float4*>(cs.qhot0) + base;
    a.h1 = reinterpret_cast<const float4*>(cs.qhot1) + int64_t(mo) * cs.NR;
    a.qcount = cnt + cntQueue(b);
    a.itemctr = cnt + cntTile(b);
    a.prectr = cnt + cntPre(b);
    a.workctr = cnt + cntWork(b);
    a.subctr = cnt + cntSub(b);
    a.bndctr = cnt + cntBnd(b);
    a.pairs = reinterpret_cast<uint2*>(cs.pairs);
    a.runc = reinterpret_cast<float4*>(cs.runc);
    a.pairCap = uint32_t(std::min<int64_t>(cs.pairCap, 0xFFFFFFFFll));
    a.preRay = cs.preRay; a.preRec = cs.preRec;
    a.preCap = uint32_t(std::min<int64_t>(cs.preCap, 0xFFFFFFFFll));
    a.cull = cull ? 1u : 0u;
    a.splitBelow = uint32_t(splitBelow);
    if (filterUsed == filterEvents.size()) {
      cudaEvent_t e0, e1;
      NRT_CUDA(cudaEventCreate(&e0)); NRT_CUDA(cudaEventCreate(&e1));
      filterEvents.push_back({e0, e1});
      filterModes.push_back(0);
    }
    auto& ev = filterEvents[filterUsed];
    filterModes[filterUsed++] = mode;
    Timed tm(this, KC_PREFILTER);
    NRT_CUDA(cudaEventRecord(ev.first, stream));
    // per-warp chunk slice + survivor buffer: 2-D bundles 8 rays/lane (40 KiB/CTA, 3 CTAs/SM); GENERAL 4 rays/lane (48 KiB/CTA, 2 CTAs/SM)
    constexpr size_t smBuf = size_t(FT_WARPS) * FT_WB * sizeof(uint2) + FT_WARPS * sizeof(uint32_t);
    constexpr size_t sm2d = size_t(FT_WARPS) * ((FT_TC / 4) * 3 + kSubPerChunk) * sizeof(float4) + smBuf;
    constexpr size_t smGen = size_t(FT_WARPS) * ((FT_TC / 4) * 4 + kSubPerChunk) * sizeof(float4) + smBuf;
    if (!smemOptIn) {
      NRT_CUDA(cudaFuncSetAttribute(k_mesh_prefilter<FM_GENERAL, 4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smGen)));
      NRT_CUDA(cudaFuncSetAttribute(k_mesh_prefilter<FM_ORIGIN, 8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm2d)));
      NRT_CUDA(cudaFuncSetAttribute(k_mesh_prefilter<FM_DIR, 8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm2d)));
      smemOptIn = true;
    }
    const unsigned gb = unsigned(sms * 8);
    if (mode == FM_GENERAL) {
      k_prefilter_bounds<FM_GENERAL, 4><<<gb, FT_THREADS, 0, stream>>>(a);
      k_mesh_prefilter<FM_GENERAL, 4, 2><<<unsigned(sms * NRT_OCC_PRE_GEN), FT_THREADS, smGen, stream>>>(a);
    } else if (mode == FM_ORIGIN) {
      k_prefilter_bounds<FM_ORIGIN, 8><<<gb, FT_THREADS, 0, stream>>>(a);
      k_mesh_prefilter<FM_ORIGIN, 8, 1><<<unsigned(sms * NRT_OCC_PRE_2D), FT_THREADS, sm2d, stream>>>(a);
    } else {
      k_prefilter_bounds<FM_DIR, 8><<<gb, FT_THREADS, 0, stream>>>(a);
      k_mesh_prefilter<FM_DIR, 8, 1><<<unsigned(sms * NRT_OCC_PRE_2D), FT_THREADS, sm2d, stream>>>(a);
    }
    launches += 2;
    NRT_CUDA(cudaGetLastError()); ++launches;
    NRT_CUDA(cudaEventRecord(ev.second, stream));
  }
  // call after a stream sync
  double filterMs(int64_t* n, double* byMode, std::vector<float>* each = nullptr) {
    double ms = 0;
    for (size_t i = 0; i < filterUsed; ++i) {
      float t = 0;
      if (cudaEventElapsedTime(&t, filterEvents[i].first, filterEvents[i].second) == cudaSuccess) { ms += t; byMode[filterModes[i]] += t; }
      if (each) each->push_back(t);
    }
    *n = int64_t(filterUsed);
    filterUsed = 0;
    return ms;
  }
  void destroy() {
    cudaSetDevice(device);
    for (auto& e : filterEvents) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    filterEvents.clear();
    for (auto& e : tEvents) cudaEventDestroy(e);
    tEvents.clear(); tCats.clear(); tUsed = 0;
    for (int k = 0; k < 2; ++k) { if (scratchPtr[k]) cudaFree(scratchPtr[k]); scratchPtr[k] = nullptr; scratchBytes[k] = 0; }
    if (pinned) cudaFreeHost(pinned);
    pinned = nullptr;
    if (dDiff) cudaFree(dDiff);
    dDiff = nullptr;
    if (sHi) cudaStreamDestroy(sHi);
    if (sLo) cudaStreamDestroy(sLo);
    stream = sHi = sLo = nullptr;
  }
};

// ------------------------------------------------------------ global state ----
// Lanes: a GPU's share of a frame is rendered as up to kMaxLanes INDEPENDENT pipelines (pixels are independent:
// lane k of K takes every K-th of the worker's scanlines), each with its own stream, buffers and host thread.
// A pipeline is a chain of dependent launches with a few host round trips; late bounces and small frames (a
// 1/8 frame on each of eight GPUs) leave the GPU mostly idle inside one chain — several chains in flight fill it.
static constexpr int kMaxLanes = 8;
struct DeviceCtx {
  CudaBackend be;                              // lane 0 (also scene builds, output stage, timers)
  std::vector<CudaBackend*> extra;             // lanes 1 ..
  CudaBackend& lane(int k) { return k == 0 ? be : *extra[size_t(k - 1)]; }
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;   // frame bracket (nrt_profile.total_ms)
  cudaEvent_t tb0 = nullptr, tb1 = nullptr;   // user bracket (nrt_timer_begin/end)
  std::vector<cudaEvent_t> laneDone;           // per extra lane: its part of the frame is complete
  // CPUs close to the GPU (sysfs local_cpulist of its PCI device): the library's own host threads — lanes, helpers —
  // run there.  A render chain is a few dozen dependent launches with host round trips; on a two-socket host a thread on
  // the far socket pays the inter-socket hop on every one of them (measured on 8 GPUs: ranks whose threads the OS had
  // placed far away needed 3.3-3.4 ms per 1/8 frame against 2.9 ms).
  std::string cpuList;
  cpu_set_t cpus;
  bool haveCpus = false;
  std::vector<CudaBackend*> subBe;             // per lane: the backend (streams, scratch) of its helper pipeline (the fork)
  std::vector<Helper*> helpers;                // per lane: the helper's host thread
  float* thr[17] = {nullptr};                  // output stage: cut points of the sRGB pow branch per bit depth (device)
  float* outFb = nullptr; unsigned char* outQ = nullptr; int64_t outFbN = 0, outQN = 0;   // nrt_framebuf_* staging (grow-only)
};

static std::mutex g_mu;
static std::vector<DeviceCtx*> g_devs;
static int g_part_index = 0, g_part_count = 1;

// persistent host threads for the lanes / devices beyond the calling thread's
class HostPool {
 public:
  void run(std::vector<std::function<void()>>& jobs) {   // jobs[0] runs on the caller
    if (jobs.empty()) return;
    {
      std::unique_lock<std::mutex> lk(mu_);
      while (threads_.size() + 1 < jobs.size()) threads_.emplace_back([this] { loop(); });
      pending_ = int(jobs.size()) - 1;
      for (size_t i = 1; i < jobs.size(); ++i) q_.push_back(&jobs[i]);
    }
    cv_.notify_all();
    jobs[0]();
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [this] { return pending_ == 0; });
  }
  ~HostPool() {
    { std::unique_lock<std::mutex> lk(mu_); stop_ = true; }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
  }
 private:
  void loop() {
    for (;;) {
      std::function<void()>* job = nullptr;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [this] { return stop_ || !q_.empty(); });
        if (stop_ && q_.empty()) return;
        job = q_.front(); q_.pop_front();
      }
      (*job)();
      { std::unique_lock<std::mutex> lk(mu_); if (--pending_ == 0) done_.notify_all(); }
    }
  }
  std::mutex mu_;
  std::condition_variable cv_, done_;
  std::deque<std::function<void()>*> q_;
  std::vector<std::thread> threads_;
  int pending_ = 0;
  bool stop_ = false;
};
static HostPool g_pool;

// What the last frame of a (scene, device) found, band by band: the number of samples FusedBounce handed to the
// wavefront at bounce 0.  The next frame with the same geometry of bands deals the bands WITH mesh work to the
// high-priority lanes and the others to the low-priority ones (renderImpl: planLanes).  Scheduling only: every band
// is rendered by the same kernels whichever lane takes it.
struct BandFeedback {
  std::vector<int64_t> key;          // frame geometry the numbers belong to
  std::vector<int32_t> y;            // first row of every band of this device, in order
  std::vector<uint32_t> hard;        // per band
};
struct PerDevice {
  SceneData<CudaBackend> sd;
  Renderer<CudaBackend> rn[kMaxLanes];
  Renderer<CudaBackend> rnSub[kMaxLanes];   // the lanes' helper pipelines (the fork)
  BandFeedback fbk;
  float* fbStage = nullptr; int64_t fbStageN = 0;
  unsigned char* qStage = nullptr; int64_t qStageN = 0;   // nrt_render_quantized: the integer image
  int32_t* aovObj = nullptr; int32_t* aovTri = nullptr; double* aovT = nullptr; int64_t aovN = 0;
};

}  // namespace nrt

struct nrt_scene {
  std::vector<nrt::PerDevice> dev;
  nrt_profile prof{};
  nrt_kernel_times ktimes{};   // last frame rendered with kernel timing on (device 0 of the group)
};

namespace nrt {

static int initLocked(int ngpu, const int* ids) {
  if (!g_devs.empty()) return NRT_OK;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0)
    return fail(NRT_ERR_NO_DEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU fallback)");
  std::vector<int> use;
  if (ids && ngpu > 0) use.assign(ids, ids + ngpu);
  else if (ngpu == 0) for (int i = 0; i < count; ++i) use.push_back(i);
  else for (int i = 0; i < std::min(std::max(ngpu, 1), count); ++i) use.push_back(i);
  for (int id : use)
    if (id < 0 || id >= count) return fail(NRT_ERR_INVALID, "device id out of range");
  try {
    for (int id : use) {
      cudaDeviceProp p;
      NRT_CUDA(cudaGetDeviceProperties(&p, id));
      if (p.major < 10) return fail(NRT_ERR_NO_DEVICE, std::string("device ") + p.name + " is not sm_100 class; libnrt.so is built for sm_100a only");
      auto* d = new DeviceCtx();
      d->be.device = id;
      d->be.sms = p.multiProcessorCount;
      if (const char* e = std::getenv("NRT_PREFILTER_CULL")) d->be.cull = std::atoi(e) != 0;
      NRT_CUDA(cudaSetDevice(id));
      d->be.createStreams();
      {   // the device's local CPUs
        char bus[32] = {0};
        if (cudaDeviceGetPCIBusId(bus, sizeof(bus), id) == cudaSuccess) {
          for (char* c = bus; *c; ++c) *c = char(std::tolower(*c));
          std::string path = std::string("/sys/bus/pci/devices/") + bus + "/local_cpulist";
          if (FILE* f = std::fopen(path.c_str(), "r")) {
            char line[1024] = {0};
            if (std::fgets(line, sizeof(line), f)) {
              d->cpuList = line;
              while (!d->cpuList.empty() && (d->cpuList.back() == '\n' || d->cpuList.back() == ' ')) d->cpuList.pop_back();
              CPU_ZERO(&d->cpus);
              int n = 0;
              const char* p = d->cpuList.c_str();
              while (*p) {
                char* e = nullptr;
                long a = std::strtol(p, &e, 10), b = a;
                if (e == p) break;
                if (*e == '-') { p = e + 1; b = std::strtol(p, &e, 10); }
                for (long c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET(int(c), &d->cpus); ++n; }
                p = (*e == ',') ? e + 1 : e;
                if (*e != ',' ) break;
              }
              d->haveCpus = n > 0;
            }
            std::fclose(f);
          }
        }
        cudaGetLastError();
      }
      NRT_CUDA(cudaEventCreate(&d->ev0)); NRT_CUDA(cudaEventCreate(&d->ev1));
      NRT_CUDA(cudaEventCreate(&d->tb0)); NRT_CUDA(cudaEventCreate(&d->tb1));
      g_devs.push_back(d);
    }
    // peer access between the selected devices (NVLink): finalize stores go straight to device 0
    for (size_t i = 0; i < g_devs.size(); ++i)
      for (size_t j = 0; j < g_devs.size(); ++j) {
        if (i == j) continue;
        int can = 0;
        NRT_CUDA(cudaDeviceCanAccessPeer(&can, g_devs[i]->be.device, g_devs[j]->be.device));
        if (can) {
          NRT_CUDA(cudaSetDevice(g_devs[i]->be.device));
          cudaError_t pe = cudaDeviceEnablePeerAccess(g_devs[j]->be.device, 0);
          if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) NRT_CUDA(pe);
          cudaGetLastError();
        }
      }
  } catch (const std::exception& ex) {
    return fail(NRT_ERR_CUDA, ex.what());
  }
  return NRT_OK;
}

// ---- output stage: the cut points of linearToSRGB's pow branch (nrt_pipeline.h: OutStage) ----
// utils/color.nim:17-22 with the literals in the float32 type of `v` and `a` a float64 (see oracle/ref_cpu.cpp);
// powf is the libm call the reference's C back-end makes
static float hostLinearToSRGBPow(float v) {
  const double a = 0.055;
  return float((1 + a) * double(powf(v, float(1 / 2.4))) - a);
}
static uint32_t hostLevel(float v, float maxval) { return uint32_t(roundf(hostLinearToSRGBPow(v) * maxval)); }
static uint32_t fbitsHost(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static float bitsfHost(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
// thr[k - 1] = the smallest float32 in (0.0031308, 1] whose sample is >= k.  Positive floats order like their bit
// patterns, so this is a binary search over patterns; the conversion's monotonicity — which the search and the
// device's counting rely on — is checked on 32 neighbours either side of every cut point.
static bool buildCutPoints(int bits, std::vector<float>& thr) {
  const uint32_t maxv = (1u << bits) - 1u;
  const float maxval = float(maxv);
  const uint32_t lo0 = fbitsHost(0.0031308f) + 1, hi0 = fbitsHost(1.0f);
  thr.assign(maxv, 0.f);
  for (uint32_t k = 1; k <= maxv; ++k) {
    uint32_t lo = lo0, hi = hi0;            // invariant: level(hi) >= k (level(1.0) == maxval)
    if (hostLevel(bitsfHost(lo), maxval) >= k) hi = lo;
    while (lo < hi) {
      const uint32_t mid = lo + (hi - lo) / 2;
      if (hostLevel(bitsfHost(mid), maxval) >= k) hi = mid; else lo = mid + 1;
    }
    thr[k - 1] = bitsfHost(hi);
    for (uint32_t d = 1; d <= 32; ++d) {
      if (hi >= lo0 + d && hostLevel(bitsfHost(hi - d), maxval) >= k) return false;
      if (hi + d <= hi0 && hostLevel(bitsfHost(hi + d), maxval) < k) return false;
    }
  }
  return true;
}
// device copy of the table for `bits` on device context dc (built once)
static const float* cutPoints(DeviceCtx* dc, int bits) {
  if (dc->thr[bits]) return dc->thr[bits];
  std::vector<float> t;
  if (!buildCutPoints(bits, t)) throw std::runtime_error("output stage: powf is not monotonic around a cut point on this host");
  float* d = static_cast<float*>(dc->be.dalloc(sizeof(float) * t.size()));
  dc->be.upload(d, t.data(), sizeof(float) * t.size());
  dc->be.sync();
  dc->thr[bits] = d;
  return d;
}
struct OutSpec { int bits, srgb, rgba, alpha; };
static int64_t outBytesPerPixel(const OutSpec& q) { return q.rgba ? 4 : (q.bits <= 8 ? 3 : 6); }

// The rendered rows of [y0, y1) — (y - y0) % step == 0, numbered i = (y - y0) / step — owned by lane `lane` of
// partition `part`: i % nparts == part (scanline interleave over the GPUs / ranks: the reference's work items,
// raytracer.nim:67-70; counted among the RENDERED rows, so a progressive pass with step >= nparts still spreads
// over every GPU), and among a partition's rows every nlanes-th goes to the same lane.
// With T > 1 (tile order, step == 1) the units dealt out are BANDS of T scanlines — rows of T x T tiles; the vector
// holds their first rows.
// Unit i of a pass goes to partition i mod n in even rounds of n units and to n - 1 - (i mod n) in odd rounds (a
// serpentine deal): every partition's units then have the same mean position inside a round, so a cost that varies
// smoothly down the image — the bunny's rows — no longer favours the partitions of one end of the round (measured on 8
// GPUs with the plain i mod n deal: 2.90 ms on the lightest rank, 3.39 ms on the heaviest, the same ranks every frame).
static inline int unitOwner(int64_t i, int n) {
  const int r = int(i % n);
  return ((i / n) & 1) ? n - 1 - r : r;
}
static std::vector<int32_t> rowsFor(int height, int y0, int y1, int step, int part, int nparts, int lane, int nlanes, int T = 1) {
  std::vector<int32_t> r;
  const int unit = step * T;
  for (int y = std::max(0, y0); y < std::min(y1, height); ++y) {
    if ((y - y0) % unit != 0) continue;
    const int i = (y - y0) / unit;
    if (unitOwner(i, nparts) == part && (i / nparts) % nlanes == lane) r.push_back(y);
  }
  return r;
}

// `qspec` != null: `fb` is the caller's INTEGER image (host memory): Finalize's epilogue converts, only those bytes travel
static int renderImpl(nrt_scene* sc, const nrt_options* o, int y0, int y1, int step, int max_step, float* fb,
                      nrt_stats* stats, const nrt_aov* aov, bool deviceOut, const OutSpec* qspec = nullptr) {
  if (!sc || !o || !fb) return fail(NRT_ERR_INVALID, "null scene, options or framebuffer");
  if (qspec && (qspec->bits < 1 || qspec->bits > 16)) return fail(NRT_ERR_INVALID, "bits must be in 1..16 (framebuf.nim:56)");
  if (o->width <= 0 || o->height <= 0) return fail(NRT_ERR_INVALID, "non-positive image size");
  if (!isPow2(step) || !isPow2(max_step) || max_step < step)
    return fail(NRT_ERR_UNSUPPORTED, "step and maxStep must be powers of two with maxStep >= step (renderer.nim:166-168)");
  if (o->aa_kind < NRT_AA_NONE || o->aa_kind > NRT_AA_CORRELATED_MULTI_JITTERED) return fail(NRT_ERR_INVALID, "bad antialias kind");
  if (o->aa_kind != NRT_AA_NONE && (o->grid_size <= 0 || o->grid_size > 64)) return fail(NRT_ERR_INVALID, "gridSize must be in 1..64");
  if (o->aa_kind >= NRT_AA_JITTERED && o->grid_size > kMaxJitterGrid) return fail(NRT_ERR_UNSUPPORTED, "jittered kinds support gridSize <= 16");
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_devs.empty()) return fail(NRT_ERR_NOT_INIT, "nrt_init() has not been called");
  const int nd = int(sc->dev.size());
  const int64_t npx = int64_t(o->width) * o->height;
  // lanes per device: NRT_LANES, else by the device's share of the frame (a lane should keep >= ~1 M samples;
  // kernel timing wants one chain so that the per-launch events do not overlap)
  int nlanes = 4;
  bool smallShare = false;
  if (const char* e = std::getenv("NRT_LANES")) nlanes = std::atoi(e);
  {
    const int spp = o->aa_kind == NRT_AA_NONE ? 1 : o->grid_size * o->grid_size;
    const int64_t rowsAll = (std::min(y1, o->height) - std::max(0, y0) + step - 1) / std::max(step, 1);
    const int64_t samples = std::max<int64_t>(0, rowsAll) * ((o->width + step - 1) / step) * spp / std::max(1, nd * g_part_count);
    // (small shares — a 1/8 frame — are latency-bound in the wavefront chains: three lanes, one of them taking the bands
    // with mesh work at high priority, measured 3.30 ms against 3.55 ms for four equal lanes; a whole frame is
    // throughput-bound and does best with four equal lanes: 19.8 ms against 21-22 ms with the heavy / light plan)
    smallShare = samples < (int64_t(24) << 20);
    if (!std::getenv("NRT_LANES")) nlanes = int(std::min<int64_t>(smallShare ? 3 : 4, std::max<int64_t>(1, samples / (int64_t(1) << 20))));
    if (g_devs[0]->be.timing && !std::getenv("NRT_TIMELINE")) nlanes = 1;
  }
  nlanes = std::max(1, std::min(nlanes, kMaxLanes));
  const int tshift = tileShiftFor(*o, step, max_step), T = 1 << tshift;   // T > 1: bands of T scanlines, tile order inside
  const int yEnd = std::min(y1, o->height);
  const int nunits = nd * nlanes;
  // ---- the lanes' bands.  Default: band i of the device goes to lane i % nlanes.  With the last frame's per-band
  // counts (BandFeedback) for the same band geometry: the bands are sorted by their count, the longest prefix whose
  // counts fit the light lanes' budget (a wavefront list that short is finished by ONE PathTail launch, nrt_renderer.h)
  // becomes the LIGHT set, the rest the HEAVY set; heavy lanes launch on the high-priority stream, light lanes on the
  // low-priority one.  A heavy lane's frame is FusedBounce over its bands + the latency-bound chain of small
  // wavefront kernels; the light lanes' FusedBounce work fills the SMs under that chain instead of preceding it.
  const size_t ndz = static_cast<size_t>(nd), nlz = static_cast<size_t>(nlanes);
  std::vector<std::vector<std::vector<int32_t>>> laneRows(ndz, std::vector<std::vector<int32_t>>(nlz));
  std::vector<std::vector<int32_t>> devUnits(ndz);
  std::vector<std::vector<char>> laneLow(ndz, std::vector<char>(nlz, 0));
  const std::vector<int64_t> bandKey = {o->width, o->height, y0, y1, step, max_step, T, g_part_index, g_part_count, nd,
                                        o->aa_kind, o->aa_kind == NRT_AA_NONE ? 1 : o->grid_size, nlanes};
  bool useFbk = false;
  {
    const int64_t hardTail = Renderer<CudaBackend>::envInt("NRT_HARD_TAIL_BELOW", 16384);
    useFbk = Renderer<CudaBackend>::envInt("NRT_LANE_FEEDBACK", smallShare ? 1 : 0) != 0 && nlanes >= 2 && hardTail > 0;
    int nHeavy = int(Renderer<CudaBackend>::envInt("NRT_HEAVY_LANES", 1));
    nHeavy = std::max(1, std::min(nHeavy, nlanes - 1));
    for (int di = 0; di < nd; ++di) {
      auto& units = devUnits[size_t(di)];
      units = rowsFor(o->height, y0, y1, step, g_part_index * nd + di, nd * g_part_count, 0, 1, T);
      const BandFeedback& f = sc->dev[size_t(di)].fbk;
      bool planned = false;
      if (useFbk && f.key == bandKey && f.y == units && !units.empty()) {
        std::vector<uint32_t> idx(units.size());
        for (size_t j = 0; j < idx.size(); ++j) idx[j] = uint32_t(j);
        std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return f.hard[a] < f.hard[b]; });
        const int nLight = nlanes - nHeavy;
        const int64_t budget = int64_t(nLight) * hardTail * 3 / 4;
        std::vector<char> light(units.size(), 0);
        int64_t sum = 0; size_t nl = 0;
        for (; nl < idx.size() && sum + f.hard[idx[nl]] <= budget; ++nl) { sum += f.hard[idx[nl]]; light[idx[nl]] = 1; }
        if (nl >= units.size() / 8 && nl < units.size()) {
          planned = true;
          int ih = 0, il = 0;
          for (size_t j = 0; j < units.size(); ++j) {
            if (light[j]) laneRows[size_t(di)][size_t(nHeavy + (il++ % nLight))].push_back(units[j]);
            else laneRows[size_t(di)][size_t(ih++ % nHeavy)].push_back(units[j]);
          }
          for (int ln = nHeavy; ln < nlanes; ++ln) laneLow[size_t(di)][size_t(ln)] = 1;
          if (std::getenv("NRT_TRACE_LANES"))
            fprintf(stderr, "[lanes] dev %d: %zu bands, %zu light (%lld wavefront samples last frame), %d heavy + %d light lanes\n",
                    di, units.size(), nl, (long long)sum, nHeavy, nLight);
        }
      }
      if (!planned && useFbk && std::getenv("NRT_TRACE_LANES"))
        fprintf(stderr, "[lanes] dev %d: no plan (feedback key %s, rows %s, %zu bands)\n", di, f.key == bandKey ? "same" : "differs",
                f.y == units ? "same" : "differ", units.size());
      if (!planned)
        for (size_t j = 0; j < units.size(); ++j) laneRows[size_t(di)][j % size_t(nlanes)].push_back(units[j]);
    }
  }
  std::vector<int> rc(nunits, NRT_OK);
  std::vector<std::string> errs(nunits);
  std::vector<std::vector<unsigned long long>> st(nunits, std::vector<unsigned long long>(ST_COUNT, 0));
  const bool wantAov = aov && (aov->obj_id || aov->tri_id || aov->t_hit);

  // per device, before the lanes start: staging buffers, the extra lanes' backends, the frame's start event
  try {
    for (int di = 0; di < nd; ++di) {
      PerDevice& pd = sc->dev[di];
      DeviceCtx* dc = g_devs[di];
      CudaBackend& be = dc->be;
      be.use();
      for (int k = 0; k < nlanes && k <= int(dc->extra.size()); ++k) {
        CudaBackend& lb = dc->lane(k);
        lb.stream = laneLow[size_t(di)][size_t(k)] ? lb.sLo : lb.sHi;
      }
      while (int(dc->extra.size()) + 1 < nlanes) {
        auto* b = new CudaBackend();
        b->device = be.device; b->sms = be.sms;
        b->createStreams();
        if (laneLow[size_t(di)][dc->extra.size() + 1]) b->stream = b->sLo;
        dc->extra.push_back(b);
        cudaEvent_t e; NRT_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        dc->laneDone.push_back(e);
      }
      while (int(dc->subBe.size()) < nlanes) {
        auto* b = new CudaBackend();
        b->device = be.device; b->sms = be.sms;
        b->createStreams();
        dc->subBe.push_back(b);
        dc->helpers.push_back(new Helper());
      }
      for (int k = 0; k < nlanes; ++k) {
        CudaBackend& sb = *dc->subBe[size_t(k)];
        sb.stream = laneLow[size_t(di)][size_t(k)] ? sb.sLo : sb.sHi;
        dc->lane(k).helper = dc->helpers[size_t(k)];
      }
      if (qspec) {
        const int64_t need = npx * outBytesPerPixel(*qspec);
        if (pd.qStageN < need) { be.dfree(pd.qStage); pd.qStage = static_cast<unsigned char*>(be.dalloc(size_t(need))); pd.qStageN = need; }
        if (!qspec->rgba && qspec->srgb) cutPoints(dc, qspec->bits);
      }
      if (!deviceOut) {
        if (!qspec && pd.fbStageN < npx * 3) {
          be.dfree(pd.fbStage);
          pd.fbStage = static_cast<float*>(be.dalloc(sizeof(float) * npx * 3));
          pd.fbStageN = npx * 3;
        }
        if (wantAov && pd.aovN < npx) {
          be.dfree(pd.aovObj); be.dfree(pd.aovTri); be.dfree(pd.aovT);
          pd.aovObj = static_cast<int32_t*>(be.dalloc(sizeof(int32_t) * npx));
          pd.aovTri = static_cast<int32_t*>(be.dalloc(sizeof(int32_t) * npx));
          pd.aovT = static_cast<double*>(be.dalloc(sizeof(double) * npx));
          pd.aovN = npx;
        }
      }
      NRT_CUDA(cudaEventRecord(dc->ev0, be.stream));
      for (int k = 1; k < nlanes; ++k) NRT_CUDA(cudaStreamWaitEvent(dc->lane(k).stream, dc->ev0, 0));
    }
  } catch (const std::exception& ex) {
    return fail(NRT_ERR_CUDA, ex.what());
  }

  const bool pinThreads = Renderer<CudaBackend>::envInt("NRT_PIN_THREADS", 1) != 0;
  const bool pinCaller = Renderer<CudaBackend>::envInt("NRT_PIN_CALLER", 0) != 0;
  auto work = [&](int unit) {
    const int di = unit / nlanes, ln = unit % nlanes;
    PerDevice& pd = sc->dev[di];
    DeviceCtx* dc = g_devs[di];
    CudaBackend& be = dc->lane(ln);
    pd.rn[ln].be = &be;
    // the library's own threads run on the device's local CPUs (the calling thread — unit 0 — is the host's to place:
    // nrt_device_local_cpus tells it where); NRT_PIN_THREADS=0: leave them to the OS
    if ((unit != 0 || pinCaller) && pinThreads && dc->haveCpus) {
      static thread_local const DeviceCtx* pinnedTo = nullptr;
      if (pinnedTo != dc) { pthread_setaffinity_np(pthread_self(), sizeof(cpu_set_t), &dc->cpus); pinnedTo = dc; }
    }
    CudaBackend& sbe = *dc->subBe[size_t(ln)];
    pd.rnSub[ln].be = &sbe;
    pd.rn[ln].sub = &pd.rnSub[ln];
    pd.rn[ln].wantBandCounts = useFbk;
    try {
      be.use();
      const std::vector<int32_t>& rows = laneRows[size_t(di)][size_t(ln)];
      float* target = fb;
      int32_t *aObj = nullptr, *aTri = nullptr; double* aT = nullptr;
      OutStage qs{};
      if (qspec) {
        qs.out = pd.qStage; qs.bits = qspec->bits; qs.srgb = qspec->srgb; qs.rgba = qspec->rgba; qs.alpha = qspec->alpha;
        qs.thr = (!qspec->rgba && qspec->srgb) ? dc->thr[qspec->bits] : nullptr;
      }
      const size_t fbElem = qspec ? size_t(outBytesPerPixel(*qspec)) : 3 * sizeof(float);
      void* const fbDev = qspec ? static_cast<void*>(pd.qStage) : static_cast<void*>(pd.fbStage);
      if (!deviceOut) {
        target = pd.fbStage;
        if (wantAov) { aObj = aov->obj_id ? pd.aovObj : nullptr; aTri = aov->tri_id ? pd.aovTri : nullptr; aT = aov->t_hit ? pd.aovT : nullptr; }
      } else if (wantAov) {
        aObj = aov->obj_id; aTri = aov->tri_id; aT = aov->t_hit;
      }
      be.launches = 0; sbe.launches = 0;
      be.timing = sbe.timing = dc->be.timing;
      if (const char* e = std::getenv("NRT_PREFILTER_CULL")) be.cull = std::atoi(e) != 0; else be.cull = true;
      if (const char* e = std::getenv("NRT_PREFILTER_SPLIT")) be.splitBelow = std::max(0, std::atoi(e));
      if (const char* e = std::getenv("NRT_PREFETCH_AHEAD")) be.prefetchAhead = std::max<int64_t>(0, std::atoll(e));
      else be.prefetchAhead = int64_t(be.sms) * 512;   // measured on B200, config 4: off 25.23 ms; 256..2048 per SM 24.75-24.81; 4096 per SM 25.28
      sbe.cull = be.cull; sbe.splitBelow = be.splitBelow; sbe.prefetchAhead = be.prefetchAhead;
      // Host <-> staging copies of exactly the rows this worker renders (and their step x step
      // fill rows); equally spaced rows (scanline interleave) go out as one 2D copy.
      auto copyRows = [&](bool toHost) {
        const cudaMemcpyKind kind = toHost ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice;
        auto one = [&](void* host, void* devp, size_t elem, size_t firstRow, int stride, int fill, size_t cnt) {
          const size_t rowB = size_t(o->width) * elem;
          char* h = static_cast<char*>(host) + firstRow * rowB;
          char* d = static_cast<char*>(devp) + firstRow * rowB;
          void* dst = toHost ? static_cast<void*>(h) : static_cast<void*>(d);
          const void* src = toHost ? static_cast<const void*>(d) : static_cast<const void*>(h);
          if (cnt == 1) NRT_CUDA(cudaMemcpyAsync(dst, src, rowB * fill, kind, be.stream));
          else NRT_CUDA(cudaMemcpy2DAsync(dst, rowB * stride, src, rowB * stride, rowB * fill, cnt, kind, be.stream));
        };
        size_t i = 0;
        while (i < rows.size()) {
          size_t j = i + 1;
          // rows a unit covers: its step x step fill block (progressive) or its band of T scanlines, clipped
          auto fillOf = [&](int row) { return T > 1 ? std::min(T, yEnd - row) : std::min(step, o->height - row); };
          const int fill = fillOf(rows[i]);
          const int stride = (j < rows.size()) ? rows[j] - rows[i] : 0;
          while (stride > 0 && j < rows.size() && rows[j] - rows[j - 1] == stride && fillOf(rows[j]) == fill) ++j;
          const size_t cnt = j - i;
          one(fb, fbDev, fbElem, size_t(rows[i]), stride, fill, cnt);
          if (aObj) one(aov->obj_id, aObj, sizeof(int32_t), size_t(rows[i]), stride, fill, cnt);
          if (aTri) one(aov->tri_id, aTri, sizeof(int32_t), size_t(rows[i]), stride, fill, cnt);
          if (aT) one(aov->t_hit, aT, sizeof(double), size_t(rows[i]), stride, fill, cnt);
          i = j;
        }
      };
      // Progressive passes leave some pixels of the touched rows untouched (renderer.nim:175-178,
      // and AOVs exist only at rendered pixels): round-trip the caller's current content.
      if (!deviceOut && (step > 1 || step < max_step)) copyRows(false);
      rc[unit] = pd.rn[ln].render(pd.sd, *o, rows, step, max_step, target, aObj, aTri, aT, st[unit].data(), errs[unit], qspec ? &qs : nullptr, tshift, yEnd);
      if (rc[unit] == NRT_OK && !deviceOut) copyRows(true);
      if (ln > 0) NRT_CUDA(cudaEventRecord(dc->laneDone[size_t(ln - 1)], be.stream));
      NRT_CUDA(cudaStreamSynchronize(be.stream));
    } catch (const std::exception& ex) {
      rc[unit] = NRT_ERR_CUDA;
      errs[unit] = ex.what();
    }
  };

  if (nunits == 1) work(0);
  else {
    std::vector<std::function<void()>> jobs;
    for (int u = 0; u < nunits; ++u) jobs.emplace_back([&work, u] { work(u); });
    g_pool.run(jobs);
  }
  // the frame ends when every lane of the device has ended: ev1 on lane 0's stream behind the other lanes' events
  for (int di = 0; di < nd; ++di) {
    DeviceCtx* dc = g_devs[di];
    cudaSetDevice(dc->be.device);
    for (int k = 1; k < nlanes; ++k) cudaStreamWaitEvent(dc->be.stream, dc->laneDone[size_t(k - 1)], 0);
    cudaEventRecord(dc->ev1, dc->be.stream);
    cudaStreamSynchronize(dc->be.stream);
  }
  for (int u = 0; u < nunits; ++u)
    if (rc[u] != NRT_OK) return fail(rc[u], errs[u]);

  // The automatic path choice (nrt_renderer.h: meshShare) is a property of the device's share of the frame, not of a
  // lane: with the heavy / light lane plan the heavy lane alone would cross the threshold, fall back to the wavefront
  // for every bounce, deliver no band counts, lose the plan for the next frame, cross back ... (seen on five of eight
  // ranks: 3.3 ms and 2.9 ms frames in turn).  Every lane gets the device-wide figure.
  for (int di = 0; di < nd; ++di) {
    PerDevice& pd = sc->dev[size_t(di)];
    double wave0 = 0, act0 = 0, meshRays = 0, primary = 0;
    bool anyFused = false, anyWave = false;
    for (int ln = 0; ln < nlanes; ++ln) {
      if (laneRows[size_t(di)][size_t(ln)].empty()) continue;
      const Renderer<CudaBackend>& r = pd.rn[ln];
      if (r.pathMode == 1) { anyFused = true; wave0 += double(r.prof.wavefront[0]); act0 += double(r.prof.active[0]); }
      else if (r.pathMode == 0) { anyWave = true; meshRays += double(r.prof.mesh_rays); primary += double(st[size_t(di * nlanes + ln)][ST_PRIMARY]); }
    }
    double share = -1.0;
    if (anyFused && !anyWave && act0 > 0) share = wave0 / act0;
    else if (anyWave && !anyFused && primary > 0) share = std::min(1.0, meshRays / primary);
    if (share >= 0.0)
      for (int ln = 0; ln < kMaxLanes; ++ln) pd.rn[ln].meshShare = share;
  }

  // the per-band counts of this frame, for the next one's lane plan
  for (int di = 0; di < nd; ++di) {
    PerDevice& pd = sc->dev[size_t(di)];
    BandFeedback& f = pd.fbk;
    const auto& units = devUnits[size_t(di)];
    std::vector<uint32_t> hard(units.size(), 0);
    bool ok = !units.empty();
    for (int ln = 0; ln < nlanes && ok; ++ln) {
      const auto& rows = laneRows[size_t(di)][size_t(ln)];
      const auto& bh = pd.rn[ln].bandHard;
      if (rows.empty()) continue;
      if (bh.size() != rows.size()) { ok = false; break; }
      for (size_t k = 0; k < rows.size(); ++k) {
        const auto it = std::lower_bound(units.begin(), units.end(), rows[k]);
        if (it == units.end() || *it != rows[k]) { ok = false; break; }
        hard[size_t(it - units.begin())] = bh[k];
      }
    }
    if (ok) { f.key = bandKey; f.y = units; f.hard.swap(hard); }
    else {
      if (useFbk && std::getenv("NRT_TRACE_LANES")) {
        fprintf(stderr, "[lanes] dev %d: feedback dropped:", di);
        for (int ln = 0; ln < nlanes; ++ln) fprintf(stderr, " lane %d rows %zu counts %zu", ln, laneRows[size_t(di)][size_t(ln)].size(), pd.rn[ln].bandHard.size());
        fprintf(stderr, "\n");
      }
      f.key.clear(); f.y.clear(); f.hard.clear();
    }
  }

  nrt_profile& p = sc->prof;
  p = nrt_profile{};
  unsigned long long tot[ST_COUNT] = {0};
  for (int u = 0; u < nunits; ++u)
    for (int k = 0; k < ST_COUNT; ++k) tot[k] += st[u][k];
  for (int d = 0; d < nd; ++d) {
    DeviceCtx* dc = g_devs[d];
    float ms = 0;
    cudaSetDevice(dc->be.device);
    cudaEventElapsedTime(&ms, dc->ev0, dc->ev1);
    p.total_ms = std::max(p.total_ms, double(ms));
   double fmsDev = 0;
   for (int lb = 0; lb < 2 * nlanes; ++lb) {
    const int ln = lb >> 1;
    const bool isSub = (lb & 1) != 0;         // the lane's helper pipeline: its launches, prefilter times and timeline
    CudaBackend& lbe = isSub ? *dc->subBe[size_t(ln)] : dc->lane(ln);
    int64_t nl = 0;
    double byMode[4] = {0, 0, 0, 0};
    std::vector<float> each;
    const bool traceP = std::getenv("NRT_TRACE_PREFILTER") != nullptr;
    const double fms = lbe.filterMs(&nl, byMode, traceP ? &each : nullptr);
    if (traceP) {
      const auto& lg = isSub ? sc->dev[d].rnSub[ln].preLog : sc->dev[d].rn[ln].preLog;
      for (size_t i = 0; i < lg.size() && i < each.size(); ++i)
        fprintf(stderr, "[prefilter] dev %d wave %2d mo %d bundle %d mode %d rays %9lld runs %7lld chunks %4lld work %9lld (%.2f per run) pre %9lld  %8.1f us\n",
                d, lg[i].wave, lg[i].mo, lg[i].b, lg[i].mode, (long long)lg[i].rays,
                (long long)((lg[i].rays + prefilterRunRays(lg[i].mode) - 1) / prefilterRunRays(lg[i].mode)), (long long)lg[i].nch,
                (long long)lg[i].work, double(lg[i].work) / std::max<double>(1.0, double((lg[i].rays + prefilterRunRays(lg[i].mode) - 1) / prefilterRunRays(lg[i].mode))),
                (long long)lg[i].pre, each[i] * 1e3);
    }
    fmsDev += fms;   // (the lanes' prefilter launches of one device: summed CUDA-event time; they may overlap)
    p.mesh_filter_launches += nl;
    p.kernel_launches += lbe.launches;
    if (lbe.timing)
      if (const char* tl = std::getenv("NRT_TIMELINE"))
        if (FILE* f = std::fopen(tl, "a")) { lbe.dumpTimeline(f, dc->ev0, d, ln + (isSub ? 10 : 0)); std::fclose(f); }
    if (lbe.timing && d == 0 && ln == 0) {
      if (!isSub) sc->ktimes = nrt_kernel_times{};
      lbe.collectTimes(sc->ktimes.ms, sc->ktimes.launches, sc->ktimes.max_ms);
    } else if (lbe.timing) {
      nrt_kernel_times scratchT{};
      lbe.collectTimes(scratchT.ms, scratchT.launches, scratchT.max_ms);
    }
    for (int m = 0; m < 3; ++m) p.mesh_ms_by_mode[m] += byMode[m];
    if (isSub) continue;                       // (the helper's counters are merged into its lane's profile)
    const ProfileAcc& a = sc->dev[d].rn[ln].prof;
    p.mesh_tests += a.mesh_tests; p.mesh_tests_ref += a.mesh_tests_ref; p.mesh_rays += a.mesh_rays;
    p.candidates += a.candidates;
    p.pre_candidates += a.pre_candidates;
    for (int b = 0; b < 8; ++b) { p.active_samples[b] += a.active[b]; p.wavefront_samples[b] += a.wavefront[b]; }
    p.tail_samples += a.tail;
    p.lanes = nlanes;
    for (int m = 0; m < 3; ++m) {
      p.mesh_tests_by_mode[m] += a.tests_by_mode[m];
      // executed float32 flops per prefilter test (FFMA = 2): nrt_filter.h prefilterFlops()
      p.fp32_flops += double(a.tests_by_mode[m]) * prefilterFlops(m);
    }
   }
   p.mesh_filter_ms = std::max(p.mesh_filter_ms, fmsDev);
  }
  if (stats) {
    stats->num_primary_rays = int64_t(tot[ST_PRIMARY]);
    stats->num_intersection_tests = int64_t(tot[ST_TESTS]);
    stats->num_intersection_hits = int64_t(tot[ST_HITS]);
    stats->num_rays = int64_t(tot[ST_RAYS]);
    stats->num_capped_samples = int64_t(tot[ST_CAPPED]);
  }
  return NRT_OK;
}

}  // namespace nrt

using namespace nrt;

// ---------------------------------------------------------------- C ABI ------
extern "C" {

int nrt_abi_version(void) { return NRT_ABI_VERSION; }
const char* nrt_last_error(void) { return g_err.c_str(); }
int nrt_band_rows(void) { return 1; }
int nrt_band_rows_for(const nrt_options* opts, int step, int max_step) { return opts ? (1 << tileShiftFor(*opts, step, max_step)) : 1; }
int nrt_partition_rows(int height, int y0, int y1, int step, int band, int index, int count, int* rows, int cap) {
  if (height <= 0 || step <= 0 || band <= 0 || count <= 0 || index < 0 || index >= count || cap < 0 || (cap > 0 && !rows)) return -1;
  const std::vector<int32_t> r = rowsFor(height, y0, y1, step, index, count, 0, 1, band);   // (the enumeration renderImpl deals out)
  for (size_t k = 0; k < r.size() && k < size_t(cap); ++k) rows[k] = r[k];
  return int(r.size());
}

int nrt_init(int ngpu, const int* dev_ids) {
  std::lock_guard<std::mutex> lk(g_mu);
  return initLocked(ngpu, dev_ids);
}

void nrt_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto* d : g_devs) {
    cudaSetDevice(d->be.device);
    if (d->ev0) cudaEventDestroy(d->ev0);
    if (d->ev1) cudaEventDestroy(d->ev1);
    if (d->tb0) cudaEventDestroy(d->tb0);
    if (d->tb1) cudaEventDestroy(d->tb1);
    for (auto e : d->laneDone) cudaEventDestroy(e);
    for (auto*& t : d->thr) { if (t) cudaFree(t); t = nullptr; }
    if (d->outFb) cudaFree(d->outFb);
    if (d->outQ) cudaFree(d->outQ);
    for (auto* h : d->helpers) delete h;
    for (auto* b : d->subBe) { b->destroy(); delete b; }
    for (auto* b : d->extra) { b->destroy(); delete b; }
    d->be.destroy();
    delete d;
  }
  g_devs.clear();
  g_part_index = 0; g_part_count = 1;
}

int nrt_device_count(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  return int(g_devs.size());
}

int nrt_device_local_cpus(int index, char* buf, int buflen) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (index < 0 || index >= int(g_devs.size()) || !buf || buflen <= 0) return fail(NRT_ERR_INVALID, "bad device index or buffer");
  const std::string& c = g_devs[size_t(index)]->cpuList;
  if (int(c.size()) + 1 > buflen) return fail(NRT_ERR_INVALID, "buffer too small");
  std::memcpy(buf, c.c_str(), c.size() + 1);
  return NRT_OK;
}

int nrt_unit_owner(long long unit, int count) { return (count > 0 && unit >= 0) ? unitOwner(unit, count) : -1; }

int nrt_set_partition(int index, int count) {
  if (count <= 0 || index < 0 || index >= count) return fail(NRT_ERR_INVALID, "bad partition");
  std::lock_guard<std::mutex> lk(g_mu);
  g_part_index = index; g_part_count = count;
  return NRT_OK;
}

static int sceneBuild(nrt_scene* s, const nrt_scene_desc* desc, bool reuse) {
  for (size_t d = 0; d < s->dev.size(); ++d) {
    std::string err;
    try {
      PerDevice& pd = s->dev[d];
      for (int k = 0; k < kMaxLanes; ++k) if (!pd.rn[k].be) pd.rn[k].be = &g_devs[d]->be;
      const int rc = pd.sd.build(&g_devs[d]->be, desc, reuse, err);
      if (rc != NRT_OK) return fail(rc, err);
      g_devs[d]->be.sync();
    } catch (const std::exception& ex) {
      return fail(NRT_ERR_CUDA, ex.what());
    }
  }
  return NRT_OK;
}

int nrt_scene_create(const nrt_scene_desc* desc, nrt_scene** out) {
  if (!desc || !out) return fail(NRT_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_devs.empty()) return fail(NRT_ERR_NOT_INIT, "nrt_init() has not been called");
  auto* s = new nrt_scene();
  s->dev.resize(g_devs.size());
  const int rc = sceneBuild(s, desc, false);
  if (rc != NRT_OK) {
    for (auto& pd : s->dev) pd.sd.destroy();
    delete s;
    return rc;
  }
  *out = s;
  return NRT_OK;
}

int nrt_scene_update(nrt_scene* scene, const nrt_scene_desc* desc) {
  if (!scene || !desc) return fail(NRT_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_devs.empty()) return fail(NRT_ERR_NOT_INIT, "nrt_init() has not been called");
  return sceneBuild(scene, desc, true);
}

void nrt_scene_destroy(nrt_scene* scene) {
  if (!scene) return;
  std::lock_guard<std::mutex> lk(g_mu);
  for (size_t d = 0; d < scene->dev.size() && d < g_devs.size(); ++d) {
    PerDevice& pd = scene->dev[d];
    CudaBackend& be = g_devs[d]->be;
    pd.sd.destroy();
    for (int k = 0; k < kMaxLanes; ++k) if (pd.rn[k].be) pd.rn[k].freeAll();
    for (int k = 0; k < kMaxLanes; ++k) if (pd.rnSub[k].be) pd.rnSub[k].freeAll();
    be.dfree(pd.fbStage); be.dfree(pd.qStage); be.dfree(pd.aovObj); be.dfree(pd.aovTri); be.dfree(pd.aovT);
  }
  delete scene;
}

int nrt_render(nrt_scene* scene, const nrt_options* opts, int y0, int y1, int step, int max_step, float* fb,
               nrt_stats* stats, const nrt_aov* aov) {
  return renderImpl(scene, opts, y0, y1, step, max_step, fb, stats, aov, false);
}

int nrt_render_device(nrt_scene* scene, const nrt_options* opts, int y0, int y1, int step, int max_step, float* fb_dev,
                      nrt_stats* stats, const nrt_aov* aov_dev) {
  return renderImpl(scene, opts, y0, y1, step, max_step, fb_dev, stats, aov_dev, true);
}

int nrt_set_kernel_timing(int enable) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_devs.empty()) return fail(NRT_ERR_NOT_INIT, "nrt_init() has not been called");
  for (auto* d : g_devs) { d->be.timing = enable != 0; d->be.tUsed = 0; }
  return NRT_OK;
}
int nrt_get_kernel_times(const nrt_scene* scene, nrt_kernel_times* out) {
  if (!scene || !out) return fail(NRT_ERR_INVALID, "null argument");
  *out = scene->ktimes;
  return NRT_OK;
}
const char* nrt_kernel_category_name(int category) {
  return (category >= 0 && category < KC_COUNT) ? kKernelCatNames[category] : "";
}

int nrt_get_profile(const nrt_scene* scene, nrt_profile* out) {
  if (!scene || !out) return fail(NRT_ERR_INVALID, "null argument");
  *out = scene->prof;
  return NRT_OK;
}

int nrt_output_cut_points(int bits, float* thr_host) {
  if (bits < 1 || bits > 16 || !thr_host) return fail(NRT_ERR_INVALID, "bad argument");
  std::vector<float> t;
  if (!buildCutPoints(bits, t)) return fail(NRT_ERR_UNSUPPORTED, "powf is not monotonic around a cut point on this host");
  std::memcpy(thr_host, t.data(), sizeof(float) * t.size());
  return NRT_OK;
}

int nrt_render_quantized(nrt_scene* scene, const nrt_options* opts, int y0, int y1, int step, int max_step,
                         int format, int bits, int srgb, int alpha, void* image, nrt_stats* stats) {
  if (format != NRT_OUT_RGB && format != NRT_OUT_RGBA8) return fail(NRT_ERR_INVALID, "bad output format");
  const OutSpec q{format == NRT_OUT_RGBA8 ? 8 : bits, format == NRT_OUT_RGBA8 ? 0 : (srgb != 0), format == NRT_OUT_RGBA8, alpha & 0xFF};
  return renderImpl(scene, opts, y0, y1, step, max_step, static_cast<float*>(image), stats, nullptr, false, &q);
}

// fb / out: device pointers when `device`, else host memory staged through grow-only device buffers
static int quantizeImpl(const float* fb, int width, int height, int bits, int srgb, void* out, bool rgba, unsigned char alpha, bool device) {
  if (!fb || !out || width <= 0 || height <= 0 || bits < 1 || bits > 16) return fail(NRT_ERR_INVALID, "bad argument");
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_devs.empty()) return fail(NRT_ERR_NOT_INIT, "nrt_init() has not been called");
  DeviceCtx* dc = g_devs[0];
  CudaBackend& be = dc->be;
  const int64_t npix = int64_t(width) * height, n = npix * 3;
  const OutSpec spec{bits, srgb != 0, rgba, alpha};
  const int64_t outBytes = npix * outBytesPerPixel(spec);
  try {
    be.use();
    OutStage q{};
    q.bits = bits; q.srgb = srgb != 0; q.rgba = rgba; q.alpha = alpha;
    q.thr = (!rgba && srgb) ? cutPoints(dc, bits) : nullptr;
    const float* src = fb;
    if (device) q.out = static_cast<unsigned char*>(out);
    else {
      if (dc->outFbN < n) { be.dfree(dc->outFb); dc->outFb = static_cast<float*>(be.dalloc(size_t(n) * sizeof(float))); dc->outFbN = n; }
      if (dc->outQN < outBytes) { be.dfree(dc->outQ); dc->outQ = static_cast<unsigned char*>(be.dalloc(size_t(outBytes))); dc->outQN = outBytes; }
      be.upload(dc->outFb, fb, size_t(n) * sizeof(float));
      src = dc->outFb; q.out = dc->outQ;
    }
    k_out_stage<<<CudaBackend::blocksFor(npix), kBlock, 0, be.stream>>>(src, q, npix);
    NRT_CUDA(cudaGetLastError());
    if (device) be.sync();
    else be.download(out, dc->outQ, size_t(outBytes));
  } catch (const std::exception& ex) {
    return fail(NRT_ERR_CUDA, ex.what());
  }
  return NRT_OK;
}
int nrt_framebuf_to_srgb8(const float* fb_host, int width, int height, int srgb, unsigned char* rgb8) {
  return quantizeImpl(fb_host, width, height, 8, srgb, rgb8, false, 0, false);
}
int nrt_framebuf_quantize(const float* fb_host, int width, int height, int bits, int srgb, void* out) {
  return quantizeImpl(fb_host, width, height, bits, srgb, out, false, 0, false);
}
int nrt_framebuf_to_rgba8(const float* fb_host, int width, int height, unsigned char alpha, unsigned char* rgba8) {
  return quantizeImpl(fb_host, width, height, 8, 0, rgba8, true, alpha, false);
}
int nrt_framebuf_quantize_device(const float* fb_dev, int width, int height, int bits, int srgb, void* out_dev) {
  return quantizeImpl(fb_dev, width, height, bits, srgb, out_dev, false, 0, true);
}
int nrt_framebuf_to_rgba8_device(const float* fb_dev, int width, int height, unsigned char alpha, unsigned char* rgba8_dev) {
  return quantizeImpl(fb_dev, width, height, 8, 0, rgba8_dev, true, alpha, true);
}

#define NRT_NEED_DEV0()                                                               \
  std::lock_guard<std::mutex> lk(g_mu);                                               \
  if (g_devs.empty()) return fail(NRT_ERR_NOT_INIT, "nrt_init() has not been called"); \
  CudaBackend& be = g_devs[0]->be;

int nrt_device_alloc(int64_t bytes, void** dev_ptr) {
  if (!dev_ptr || bytes < 0) return fail(NRT_ERR_INVALID, "bad argument");
  NRT_NEED_DEV0();
  try { *dev_ptr = be.dalloc(size_t(bytes)); } catch (const std::exception& ex) { return fail(NRT_ERR_CUDA, ex.what()); }
  return NRT_OK;
}
int nrt_device_free(void* dev_ptr) {
  NRT_NEED_DEV0();
  be.dfree(dev_ptr);
  return NRT_OK;
}
int nrt_device_memset(void* dev_ptr, int value, int64_t bytes) {
  NRT_NEED_DEV0();
  try { be.use(); NRT_CUDA(cudaMemsetAsync(dev_ptr, value, size_t(bytes), be.stream)); be.sync(); }
  catch (const std::exception& ex) { return fail(NRT_ERR_CUDA, ex.what()); }
  return NRT_OK;
}
int nrt_copy_to_device(void* dev_dst, const void* host_src, int64_t bytes) {
  if (!dev_dst || !host_src || bytes < 0) return fail(NRT_ERR_INVALID, "bad argument");
  NRT_NEED_DEV0();
  try { be.upload(dev_dst, host_src, size_t(bytes)); be.sync(); } catch (const std::exception& ex) { return fail(NRT_ERR_CUDA, ex.what()); }
  return NRT_OK;
}
int nrt_copy_to_host(void* host_dst, const void* dev_src, int64_t bytes) {
  NRT_NEED_DEV0();
  try { be.download(host_dst, dev_src, size_t(bytes)); } catch (const std::exception& ex) { return fail(NRT_ERR_CUDA, ex.what()); }
  return NRT_OK;
}
int nrt_device_synchronize(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  try { for (auto* d : g_devs) d->be.sync(); } catch (const std::exception& ex) { return fail(NRT_ERR_CUDA, ex.what()); }
  return NRT_OK;
}
int nrt_ipc_export(void* dev_ptr, nrt_ipc_handle* out) {
  if (!dev_ptr || !out) return fail(NRT_ERR_INVALID, "null argument");
  NRT_NEED_DEV0();
  static_assert(sizeof(cudaIpcMemHandle_t) <= sizeof(nrt_ipc_handle), "handle size");
  cudaIpcMemHandle_t h;
  be.use();
  cudaError_t e = cudaIpcGetMemHandle(&h, dev_ptr);
  if (e != cudaSuccess) return fail(NRT_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
  std::memset(out, 0, sizeof(*out));
  std::memcpy(out->bytes, &h, sizeof(h));
  return NRT_OK;
}
int nrt_ipc_open(const nrt_ipc_handle* handle, void** dev_ptr) {
  if (!handle || !dev_ptr) return fail(NRT_ERR_INVALID, "null argument");
  NRT_NEED_DEV0();
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle->bytes, sizeof(h));
  be.use();
  cudaError_t e = cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return fail(NRT_ERR_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
  return NRT_OK;
}
int nrt_ipc_close(void* dev_ptr) {
  NRT_NEED_DEV0();
  be.use();
  cudaError_t e = cudaIpcCloseMemHandle(dev_ptr);
  if (e != cudaSuccess) return fail(NRT_ERR_CUDA, std::string("cudaIpcCloseMemHandle: ") + cudaGetErrorString(e));
  return NRT_OK;
}

int nrt_measure_fp32_peak(double* tflops, double* sm_clock_mhz_hint) {
  if (!tflops) return fail(NRT_ERR_INVALID, "null argument");
  NRT_NEED_DEV0();
  try {
    be.use();
    float* out = static_cast<float*>(be.dalloc(16));
    cudaEvent_t e0, e1;
    NRT_CUDA(cudaEventCreate(&e0)); NRT_CUDA(cudaEventCreate(&e1));
    const int iters = 4096, blocks = be.sms * 8;
    double best = 0;
    for (int rep = 0; rep < 6; ++rep) {
      NRT_CUDA(cudaEventRecord(e0, be.stream));
      k_ffma_peak<<<blocks, 256, 0, be.stream>>>(out, iters, 1.0001f, 0.0001f);
      NRT_CUDA(cudaEventRecord(e1, be.stream));
      NRT_CUDA(cudaStreamSynchronize(be.stream));
      float ms = 0;
      NRT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      const double fl = double(blocks) * 256.0 * iters * 16.0 * 2.0;
      if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    be.dfree(out);
    *tflops = best;
    if (sm_clock_mhz_hint) {
      int khz = 0;
      cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, be.device);
      *sm_clock_mhz_hint = khz / 1000.0;
    }
  } catch (const std::exception& ex) { return fail(NRT_ERR_CUDA, ex.what()); }
  return NRT_OK;
}

// register-resident DFMA loop: the float64 pipe's roofline denominator (FusedBounce is bound by it and by issue slots)
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters, double a, double b) {
  double x[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] = double(threadIdx.x + k);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = fma(x[k], a, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += x[k];
  if (s == 12345.678) out[0] = s;
}
int nrt_measure_fp64_peak(double* tflops) {
  if (!tflops) return fail(NRT_ERR_INVALID, "null argument");
  NRT_NEED_DEV0();
  try {
    be.use();
    double* out = static_cast<double*>(be.dalloc(16));
    cudaEvent_t e0, e1;
    NRT_CUDA(cudaEventCreate(&e0)); NRT_CUDA(cudaEventCreate(&e1));
    const int iters = 2048, blocks = be.sms * 8;
    double best = 0;
    for (int rep = 0; rep < 6; ++rep) {
      NRT_CUDA(cudaEventRecord(e0, be.stream));
      k_dfma_peak<<<blocks, 256, 0, be.stream>>>(out, iters, 1.0001, 0.0001);
      NRT_CUDA(cudaEventRecord(e1, be.stream));
      NRT_CUDA(cudaStreamSynchronize(be.stream));
      float ms = 0;
      NRT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      const double fl = double(blocks) * 256.0 * iters * 8.0 * 2.0;
      if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    be.dfree(out);
    *tflops = best;
  } catch (const std::exception& ex) { return fail(NRT_ERR_CUDA, ex.what()); }
  return NRT_OK;
}

int nrt_timer_begin(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_devs.empty()) return fail(NRT_ERR_NOT_INIT, "nrt_init() has not been called");
  try {
    for (auto* d : g_devs) { d->be.use(); NRT_CUDA(cudaEventRecord(d->tb0, d->be.stream)); }
  } catch (const std::exception& ex) { return fail(NRT_ERR_CUDA, ex.what()); }
  return NRT_OK;
}

int nrt_timer_end(double* ms) {
  if (!ms) return fail(NRT_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_devs.empty()) return fail(NRT_ERR_NOT_INIT, "nrt_init() has not been called");
  try {
    double worst = 0;
    for (auto* d : g_devs) { d->be.use(); NRT_CUDA(cudaEventRecord(d->tb1, d->be.stream)); }
    for (auto* d : g_devs) {
      d->be.use();
      NRT_CUDA(cudaEventSynchronize(d->tb1));
      float t = 0;
      NRT_CUDA(cudaEventElapsedTime(&t, d->tb0, d->tb1));
      worst = std::max(worst, double(t));
    }
    *ms = worst;
  } catch (const std::exception& ex) { return fail(NRT_ERR_CUDA, ex.what()); }
  return NRT_OK;
}

int nrt_host_alloc_pinned(int64_t bytes, void** host_ptr) {
  if (!host_ptr || bytes < 0) return fail(NRT_ERR_INVALID, "bad argument");
  cudaError_t e = cudaHostAlloc(host_ptr, size_t(bytes ? bytes : 16), cudaHostAllocPortable);
  if (e != cudaSuccess) return fail(NRT_ERR_CUDA, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
  return NRT_OK;
}
int nrt_host_free_pinned(void* host_ptr) {
  if (host_ptr) cudaFreeHost(host_ptr);
  return NRT_OK;
}
int nrt_host_register(void* host_ptr, int64_t bytes) {
  if (!host_ptr || bytes <= 0) return fail(NRT_ERR_INVALID, "bad argument");
  cudaError_t e = cudaHostRegister(host_ptr, size_t(bytes), cudaHostRegisterPortable);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(NRT_ERR_CUDA, std::string("cudaHostRegister: ") + cudaGetErrorString(e)); }
  return NRT_OK;
}
int nrt_host_unregister(void* host_ptr) {
  if (!host_ptr) return NRT_OK;
  cudaError_t e = cudaHostUnregister(host_ptr);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(NRT_ERR_CUDA, std::string("cudaHostUnregister: ") + cudaGetErrorString(e)); }
  return NRT_OK;
}

}  // extern "C"
