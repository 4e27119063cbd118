// nrt_pipeline.h — wavefront formulation of the reference's per-pixel path.
//
// The reference evaluates  renderLine -> calcPixel* -> castPrimaryRay / trace /
// shade  recursively per pixel on CPU threads (renderer.nim:31-211).  Here the
// same computation is organised as waves of rays over a chunk of samples:
//
//   gen      : castPrimaryRay for every sample + the gate of its mesh boxes  (renderer.nim:31-44,132-159)
//   per bounce:
//     gate   : per (ray, mesh object) AABB gate of TriangleMesh.intersect
//              (geom.nim:340); rays that enter a box go, IN WAVE ORDER, into one
//              queue of float32 filter rays per ray bundle
//     prefilter / refine : float32 ray x face filter over the queues (nrt_filter.h):
//              chunk bounds -> sub-chunk bounds -> bounding circles -> sign test
//     verify : float64 exact re-evaluation of the candidates, nearest hit with
//              first-index-wins ties                        (geom.nim:346-358)
//     shade  : trace()'s in-order object scan + hit point/normal (renderer.nim:47-88)
//     gate/prefilter/refine/verify again for the shadow rays (renderer.nim:93-104)
//     shadow trace : trace() of the shadow rays -> occlusion flags (renderer.nim:101-103)
//     resolve: shadeDiffuse of the unoccluded lights, reflection ray (renderer.nim:90-127)
//   finalize : per-pixel sample sum, * 1/N, float32 store   (renderer.nim:147-159,204-209)
//
// The functors below are the per-element bodies; a backend (CUDA in nrt.cu,
// plain loops in the test-only emulation) supplies launch/alloc primitives.
#pragma once

#include "nrt_core.h"

namespace nrt {

enum WaveKind { WAVE_PATH = 0, WAVE_SHADOW = 1 };

// Counter block per (wave, mesh object): [EXACT, CAND, NE, then (QUEUE_b, TILE_b, PRE_b, WORK_b, SUB_b, BND_b) per ray bundle b].
// Bundle 0 holds arbitrary rays (GENERAL mode; ORIGIN mode for the primary wave, whose rays share
// the camera origin); bundle 1 + l holds the shadow rays of DistantLight l (DIR mode).
// NE (mesh object 0's block only): 256-ray blocks of the wave with at least one ray entering a mesh box.
enum { CNT_EXACT = 0, CNT_CAND = 1, CNT_NE = 2, CNT_BUNDLE0 = 3 };
NRT_HD int cntStride(int nL) { return CNT_BUNDLE0 + 6 * (1 + nL); }
NRT_HD int cntQueue(int b) { return CNT_BUNDLE0 + 6 * b; }      // rays queued for the bundle
NRT_HD int cntTile(int b) { return CNT_BUNDLE0 + 6 * b + 1; }   // prefilter work-item counter
NRT_HD int cntPre(int b) { return CNT_BUNDLE0 + 6 * b + 2; }    // pre-candidates (prefilter survivors)
NRT_HD int cntWork(int b) { return CNT_BUNDLE0 + 6 * b + 3; }   // (ray run, chunk) pairs admitted by the chunk bounds
NRT_HD int cntSub(int b) { return CNT_BUNDLE0 + 6 * b + 4; }    // (ray run, sub-chunk) pairs evaluated in full
NRT_HD int cntBnd(int b) { return CNT_BUNDLE0 + 6 * b + 5; }    // (ray run, chunk) pairs whose bound was tested ray by ray
// Stats slots
enum { ST_PRIMARY = 0, ST_TESTS = 1, ST_HITS = 2, ST_RAYS = 3, ST_CAPPED = 4, ST_CONT = 5, ST_COUNT = 8 };

struct FrameParams {
  int32_t width, height;
  int32_t aa_kind, grid, spp;
  int32_t step, max_step;
  int32_t depth_mode, max_ray_depth, bounce_cap;
  int32_t nx;              // x positions per row = ceil(width / step)
  // Pixel order of a worker's samples.  tshift == 0: scanline by scanline (cs.rows = the rows).  tshift > 0
  // (whole-resolution passes): cs.rows are the first rows of BANDS of T = 2^tshift scanlines, and inside a band
  // the pixels are enumerated tile by tile (T x T pixels, left to right; row-major inside a tile) — so that the 256
  // consecutive queue entries one prefilter warp works on are a compact 2-D patch of the image (4 x 4 pixels x 16
  // samples) instead of a 16 x 1 strip: smaller patches admit fewer chunks and sub-chunks of the mesh.
  int32_t tshift;
  int32_t band_pix;        // pixel slots per band = ceil(width / T) * T * T (slots right of the image are dead)
  int32_t y_end;           // rows >= y_end are dead (a band clipped by the end of the requested range)
  double bias;
  uint64_t seed;
  double aspect;           // double(width) / double(height)   (renderer.nim:36), the same division done once on the host
  double inv_grid;         // 1.0 / double(grid)               (sampling.nim:6-7)
};

// ---- output stage: the reference's float32 -> integer sample conversions -------------
//   writePpm's outvalue   utils/framebuf.nim:74-78: clamp(v, 0, 1) -> linearToSRGB (utils/color.nim:17-22) ->
//                         Natural(round(c * maxval)), maxval = float32(2^bits - 1); 8-bit samples for bits <= 8,
//                         big-endian 16-bit above (framebuf.nim:67-71)
//   ImageRGBA.copyFrom    utils/image.nim:45-54: round(v * 0xff).uint8 per channel + a constant alpha
// Byte work must equal the reference's bit for bit.  Everything except the pow() of linearToSRGB is a handful of
// exactly rounded float32 operations, evaluated here as written.  For the pow branch the library builds, on the host
// with the SAME libm call the reference's C back-end makes (powf), the table of cut points
//   thr[k - 1] = the smallest float32 input in (0.0031308, 1] whose sample is >= k,   k = 1 .. maxval
// (the conversion is monotonic; nrt.cu verifies that around every cut point) and the device counts the cut points
// <= v with a binary search: exact by construction, no device pow.  NaN samples (Natural(NaN) is undefined in the
// reference) are defined as 0.
struct OutStage {
  unsigned char* out;     // null: no integer output
  const float* thr;       // cut points of the sRGB pow branch for `bits` (device), maxval entries
  int32_t bits, srgb;     // RGB samples: bits in 1..16
  int32_t rgba, alpha;    // rgba != 0: ImageRGBA bytes (r, g, b, alpha) instead
};
NRT_HD uint32_t outvalueQ(float v, int bits, int srgb, const float* thr) {
  if (v != v) return 0u;
  const float c = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
  const uint32_t maxv = (1u << bits) - 1u;
  const float maxval = float(maxv);
  if (!srgb) return uint32_t(roundf(c * maxval));
  if (c <= 0.0031308f) return uint32_t(roundf((12.92f * c) * maxval));
  uint32_t lo = 0, hi = maxv;          // number of cut points <= c
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (thr[mid] <= c) lo = mid + 1; else hi = mid;
  }
  return lo;
}
NRT_HD unsigned char rgbaByteQ(float v) {
  const float r = roundf(v * 255.0f);
  return (unsigned char)(r != r ? 0.0f : (r < 0.0f ? 0.0f : (r > 255.0f ? 255.0f : r)));
}
// pixel `pi` of an integer image
NRT_HD void storeQ(const OutStage& q, int64_t pi, float r, float g, float b) {
  if (q.rgba) {
    unsigned char* p = q.out + 4 * pi;
    p[0] = rgbaByteQ(r); p[1] = rgbaByteQ(g); p[2] = rgbaByteQ(b); p[3] = (unsigned char)q.alpha;
    return;
  }
  const uint32_t vr = outvalueQ(r, q.bits, q.srgb, q.thr), vg = outvalueQ(g, q.bits, q.srgb, q.thr), vb = outvalueQ(b, q.bits, q.srgb, q.thr);
  if (q.bits <= 8) {
    unsigned char* p = q.out + 3 * pi;
    p[0] = (unsigned char)vr; p[1] = (unsigned char)vg; p[2] = (unsigned char)vb;
  } else {
    unsigned char* p = q.out + 6 * pi;
    p[0] = (unsigned char)(vr >> 8); p[1] = (unsigned char)(vr & 0xFFu);
    p[2] = (unsigned char)(vg >> 8); p[3] = (unsigned char)(vg & 0xFFu);
    p[4] = (unsigned char)(vb >> 8); p[5] = (unsigned char)(vb & 0xFFu);
  }
}

// Device-resident state of one chunk of samples (component-major SoA).
struct ChunkState {
  int64_t S;         // sample capacity
  int64_t NR;        // wave-ray capacity = S * max(1, nlights)
  int32_t nMO;       // mesh objects in the scene
  int32_t nL;        // lights
  int64_t candCap;
  // per sample
  double* rayO;      // 4*S   (bounce >= 1 only: every primary ray starts at the camera origin)
  double* rayD;      // 4*S
  double* hitW;      // 4*S
  double* nrm;       // 4*S
  double* accum;     // 3*S   (first written, not accumulated, at bounce 0)
  double* weight;    // S     (bounce >= 1 only: 1 at bounce 0)
  int32_t* hitObj;   // S (-1 none)
  uint8_t* active;   // S
  // per wave ray and mesh object
  uint64_t* tBest;   // nMO*NR   (bit pattern of a float64)
  uint32_t* triBest; // nMO*NR
  // filter queues: per mesh object QCAP = NR + nL*S entries; bundle 0 at [0, NR), bundle 1+l at NR + l*S
  int64_t QCAP;
  uint32_t* qref;    // nMO*QCAP   wave-ray index
  float* qray0;      // nMO*QCAP*4 plane 0 (float4): (d | o', rr)
  float* qray1;      // nMO*NR*4   plane 1 (float4), bundle 0 only: (m, 0)
  float* qhot0;      // nMO*QCAP*4 prefilter plane 0 (float4): (x, y, q, 0) | (dh, q)
  float* qhot1;      // nMO*NR*4   prefilter plane 1 (float4), bundle 0 only: (2 p0, 0)
  // pre-candidates of the current bundle: (queue index, record position)
  int64_t preCap;
  uint32_t* preRay;  // preCap
  uint32_t* preRec;  // preCap
  uint32_t* xref;    // nMO*NR     exact (float64 brute force) queue
  // (run, chunk) pairs admitted by the chunk bounds of the current bundle (CUDA backend)
  int64_t pairCap;
  uint32_t* pairs;   // 2*pairCap
  float* runc;       // 4 floats per ray run of the current bundle: the run's bounding circle (2-D bundles; CUDA backend)
  uint8_t* occ;      // NR         shadow ray (sample, light) found an occluder (ShadowTrace -> Resolve)
  // ordered queue compaction (CUDA backend): gate pass 1 writes a code per (mesh object, wave
  // position) and per-block counts; a scan turns the counts into queue offsets; pass 2 writes the
  // rays, so queue order == wave order (scanline order): consecutive queue entries are neighbours.
  uint8_t* gflag;    // nMO*NR     by wave position: 0 none, 1 + b filter bundle b, 255 exact
  uint32_t* gcnt;    // nMO*(2+nL) rows x gvb: rays of each 256-ray block that go to the row's queue
  uint32_t* gseg;    // rows x gsn: the same summed per segment of 256 blocks (zero between gates)
  uint32_t* gsegBase;// rows x gsn: exclusive scan of gseg = queue offset of the segment
  uint32_t* gne;     // gvb: blocks with at least one entering ray (unordered)
  int64_t gvb;       // blocks per row (capacity)
  int64_t gsn;       // segments per row (capacity)
  // candidates of the current (wave, mesh object)
  uint32_t* candRef; // candCap
  uint32_t* candTri; // candCap
  double* candT;     // candCap
  uint32_t* counters;  // maxWaves*nMO*cntStride(nL)
  uint32_t* alist;     // 2*S: compacted sample indices of the continuing paths (ping-pong per bounce)
  uint32_t* hlist;     // S: the samples of the current bounce that go through the wavefront (fused path)
  uint32_t* flist;     // S: the samples handed to the helper pipeline at bounce 0 (the fork)
  uint32_t* acount;    // per bounce: entries of the list consumed by that bounce
  unsigned long long* stats;  // ST_COUNT
  // pixel list of this worker
  const int32_t* rows; // device array of row indices
  int64_t p0;          // first pixel (in the worker's pixel list) of this chunk
  int64_t npix;        // pixels in this chunk
  // outputs
  OutStage q;          // q.out != null: Finalize stores writePpm / ImageRGBA samples there instead of floats into fb
  float* fb;           // width*height*3 (may be peer memory)
  int32_t* aovObj;     // may be null
  int32_t* aovTri;
  double* aovT;
};

// The samples a wave works on: all `n` samples of the chunk (bounce 0), or the compacted list of
// the samples whose path continued (written by Resolve of the previous bounce, count on the device).
struct ActiveSet { const uint32_t* list; const uint32_t* count; int64_t n; };
NRT_HD int64_t activeN(const ActiveSet& a) { return a.list ? int64_t(*a.count) : a.n; }
NRT_HD int64_t sampleOf(const ActiveSet& a, int64_t i) { return a.list ? int64_t(a.list[i]) : i; }

NRT_HD int64_t queueBase(const ChunkState& cs, int mo, int b) {
  return int64_t(mo) * cs.QCAP + (b == 0 ? 0 : cs.NR + int64_t(b - 1) * cs.S);
}

NRT_HD V4 ld4(const double* a, int64_t n, int64_t i) { return v4(a[i], a[n + i], a[2 * n + i], a[3 * n + i]); }
// L2 prefetch of the four planes of SoA vector i (see k_for_each_stats: the streaming kernels request the
// inputs of the element `ahead` positions later, which a CTA scheduled about one wave later will read)
NRT_HD void pf4(const double* a, int64_t n, int64_t i) {
  NRT_PREFETCH_L2(a + i); NRT_PREFETCH_L2(a + n + i); NRT_PREFETCH_L2(a + 2 * n + i); NRT_PREFETCH_L2(a + 3 * n + i);
}
NRT_HD void st4(double* a, int64_t n, int64_t i, V4 v) { a[i] = v.x; a[n + i] = v.y; a[2 * n + i] = v.z; a[3 * n + i] = v.w; }

// 64-bit integer division is a long instruction sequence on the GPU; every index on this path fits 32 bits
NRT_HD int64_t divFast(int64_t a, int32_t b) {
  if ((b & (b - 1)) == 0) {           // power of two (a >= 0 on every call site)
#if defined(__CUDA_ARCH__)
    return a >> (__ffs(b) - 1);
#else
    return a >> __builtin_ctz(unsigned(b));
#endif
  }
  return ((uint64_t(a) >> 32) == 0) ? int64_t(uint32_t(a) / uint32_t(b)) : a / b;
}
NRT_HD void pixelOf(const FrameParams& fp, const ChunkState& cs, int64_t p, int& x, int& y) {
  if (fp.tshift > 0) {
    const int64_t u = divFast(p, fp.band_pix);
    const int32_t r = int32_t(p - u * fp.band_pix);
    const int tt = 2 * fp.tshift, q = r & ((1 << tt) - 1);
    x = ((r >> tt) << fp.tshift) + (q & ((1 << fp.tshift) - 1));
    y = cs.rows[u] + (q >> fp.tshift);
    return;
  }
  const int64_t ri = divFast(p, fp.nx);
  x = int(p - ri * fp.nx) * fp.step;
  y = cs.rows[ri];
}
NRT_HD bool pixelSkipped(const FrameParams& fp, int x, int y) {  // renderer.nim:175-178
  if (x >= fp.width || y >= fp.y_end) return true;   // dead slot of a tile that overlaps the edge of the image / of the range
  if (fp.step < fp.max_step) {
    const int mask = fp.step * 2 - 1;
    if (((x & mask) == 0) && ((y & mask) == 0)) return true;
  }
  return false;
}

// sampling.nim:5-113 for one pixel; px/py have m*m entries.  Max grid 16 for jittered kinds.
static constexpr int kMaxJitterGrid = 16;
NRT_HD void makeSamples(int kind, int m, PixelRng& rng, double* px, double* py) {
  const int n = m;
  if (kind == AA_GRID) {
    const double xs = 1.0 / double(n), ys = 1.0 / double(m);
    const double xoffs = xs * 0.5, yoffs = xs * 0.5;
    for (int j = 0; j < m; ++j)
      for (int i = 0; i < n; ++i) { px[j * m + i] = double(i) * xs + xoffs; py[j * m + i] = double(j) * ys + yoffs; }
  } else if (kind == AA_JITTERED) {
    const double xs = 1.0 / double(n), ys = 1.0 / double(m);
    for (int j = 0; j < m; ++j)
      for (int i = 0; i < n; ++i) {
        const double rx = rng.random(xs), ry = rng.random(ys);
        px[j * m + i] = double(i) * xs + rx;
        py[j * m + i] = double(j) * ys + ry;
      }
  } else {
    const double xs = 1.0 / double(n), ys = 1.0 / double(m);
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < m; ++i) {
        const double jj = j, ii = i;
        const double r1 = rng.random(1.0), r2 = rng.random(1.0);
        px[j * m + i] = (ii + (jj + r1) * xs) * ys;
        py[j * m + i] = (jj + (ii + r2) * ys) * xs;
      }
    if (kind == AA_MULTI_JITTERED) {
      for (int j = 0; j < n; ++j)
        for (int i = 0; i < m; ++i) {
          const int k = j + int(rng.random(1.0) * double(n - j));
          const double t = px[j * m + i]; px[j * m + i] = px[k * m + i]; px[k * m + i] = t;
        }
      for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) {
          const int k = i + int(rng.random(1.0) * double(m - i));
          const double t = py[j * m + i]; py[j * m + i] = py[j * m + k]; py[j * m + k] = t;
        }
    } else {
      for (int j = 0; j < n; ++j) {
        const int k = j + int(rng.random(1.0) * double(n - j));
        for (int i = 0; i < m; ++i) { const double t = px[j * m + i]; px[j * m + i] = px[k * m + i]; px[k * m + i] = t; }
      }
      for (int i = 0; i < m; ++i) {
        const int k = i + int(rng.random(1.0) * double(m - i));
        for (int j = 0; j < n; ++j) { const double t = py[j * m + i]; py[j * m + i] = py[j * m + k]; py[j * m + k] = t; }
      }
    }
  }
}

// Bounce-0 state that every sample shares is not stored: the ray origin is the camera origin
// (primaryOrigin), the path weight is 1, the colour accumulator starts at 0 and the bounce number
// is the host loop's — Gen writes 33 bytes per sample instead of 105.
NRT_HD V4 primaryOrigin(const DScene& sc) { return v4(sc.cam_orig[0], sc.cam_orig[1], sc.cam_orig[2], sc.cam_orig[3]); }   // renderer.nim:42
NRT_HD void initSample(const ChunkState& cs, int64_t s, bool alive, V4 d) {
  st4(cs.rayD, cs.S, s, d);
  cs.active[s] = alive ? 1 : 0;
}

// How a kernel functor carries the scene header.  By value (the default): the header lives in kernel-parameter space
// (constant bank), so the camera, the counts, the table pointers and — DScene::hotOk — the lights, grid headers and
// gate records are uniform constant loads / operands instead of chains of dependent global loads through a pointer
// (FusedBounce 9.44 -> 8.72 ms on BASELINE config 4).  `sc->x`, `*sc` and passing `sc` as a pointer read the same.
#ifndef NRT_SCENE_BYVAL
#define NRT_SCENE_BYVAL 1
#endif
#if NRT_SCENE_BYVAL
struct SceneArg {
  DScene v;
  NRT_HD const DScene* operator->() const { return &v; }
  NRT_HD const DScene& operator*() const { return v; }
  NRT_HD operator const DScene*() const { return &v; }
};
#define NRT_SCENE_ARG(sd) (sd).h
#else
struct SceneArg {
  const DScene* p;
  NRT_HD const DScene* operator->() const { return p; }
  NRT_HD const DScene& operator*() const { return *p; }
  NRT_HD operator const DScene*() const { return p; }
};
#define NRT_SCENE_ARG(sd) (sd).d
#endif

// ---- gen: one element per SAMPLE (akNone / akGrid) ---------------------------
struct GenOut { bool alive; V4 o, d; double cx, cy; };
// sample position of grid sample (i, j) inside its pixel (sampling.nim:5-18) / the pixel corner for akNone (renderer.nim:135)
NRT_HD double sampleX(const FrameParams& fp, int x, int i) {
  if (fp.aa_kind == AA_NONE) return double(x);
  const double xs = fp.inv_grid, xoffs = xs * 0.5;
  return double(x) + (double(i) * xs + xoffs);
}
NRT_HD double sampleY(const FrameParams& fp, int y, int j) {
  if (fp.aa_kind == AA_NONE) return double(y);
  const double xs = fp.inv_grid, ys = fp.inv_grid, yoffs = xs * 0.5;
  return double(y) + (double(j) * ys + yoffs);
}
struct GenSimple {
  const DScene* sc; FrameParams fp; ChunkState cs;
  NRT_HD void operator()(int64_t s) const { GenOut out; run(s, out); }
  NRT_HD void run(int64_t s, GenOut& out) const {
    compute(s, out);
    initSample(cs, s, out.alive, out.d);
  }
  NRT_HD void compute(int64_t s, GenOut& out) const {   // the ray of sample s, nothing stored
    const int64_t sp = divFast(s, fp.spp);
    const int64_t p = cs.p0 + sp;
    const int k = int(s - sp * fp.spp);
    int x, y; pixelOf(fp, cs, p, x, y);
    const bool alive = !pixelSkipped(fp, x, y);
    int i = 0, j = 0;
    if (fp.aa_kind == AA_GRID) { j = int(divFast(k, fp.grid)); i = k - j * fp.grid; }   // sampling.nim:5-18, element j*m+i
    V4 o, d;
    // akNone: (x.float, y.float) — the pixel corner (renderer.nim:135); else x.float + sample
    // (per-column / per-row tables of cx and cy — two loads instead of two float64 divisions per sample — were
    // measured and dropped: FusedBounce 9.42 ms with them, 9.41 ms without; the kernel waits on latency, not on issue)
    const double cx = primaryCx(*sc, fp.aspect, fp.width, sampleX(fp, x, i));
    const double cy = primaryCy(*sc, fp.height, sampleY(fp, y, j));
    castPrimaryRayC(*sc, cx, cy, o, d);
    out.alive = alive; out.o = o; out.d = d; out.cx = cx; out.cy = cy;
  }
};

// ---- gen: one element per PIXEL (jittered kinds need the whole pattern) ------
struct GenJittered {
  const DScene* sc; FrameParams fp; ChunkState cs;
  NRT_HD void operator()(int64_t pl) const {
    const int64_t p = cs.p0 + pl;
    int x, y; pixelOf(fp, cs, p, x, y);
    const bool alive = !pixelSkipped(fp, x, y);
    double px[kMaxJitterGrid * kMaxJitterGrid], py[kMaxJitterGrid * kMaxJitterGrid];
    PixelRng rng = pixelRng(fp.seed, fp.width, x, y);
    makeSamples(fp.aa_kind, fp.grid, rng, px, py);
    for (int k = 0; k < fp.spp; ++k) {
      V4 o, d;
      castPrimaryRay(*sc, fp.aspect, fp.width, fp.height, double(x) + px[k], double(y) + py[k], o, d);
      initSample(cs, pl * fp.spp + k, alive, d);
    }
  }
};

// World-space ray `i` of a wave.  PATH: the sample's current ray.  SHADOW: ray
// (sample = i / nL, light = i % nL) rebuilt from the hit record (renderer.nim:93-99).
NRT_HD bool waveRay(const DScene& sc, const FrameParams& fp, const ChunkState& cs, int kind, int bounce, int64_t i, V4& o, V4& d) {
  if (kind == WAVE_PATH) {
    if (!cs.active[i]) return false;
    o = (bounce == 0) ? primaryOrigin(sc) : ld4(cs.rayO, cs.S, i);
    d = ld4(cs.rayD, cs.S, i);
    return true;
  }
  const int64_t s = divFast(i, cs.nL);
  const int l = int(i - s * cs.nL);
  if (cs.hitObj[s] < 0) return false;
  const V4 hitW = ld4(cs.hitW, cs.S, s), n = ld4(cs.nrm, cs.S, s);
  const ShadingInfo si = getShadingInfo(lightOf(sc, l), hitW);
  o = add(hitW, scale(n, fp.bias));
  d = scale(si.lightDir, -1.0);
  return true;
}

// Conservative float32 pre-test of the AABB gate (geom.nim:340): true only if the ray (t >= 0)
// certainly stays outside the bounding sphere of the mesh box — its line misses the sphere, or its
// origin is outside and it moves away from the centre.  The reference's slab test (geom.nim:76-96)
// then returns NegInf or a negative tmin, both "no hit", so the float64 evaluation (three
// divisions in initRay) is skipped.  Slack: 1e-5 |w|^2 |d|^2 >> the float32 cancellation error
// 12u |w|^2 |d|^2 of |w x d|^2; the sphere is inflated by 0.1 %.  NaN / overflow compare false.
NRT_HD bool boxCertainMiss(const DMesh& m, V4 oo, V4 dd) {
  const float wx = float(oo.x - m.center[0]), wy = float(oo.y - m.center[1]), wz = float(oo.z - m.center[2]);
  const float dx = float(dd.x), dy = float(dd.y), dz = float(dd.z);
  const float w2 = fmaf(wx, wx, fmaf(wy, wy, wz * wz)), d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
  if (!(w2 < 1e30f) || !(d2 < 1e30f) || !(d2 > 1e-30f) || !(m.rb2f < 1e30f)) return false;
  const float cx = wy * dz - wz * dy, cy = wz * dx - wx * dz, cz = wx * dy - wy * dx;
  const float c2 = fmaf(cx, cx, fmaf(cy, cy, cz * cz));
  const float slack = 1e-5f * (w2 * d2);
  if (c2 > fmaf(m.rb2f, d2, slack)) return true;
  const float wd = fmaf(wx, dx, fmaf(wy, dy, wz * dz));
  return (w2 > m.rb2f * 1.001f) && (wd > 0.f) && (wd * wd > slack);
}

NRT_HD Ray objectRay(const DObject& ob, V4 o, V4 d) {  // renderer.nim:54-55
  V4 oo, dd;
  toObject(ob, o, d, oo, dd);
  return initRay(oo, dd);
}

// ---- gate: TriangleMesh.intersect's AABB test (geom.nim:340) per (ray, mesh object)
struct GateOut { bool pass, safe; int bundle; uint32_t wi; FilterRay fr; HotRay hr; };
struct Gate {
  SceneArg sc; FrameParams fp; ChunkState cs; int kind; ActiveSet act; int force_exact;
  int path_mode;   // FM_ORIGIN for the primary wave (all rays share the camera origin), else FM_GENERAL
  int bounce;      // host loop's bounce number of the wave
  // wave-ray index (the slot of the ray's mesh results) of wave position idx < wave size
  NRT_HD uint32_t waveIndex(int64_t idx) const {
    if (kind == WAVE_SHADOW) {
      const int64_t si = divFast(idx, cs.nL);
      return uint32_t(sampleOf(act, si) * cs.nL + (idx - si * cs.nL));
    }
    return uint32_t(sampleOf(act, idx));
  }
  // i-th ray of the wave: PATH = i-th active sample; SHADOW = (active sample i / nL, light i % nL)
  NRT_HD GateOut operator()(int64_t idx, int mo) const {
    GateOut g; g.pass = false; g.safe = false; g.bundle = 0; g.wi = kInvalidRef;
    V4 o, d;
    const int64_t nS = activeN(act);
    int64_t i;
    int lsh = 0;
    if (kind == WAVE_SHADOW) {
      if (idx >= nS * cs.nL) return g;
      const int64_t si = divFast(idx, cs.nL);
      lsh = int(idx - si * cs.nL);
      i = sampleOf(act, si) * cs.nL + lsh;
    } else {
      if (idx >= nS) return g;
      i = sampleOf(act, idx);
    }
    const bool valid = waveRay(*sc, fp, cs, kind, bounce, i, o, d);
    if (!valid) return g;
    return evalRay(o, d, uint32_t(i), mo, lsh);
  }
  // the gate of one world-space ray (o, d) with wave-ray index wi (lsh = light of a shadow ray)
  NRT_HD GateOut evalRay(V4 o, V4 d, uint32_t wi, int mo, int lsh) const {
    GateOut g; g.pass = false; g.safe = false; g.bundle = 0;
    g.wi = wi;
    const DObject& ob = sc->objects[sc->mesh_obj_index[mo]];
    const DMesh& m = sc->meshes[ob.mesh];
    V4 oo, dd;
    toObject(ob, o, d, oo, dd);                      // renderer.nim:54-55
    if (boxCertainMiss(m, oo, dd)) return g;         // float32: the ray stays clear of the box's bounding sphere
    const Ray r = initRay(oo, dd);
    const double tmin = aabbIntersect(m.bmin, m.bmax, r);
    g.pass = !(tmin < 0);
    if (g.pass && !force_exact) {
      int mode = path_mode, l = 0;
      if (kind == WAVE_SHADOW) {
        l = lsh;
        mode = (lightOf(*sc, l).kind == LIGHT_DISTANT) ? FM_DIR : FM_GENERAL;
      }
      if (mode != FM_GENERAL && !(sc->frames[frameIndex(sc->nlights, mo, mode, l)].valid > 0)) mode = FM_GENERAL;
      g.bundle = (mode == FM_DIR) ? 1 + l : 0;
      const BundleFrame& fr = sc->frames[frameIndex(sc->nlights, mo, mode, l)];
      g.safe = makeFilterRay(mode, m, r, g.fr) && makeHotRay(mode, fr, r, g.hr);
      if (!g.safe) g.bundle = 0;
    }
    return g;
  }
};
NRT_HD uint8_t gateCode(const GateOut& o) { return o.pass ? (o.safe ? uint8_t(1 + o.bundle) : uint8_t(255)) : uint8_t(0); }

// ---- exact: float64 brute force over ALL faces for rays the filter cannot take
struct ExactMesh {
  SceneArg sc; FrameParams fp; ChunkState cs; int kind; int mo; const uint32_t* count; int bounce;
  NRT_HD void operator()(int64_t qi) const {
    const uint32_t ref = cs.xref[int64_t(mo) * cs.NR + qi];
    V4 o, d;
    if (!waveRay(*sc, fp, cs, kind, bounce, ref, o, d)) return;
    const DObject& ob = sc->objects[sc->mesh_obj_index[mo]];
    const DMesh& m = sc->meshes[ob.mesh];
    const Ray r = objectRay(ob, o, d);
    double tMin = NRT_INF; uint32_t tri = kNoTri;
    for (int64_t f = 0; f < m.nfaces; ++f) {  // geom.nim:346-356
      const double t = rayTriangleExact(r, m.verts + 4 * m.vidx[3 * f], m.verts + 4 * m.vidx[3 * f + 1],
                                        m.verts + 4 * m.vidx[3 * f + 2]);
      if (t >= 0 && t < tMin) { tMin = t; tri = uint32_t(f); }
    }
    cs.tBest[int64_t(mo) * cs.NR + ref] = dbits(tMin == 0 ? 0.0 : tMin);
    cs.triBest[int64_t(mo) * cs.NR + ref] = tri;
  }
};

// ---- refine: float32 sign test (nrt_filter.h: filterTest) on the prefilter survivors ----
template <class A>
struct Refine {
  ChunkState cs; int mode; const float* recs; const uint32_t* ids; int mo; int b; uint32_t* candCount;
  NRT_HD void operator()(int64_t i) const {
    const uint32_t rq = cs.preRay[i], pos = cs.preRec[i];
    const int nc = recFloats(mode);
    float q[16];
#if defined(__CUDA_ARCH__)
    {
      const float4* p4 = reinterpret_cast<const float4*>(recs + fullIndex(pos, 0));
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        if (4 * k4 < nc) { const float4 v = __ldg(p4 + k4); q[4 * k4] = v.x; q[4 * k4 + 1] = v.y; q[4 * k4 + 2] = v.z; q[4 * k4 + 3] = v.w; }
      }
    }
#else
    for (int k = 0; k < nc; ++k) q[k] = recs[fullIndex(pos, k)];
#endif
    const int64_t at = queueBase(cs, mo, b) + rq;
    const float* p0 = cs.qray0 + 4 * at;
    const float* p1 = cs.qray1 + 4 * (int64_t(mo) * cs.NR + rq);   // read only in GENERAL mode (b == 0)
    const uint32_t x = filterTest(mode, q, p0, p1, p0[3]);
    if (int32_t(x) < 0) return;
    uint32_t tri = ids ? ids[pos] : pos;   // GENERAL: record position -> face through the Morton order
    if (recSlotId(mode) >= 0) tri = fbits(q[recSlotId(mode)]);
    const uint32_t slot = A::add32(candCount, 1u);
    if (slot < cs.candCap) { cs.candRef[slot] = cs.qref[at]; cs.candTri[slot] = tri; }
  }
};

// atomics supplied by the backend
template <class A>
struct Verify1 {  // float64 re-evaluation of candidate c; running minimum of t per ray
  SceneArg sc; FrameParams fp; ChunkState cs; int kind; int mo; int bounce;
  NRT_HD void operator()(int64_t c) const {
    const uint32_t ref = cs.candRef[c], tri = cs.candTri[c];
    V4 o, d;
    double t = NRT_NEG_INF;
    if (waveRay(*sc, fp, cs, kind, bounce, ref, o, d)) {
      const DObject& ob = sc->objects[sc->mesh_obj_index[mo]];
      const DMesh& m = sc->meshes[ob.mesh];
      const Ray r = objectRay(ob, o, d);
      t = rayTriangleExact(r, m.verts + 4 * m.vidx[3 * int64_t(tri)], m.verts + 4 * m.vidx[3 * int64_t(tri) + 1],
                           m.verts + 4 * m.vidx[3 * int64_t(tri) + 2]);
    }
    if (t >= 0) {             // geom.nim:354 `tHit >= 0` (NaN and NegInf fail)
      if (t == 0) t = 0.0;    // -0.0 -> +0.0 so that bit patterns order like values
      A::min64(&cs.tBest[int64_t(mo) * cs.NR + ref], dbits(t));
    }
    cs.candT[c] = t;
  }
};
template <class A>
struct Verify2 {  // among candidates attaining the minimum, the lowest face index wins (geom.nim:354 strict <)
  ChunkState cs; int mo;
  NRT_HD void operator()(int64_t c) const {
    const double t = cs.candT[c];
    if (!(t >= 0)) return;
    const uint32_t ref = cs.candRef[c];
    if (dbits(t) == cs.tBest[int64_t(mo) * cs.NR + ref]) A::min32(&cs.triBest[int64_t(mo) * cs.NR + ref], cs.candTri[c]);
  }
};

// ---- TriangleMesh.intersect for ONE ray by ONE thread (the path kernels below) -------------------
// The AABB gate (geom.nim:340), then a walk of the same flattened hierarchy the wavefront prefilter uses
// (chunk bounds -> sub-chunk bounds -> the face's bounding circle / sphere -> float32 sign test), and the
// reference's float64 evaluation of the survivors.  Every float32 stage is conservative, so the result —
// nearest accepted t, lowest face index on ties (geom.nim:346-356) — is the reference's whatever the
// bundle `mode` used for the walk.  Rays the filter cannot take go through all faces in float64.
struct MeshHit { double t; uint32_t tri; };
NRT_HD bool meshGatePass(const DScene& sc, int mo, V4 o, V4 d) {
  const DObject& ob = sc.objects[sc.mesh_obj_index[mo]];
  const DMesh& m = sc.meshes[ob.mesh];
  V4 oo, dd;
  toObject(ob, o, d, oo, dd);                      // renderer.nim:54-55
  if (boxCertainMiss(m, oo, dd)) return false;
  const double tmin = aabbIntersect(m.bmin, m.bmax, initRay(oo, dd));
  return !(tmin < 0);
}
// What a walk of mesh object `mo` needs for one ray: the gate's verdict, the object-space ray (orig / dir are
// all rayTriangleExact reads) and, when the float32 filter can take the ray (`safe`), its two filter forms.
struct WalkRay {
  double ox, oy, oz, dx, dy, dz;   // object-space ray
  float f[8];                      // FilterRay: (a.xyz, rr | m.xyz, 0)
  float h[8];                      // HotRay:    (a0..a3 | b0..b2, 0)
};
NRT_HD Ray walkRayAsRay(const WalkRay& w) {
  Ray r; r.orig = v4(w.ox, w.oy, w.oz, 1.0); r.dir = v4(w.dx, w.dy, w.dz, 0.0);
  r.ix = r.iy = r.iz = 0.0; r.sx = r.sy = r.sz = 0;   // (invDir / sign are only read by the AABB test)
  return r;
}
NRT_HD HotRay walkRayHot(const WalkRay& w) {
  HotRay h; h.a0 = w.h[0]; h.a1 = w.h[1]; h.a2 = w.h[2]; h.a3 = w.h[3]; h.b0 = w.h[4]; h.b1 = w.h[5]; h.b2 = w.h[6]; h.b3 = 0.f;
  return h;
}
// the bundle a ray of hint `mode` is walked in (a bundle without a usable frame falls back to GENERAL)
NRT_HD int walkMode(const DScene& sc, int mo, int mode, int l) {
  if (mode != FM_GENERAL && !(sc.frames[frameIndex(sc.nlights, mo, mode, l)].valid > 0)) return FM_GENERAL;
  return mode;
}
// returns pass (the ray enters the box, geom.nim:340); `safe`: the filter walk applies, else all faces in float64
NRT_HD bool walkPrep(const DScene& sc, int mo, int mode, int l, V4 o, V4 d, int force_exact, WalkRay& w, bool& safe) {
  safe = false;
  const DObject& ob = sc.objects[sc.mesh_obj_index[mo]];
  const DMesh& m = sc.meshes[ob.mesh];
  V4 oo, dd;
  toObject(ob, o, d, oo, dd);                      // renderer.nim:54-55
  if (boxCertainMiss(m, oo, dd)) return false;
  const Ray r = initRay(oo, dd);
  if (aabbIntersect(m.bmin, m.bmax, r) < 0) return false;   // geom.nim:340 (NegInf: no box hit)
  w.ox = oo.x; w.oy = oo.y; w.oz = oo.z; w.dx = dd.x; w.dy = dd.y; w.dz = dd.z;
  const int fi = frameIndex(sc.nlights, mo, mode, l);
  FilterRay fr; HotRay hr;
  safe = !force_exact && sc.recsets[fi].usable && makeFilterRay(mode, m, r, fr) && makeHotRay(mode, sc.frames[fi], r, hr);
  if (safe) {
    w.f[0] = fr.ax; w.f[1] = fr.ay; w.f[2] = fr.az; w.f[3] = fr.rr; w.f[4] = fr.mx; w.f[5] = fr.my; w.f[6] = fr.mz; w.f[7] = 0.f;
    w.h[0] = hr.a0; w.h[1] = hr.a1; w.h[2] = hr.a2; w.h[3] = hr.a3; w.h[4] = hr.b0; w.h[5] = hr.b1; w.h[6] = hr.b2; w.h[7] = 0.f;
  }
  return true;
}
// record `rec` of set `rs` against the ray: bounding circle / sphere, float32 sign test, float64 (geom.nim:283-336);
// keeps the nearest accepted t, the lowest face index on ties
NRT_HD void walkRecord(const DMesh& m, const RecSet& rs, int mode, int64_t rec, const WalkRay& w, const Ray& r, const HotRay& hr, double& best, uint32_t& bt) {
  const int nh = hotFloats(mode), nc = recFloats(mode);
  float hh[4] = {0.f, 0.f, 0.f, 0.f};
  for (int j = 0; j < nh; ++j) hh[j] = rs.hot[recIndex(rec, j, nh)];
  if (!prefilterTest(mode, hh, hr)) return;
  float q[16];
  for (int j = 0; j < nc; ++j) q[j] = rs.recs[fullIndex(rec, j)];
  if (int32_t(filterTest(mode, q, w.f, w.f + 4, w.f[3])) < 0) return;
  uint32_t tri = rs.ids ? rs.ids[rec] : uint32_t(rec);
  if (recSlotId(mode) >= 0) tri = fbits(q[recSlotId(mode)]);
  const double t = rayTriangleExact(r, m.verts + 4 * m.vidx[3 * int64_t(tri)], m.verts + 4 * m.vidx[3 * int64_t(tri) + 1],
                                    m.verts + 4 * m.vidx[3 * int64_t(tri) + 2]);
  if (t >= 0 && (t < best || (t == best && tri < bt))) { best = t; bt = tri; }
}
NRT_HD MeshHit walkScalar(const DScene& sc, int mo, int mode, int l, const WalkRay& w, bool safe) {
  const DMesh& m = sc.meshes[sc.objects[sc.mesh_obj_index[mo]].mesh];
  const Ray r = walkRayAsRay(w);
  double best = NRT_INF; uint32_t bt = kNoTri;     // geom.nim:343
  if (!safe) {
    for (int64_t f = 0; f < m.nfaces; ++f) {       // geom.nim:346-356
      const double t = rayTriangleExact(r, m.verts + 4 * m.vidx[3 * f], m.verts + 4 * m.vidx[3 * f + 1], m.verts + 4 * m.vidx[3 * f + 2]);
      if (t >= 0 && t < best) { best = t; bt = uint32_t(f); }
    }
  } else {
    const RecSet rs = sc.recsets[frameIndex(sc.nlights, mo, mode, l)];
    const HotRay hr = walkRayHot(w);
    const int64_t nch = paddedFaces(int64_t(rs.nrec)) / kRecPad;
    for (int64_t ch = 0; ch < nch; ++ch) {
      if (!prefilterTest(mode, rs.bounds + 4 * ch, hr)) continue;
      for (int sb = 0; sb < kSubPerChunk; ++sb) {
        const int64_t sub = ch * kSubPerChunk + sb;
        if (!prefilterTest(mode, rs.sub + 4 * sub, hr)) continue;
        for (int k = 0; k < kSubRecs; ++k) walkRecord(m, rs, mode, sub * kSubRecs + k, w, r, hr, best, bt);
      }
    }
  }
  MeshHit h; h.t = (best == 0) ? 0.0 : best; h.tri = bt;   // -0.0 -> +0.0 (as Verify1 / ExactMesh)
  return h;
}
NRT_HD MeshHit meshIntersectWalk(const DScene& sc, int mo, int mode, int l, V4 o, V4 d, int force_exact) {
  MeshHit h; h.t = NRT_NEG_INF; h.tri = kNoTri;
  mode = walkMode(sc, mo, mode, l);
  WalkRay w; bool safe;
  if (!walkPrep(sc, mo, mode, l, o, d, force_exact, w, safe)) return h;
  return walkScalar(sc, mo, mode, l, w, safe);
}

// How trace() obtains TriangleMesh.intersect of the current ray for mesh object `mo`:
//   WaveMesh  the wavefront: gate code of the ray's wave position + the results of the mesh wave
//   NoMesh    the caller has shown that the ray enters no mesh box (FusedBounce): NegInf
//   WalkMesh  the thread walks the mesh itself (PathTail)
struct WaveMesh {
  const ChunkState& cs; int64_t wi, pos; uint8_t code0;   // code0 = gate code for mesh object 0, loaded by the caller with its other inputs
  NRT_HD bool miss(uint32_t mo) const { return (mo == 0 ? code0 : cs.gflag[int64_t(mo) * cs.NR + pos]) == 0; }
  NRT_HD void eval(int mo, V4, V4, double& t, uint32_t& tri) const {
    t = NRT_NEG_INF; tri = kNoTri;
    if (cs.gflag[int64_t(mo) * cs.NR + pos]) {
      t = bitsd(cs.tBest[int64_t(mo) * cs.NR + wi]);
      tri = cs.triBest[int64_t(mo) * cs.NR + wi];
    }
  }
};
struct NoMesh {
  NRT_HD bool miss(uint32_t) const { return true; }
  NRT_HD void eval(int, V4, V4, double& t, uint32_t& tri) const { t = NRT_NEG_INF; tri = kNoTri; }
};
struct WalkMesh {
  const DScene* sc; int mode, l, force_exact;
  NRT_HD bool miss(uint32_t) const { return false; }
  NRT_HD void eval(int mo, V4 o, V4 d, double& t, uint32_t& tri) const {
    const MeshHit h = meshIntersectWalk(*sc, mo, mode, l, o, d, force_exact);
    t = h.t; tri = h.tri;
  }
};

// The mesh results of the current ray were computed before trace() is entered (PathWarp: by the whole warp);
// mesh objects beyond the kMaxWalkMO slots are walked by the thread itself.
static constexpr int kMaxWalkMO = 2;
struct MeshRes { double t[kMaxWalkMO]; uint32_t tri[kMaxWalkMO]; };
struct PreMesh {
  const MeshRes& res; WalkMesh rest;
  NRT_HD bool miss(uint32_t mo) const { return mo < uint32_t(kMaxWalkMO) && !(res.t[mo] >= 0); }
  NRT_HD void eval(int mo, V4 o, V4 d, double& t, uint32_t& tri) const {
    if (mo < kMaxWalkMO) { t = res.t[mo]; tri = res.tri[mo]; }
    else rest.eval(mo, o, d, t, tri);
  }
};

// trace(): renderer.nim:47-67; the mesh results come from the policy `mp` (above).
struct TraceOut { int obj; double t; uint32_t tri; int tests, hits; };

// float64 evaluation of object i for trace(): the reference's intersect() (or the mesh result), then the
// running-minimum update of renderer.nim:60-65
template <class MP>
NRT_HD void evalObject(const DScene& sc, const MP& mp, int i, V4 o, V4 d, bool fastRay, TraceOut& r) {
  const CObj c = loadCObj(sc.cobjs + i);
  double t; uint32_t tri = kNoTri;
  if (c.kind == GEOM_MESH) {
    mp.eval(c.mesh_obj, o, d, t, tri);
  } else {
    V4 oo, dd;
    // A Plane reads only the y components (geom.nim:240-248 with n = (0,1,0,0)): with dir.y != 0 and a nonzero
    // height, denom = ((0*dx + 1*dy) + 0*dz) + 0*dw = dy and dot(orig, n) = oy exactly, whatever the signs of the
    // zero products — so zero x / z components (a camera at x = 0 over a ground plane at the origin) do not need
    // toObject()'s literal evaluation.
    const bool planeY = c.kind == GEOM_PLANE && (o.y + c.t[1]) != 0.0;
    if (c.xlate_only && fastRay && (planeY || ((o.x != 0.0 || c.t[0] != 0.0) && (o.y != 0.0 || c.t[1] != 0.0) && (o.z != 0.0 || c.t[2] != 0.0)))) {
      oo = v4(o.x + c.t[0], o.y + c.t[1], o.z + c.t[2], 1.0);
      dd = v4(d.x, d.y, d.z, 0.0);
    } else {
      toObject(sc.objects[i], o, d, oo, dd);
    }
    // initRay's 1/dir (geom.nim:42-47) is only read by the AABB test: built for boxes only
    // spheres: a tighter float32 discriminant bound on the float64 object-space ray first; planes: orig.y and
    // dir.y of equal sign give t < 0, which trace() rejects (renderer.nim:60) whatever its value
    if (c.kind == GEOM_SPHERE) t = sphereCertainMiss(c.radius, oo, dd) ? NRT_NEG_INF : sphereIntersect(c.radius, oo, dd);
    else if (c.kind == GEOM_PLANE) {
      // (bounded magnitudes: the quotient cannot underflow to -0.0, which `t >= 0` would accept)
      const bool neg = ((oo.y > 1e-150 && dd.y > 1e-6) || (oo.y < -1e-150 && dd.y < -1e-6)) && fabs(oo.y) < 1e150 && fabs(dd.y) < 1e150;
      t = neg ? NRT_NEG_INF : planeIntersect(oo, dd);
    }
    else if (c.kind == GEOM_BOX) t = aabbIntersect(sc.objects[i].bmin, sc.objects[i].bmax, initRay(oo, dd));
    else t = NRT_NEG_INF;
  }
  if (t >= 0 && t < r.t) { r.t = t; r.obj = i; r.tri = tri; r.hits++; }
}

// float32 first look at object i (record c): true = the reference's intersect() certainly returns a
// value trace() rejects (NegInf or t < 0), so the object is skipped (it is still counted in Stats)
template <class MP>
NRT_HD bool firstLookMiss(const MP& mp, const CObjF& c, const RayF& rf, bool f32ok) {
  if (c.r2m < 3.0e38f) return f32ok && certainMissF(c, rf);
  const uint32_t tag = fbits(c.tx);
  if (tag == COF_MESH) return mp.miss(fbits(c.ty));   // (wavefront: gate code 0 = the ray did not enter the mesh's box)
  return (tag == COF_PLANE) && f32ok && planeMissF(c, rf);
}

// `pos` = wave position of the ray (index of its gate code): a ray that did not enter a mesh's box
// (code 0) has t = NegInf for that mesh without touching the per-ray mesh results.
// `code0` = gate code of the ray for mesh object 0, loaded by the caller together with its other inputs.
// CL: the scene has sphere clusters (the host picks the kernel variant, so that the flat scan of small
// scenes keeps its register budget).
// Per-ray facts shared by every float32 first look at the ray (object scan, mesh gates)
// gridOk: `gridMask` lists (bit i <-> object i) every object of a mask-grid scene the ray can possibly hit — spheres the
// float32 test applies to and mesh boxes with a float32 gate record; the other objects are in DScene.slowMask
struct RayPre { RayF rf; bool f32ok, fastRay, gridOk; uint32_t gridMask; };
NRT_HD RayPre makeRayPre(V4 o, V4 d) {
  RayPre p;
  p.gridOk = false; p.gridMask = 0xFFFFFFFFu;
  // the exact shortcut of toObject() for [I | t] matrices applies to this ray?  (zero components
  // need toObject()'s per-component treatment)
  p.f32ok = (o.w == 1.0) && (d.w == 0.0) && finite3(o) && finite3(d);
  p.fastRay = p.f32ok && d.x != 0.0 && d.y != 0.0 && d.z != 0.0;
  p.rf = makeRayF(o, d);
  return p;
}
static constexpr int kPrimaryRay = -2;
// makeRayPre + the ray's cell of the scene's mask grid: `sl` >= 0: a shadow ray towards light sl (its direction is that
// light's), kPrimaryRay: a primary ray (its origin is the camera's), anything else: no grid applies
NRT_HD RayPre makeRayPreGrid(const DScene& sc, V4 o, V4 d, int sl) {
  RayPre p = makeRayPre(o, d);
  if (sc.maskGrids && sc.sgrid && p.f32ok && (sl >= 0 || sl == kPrimaryRay)) {
    const ShadowGridF g = sgridOf(sc, sl >= 0 ? sl : sc.nlights);
    uint32_t m = 0, e = 0;
    if (g.G > 0 && shadowGridCell(g, p.rf, m, e)) { p.gridOk = true; p.gridMask = m | sc.slowMask; }
  }
  return p;
}
// the same for a primary ray whose (cx, cy) of castPrimaryRay is known: that IS its point on the camera grid's plane
// (the direction is c2w * normalize(cx, cy, -1)), so the three dot products and two divisions of the lookup are skipped
NRT_HD RayPre makeRayPrePrimary(const DScene& sc, V4 o, V4 d, double cx, double cy) {
  RayPre p = makeRayPre(o, d);
  if (sc.maskGrids && sc.sgrid && p.f32ok) {
    const ShadowGridF g = sgridOf(sc, sc.nlights);
    if (g.G > 0 && g.persp) {
      const float f1 = (float(cx) - g.lo1) * g.invh, f2 = (float(cy) - g.lo2) * g.invh;
      if (f1 >= 0.f && f2 >= 0.f && f1 < float(g.G) && f2 < float(g.G)) {
        int c1 = int(f1), c2 = int(f2);
        if (c1 > g.G - 1) c1 = g.G - 1;
        if (c2 > g.G - 1) c2 = g.G - 1;
        p.gridOk = true;
        p.gridMask = g.items[uint32_t(c2) * uint32_t(g.G) + uint32_t(c1)] | sc.slowMask;
      }
    }
  }
  return p;
}
// the AABB gate of mesh object mo for a world-space ray: float32 first look, then the reference's evaluation
NRT_HD bool meshGatePassPre(const DScene& sc, int mo, V4 o, V4 d, const RayPre& pre) {
  if (sc.hotOk) {   // (the header's own copies: mo < kHotMO)
    if (pre.gridOk && !((pre.gridMask >> (uint32_t(sc.moIndexv[mo]) & 31u)) & 1u)) return false;
    const MeshGateF gh = sc.mgatev[mo];
    if (pre.f32ok && gh.valid > 0.f && meshGateMissF(gh, pre.rf)) return false;
    return meshGatePass(sc, mo, o, d);
  }
  if (pre.gridOk && !((pre.gridMask >> (uint32_t(sc.mesh_obj_index[mo]) & 31u)) & 1u)) return false;   // the ray's line misses the box's bounding sphere
#if defined(__CUDA_ARCH__)
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(sc.mgate + mo)), g1 = __ldg(reinterpret_cast<const float4*>(sc.mgate + mo) + 1);
  MeshGateF g; g.cx = g0.x; g.cy = g0.y; g.cz = g0.z; g.r2m = g0.w; g.mm = g1.x; g.valid = g1.y; g.pad0 = g.pad1 = 0.f;
#else
  const MeshGateF g = sc.mgate[mo];
#endif
  if (pre.f32ok && g.valid > 0.f && meshGateMissF(g, pre.rf)) return false;
  return meshGatePass(sc, mo, o, d);
}

// `sl` >= 0: the ray is a shadow ray towards light sl (its direction is that light's: the light-space grid applies);
// kPrimaryRay: a primary ray (its origin is the camera's: the camera grid applies)
template <bool CL, class MP>
NRT_HD TraceOut traceObjectsPre(const DScene& sc, const MP& mp, V4 o, V4 d, double tNear, const RayPre& pre, int sl = -1);
template <bool CL, class MP>
NRT_HD TraceOut traceObjects(const DScene& sc, const MP& mp, V4 o, V4 d, double tNear, int sl = -1) {
  return traceObjectsPre<CL>(sc, mp, o, d, tNear, makeRayPreGrid(sc, o, d, sl), sl);
}
template <bool CL, class MP>
NRT_HD TraceOut traceObjectsPre(const DScene& sc, const MP& mp, V4 o, V4 d, double tNear, const RayPre& pre, int sl) {
  TraceOut r; r.obj = -1; r.t = tNear; r.tri = kNoTri; r.tests = 0; r.hits = 0;
  const bool f32ok = pre.f32ok, fastRay = pre.fastRay;
  const RayF rf = pre.rf;
  if (CL && sc.ncl1 > 0 && (sl >= 0 || sl == kPrimaryRay) && sc.sgrid && f32ok) {
    // a DistantLight's shadow ray / a primary ray: the clustered spheres it can hit are listed in ONE cell of the grid
    const ShadowGridF g = sgridOf(sc, sl >= 0 ? sl : sc.nlights);
    uint32_t gb = 0, ge = 0;
    if (g.G > 0 && shadowGridCell(g, rf, gb, ge) && ge - gb + uint32_t(sc.nslow) <= uint32_t(kSurvivorCap)) {
      uint32_t surv[kSurvivorCap] = {0};
      int ns = 0;
      auto push = [&](uint32_t i) {   // (cell lists and slowIdx are ascending; merged in list order)
        int k = ns++;
        while (k > 0 && surv[k - 1] > i) { surv[k] = surv[k - 1]; --k; }
        surv[k] = i;
      };
      for (uint32_t k = gb; k < ge; ++k) {
        const uint32_t i = g.items[k];
        if (!certainMissF(loadCObjF(sc.cobjf + i), rf)) push(i);
      }
      for (int k = 0; k < sc.nslow; ++k) {
        const uint32_t i = sc.slowIdx[k];
        if (!firstLookMiss(mp, loadCObjF(sc.cobjf + i), rf, f32ok)) push(i);
      }
      r.tests = sc.nobjects;
      for (int k = 0; k < ns; ++k) evalObject(sc, mp, int(surv[k]), o, d, fastRay, r);
      return r;
    }
  }
#if defined(__CUDA_ARCH__)
  // The warp's rays as one bundle (nrt_core.h: RayBundle): whole objects and whole sphere clusters are dropped for
  // every lane after one test in one lane; the per-ray first look and the float64 evaluation see only the rest.
  // (clustered scenes only: a bundle costs ~100 instructions per trace — on BASELINE config 4's nine objects the flat
  // per-ray scan is cheaper: FusedBounce 11.0 ms without the bundle, 13.1 ms with it, measured)
  RayBundle B;
  B.on = false;
  if (CL && sc.ncl1 > 0) B = makeRayBundle(rf, f32ok);
  if (CL && sc.ncl1 > 0 && B.on) {
    // A bundle that is fat where the spheres are (shadow rays leaving a triangle soup: origins spread along the view
    // rays) admits almost every cluster and is worse than the per-ray traversal: the leader's own ray is tested next to
    // the bundle at level 1, and a bundle that admits more than twice what the one ray admits (+2) is given up.
    int nbu = 0, nld = 0;
    RayF lf; lf.ox = B.ox; lf.oy = B.oy; lf.oz = B.oz; lf.dx = B.ux; lf.dy = B.uy; lf.dz = B.uz; lf.a = B.a; lf.mray = B.mray;
    for (int base = 0; base < sc.ncl1; base += B.nact) {
      const int k = base + B.rank;
      bool cb = false, cl = false;
      if (k < sc.ncl1) { const CObjF c = loadCObjF(sc.cl1 + k); cb = !bundleMiss(c, B); cl = !certainMissF(c, lf); }
      nbu += __popc(__ballot_sync(B.wm, cb)); nld += __popc(__ballot_sync(B.wm, cl));
    }
    if (nbu > 2 * nld + 2) B.on = false;
  }
  if (CL && sc.ncl1 > 0 && B.on) {
    uint32_t surv[kSurvivorCap] = {0};
    int ns = 0;
    bool full = false;
    auto push = [&](uint32_t i) {
      if (i == kInvalidRef) return;   // padding slot of a cluster
      if (ns == kSurvivorCap) { full = true; return; }
      int k = ns++;
      while (k > 0 && surv[k - 1] > i) { surv[k] = surv[k - 1]; --k; }
      surv[k] = i;
    };
    bundleScan(B, sc.cl1, sc.ncl1, [&](int a1) {
      bundleScan(B, sc.cl2 + a1 * kClusterSize, kClusterSize, [&](int c2) {
        const int a2 = a1 * kClusterSize + c2;
        bundleScan(B, sc.clm + a2 * kClusterSize, kClusterSize, [&](int cm) {
          const int m = a2 * kClusterSize + cm;
          if (!certainMissF(loadCObjF(sc.clm + m), rf)) push(sc.clmIdx[m]);
        });
      });
    });
    for (int k = 0; k < sc.nslow; ++k) {
      const uint32_t i = sc.slowIdx[k];
      if (!firstLookMiss(mp, loadCObjF(sc.cobjf + i), rf, f32ok)) push(i);
    }
    // (a full list — more than kSurvivorCap objects along one ray — falls back to the flat scan below; a lane that
    // leaves here no longer takes part in the warp's later bundle steps, which only makes those bundles smaller)
    if (!full) {
      r.tests = sc.nobjects;
      for (int k = 0; k < ns; ++k) evalObject(sc, mp, int(surv[k]), o, d, fastRay, r);
      return r;
    }
  }
  if (!(CL && sc.ncl1 > 0 && B.on))
#endif
  if (CL && sc.ncl1 > 0 && f32ok) {
    // Scenes with many spheres: a flattened two-level traversal of the sphere clusters.  A ray that
    // certainly misses a cluster's bounding sphere certainly misses every member, so whole groups of
    // 256 and 16 spheres are skipped; the objects that are not certain misses are collected in a small
    // list kept in LIST ORDER (renderer.nim:53: the in-order scan decides ties and Stats) and evaluated
    // in float64 afterwards.  A full list falls back to the flat in-order scan below.
    uint32_t surv[kSurvivorCap] = {0};
    int ns = 0;
    bool full = false;
    auto push = [&](uint32_t i) {
      if (i == kInvalidRef) return;   // padding slot of a cluster
      if (ns == kSurvivorCap) { full = true; return; }
      int k = ns++;
      while (k > 0 && surv[k - 1] > i) { surv[k] = surv[k - 1]; --k; }
      surv[k] = i;
    };
    for (int a1 = 0; a1 < sc.ncl1 && !full; ++a1) {
      if (certainMissF(loadCObjF(sc.cl1 + a1), rf)) continue;
      for (int a2 = a1 * kClusterSize; a2 < (a1 + 1) * kClusterSize && !full; ++a2) {
        if (certainMissF(loadCObjF(sc.cl2 + a2), rf)) continue;
        for (int m = a2 * kClusterSize; m < (a2 + 1) * kClusterSize; ++m)
          if (!certainMissF(loadCObjF(sc.clm + m), rf)) push(sc.clmIdx[m]);
      }
    }
    for (int k = 0; k < sc.nslow && !full; ++k) {
      const uint32_t i = sc.slowIdx[k];
      if (!firstLookMiss(mp, loadCObjF(sc.cobjf + i), rf, f32ok)) push(i);
    }
    if (!full) {
      r.tests = sc.nobjects;
      for (int k = 0; k < ns; ++k) evalObject(sc, mp, int(surv[k]), o, d, fastRay, r);
      return r;
    }
  }
  // Two phases per batch of 32 objects, both in list order: a branch-free float32 pass marks the
  // objects that are not certain misses (most (ray, sphere) pairs miss by far); the float64
  // evaluation of the reference then runs for the marked ones only.
  for (int base = 0; base < sc.nobjects; base += 32) {
    const int nb = (sc.nobjects - base < 32) ? sc.nobjects - base : 32;
    uint32_t need = (nb == 32) ? 0xFFFFFFFFu : ((1u << nb) - 1u);
    if (pre.gridOk) {
      // mask-grid scene (<= 32 objects: one batch): the ray's cell already lists what it can hit
      need &= pre.gridMask;
    } else {
      // branch-free: EVERY record goes through the sphere formula — a record that is not a float32 sphere has
      // r2m = +Inf, for which the test is never true, and is looked at again (by tag) below
      uint32_t miss = 0;
#pragma unroll 4
      for (int j = 0; j < nb; ++j)   // the same record for every lane
        miss |= uint32_t(certainMissF(loadCObjF(sc.cobjf + base + j), rf)) << j;
      if (f32ok) need &= ~miss;
    }
    r.tests += nb;
    while (need) {
      const int i = base + (__builtin_ffs(int(need)) - 1);
      need &= need - 1;
      const CObjF c = loadCObjF(sc.cobjf + i);
      if (!(c.r2m < 3.0e38f) && firstLookMiss(mp, c, rf, f32ok)) continue;   // plane below / above the ray, mesh box not entered
      if (pre.gridOk && c.r2m < 3.0e38f && certainMissF(c, rf)) continue;     // (the cell's spheres: the per-ray first look)
      evalObject(sc, mp, i, o, d, fastRay, r);
    }
  }
  return r;
}

struct StatDelta { unsigned long long v[ST_COUNT]; };
NRT_HD StatDelta zeroStats() { StatDelta s; for (int i = 0; i < ST_COUNT; ++i) s.v[i] = 0; return s; }

// ---- shade: nearest hit of the path ray, hit point and normal (renderer.nim:71-88)
struct ShadeOut { bool hit; int64_t s; V4 hitW, n; };
template <bool CL>
struct ShadeT {
  SceneArg sc; FrameParams fp; ChunkState cs; ActiveSet act; int bounce;
  NRT_HD StatDelta operator()(int64_t idx) const { ShadeOut out; return run(idx, out); }
  NRT_HD void prefetch(int64_t idx) const {   // inputs of wave position idx (identity active set only)
    if (act.list) return;
    pf4(cs.rayD, cs.S, idx);
    NRT_PREFETCH_L2(cs.active + idx);
    if (cs.nMO > 0) NRT_PREFETCH_L2(cs.gflag + idx);
    if (bounce != 0) pf4(cs.rayO, cs.S, idx);
  }
  NRT_HD StatDelta run(int64_t idx, ShadeOut& out) const {
    StatDelta st = zeroStats();
    out.hit = false; out.s = 0;
    if (idx >= activeN(act)) return st;
    const int64_t s = sampleOf(act, idx);
    // (the ray is loaded before the activity flag is looked at: both requests are in flight together)
    const V4 d = ld4(cs.rayD, cs.S, s);
    const uint8_t alive = cs.active[s], code0 = (cs.nMO > 0) ? cs.gflag[idx] : uint8_t(0);
    const V4 o = (bounce == 0) ? primaryOrigin(*sc) : ld4(cs.rayO, cs.S, s);
    if (!alive) { cs.hitObj[s] = -1; return st; }
    const TraceOut tr = traceObjects<CL>(*sc, WaveMesh{cs, s, idx, code0}, o, d, NRT_INF, bounce == 0 ? kPrimaryRay : -1);
    st.v[ST_RAYS] = 1; st.v[ST_TESTS] = tr.tests; st.v[ST_HITS] = tr.hits;
    if (bounce == 0) {
      st.v[ST_PRIMARY] = 1;
      if ((cs.aovObj || cs.aovTri || cs.aovT) && s == divFast(s, fp.spp) * fp.spp) {
        int x, y; pixelOf(fp, cs, cs.p0 + divFast(s, fp.spp), x, y);
        const int64_t pi = int64_t(y) * fp.width + x;
        if (cs.aovObj) cs.aovObj[pi] = tr.obj;
        if (cs.aovTri) cs.aovTri[pi] = (tr.obj >= 0 && tr.tri != kNoTri) ? int32_t(tr.tri) : -1;
        if (cs.aovT) cs.aovT[pi] = tr.t;
      }
    }
    if (tr.obj < 0) {  // renderer.nim:74-75 (or :123-124 for a reflection ray): background
      const double w = (bounce == 0) ? 1.0 : cs.weight[s];
      const double a0 = (bounce == 0) ? 0.0 : cs.accum[s], a1 = (bounce == 0) ? 0.0 : cs.accum[cs.S + s],
                   a2 = (bounce == 0) ? 0.0 : cs.accum[2 * cs.S + s];
      cs.accum[s] = a0 + sc->bg[0] * w;
      cs.accum[cs.S + s] = a1 + sc->bg[1] * w;
      cs.accum[2 * cs.S + s] = a2 + sc->bg[2] * w;
      cs.active[s] = 0;
      cs.hitObj[s] = -1;
      return st;
    }
    const DObject& ob = sc->objects[tr.obj];
    const V4 hitW = add(o, scale(d, tr.t));
    V4 n;
    if (tr.tri == kNoTri) {
      if (ob.kind == GEOM_PLANE) {   // constant normal: the product was done at scene build
        n = v4(ob.plane_nw[0], ob.plane_nw[1], ob.plane_nw[2], ob.plane_nw[3]);
      } else {
        const V4 hitO = mulm(ob.w2o, hitW);
        n = mulm(ob.o2w, geomNormal(ob, hitO));
      }
    } else {
      const DMesh& m = sc->meshes[ob.mesh];
      const double* nn = m.normals + 4 * m.nidx[3 * int64_t(tr.tri)];
      n = mulm(ob.o2w, v4(nn[0], nn[1], nn[2], nn[3]));
    }
    st4(cs.hitW, cs.S, s, hitW);
    st4(cs.nrm, cs.S, s, n);
    cs.hitObj[s] = tr.obj;
    out.hit = true; out.s = s; out.hitW = hitW; out.n = n;
    return st;
  }
};

using Shade = ShadeT<false>;
using ShadeClustered = ShadeT<true>;

// ---- fused producer: the kernel that creates the primary rays also evaluates their gate codes while
// the rays are in registers (the CUDA backend counts the codes per 256-position block in the same
// kernel, so the separate flags pass over the sample state disappears).  `emit(k, mo, code)` receives
// the code of wave position i * mult + k for mesh object mo.  (The same fusion of Shade with the
// shadow-ray gate was measured slower on B200: 98 registers, two shadow rays per thread in sequence.)
struct GenGate {      // primary rays: position == sample, mult == 1
  GenSimple gen; Gate gate; int nMO;
  NRT_HD void prefetch(int64_t) const {}   // (no inputs in memory)
  template <class E> NRT_HD void operator()(int64_t s, E& emit) const {
    GenOut out; gen.run(s, out);
    for (int mo = 0; mo < nMO; ++mo)
      emit(0, mo, out.alive ? gateCode(gate.evalRay(out.o, out.d, uint32_t(s), mo, 0)) : uint8_t(0));
  }
};

// Shadow rays: one element per ACTIVE SAMPLE evaluates the gate of its nL shadow rays (positions
// idx * nL + l): the hit record is loaded once and the shared origin hitW + n * bias is formed once.
struct ShadowGate {   // gate.kind == WAVE_SHADOW, mult == nL
  Gate gate; int nMO;
  NRT_HD void prefetch(int64_t idx) const {   // the hit record of wave position idx (identity active set only)
    if (gate.act.list) return;
    const ChunkState& cs = gate.cs;
    pf4(cs.hitW, cs.S, idx); pf4(cs.nrm, cs.S, idx);
    NRT_PREFETCH_L2(cs.hitObj + idx);
  }
  template <class E> NRT_HD void operator()(int64_t idx, E& emit) const {
    const ChunkState& cs = gate.cs;
    const int64_t s = sampleOf(gate.act, idx);
    // (hit record requested together with the hit flag: one memory round trip instead of two)
    V4 hitW = ld4(cs.hitW, cs.S, s), nrm = ld4(cs.nrm, cs.S, s);
    const bool hit = cs.hitObj[s] >= 0;
    NRT_KEEP_D(hitW.x); NRT_KEEP_D(hitW.y); NRT_KEEP_D(hitW.z); NRT_KEEP_D(hitW.w);
    NRT_KEEP_D(nrm.x); NRT_KEEP_D(nrm.y); NRT_KEEP_D(nrm.z); NRT_KEEP_D(nrm.w);
    V4 o = hitW;
    if (hit) o = add(hitW, scale(nrm, gate.fp.bias));   // renderer.nim:98
    for (int l = 0; l < cs.nL; ++l) {
      V4 d = o;
      if (hit) d = scale(getShadingInfo(lightOf(*gate.sc, l), hitW).lightDir, -1.0);   // renderer.nim:99
      for (int mo = 0; mo < nMO; ++mo)
        emit(l, mo, hit ? gateCode(gate.evalRay(o, d, uint32_t(s * cs.nL + l), mo, l)) : uint8_t(0));
    }
  }
};

// ---- shadow trace: one element per shadow ray (active sample i / nL, light i % nL): the trace()
// call of renderer.nim:101-102; only "some object hit before the light" is kept (renderer.nim:103)
struct ShadowTrace {
  SceneArg sc; FrameParams fp; ChunkState cs; ActiveSet act;
  NRT_HD void prefetch(int64_t) const {}
  NRT_HD StatDelta operator()(int64_t idx) const {
    StatDelta st = zeroStats();
    const int64_t si = divFast(idx, cs.nL);
    if (si >= activeN(act)) return st;
    const int l = int(idx - si * cs.nL);
    const int64_t s = sampleOf(act, si);
    if (cs.hitObj[s] < 0) return st;
    const V4 hitW = ld4(cs.hitW, cs.S, s), n = ld4(cs.nrm, cs.S, s);
    const ShadingInfo li = getShadingInfo(lightOf(*sc, l), hitW);
    const V4 so = add(hitW, scale(n, fp.bias)), sd = scale(li.lightDir, -1.0);   // renderer.nim:98-99
    const TraceOut tr = traceObjects<false>(*sc, WaveMesh{cs, s * cs.nL + l, idx, (cs.nMO > 0) ? cs.gflag[idx] : uint8_t(0)}, so, sd, li.lightDistance);
    st.v[ST_RAYS] = 1; st.v[ST_TESTS] = tr.tests; st.v[ST_HITS] = tr.hits;
    cs.occ[s * cs.nL + l] = tr.obj >= 0 ? 1 : 0;
    return st;
  }
};

// The same per ACTIVE SAMPLE: its nL shadow rays share the loads of the hit record and the origin.
template <bool CL>
struct ShadowTraceSampleT {
  SceneArg sc; FrameParams fp; ChunkState cs; ActiveSet act;
  NRT_HD void prefetch(int64_t idx) const {   // the hit record and gate codes of wave position idx (identity active set only)
    if (act.list) return;
    pf4(cs.hitW, cs.S, idx); pf4(cs.nrm, cs.S, idx);
    NRT_PREFETCH_L2(cs.hitObj + idx);
    if (cs.nMO > 0) NRT_PREFETCH_L2(cs.gflag + idx * cs.nL);
  }
  NRT_HD StatDelta operator()(int64_t idx) const {
    StatDelta st = zeroStats();
    if (idx >= activeN(act)) return st;
    const int64_t s = sampleOf(act, idx);
    // (hit record requested together with the hit flag: one memory round trip instead of two)
    V4 hitW = ld4(cs.hitW, cs.S, s), n = ld4(cs.nrm, cs.S, s);
    const int32_t ho = cs.hitObj[s];
    uint32_t codes = 0;   // gate codes (mesh object 0) of the first four lights, requested with the hit record
    if (cs.nMO > 0)
      for (int l = 0; l < cs.nL && l < 4; ++l) codes |= uint32_t(cs.gflag[idx * cs.nL + l]) << (8 * l);
    NRT_KEEP_D(hitW.x); NRT_KEEP_D(hitW.y); NRT_KEEP_D(hitW.z); NRT_KEEP_D(hitW.w);
    NRT_KEEP_D(n.x); NRT_KEEP_D(n.y); NRT_KEEP_D(n.z); NRT_KEEP_D(n.w);
    if (ho < 0) return st;
    const V4 so = add(hitW, scale(n, fp.bias));                                   // renderer.nim:98
    for (int l = 0; l < cs.nL; ++l) {
      const ShadingInfo li = getShadingInfo(lightOf(*sc, l), hitW);
      const V4 sd = scale(li.lightDir, -1.0);                                      // renderer.nim:99
      const uint8_t code0 = (l < 4) ? uint8_t(codes >> (8 * l)) : ((cs.nMO > 0) ? cs.gflag[idx * cs.nL + l] : uint8_t(0));
      const TraceOut tr = traceObjects<CL>(*sc, WaveMesh{cs, s * cs.nL + l, idx * cs.nL + l, code0}, so, sd, li.lightDistance, l);
      st.v[ST_RAYS] += 1; st.v[ST_TESTS] += tr.tests; st.v[ST_HITS] += tr.hits;
      cs.occ[s * cs.nL + l] = tr.obj >= 0 ? 1 : 0;
    }
    return st;
  }
};

using ShadowTraceSample = ShadowTraceSampleT<false>;
using ShadowTraceSampleClustered = ShadowTraceSampleT<true>;

// ---- resolve: shadeDiffuse of the unoccluded lights + reflection set-up (renderer.nim:90-127)
// Samples whose path continues keep active == 1; the backend compacts them IN SAMPLE ORDER into
// the next bounce's active list (compactActive), so reflection rays of neighbouring pixels stay
// neighbours in the queues.
// The work of one hit sample after its shadow rays are known.  `occluded(l)`: the shadow ray towards
// light l hit something (renderer.nim:103).  `hitW` is only valid with point lights (a DistantLight
// ignores it); without them it is fetched here for the (few) continuing samples only.
template <class OCC>
NRT_HD void resolveSample(const DScene* sc, const FrameParams& fp, const ChunkState& cs, int bounce, int pointLights,
                          int64_t s, int objHit, V4 hitW, const V4& n, const OCC& occluded, StatDelta& st) {
  const DObject& ob = sc->objects[objHit];
  V3 local = v3(0.0, 0.0, 0.0);
  for (int l = 0; l < cs.nL; ++l) {
    if (occluded(l)) continue;
    const ShadingInfo si = getShadingInfo(lightOf(*sc, l), hitW);
    local = add(local, shadeDiffuse(ob, si, n));
  }
  const double k = ob.reflection, w = (bounce == 0) ? 1.0 : cs.weight[s];
  const int depth = (fp.depth_mode == DEPTH_INTENDED) ? (1 + bounce) : 0;  // renderer.nim:108 + depth bug
  bool cont = false;
  double wl = w;  // weight of `local` in the pixel
  if (k > 0.0 && depth <= fp.max_ray_depth) {
    if (bounce >= fp.bounce_cap) {
      st.v[ST_CAPPED] = 1;
    } else {
      cont = true;
      wl = w * (1.0 - k);  // result = (1-k)*result + k*reflColor (renderer.nim:126-127)
    }
  }
  const double a0 = (bounce == 0) ? 0.0 : cs.accum[s], a1 = (bounce == 0) ? 0.0 : cs.accum[cs.S + s],
               a2 = (bounce == 0) ? 0.0 : cs.accum[2 * cs.S + s];
  cs.accum[s] = a0 + local.x * wl;
  cs.accum[cs.S + s] = a1 + local.y * wl;
  cs.accum[2 * cs.S + s] = a2 + local.z * wl;
  if (cont) {
    if (!pointLights) hitW = ld4(cs.hitW, cs.S, s);
    const V4 i = ld4(cs.rayD, cs.S, s);
    const V4 r = sub(i, scale(n, 2 * dot(n, i)));  // renderer.nim:112
    st4(cs.rayO, cs.S, s, add(hitW, scale(r, fp.bias)));
    st4(cs.rayD, cs.S, s, r);
    cs.weight[s] = w * k;
    cs.active[s] = 1;
    st.v[ST_CONT] = 1;
  } else {
    cs.active[s] = 0;
  }
}

struct Resolve {
  SceneArg sc; FrameParams fp; ChunkState cs; ActiveSet act; int bounce;
  int pointLights;   // some light is a PointLight: getShadingInfo needs the hit point (a DistantLight ignores it)
  NRT_HD void prefetch(int64_t) const {}
  NRT_HD StatDelta operator()(int64_t idx) const {
    StatDelta st = zeroStats();
    if (idx >= activeN(act)) return st;
    const int64_t s = sampleOf(act, idx);
    // (hit record requested together with the hit flag: one memory round trip instead of two)
    // The hit point is only read by point lights and by the reflection set-up: without point lights it is
    // fetched for the (few) continuing samples only, 32 bytes less per sample for everyone else.
    V4 hitW = v4(0.0, 0.0, 0.0, 1.0), n = ld4(cs.nrm, cs.S, s);
    if (pointLights) hitW = ld4(cs.hitW, cs.S, s);
    const int objHit = cs.hitObj[s];
    NRT_KEEP_D(hitW.x); NRT_KEEP_D(hitW.y); NRT_KEEP_D(hitW.z); NRT_KEEP_D(hitW.w);
    NRT_KEEP_D(n.x); NRT_KEEP_D(n.y); NRT_KEEP_D(n.z); NRT_KEEP_D(n.w);
    if (objHit < 0) return st;
    const uint8_t* occ = cs.occ + s * cs.nL;
    resolveSample(sc, fp, cs, bounce, pointLights, s, objHit, hitW, n, [occ](int l) { return occ[l] != 0; }, st);
    return st;
  }
};

// ---- shadow trace + resolve in one pass over the hit samples (at most 32 lights): the occlusion flags
// stay in a register instead of going through cs.occ, the hit id is read once, and the normal (and the
// hit point, where Resolve reads it) is requested again after the shadow rays — the line was fetched a
// few microseconds earlier by the same thread, so the second request is served on chip — instead of
// being kept in registers across the object scans (which would cost the scans their occupancy).
// Per sample of HBM traffic this drops Resolve's 38 bytes of reads and the 2 flag bytes written.
template <bool CL>
struct ShadowResolveT {
  SceneArg sc; FrameParams fp; ChunkState cs; ActiveSet act; int bounce; int pointLights;
  NRT_HD void prefetch(int64_t idx) const {   // the hit record and gate codes of wave position idx (identity active set only)
    if (act.list) return;
    pf4(cs.hitW, cs.S, idx); pf4(cs.nrm, cs.S, idx);
    NRT_PREFETCH_L2(cs.hitObj + idx);
    if (cs.nMO > 0) NRT_PREFETCH_L2(cs.gflag + idx * cs.nL);
  }
  NRT_HD StatDelta operator()(int64_t idx) const {
    StatDelta st = zeroStats();
    if (idx >= activeN(act)) return st;
    const int64_t s = sampleOf(act, idx);
    uint32_t occ = 0;
    int32_t ho;
    {
      V4 hitW = ld4(cs.hitW, cs.S, s), n = ld4(cs.nrm, cs.S, s);
      ho = cs.hitObj[s];
      uint32_t codes = 0;   // gate codes (mesh object 0) of the first four lights, requested with the hit record
      if (cs.nMO > 0)
        for (int l = 0; l < cs.nL && l < 4; ++l) codes |= uint32_t(cs.gflag[idx * cs.nL + l]) << (8 * l);
      NRT_KEEP_D(hitW.x); NRT_KEEP_D(hitW.y); NRT_KEEP_D(hitW.z); NRT_KEEP_D(hitW.w);
      NRT_KEEP_D(n.x); NRT_KEEP_D(n.y); NRT_KEEP_D(n.z); NRT_KEEP_D(n.w);
      if (ho < 0) return st;
      const V4 so = add(hitW, scale(n, fp.bias));                                   // renderer.nim:98
      for (int l = 0; l < cs.nL; ++l) {
        const ShadingInfo li = getShadingInfo(lightOf(*sc, l), hitW);
        const V4 sd = scale(li.lightDir, -1.0);                                      // renderer.nim:99
        const uint8_t code0 = (l < 4) ? uint8_t(codes >> (8 * l)) : ((cs.nMO > 0) ? cs.gflag[idx * cs.nL + l] : uint8_t(0));
        const TraceOut tr = traceObjects<CL>(*sc, WaveMesh{cs, s * cs.nL + l, idx * cs.nL + l, code0}, so, sd, li.lightDistance, l);
        st.v[ST_RAYS] += 1; st.v[ST_TESTS] += tr.tests; st.v[ST_HITS] += tr.hits;
        occ |= uint32_t(tr.obj >= 0 ? 1u : 0u) << l;
      }
    }
    NRT_COMPILER_FENCE();   // Resolve's inputs are requested here, not before / across the object scans
    const V4 n = ld4(cs.nrm, cs.S, s);
    V4 hitW = v4(0.0, 0.0, 0.0, 1.0);
    if (pointLights) hitW = ld4(cs.hitW, cs.S, s);
    resolveSample(sc, fp, cs, bounce, pointLights, s, ho, hitW, n, [occ](int l) { return ((occ >> l) & 1u) != 0; }, st);
    return st;
  }
};
using ShadowResolve = ShadowResolveT<false>;
using ShadowResolveClustered = ShadowResolveT<true>;

// ---- fused path kernels -----------------------------------------------------------------------------
// The wavefront above streams every sample's float64 state through HBM five times per bounce.  Most samples
// never need it: in a frame of BASELINE config 4 about 88 % of the samples have NO ray (primary or shadow)
// that enters a mesh box, so their whole bounce — castPrimaryRay, trace, shade's shadow rays, shadeDiffuse,
// the reflection set-up (renderer.nim:31-127) — is done by ONE thread in registers:
//   FusedBounce   one bounce of every active sample (bounce 0: from castPrimaryRay; later: from the stored ray).  A sample one of whose rays passes a mesh's AABB gate is handed
//                 to the wavefront instead (its ray and active = 1 are stored, nothing else, and it adds
//                 nothing to Stats: the wavefront redoes it from the ray); every other sample is finished
//                 here: accumulator written once, and, if its path continues, the reflection ray is stored
//                 (flag kFlagContinues).
//   PathTail      the remaining bounces of a SMALL active list (below NRT_TAIL_BELOW samples), one lane per
//                 sample to the end of its path, the warp walking the mesh hierarchy for the rays that enter a
//                 box: one launch instead of ~25 dependent launches and a host round trip per bounce for waves
//                 of a few thousand samples.  (Large waves stay with the wavefront: its prefilter shares every
//                 record it loads among the 256 rays of a run — measured 10x the walk's throughput per ray.)
//   PathMega      every sample start to end in one launch (NRT_PATH=2; for comparison and very small frames).
// The arithmetic is the wavefront's, operation by operation (Shade, ShadowResolve, resolveSample).
enum PathKind { PATH_PRIMARY = 0, PATH_TAIL = 1, PATH_MEGA = 2 };
NRT_HD void writeAovOf(const FrameParams& fp, const ChunkState& cs, int64_t s, const TraceOut& tr) {
  if ((cs.aovObj || cs.aovTri || cs.aovT) && s == divFast(s, fp.spp) * fp.spp) {
    int x, y; pixelOf(fp, cs, cs.p0 + divFast(s, fp.spp), x, y);
    const int64_t pi = int64_t(y) * fp.width + x;
    if (cs.aovObj) cs.aovObj[pi] = tr.obj;
    if (cs.aovTri) cs.aovTri[pi] = (tr.obj >= 0 && tr.tri != kNoTri) ? int32_t(tr.tri) : -1;
    if (cs.aovT) cs.aovT[pi] = tr.t;
  }
}
// normal at the hit (renderer.nim:76-88)
NRT_HD V4 hitNormal(const DScene& sc, const DObject& ob, const TraceOut& tr, V4 hitW) {
  if (tr.tri == kNoTri) {
    if (ob.kind == GEOM_PLANE) return v4(ob.plane_nw[0], ob.plane_nw[1], ob.plane_nw[2], ob.plane_nw[3]);   // the product was done at scene build
    return mulm(ob.o2w, geomNormal(ob, mulm(ob.w2o, hitW)));
  }
  const DMesh& m = sc.meshes[ob.mesh];
  const double* nn = m.normals + 4 * m.nidx[3 * int64_t(tr.tri)];
  return mulm(ob.o2w, v4(nn[0], nn[1], nn[2], nn[3]));
}

// cs.active after FusedBounce: 0 = the sample is finished, kFlagWavefront = bounce 0 is the wavefront's,
// kFlagContinues = bounce 0 done here, the reflection ray is stored.  After the wavefront's bounce 0 every
// nonzero flag is a sample whose path continues (Resolve writes 0 / 1 for its samples).
static constexpr uint8_t kFlagWavefront = 1, kFlagContinues = 2;
template <bool CL>
struct FusedBounceT {
  // The scene header travels BY VALUE (kernel-parameter space = constant bank): camera matrix and origin, tan(fov/2),
  // background, the table pointers and counts are operands or uniform constant loads instead of ~30 generic 64-bit loads
  // per sample through a pointer (r02 ncu: 47 LD + 46 LDG per sample, long-scoreboard the top stall of this kernel).
  SceneArg sc; FrameParams fp; ChunkState cs; int force_exact;
  int genFromState;   // bounce 0: the primary rays are in cs.rayD / cs.active already (jittered kinds: GenJittered)
  ActiveSet act;      // the samples of this bounce (bounce 0: every sample of the chunk)
  int bounce;
  NRT_HD void prefetch(int64_t) const {}
  NRT_HD StatDelta operator()(int64_t idx) const {
    StatDelta st = zeroStats();
    const int64_t s = sampleOf(act, idx);
    V4 o, d;
    double w = 1.0, a0 = 0.0, a1 = 0.0, a2 = 0.0;
    double pcx = 0.0, pcy = 0.0;   // castPrimaryRay's (cx, cy) when this thread generated the ray
    bool havePc = false;
    if (bounce == 0) {
      bool alive;
      if (genFromState) {
        o = primaryOrigin(*sc); d = ld4(cs.rayD, cs.S, s); alive = cs.active[s] != 0;
      } else {
        GenOut g; GenSimple{sc, fp, cs}.compute(s, g);
        o = g.o; d = g.d; alive = g.alive; pcx = g.cx; pcy = g.cy; havePc = true;
      }
      if (!alive) { cs.active[s] = 0; return st; }   // (skipped pixel of a progressive pass)
    } else {
      o = ld4(cs.rayO, cs.S, s); d = ld4(cs.rayD, cs.S, s);
      w = cs.weight[s];
      a0 = cs.accum[s]; a1 = cs.accum[cs.S + s]; a2 = cs.accum[2 * cs.S + s];
    }
    const int nMO = cs.nMO, nL = cs.nL;
    const RayPre pre = havePc ? makeRayPrePrimary(*sc, o, d, pcx, pcy) : makeRayPreGrid(*sc, o, d, bounce == 0 ? kPrimaryRay : -1);
    for (int mo = 0; mo < nMO; ++mo)
      if (meshGatePassPre(*sc, mo, o, d, pre)) return toWavefront(s, d);
    const TraceOut tr = traceObjectsPre<CL>(*sc, NoMesh{}, o, d, NRT_INF, pre, bounce == 0 ? kPrimaryRay : -1);
    st.v[ST_RAYS] = 1; st.v[ST_TESTS] = tr.tests; st.v[ST_HITS] = tr.hits;
    if (bounce == 0) st.v[ST_PRIMARY] = 1;
    uint8_t flag = 0;
    if (tr.obj < 0) {   // renderer.nim:74-75 / :123-124: background
      if (bounce == 0) writeAovOf(fp, cs, s, tr);
      a0 = a0 + sc->bg[0] * w; a1 = a1 + sc->bg[1] * w; a2 = a2 + sc->bg[2] * w;
    } else {
      const DObject& ob = sc->objects[tr.obj];
      const V4 hitW = add(o, scale(d, tr.t));
      const V4 n = hitNormal(*sc, ob, tr, hitW);
      const V4 so = add(hitW, scale(n, fp.bias));                                    // renderer.nim:98
      // One pass over the lights: a shadow ray that enters a mesh box hands the sample to the wavefront (nothing has
      // been written or counted yet: `st` and `local` die with the return), any other is traced here — the ray's
      // float32 image and grid cell (RayPre) serve the gate and the object scan alike.
      V3 local = v3(0.0, 0.0, 0.0);
      for (int l = 0; l < nL; ++l) {
        const ShadingInfo li = getShadingInfo(lightOf(*sc, l), hitW);
        const V4 sdir = scale(li.lightDir, -1.0);                                    // renderer.nim:99
        const RayPre sp = makeRayPreGrid(*sc, so, sdir, l);
        for (int mo = 0; mo < nMO; ++mo)
          if (meshGatePassPre(*sc, mo, so, sdir, sp)) return toWavefront(s, d);
        const TraceOut ts = traceObjectsPre<CL>(*sc, NoMesh{}, so, sdir, li.lightDistance, sp, l);
        st.v[ST_RAYS] += 1; st.v[ST_TESTS] += ts.tests; st.v[ST_HITS] += ts.hits;
        if (ts.obj < 0) local = add(local, shadeDiffuse(ob, li, n));                 // renderer.nim:103-105
      }
      if (bounce == 0) writeAovOf(fp, cs, s, tr);
      // resolveSample's arithmetic
      const double k = ob.reflection;
      const int depth = (fp.depth_mode == DEPTH_INTENDED) ? (1 + bounce) : 0;      // renderer.nim:108 + depth bug
      bool cont = false;
      double wl = w;
      if (k > 0.0 && depth <= fp.max_ray_depth) {
        if (bounce >= fp.bounce_cap) st.v[ST_CAPPED] = 1;
        else { cont = true; wl = w * (1.0 - k); }
      }
      a0 = a0 + local.x * wl; a1 = a1 + local.y * wl; a2 = a2 + local.z * wl;
      if (cont) {
        const V4 r = sub(d, scale(n, 2 * dot(n, d)));                                // renderer.nim:112
        st4(cs.rayO, cs.S, s, add(hitW, scale(r, fp.bias))); st4(cs.rayD, cs.S, s, r);
        cs.weight[s] = w * k;
        flag = kFlagContinues;
      }
    }
    cs.accum[s] = a0; cs.accum[cs.S + s] = a1; cs.accum[2 * cs.S + s] = a2;
    cs.active[s] = flag;
    return st;
  }
  // hands sample s to the wavefront for this bounce: its ray (already stored from bounce 1 on) + the flag that puts
  // it on the wavefront's list; nothing else is written and nothing is counted (the wavefront redoes the bounce)
  NRT_HD StatDelta toWavefront(int64_t s, V4 d) const {
    if (bounce == 0 && !genFromState) st4(cs.rayD, cs.S, s, d);
    cs.active[s] = kFlagWavefront;
    return zeroStats();
  }
};
using FusedBounce = FusedBounceT<false>;
using FusedBounceClustered = FusedBounceT<true>;

// ---- PathTail / PathMega: warp-synchronous paths -----------------------------------------------------
// One lane per sample; the lanes of a warp walk through the bounces together (`while any lane alive`), so
// that the mesh intersections of the warp's current rays happen at points where all 32 lanes are present:
// there the warp takes the rays that passed an AABB gate one at a time and walks the mesh hierarchy for
// each with all its lanes (nrt.cu: WarpCoop — 32 chunk bounds per step, ballots, two sub-chunks of 16 per
// step, the float64 evaluations of the survivors side by side in different lanes).  A thread walking the
// hierarchy alone (ScalarCoop: the emulation, and mesh objects beyond kMaxWalkMO) serialises on divergence:
// measured 13.5 ms for the 0.56 M tail samples of a config-4 frame against ~1 ms of the wavefront.
// `W::any(x)`: some lane of the warp has x;  `W::meshAll(...)`: TriangleMesh.intersect for every lane's ray.
struct ScalarCoop {
  NRT_HD bool any(bool x) const { return x; }
  NRT_HD void meshAll(const DScene& sc, bool need, int mode, int l, V4 o, V4 d, int force_exact, MeshRes& mr) const {
    for (int mo = 0; mo < sc.nmesh_objs && mo < kMaxWalkMO; ++mo) {
      mr.t[mo] = NRT_NEG_INF; mr.tri[mo] = kNoTri;
      if (!need) continue;
      const MeshHit h = meshIntersectWalk(sc, mo, mode, l, o, d, force_exact);
      mr.t[mo] = h.t; mr.tri[mo] = h.tri;
    }
  }
};

template <bool CL, int KIND>   // KIND: PATH_TAIL (the samples of the tail list, from bounce 1) or PATH_MEGA (every sample, from its primary ray)
struct PathWarpT {
  SceneArg sc; FrameParams fp; ChunkState cs; int force_exact;
  int genFromState;
  ActiveSet act;     // PATH_TAIL: the samples to finish (list + count on the device)
  int bounce0;       // PATH_TAIL: the bounce their stored rays belong to
  NRT_HD void prefetch(int64_t) const {}
  NRT_HD StatDelta operator()(int64_t idx) const { return run(idx, true, ScalarCoop{}); }
  template <class W>
  NRT_HD StatDelta run(int64_t idx, bool valid, const W& coop) const {
    StatDelta st = zeroStats();
    int64_t s = 0;
    V4 o = v4(0.0, 0.0, 0.0, 1.0), d = v4(0.0, 0.0, -1.0, 0.0);
    double w = 1.0, a0 = 0.0, a1 = 0.0, a2 = 0.0;
    int bounce = (KIND == PATH_TAIL) ? bounce0 : 0;
    bool alive = valid;
    if (valid) {
      if (KIND == PATH_TAIL) {
        s = sampleOf(act, idx);
        d = ld4(cs.rayD, cs.S, s);
        if (bounce0 == 0) {   // a sample FusedBounce handed over at bounce 0: only its primary direction is stored
          o = primaryOrigin(*sc);
        } else {
          o = ld4(cs.rayO, cs.S, s);
          w = cs.weight[s];
          a0 = cs.accum[s]; a1 = cs.accum[cs.S + s]; a2 = cs.accum[2 * cs.S + s];
        }
      } else {
        s = idx;
        if (genFromState) {
          o = primaryOrigin(*sc); d = ld4(cs.rayD, cs.S, s); alive = cs.active[s] != 0;
        } else {
          GenOut g; GenSimple{sc, fp, cs}.compute(s, g);
          o = g.o; d = g.d; alive = g.alive;
        }
      }
    }
    const bool started = alive;
    const int nL = cs.nL;
    while (coop.any(alive)) {
      MeshRes mr;
      coop.meshAll(*sc, alive, bounce == 0 ? FM_ORIGIN : FM_GENERAL, 0, o, d, force_exact, mr);
      bool hit = false;
      TraceOut tr; tr.obj = -1; tr.t = 0.0; tr.tri = kNoTri; tr.tests = 0; tr.hits = 0;
      V4 hitW = o, n = d, so = o;
      if (alive) {
        tr = traceObjects<CL>(*sc, PreMesh{mr, WalkMesh{sc, bounce == 0 ? FM_ORIGIN : FM_GENERAL, 0, force_exact}}, o, d, NRT_INF, bounce == 0 ? kPrimaryRay : -1);
        st.v[ST_RAYS] += 1; st.v[ST_TESTS] += tr.tests; st.v[ST_HITS] += tr.hits;
        if (bounce == 0) { st.v[ST_PRIMARY] = 1; writeAovOf(fp, cs, s, tr); }
        if (tr.obj < 0) {   // renderer.nim:74-75 / :123-124: background
          a0 = a0 + sc->bg[0] * w; a1 = a1 + sc->bg[1] * w; a2 = a2 + sc->bg[2] * w;
          alive = false;
        } else {
          hit = true;
          hitW = add(o, scale(d, tr.t));
          n = hitNormal(*sc, sc->objects[tr.obj], tr, hitW);
          so = add(hitW, scale(n, fp.bias));                                         // renderer.nim:98
        }
      }
      V3 local = v3(0.0, 0.0, 0.0);
      for (int l = 0; l < nL; ++l) {
        ShadingInfo li; li.lightDir = d; li.lightIntensity = v3(0.0, 0.0, 0.0); li.lightDistance = NRT_INF;
        if (hit) li = getShadingInfo(lightOf(*sc, l), hitW);
        const V4 sdir = scale(li.lightDir, -1.0);                                    // renderer.nim:99
        const int smode = lightOf(*sc, l).kind == LIGHT_DISTANT ? FM_DIR : FM_GENERAL;
        coop.meshAll(*sc, hit, smode, l, so, sdir, force_exact, mr);
        if (hit) {
          const TraceOut ts = traceObjects<CL>(*sc, PreMesh{mr, WalkMesh{sc, smode, l, force_exact}}, so, sdir, li.lightDistance, l);
          st.v[ST_RAYS] += 1; st.v[ST_TESTS] += ts.tests; st.v[ST_HITS] += ts.hits;
          if (ts.obj < 0) local = add(local, shadeDiffuse(sc->objects[tr.obj], li, n));   // renderer.nim:103-105
        }
      }
      if (hit) {   // resolveSample's arithmetic
        const double k = sc->objects[tr.obj].reflection;
        const int depth = (fp.depth_mode == DEPTH_INTENDED) ? (1 + bounce) : 0;    // renderer.nim:108 + depth bug
        bool cont = false;
        double wl = w;
        if (k > 0.0 && depth <= fp.max_ray_depth) {
          if (bounce >= fp.bounce_cap) st.v[ST_CAPPED] += 1;
          else { cont = true; wl = w * (1.0 - k); }
        }
        a0 = a0 + local.x * wl; a1 = a1 + local.y * wl; a2 = a2 + local.z * wl;
        if (cont) {
          const V4 r = sub(d, scale(n, 2 * dot(n, d)));                              // renderer.nim:112
          o = add(hitW, scale(r, fp.bias)); d = r; w = w * k;
        } else {
          alive = false;
        }
      }
      ++bounce;   // (the same for every lane that is still alive)
    }
    if (started) {
      cs.accum[s] = a0; cs.accum[cs.S + s] = a1; cs.accum[2 * cs.S + s] = a2;
      if (KIND == PATH_TAIL) cs.active[s] = 0;   // finished: not on any later list (a hard list taken mid-frame)
    }
    return st;
  }
};
using PathTail = PathWarpT<false, PATH_TAIL>;
using PathTailClustered = PathWarpT<true, PATH_TAIL>;
using PathMega = PathWarpT<false, PATH_MEGA>;
using PathMegaClustered = PathWarpT<true, PATH_MEGA>;

// ---- the fork (nrt_renderer.h: Renderer::sub): the samples FusedBounce finished at bounce 0 with a reflection ray
// stored are moved into a second pipeline's sample space (positions 0 .. n-1, in list order) and come back as their
// accumulators.  `a`: the main pipeline's state, `b`: the helper's.
struct GatherPool {
  ChunkState a, b; const uint32_t* list;
  NRT_HD void operator()(int64_t i) const {
    const int64_t s = int64_t(list[i]);
    st4(b.rayO, b.S, i, ld4(a.rayO, a.S, s));
    st4(b.rayD, b.S, i, ld4(a.rayD, a.S, s));
    b.weight[i] = a.weight[s];
    b.accum[i] = a.accum[s]; b.accum[b.S + i] = a.accum[a.S + s]; b.accum[2 * b.S + i] = a.accum[2 * a.S + s];
  }
};
struct ScatterAccum {
  ChunkState a, b; const uint32_t* list;
  NRT_HD void operator()(int64_t i) const {
    const int64_t s = int64_t(list[i]);
    a.accum[s] = b.accum[i]; a.accum[a.S + s] = b.accum[b.S + i]; a.accum[2 * a.S + s] = b.accum[2 * b.S + i];
  }
};

// ---- per-band counts of a sorted sample list (the bounce-0 wavefront list): the next frame's lane partition
// puts the bands with mesh work first (nrt.cu: lanesFor).  One thread per band: two binary searches.
struct BandCount {
  const uint32_t* list; const uint32_t* count; int64_t p0, band_pix, nS; int spp; uint32_t* out;
  NRT_HD int64_t lowerBound(int64_t n, int64_t key) const {
    int64_t lo = 0, hi = n;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (int64_t(list[mid]) < key) lo = mid + 1; else hi = mid; }
    return lo;
  }
  NRT_HD void operator()(int64_t b) const {
    int64_t n = int64_t(*count);
    if (n > nS) n = nS;
    int64_t lo = (b * band_pix - p0) * spp, hi = ((b + 1) * band_pix - p0) * spp;
    if (lo < 0) lo = 0;
    if (hi > nS) hi = nS;
    if (hi <= lo || n == 0) return;
    out[b] += uint32_t(lowerBound(n, hi) - lowerBound(n, lo));
  }
};

// ---- finalize: sample sum in order, * 1/N, float32 store (+ step x step fill)
struct Finalize {
  FrameParams fp; ChunkState cs;
  NRT_HD void operator()(int64_t pl) const {
    int x, y; pixelOf(fp, cs, cs.p0 + pl, x, y);
    if (pixelSkipped(fp, x, y)) return;
    double r, g, b;
    const int64_t s0 = pl * fp.spp;
    if (fp.aa_kind == AA_NONE) {
      r = cs.accum[s0]; g = cs.accum[cs.S + s0]; b = cs.accum[2 * cs.S + s0];
    } else {
      r = 0.0; g = 0.0; b = 0.0;
      for (int k = 0; k < fp.spp; ++k) {
        r = r + cs.accum[s0 + k]; g = g + cs.accum[cs.S + s0 + k]; b = b + cs.accum[2 * cs.S + s0 + k];
      }
    }
    store(pl, r, g, b);
  }
  // (r, g, b) = the pixel's sample sums in sample order (its single sample for akNone)
  NRT_HD void store(int64_t pl, double r, double g, double b) const {
    int x, y; pixelOf(fp, cs, cs.p0 + pl, x, y);
    if (pixelSkipped(fp, x, y)) return;
    if (fp.aa_kind != AA_NONE) {
      const double inv = 1 / double(fp.spp);  // renderer.nim:159
      r = r * inv; g = g * inv; b = b * inv;
    }
    const float fr = float(r), fg = float(g), fbv = float(b);  // framebuf.nim:26-28
    const int x1 = (fp.step > 1) ? (x + fp.step < fp.width ? x + fp.step : fp.width) : x + 1;
    const int y1 = (fp.step > 1) ? (y + fp.step < fp.height ? y + fp.step : fp.height) : y + 1;
    for (int j = y; j < y1; ++j)
      for (int i = x; i < x1; ++i) {
        const int64_t pi = int64_t(j) * fp.width + i;
        if (cs.q.out) { storeQ(cs.q, pi, fr, fg, fbv); continue; }   // output stage fused into the pixel store
        float* p = cs.fb + pi * 3;
        p[0] = fr; p[1] = fg; p[2] = fbv;
      }
  }
};

// ---- filter records (nrt_filter.h), one element per face (or per padding slot) ----
struct RecOut { bool keep; float c[16]; float h[4]; };

// Morton keys of the faces (record order).
struct FaceKeys {
  DMesh m; uint32_t* keys; uint32_t* idx;
  NRT_HD void operator()(int64_t f) const {
    keys[f] = mortonKey(m, m.verts + 4 * m.vidx[3 * f], m.verts + 4 * m.vidx[3 * f + 1], m.verts + 4 * m.vidx[3 * f + 2]);
    idx[f] = uint32_t(f);
  }
};

// GENERAL: depends on the mesh only; record r holds face order[r] (no culling).
struct BuildRecsGeneral {
  DMesh m;
  NRT_HD void operator()(int64_t r) const {
    float c[16], h[4];
    if (r < m.nfaces) {
      const int64_t f = m.order[r];
      const double *p0 = m.verts + 4 * m.vidx[3 * f], *p1 = m.verts + 4 * m.vidx[3 * f + 1], *p2 = m.verts + 4 * m.vidx[3 * f + 2];
      makeRecGeneral(m, p0, p1, p2, c);
      BundleFrame fr;
      fr.org[0] = m.center[0]; fr.org[1] = m.center[1]; fr.org[2] = m.center[2];
      makeHotRec(FM_GENERAL, fr, p0, p1, p2, h);
      if (!(c[3] < 1e30f)) alwaysHot(FM_GENERAL, h);
    } else {
      neverHitRecord(FM_GENERAL, c);
      neverHitHot(FM_GENERAL, h);
    }
    for (int k = 0; k < 16; ++k) m.recs[fullIndex(r, k)] = c[k];
    for (int k = 0; k < 4; ++k) m.hot[recIndex(r, k, 4)] = h[k];
  }
};

// ORIGIN: per (mesh object, camera): culled against the shared origin.  Called with the Morton
// rank r; the backend compacts the kept records preserving that order.
struct BuildRecsOrigin {
  const DScene* sc; int mo;
  NRT_HD RecOut operator()(int64_t r) const {
    RecOut o;
    const DObject& ob = sc->objects[sc->mesh_obj_index[mo]];
    const DMesh& m = sc->meshes[ob.mesh];
    const int64_t f = m.order[r];
    const V4 ow = mulm(sc->c2w, v4(0.0, 0.0, 0.0, 1.0));   // castPrimaryRay's origin (renderer.nim:42)
    const V4 oo = mulm(ob.w2o, ow);                        // trace()'s object-space origin (renderer.nim:54)
    const double O[3] = {oo.x, oo.y, oo.z};
    const double *p0 = m.verts + 4 * m.vidx[3 * f], *p1 = m.verts + 4 * m.vidx[3 * f + 1], *p2 = m.verts + 4 * m.vidx[3 * f + 2];
    o.keep = makeRecOrigin(O, p0, p1, p2, uint32_t(f), o.c);
    if (!o.keep) return o;
    makeHotRec(FM_ORIGIN, sc->frames[frameIndex(sc->nlights, mo, FM_ORIGIN, 0)], p0, p1, p2, o.h);
    // a record float32 cannot hold (|v0 - O| or S overflow) becomes an always-candidate record
    if (!(o.c[3] < 1e30f)) {
      for (int k = 0; k < 12; ++k) o.c[k] = 0.f;
      o.c[3] = 1e30f; o.c[7] = bitsToFloat(uint32_t(f));
      alwaysHot(FM_ORIGIN, o.h);
    }
    return o;
  }
};

// DIR: per (mesh object, DistantLight l): culled by det < 1e-6 for the shared direction.
struct BuildRecsDir {
  const DScene* sc; int mo; int l;
  NRT_HD RecOut operator()(int64_t r) const {
    RecOut o;
    const DObject& ob = sc->objects[sc->mesh_obj_index[mo]];
    const DMesh& m = sc->meshes[ob.mesh];
    const int64_t f = m.order[r];
    const DLight& li = sc->lights[l];
    const V4 dw = scale(v4(li.dir[0], li.dir[1], li.dir[2], li.dir[3]), -1.0);   // renderer.nim:96
    const V4 dobj = mulm(ob.w2o, dw);                                             // renderer.nim:55
    const double D[3] = {dobj.x, dobj.y, dobj.z};
    const double *p0 = m.verts + 4 * m.vidx[3 * f], *p1 = m.verts + 4 * m.vidx[3 * f + 1], *p2 = m.verts + 4 * m.vidx[3 * f + 2];
    o.keep = makeRecDir(m, D, p0, p1, p2, uint32_t(f), o.c);
    if (!o.keep) return o;
    makeHotRec(FM_DIR, sc->frames[frameIndex(sc->nlights, mo, FM_DIR, l)], p0, p1, p2, o.h);
    if (!(o.c[8] < 1e29f)) alwaysHot(FM_DIR, o.h);
    return o;
  }
};

// Bounds of a record set: elements [0, nch) are the chunk bounds, elements [nch, nch * (1 + kSubPerChunk))
// the sub-chunk bounds, stored behind them (nch = chunks of the mesh's padded face count).
struct BuildBounds {
  int mode; const float* hot; const uint32_t* count; float* bounds; int64_t nch;
  NRT_HD void operator()(int64_t e) const {
    float b[4] = {0.f, 0.f, 0.f, 0.f};
    const int64_t np = paddedFaces(int64_t(*count));
    if (e < nch) {
      if (e * kRecPad < np) chunkBound(mode, hot, e, b); else neverHitHot(mode, b);
    } else {
      const int64_t sub = e - nch;
      if (sub * kSubRecs < np) subChunkBound(mode, hot, sub, b); else neverHitHot(mode, b);
    }
    for (int k = 0; k < 4; ++k) bounds[4 * e + k] = b[k];
  }
};

}  // namespace nrt
