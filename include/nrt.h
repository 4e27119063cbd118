/*
 * nrt.h — C ABI of the B200-native render hot path for nim-raytracer.
 *
 * This is the drop-in boundary behind the reference's `src/renderer` entry
 * points.  Every entry point cites the reference interface it replaces
 * (paths relative to the nim-raytracer tree).  Nim host code binds these with
 * `importc`; C++ hosts use nim_raytracer_b200/host/nrt_host.hpp;
 * Python uses nim_raytracer_b200/api.py (ctypes).
 *
 * Plain C: POD structs, plain pointers and sizes, no CUDA or torch types.
 * All functions return 0 on success or a negative nrt_status; none abort.
 *
 * The same POD scene description is consumed by the CPU oracle
 * (oracle/ref_cpu.cpp, test infrastructure only) so both sides see identical
 * inputs.
 */
#ifndef NRT_H
#define NRT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NRT_ABI_VERSION 2

/* ---------------------------------------------------------------- status -- */
typedef enum nrt_status {
  NRT_OK = 0,
  NRT_ERR_INVALID = -1,      /* bad argument (null pointer, bad enum, bad size)      */
  NRT_ERR_CUDA = -2,         /* a CUDA runtime call failed; see nrt_last_error()     */
  NRT_ERR_NO_DEVICE = -3,    /* no usable sm_100 device; there is NO CPU fallback    */
  NRT_ERR_NOT_INIT = -4,     /* nrt_init() has not been called                       */
  NRT_ERR_OVERFLOW = -5,     /* internal work lists still too small after five attempts, each sized from what the previous one asked for */
  NRT_ERR_UNSUPPORTED = -6   /* e.g. step/maxStep not a power of two (renderer.nim:166-168) */
} nrt_status;

/* ------------------------------------------------------------ scene types -- */

/* Geometry kinds: Sphere/Plane/Box/TriangleMesh of src/renderer/geom.nim:137-155 */
typedef enum nrt_geom_kind {
  NRT_GEOM_SPHERE = 0,
  NRT_GEOM_PLANE = 1,
  NRT_GEOM_BOX = 2,
  NRT_GEOM_MESH = 3
} nrt_geom_kind;

/* Light kinds: DistantLight/PointLight of src/renderer/light.nim:8-17 */
typedef enum nrt_light_kind {
  NRT_LIGHT_DISTANT = 0,
  NRT_LIGHT_POINT = 1
} nrt_light_kind;

/* AntialiasKind of src/renderer/renderer.nim:10-12 (same ordinal values). */
typedef enum nrt_aa_kind {
  NRT_AA_NONE = 0,
  NRT_AA_GRID = 1,
  NRT_AA_JITTERED = 2,
  NRT_AA_MULTI_JITTERED = 3,
  NRT_AA_CORRELATED_MULTI_JITTERED = 4
} nrt_aa_kind;

/* How `ray.depth <= opts.maxRayDepth` (renderer.nim:108) is evaluated.
 * REFBUG:   literal reference behaviour — initRay never stores depth
 *           (geom.nim:41-48) so ray.depth == 0 for every ray and reflection
 *           recursion ends only on a miss / non-reflective hit.  A safety cap
 *           (nrt_options.bounce_cap) bounds it; capped samples are counted.
 * INTENDED: primary depth 1 (initRay's default), child depth+1, reflect while
 *           depth <= maxRayDepth  ==> at most maxRayDepth reflection bounces. */
typedef enum nrt_depth_mode {
  NRT_DEPTH_REFBUG = 0,
  NRT_DEPTH_INTENDED = 1
} nrt_depth_mode;

/* 4x4 float64 matrices are passed as m[col*4 + row] (GLM column vectors,
 * v' = M * v); the Nim shim fills them element-wise. */

/* Object{name,geometry,material} of src/renderer/scene.nim:7-10 flattened with
 * Geometry (geom.nim:137-155) and Material (material.nim:4-7). */
typedef struct nrt_object {
  int32_t kind;               /* nrt_geom_kind                                   */
  int32_t mesh;               /* index into nrt_scene_desc.meshes for MESH, else -1 */
  double object_to_world[16]; /* Geometry.objectToWorld                           */
  double world_to_object[16]; /* Geometry.worldToObject (= inverse, geom.nim:162) */
  double radius;              /* Sphere.r                                         */
  double vmin[4];             /* Box.aabb.vmin (x,y,z,w as given by the caller)   */
  double vmax[4];             /* Box.aabb.vmax                                    */
  double albedo[3];           /* Material.albedo                                  */
  double reflection;          /* Material.reflection                              */
} nrt_object;

/* TriangleMesh of geom.nim:151-155: vertices/normals are Vec4[float64];
 * faces are Triangle{vertexIdx[3], normalIdx[3]} (geom.nim:21-24).  The AABB
 * is computed by the callee exactly as calcAABB (geom.nim:175-188). */
typedef struct nrt_mesh {
  int64_t nverts;
  const double* vertices;     /* nverts * 4                                       */
  int64_t nnormals;
  const double* normals;      /* nnormals * 4                                     */
  int64_t nfaces;
  const int64_t* vertex_idx;  /* nfaces * 3                                       */
  const int64_t* normal_idx;  /* nfaces * 3                                       */
} nrt_mesh;

/* DistantLight / PointLight of light.nim:8-17. */
typedef struct nrt_light {
  int32_t kind;               /* nrt_light_kind                                   */
  int32_t _pad;
  double color[3];
  double intensity;
  double dir[4];              /* DistantLight.dir (already normalised by caller)  */
  double pos[4];              /* PointLight.pos                                   */
} nrt_light;

/* Scene of src/renderer/scene.nim:13-18. */
typedef struct nrt_scene_desc {
  int32_t nobjects;
  int32_t nlights;
  int32_t nmeshes;
  int32_t _pad;
  const nrt_object* objects;  /* in list order (order matters: renderer.nim:53-65) */
  const nrt_light* lights;
  const nrt_mesh* meshes;
  double fov;                 /* degrees                                          */
  double camera_to_world[16];
  double bg_color[3];
} nrt_scene_desc;

/* Options + Antialias of src/renderer/renderer.nim:14-28. */
typedef struct nrt_options {
  int32_t width;
  int32_t height;
  int32_t aa_kind;            /* nrt_aa_kind                                      */
  int32_t grid_size;          /* Antialias.gridSize                               */
  double bias;
  int32_t max_ray_depth;
  int32_t depth_mode;         /* nrt_depth_mode                                   */
  int32_t bounce_cap;         /* safety cap on reflection bounces (0 -> 64)       */
  int32_t _pad;
  uint64_t seed;              /* jittered AA kinds: counter-based RNG seed        */
} nrt_options;

/* Stats of src/renderer/stats.nim:4-8 (+ build-specific extras after the
 * three reference counters). */
typedef struct nrt_stats {
  int64_t num_primary_rays;
  int64_t num_intersection_tests;
  int64_t num_intersection_hits;
  int64_t num_rays;             /* calls of trace(): primary + shadow + reflection */
  int64_t num_capped_samples;   /* samples stopped by bounce_cap                   */
} nrt_stats;

/* Optional per-pixel debug outputs ("AOVs") of the first sample's primary ray:
 * what trace() returned for it (renderer.nim:47-67).  Any pointer may be NULL.
 * obj_id = index into objects[] or -1; tri_id = face index or -1;
 * t_hit = tHit as float64 (+Inf on a miss). Arrays are width*height, row-major. */
typedef struct nrt_aov {
  int32_t* obj_id;
  int32_t* tri_id;
  double* t_hit;
} nrt_aov;

/* Device-side profile of the last nrt_render* call on a scene (this build's
 * measurement hook; no reference counterpart). */
typedef struct nrt_profile {
  double total_ms;              /* CUDA-event time of the whole frame on device 0  */
  double mesh_filter_ms;        /* summed CUDA-event time of the mesh kernel       */
  int64_t mesh_filter_launches;
  int64_t mesh_tests;           /* (ray, triangle) pairs evaluated by the prefilter */
  int64_t mesh_tests_ref;       /* pairs the reference would evaluate (geom.nim:346) */
  int64_t mesh_rays;            /* rays that passed the AABB gate (geom.nim:340)   */
  int64_t candidates;           /* pairs re-evaluated in float64                   */
  int64_t pre_candidates;       /* prefilter survivors re-tested by the float32 sign test */
  int64_t kernel_launches;      /* all kernels launched for the frame              */
  double fp32_flops;            /* FP32 flops executed by the mesh kernel (2/FFMA) */
  int64_t mesh_tests_by_mode[4];/* general / shared-origin / shared-dir / reserved */
  double mesh_ms_by_mode[4];
  /* fused path (ABI v2): per bounce b < 8, the samples FusedBounce took (active set) and, of those, the samples with a
   * ray entering a mesh box, which went through the wavefront for that bounce; samples finished by PathTail */
  int64_t active_samples[8];
  int64_t wavefront_samples[8];
  int64_t tail_samples;
  int64_t lanes;                /* concurrent pipelines per GPU used for the frame */
} nrt_profile;

/* Device time per kernel family of one frame (measurement hook, see nrt_set_kernel_timing). */
#define NRT_KERNEL_CATEGORIES 24
typedef struct nrt_kernel_times {
  double ms[NRT_KERNEL_CATEGORIES];        /* summed CUDA-event time of the category's launches */
  int64_t launches[NRT_KERNEL_CATEGORIES];
  double max_ms[NRT_KERNEL_CATEGORIES];    /* the category's longest single launch (the bounce-0 launch over every sample) */
} nrt_kernel_times;

typedef struct nrt_scene nrt_scene;

/* 64-byte opaque handle for sharing a device allocation across processes
 * (one process per GPU under torchrun); wraps cudaIpcMemHandle_t. */
typedef struct nrt_ipc_handle { unsigned char bytes[64]; } nrt_ipc_handle;

/* ------------------------------------------------------------ entry points -- */

/* Replaces `initRenderer*()` (renderer.nim:214-215) and the worker-pool set-up
 * `initRenderWorkers()` (src/raytracer.nim:35-39,61-65): selects the GPUs of
 * this process.  ngpu <= 0 or dev_ids == NULL -> device 0 only... see below.
 * ngpu == 0: all visible devices.  Idempotent; nrt_shutdown() undoes it. */
int nrt_init(int ngpu, const int* dev_ids);
void nrt_shutdown(void);
int nrt_device_count(void);           /* devices selected by nrt_init            */
/* The CPUs close to selected device `index` as a Linux cpulist ("0-31,64-95"; "" when unknown).  The library runs its
 * own host threads there; the caller — the Nim main thread that replaces raytracer.nim:61-70's worker pool — does
 * well to run there too (a frame is a chain of dependent launches with host round trips).  No reference counterpart. */
int nrt_device_local_cpus(int index, char* buf, int buflen);
const char* nrt_last_error(void);     /* thread-local message of the last failure */
int nrt_abi_version(void);

/* Multi-process sharding (one process per GPU): this process renders only the
 * row bands it owns (nrt_unit_owner; band height nrt_band_rows_for()) of every
 * nrt_render* call; other pixels are left untouched.  Default (0,1).
 * Reference counterpart: the per-scanline work items of raytracer.nim:67-70. */
int nrt_set_partition(int index, int count);
/* The partition that renders unit `unit` (band or rendered scanline, counted from y0) when `count` partitions share
 * a pass: unit mod count in even rounds of `count` units, count - 1 - (unit mod count) in odd rounds (a serpentine
 * deal: equal mean position inside a round for every partition).  Callers that assemble the frame themselves
 * (distributed.py: owned_rows / gather_rows) use it to know whose rows are whose. */
int nrt_unit_owner(long long unit, int count);
int nrt_band_rows(void);   /* 1: the band height of progressive passes (kept for ABI v1 callers) */
/* Band height T of a pass with these options: whole-resolution passes (step == max_step == 1) deal the image out in
 * bands of T scanlines — rows of T x T screen-space tiles, enumerated tile by tile inside a band — where T is the
 * largest power of two with T*T*spp <= 256 (16 spp: 4); partition `index` renders the bands nrt_unit_owner gives it,
 * counted from y0.  Progressive passes: 1. */
int nrt_band_rows_for(const nrt_options* opts, int step, int max_step);
/* The first scanlines of the units of [y0, y1) that partition `index` of `count` renders in a pass with this `step`
 * and band height `band` (nrt_band_rows_for): the library's own enumeration (a unit starts where (y - y0) mod
 * (step * band) == 0 and goes to nrt_unit_owner(its number, count)), for callers that assemble a frame from the parts
 * of several processes.  Writes at most `cap` rows (rows may be null with cap 0) and returns how many there are,
 * or -1 for arguments outside height > 0, step > 0, band > 0, 0 <= index < count.  No device needed.
 * Reference counterpart: the scanline work items of raytracer.nim:67-70. */
int nrt_partition_rows(int height, int y0, int y1, int step, int band, int index, int count, int* rows, int cap);

/* Deep-copies and flattens a Scene (scene.nim:13-18) to every selected GPU:
 * replaces building `Scene`/`Object`/`TriangleMesh` refs that renderLine reads
 * (renderer.nim:162).  Nothing in `desc` is retained after return. */
int nrt_scene_create(const nrt_scene_desc* desc, nrt_scene** out);
/* Re-sends a description with the same object/light/mesh shape into the
 * existing device buffers (host->device copies + device-side precompute). */
int nrt_scene_update(nrt_scene* scene, const nrt_scene_desc* desc);
void nrt_scene_destroy(nrt_scene* scene);

/* Replaces `renderLine*(scene, opts, fb, y, step, maxStep): Stats`
 * (renderer.nim:162-211) for all lines y in [y0, y1) with (y - y0) mod step == 0
 * (the lines a caller queues: raytracer.nim:67-70, gui.nim:113-122).
 * `fb` is the caller-owned Framebuf.data (utils/framebuf.nim:7-28):
 * width*height*3 float32, offset (y*width+x)*3, HOST memory.  Only pixels of
 * the requested lines (and their step x step fill blocks) are written.
 * `stats` (may be NULL) receives the sum of the per-line Stats.
 * Safe to call from several host threads (serialised internally). */
int nrt_render(nrt_scene* scene, const nrt_options* opts,
               int y0, int y1, int step, int max_step,
               float* fb, nrt_stats* stats, const nrt_aov* aov);

/* Same, but `fb_dev` / AOV pointers are DEVICE memory reachable from the
 * rendering GPUs (local, peer-mapped, or opened with nrt_ipc_open): the final
 * pixel-store kernel writes straight into it (over NVLink for a peer), so a
 * multi-GPU frame needs no separate gather pass.  Asynchronous w.r.t. the
 * host only until stats are read back; returns after the frame completed. */
int nrt_render_device(nrt_scene* scene, const nrt_options* opts,
                      int y0, int y1, int step, int max_step,
                      float* fb_dev, nrt_stats* stats, const nrt_aov* aov_dev);

/* Output stage of the reference (next-scope row f-2): clamp -> sRGB -> 8-bit
 * (utils/framebuf.nim:74-78, utils/color.nim:17-22) on the GPU.
 * rgb8 is host memory, width*height*3 bytes. */
int nrt_framebuf_to_srgb8(const float* fb_host, int width, int height,
                          int srgb, unsigned char* rgb8);
/* The general form of writePpm's sample conversion (utils/framebuf.nim:55-93): bits in 1..16,
 * maxval = 2^bits - 1; `out` receives width*height*3 uint8 samples (bits <= 8) or big-endian uint16
 * samples (bits > 8, the PPM byte order).  Host memory. */
int nrt_framebuf_quantize(const float* fb_host, int width, int height,
                          int bits, int srgb, void* out);
/* ImageRGBA.copyFrom (utils/image.nim:45-54, the GUI's display copy): round(v*255) per channel,
 * constant alpha; rgba8 is host memory, width*height*4 bytes. */
int nrt_framebuf_to_rgba8(const float* fb_host, int width, int height,
                          unsigned char alpha, unsigned char* rgba8);

/* The same conversions on DEVICE memory (fb_dev from nrt_render_device / nrt_device_alloc): no staging, no copies. */
int nrt_framebuf_quantize_device(const float* fb_dev, int width, int height,
                                 int bits, int srgb, void* out_dev);
int nrt_framebuf_to_rgba8_device(const float* fb_dev, int width, int height,
                                 unsigned char alpha, unsigned char* rgba8_dev);

/* renderLine* + writePpm / ImageRGBA.copyFrom in one call (renderer.nim:162-211 followed by
 * utils/framebuf.nim:55-93 or utils/image.nim:45-54): the output stage runs as the epilogue of the
 * kernel that stores the pixels, and only the integer image leaves the GPU — 3 (bits <= 8), 6 (bits > 8,
 * big-endian) or 4 (NRT_OUT_RGBA8) bytes per pixel instead of 12.  `image` is HOST memory laid out like the
 * PPM sample stream / ImageRGBA.data: offset (y*width+x) * bytes-per-pixel.  Only the pixels of the requested
 * lines (and their step x step fill blocks) are written.  srgb / bits are ignored for NRT_OUT_RGBA8. */
enum { NRT_OUT_RGB = 0, NRT_OUT_RGBA8 = 1 };
/* The table the device-side sRGB conversion searches (no device needed): thr[k-1] = the smallest float32 input of
 * linearToSRGB's pow branch (utils/color.nim:21) whose sample is >= k, k = 1 .. 2^bits - 1.  thr_host holds
 * 2^bits - 1 floats.  tests/ compares it with the oracle over EVERY float32 of the branch. */
int nrt_output_cut_points(int bits, float* thr_host);
int nrt_render_quantized(nrt_scene* scene, const nrt_options* opts,
                         int y0, int y1, int step, int max_step,
                         int format, int bits, int srgb, int alpha,
                         void* image, nrt_stats* stats);

int nrt_get_profile(const nrt_scene* scene, nrt_profile* out);

/* Per-kernel-family timing (no reference counterpart; the reference only prints a wall-clock total,
 * src/utils/progress.nim:27-28).  While enabled every kernel launch of nrt_render* is bracketed by
 * two CUDA events; nrt_get_kernel_times returns the sums of the last frame rendered on `scene`
 * (first device of the group).  nrt_kernel_category_name(i) names slot i ("" past the last one). */
int nrt_set_kernel_timing(int enable);
int nrt_get_kernel_times(const nrt_scene* scene, nrt_kernel_times* out);
const char* nrt_kernel_category_name(int category);

/* Device buffers + cross-process sharing for the one-process-per-GPU launch. */
int nrt_device_alloc(int64_t bytes, void** dev_ptr);         /* on group device 0 */
int nrt_device_free(void* dev_ptr);
int nrt_device_memset(void* dev_ptr, int value, int64_t bytes);
int nrt_copy_to_host(void* host_dst, const void* dev_src, int64_t bytes);
int nrt_copy_to_device(void* dev_dst, const void* host_src, int64_t bytes);
int nrt_ipc_export(void* dev_ptr, nrt_ipc_handle* out);
int nrt_ipc_open(const nrt_ipc_handle* handle, void** dev_ptr);
int nrt_ipc_close(void* dev_ptr);
int nrt_device_synchronize(void);

/* Measured FP32 FFMA peak of device 0 in TFLOP/s (register-resident FFMA
 * micro-kernel, CUDA-event timed) — the denominator BASELINE.md asks for. */
int nrt_measure_fp32_peak(double* tflops, double* sm_clock_mhz_hint);
/* The same for the float64 pipe (register-resident DFMA loop): the per-sample kernels do the reference's float64
 * arithmetic operation by operation, so this is the pipe that bounds them. */
int nrt_measure_fp64_peak(double* tflops);

/* CUDA-event bracket on the library's own stream(s): elapsed device time between
 * the two calls, max over the selected devices (bench.py's timed region). */
int nrt_timer_begin(void);
int nrt_timer_end(double* ms);

/* Pinned host memory helpers for callers that want async D2H (bench e2e). */
int nrt_host_alloc_pinned(int64_t bytes, void** host_ptr);
int nrt_host_free_pinned(void* host_ptr);
/* Page-locks caller-owned host memory (e.g. the Framebuf's data, or a shared-memory
 * framebuffer several worker processes write their scanlines into, as the threads of
 * src/raytracer.nim:67-70 do with one Framebuf) so that nrt_render copies at full
 * PCIe speed.  Optional: nrt_render accepts pageable memory as well. */
int nrt_host_register(void* host_ptr, int64_t bytes);
int nrt_host_unregister(void* host_ptr);

#ifdef __cplusplus
}
#endif
#endif /* NRT_H */
