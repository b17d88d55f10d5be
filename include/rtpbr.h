/*
 * rtpbr.h -- C-ABI of the B200-native path-tracing hot path (librtpbr.so).
 *
 * Drop-in boundary for HK-SHAO/RayTracingPBR's Python -> Taichi kernel calls.  The reference
 * has no FFI layer: the boundary is the `@ti.kernel` call plus module-global Taichi fields
 * (SURVEY.md 8(b)).  Every entry point below names the reference interface it replaces
 * (paths relative to the reference root).  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; rtpbr_last_error() gives the text
 *     (the reference raises Python exceptions from Taichi; the Python wrapper re-raises).
 *   - image fields are dense (W, H, C) float32, j (= y, up) fastest, i = x (right), origin
 *     bottom-left, exactly the layout of `ti.root.dense(ti.ij, image_resolution)`
 *     (src/fileds.py:11-13).
 *   - one CUDA stream per context; launches are asynchronous; rtpbr_download()/rtpbr_sync()
 *     synchronise (Taichi: implicit sync on field read).
 *   - there is NO CPU fallback: every compute entry point fails with RTPBR_ERR_CUDA when no
 *     sm_100 device is usable.
 */
#ifndef RTPBR_H_
#define RTPBR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define RTPBR_API __attribute__((visibility("default")))
#else
#define RTPBR_API
#endif

#define RTPBR_VERSION 2
#define RTPBR_MAX_OBJECTS 16
#define RTPBR_MAX_BOUNCES 1024

enum {
    RTPBR_OK = 0,
    RTPBR_ERR_ARG = -1,     /* bad argument */
    RTPBR_ERR_CUDA = -2,    /* CUDA runtime / no device */
    RTPBR_ERR_STATE = -3,   /* call order (e.g. pathtrace before set_scene) */
    RTPBR_ERR_NCCL = -4,    /* NCCL not loadable or collective failed */
    RTPBR_ERR_UNSUPPORTED = -5
};

/* src/sdf.py:12-18 `class SHAPE(IntEnum)`; 6 = neural bunny (bunny_sdf_glass.py:149-203) */
enum { RTPBR_SHAPE_NONE = 0, RTPBR_SHAPE_SPHERE = 1, RTPBR_SHAPE_BOX = 2, RTPBR_SHAPE_CYLINDER = 3,
       RTPBR_SHAPE_CONE = 4, RTPBR_SHAPE_PLANE = 5, RTPBR_SHAPE_BUNNY = 6 };

/* integrator family */
enum { RTPBR_FAMILY_A = 0,   /* examples/cornell_box/cornell_box_shortest.py:81-100 (diffuse only)      */
       RTPBR_FAMILY_B = 1,   /* examples/.../cornell_box.py:296-319, bunny_sdf_glass.py:343-366, tokyo_ibl.py:339-362 */
       RTPBR_FAMILY_C = 2 }; /* src/pathtracer.py:16-103 (one bounce per launch, ray_buffer state)      */

enum { RTPBR_MARCH_PLAIN = 0,      /* cornell_box_shortest.py:63-72, cornell_box.py:213-223 */
       RTPBR_MARCH_ENHANCED = 1,   /* cornell_box_v3/pathtracer.py:52-78, bunny_sdf_glass.py:248-267, tokyo_ibl.py:246-265 */
       RTPBR_MARCH_SRC = 2 };      /* src/scene.py:59-84 */

enum { RTPBR_SKY_BLACK = 0, RTPBR_SKY_ENVMAP = 1, RTPBR_SKY_GRADIENT = 2 };

enum { RTPBR_KERNEL_PERSISTENT = 0,  /* persistent wavefront kernel with per-warp path pools (default) */
       RTPBR_KERNEL_SIMPLE = 1 };    /* one thread per pixel, run to completion (validation)          */

/* Flattened src/dataclass.py:13-35 (Material, Transform, SDFObject). `Transform.matrix`
 * is derived by rtpbr_set_scene (replaces kernel update_all_transform, src/scene.py:99-109). */
typedef struct RtpbrObject {
    int32_t type;
    float position[3];
    float rotation[3];      /* Euler angles, degrees */
    float scale[3];
    float albedo[3];
    float emission[3];      /* multiplicative; (1,1,1) for non-emitters (src/scene.py:14) */
    float roughness, metallic, transmission, ior;
} RtpbrObject;

/* src/dataclass.py:38-46 `Camera`; passed by value to the example kernels
 * (cornell_box_shortest.py:103, bunny_sdf_glass.py:394-398). */
typedef struct RtpbrCamera {
    float lookfrom[3], lookat[3], vup[3];
    float vfov;             /* degrees */
    float aspect, aperture, focus;
} RtpbrCamera;

/* Every constant of src/config.py:7-28 plus the variant selectors of SURVEY.md 8(a). */
typedef struct RtpbrConfig {
    int32_t width, height;        /* image_resolution, src/config.py:7 */
    int32_t family;               /* RTPBR_FAMILY_* */
    int32_t max_bounces;          /* MAX_RAYTRACE, src/config.py:26; range(3) in shortest:83 */
    int32_t max_steps;            /* MAX_RAYMARCH, src/config.py:25; range(256) in shortest:66 */
    int32_t marcher;              /* RTPBR_MARCH_* */
    float t_start;                /* shortest:65 0.0005; MIN_DIS cornell_box.py:14 */
    float hit_eps;                /* shortest:70 0.00001; PRECISION; PIXEL_RADIUS */
    float t_far;                  /* MAX_DIS */
    float relax_w0;               /* enhanced sphere tracing start w (1.6 / 0.5) */
    int32_t relax_guard;          /* 1: fallback requires w > 1 (src/scene.py:68) */
    int32_t relax_reset;          /* 0: w = relax_w_reset; 1: w = 0.5 + 0.5 w (tokyo_ibl.py:256) */
    float relax_w_reset;
    float normal_h;               /* tetrahedron step 0.5773*0.005 (src/sdf.py:80) or PRECISION */
    float box_round;              /* 0.03 src/sdf.py:34; 0 shortest:45; 0.01 v2/v3 */
    float light_quality;          /* RR p_i = 1 - exp(-i/light_quality), shortest:84 */
    int32_t bsdf;                 /* 0 diffuse (A), 1 PBR examples (B), 2 PBR src/pbr.py (C) */
    int32_t f0_variant;           /* 0: 2f^2 (cornell_box.py:275); 1: (2f)^2 (tokyo_ibl.py:318, src/pbr.py:44-45) */
    float visibility_min;         /* VISIBILITY.x, src/config.py:16 */
    float visibility_max;         /* VISIBILITY.y */
    int32_t sky;                  /* RTPBR_SKY_* */
    float sky_scale;
    uint32_t seed;                /* Philox key */
    float min_dis;                /* MIN_DIS, src/config.py:22 (family C surface offset) */
    float pixel_radius;           /* PIXEL_RADIUS, src/config.py:20 */
    float quality_per_sample;     /* QUALITY_PER_SAMPLE, src/config.py:11 */
    int32_t black_background;     /* BLACK_BACKGROUND, src/config.py:13 */
    int32_t nearest_seed;         /* 0: first object seeds the minimum (shortest:48, cornell_box.py:198);
                                     1: MAX_DIS seeds it (src/scene.py:46, tokyo_ibl.py:222) */
    int32_t normal_mode;          /* 0: world-space tetrahedron offsets (shortest:55-61);
                                     1: object-space, one transform (src/sdf.py:77-87) */
    int32_t samples_per_pixel;    /* SAMPLES_PER_PIXEL per launch, src/config.py:10 (family C) */
    int32_t adaptive_sampling;    /* ADAPTIVE_SAMPLING, src/config.py:14 (family C): pathtrace() skips pixels whose running
                                     mean of tone-mapped change, diff_pixels, is <= noise_threshold (src/pathtracer.py:97-101) */
    float noise_threshold;        /* NOISE_THRESHOLD, src/config.py:17 */
    int32_t inner_spp;            /* > 0 (family B): kernel render() of examples/bunny/bunny_sdf.py:397-427 / bunny_sdf_v2.py:397-432 --
                                     SAMPLE_PER_PIXEL samples per launch in an in-kernel loop that shares ONE ti.random stream per pixel;
                                     every launch overwrites image_buffer with their sum.  Pixel-granular: runs on the one-thread-per-pixel
                                     kernel (RTPBR_KERNEL_SIMPLE is implied) */
    int32_t primary_miss;         /* a camera ray that misses everything: 0 colour *= sky (bunny_sdf_glass.py:355), 1 white
                                     (bunny_sdf_v2.py:355-358), 2 colour *= 0, then *= sky (bunny_sdf.py:352-353) */
    int32_t bunny_bob;            /* neural bunny animation: 1 rotation + bob 0.1 sin t (bunny_sdf_glass.py:213-216, v2), 0 rotation only
                                     (bunny_sdf.py:214) */
    int32_t kernel;               /* RTPBR_KERNEL_* */
    int32_t count_work;           /* 1: count scene evals / rays / lane occupancy (slower) */
} RtpbrConfig;

typedef struct RtpbrContext RtpbrContext;

/* which-buffer selectors for rtpbr_download / rtpbr_upload */
enum { RTPBR_BUF_IMAGE_BUFFER = 0,  /* image_buffer  vec4 f32 (W,H,4), src/fileds.py:8 */
       RTPBR_BUF_IMAGE_PIXELS = 1,  /* image_pixels  vec3 f32 (W,H,3), src/fileds.py:9 */
       RTPBR_BUF_RAY_BUFFER = 2,    /* ray_buffer    AOS Ray, 10 x 4 bytes (W,H,10), src/fileds.py:7 */
       RTPBR_BUF_DIFF_BUFFER = 3,   /* diff_buffer   vec2 f32 (W,H,2), src/fileds.py:21 (adaptive sampling only) */
       RTPBR_BUF_DIFF_PIXELS = 4,   /* diff_pixels   f32 (W,H,1), src/fileds.py:22 (adaptive sampling only) */
       RTPBR_BUF_DENOISE_PIXELS = 5 }; /* denoise_pixels vec3 f32 (W,H,3), examples/denoise/denoise_test_1.py:53 (after rtpbr_denoise) */

/* counters written by rtpbr_get_counters (valid when count_work = 1) */
enum { RTPBR_CNT_SCENE_EVALS = 0, RTPBR_CNT_RAYS = 1, RTPBR_CNT_NORMALS = 2, RTPBR_CNT_SAMPLES = 3,
       RTPBR_CNT_MARCH_ITERS = 4,      /* warp-level march iterations x 32 (issued lane slots) */
       RTPBR_CNT_MARCH_ACTIVE = 5,     /* lanes actually marching in those iterations         */
       RTPBR_CNT_RESOLVE_ROUNDS = 6, RTPBR_CNT_LAUNCHES = 7,
       RTPBR_CNT_RESOLVED_SLOTS = 8,   /* slots handled by resolve rounds (x / rounds / 32 = resolve occupancy) */
       RTPBR_CNT_MLP_EVALS = 9,        /* evaluations of the neural bunny's MLP (points inside its unit sphere) */
       RTPBR_CNT_COUNT = 10 };

/* replaces `ti.init(arch=ti.gpu, ...)` (src/config.py:5) + field allocation (src/fileds.py:7-13) */
RTPBR_API int rtpbr_create(const RtpbrConfig* cfg, int device, RtpbrContext** out);
RTPBR_API int rtpbr_destroy(RtpbrContext* ctx);

/* replaces `objects[i] = OBJECTS[i]` (src/scene.py:38-41) + build_scene() (src/scene.py:112-113) */
RTPBR_API int rtpbr_set_scene(RtpbrContext* ctx, const RtpbrObject* objects, int n);
/* replaces smooth.position/lookat/up + camera_* scalar fields (src/camera.py:115-129) and the
 * by-value camera arguments of the example kernels (cornell_box_shortest.py:103) */
RTPBR_API int rtpbr_set_camera(RtpbrContext* ctx, const RtpbrCamera* cam);
/* replaces Image.__init__/process (src/ibl.py:12-23): rgb is (w, h, 3) f32, j fastest, already
 * exposure/gamma adjusted on the host */
RTPBR_API int rtpbr_set_envmap(RtpbrContext* ctx, const float* rgb, int w, int h);
/* replaces `u_frame[None] = frame` (bunny_sdf_glass.py:409) */
RTPBR_API int rtpbr_set_frame(RtpbrContext* ctx, int frame);
/* sample-index base of the next rtpbr_pathtrace (Philox counter word 1) */
RTPBR_API int rtpbr_set_sample_base(RtpbrContext* ctx, uint32_t sample_base);
/* column-band sharding: this context renders columns i with (i / band) % nranks == rank */
RTPBR_API int rtpbr_set_shard(RtpbrContext* ctx, int rank, int nranks, int band);

/* replaces kernel refresh() (src/renderer.py:12-22, bunny_sdf_glass.py:418-421) */
RTPBR_API int rtpbr_refresh(RtpbrContext* ctx);
/* replaces kernel pathtrace() (src/pathtracer.py:94-103) x spp launches, kernel render()
 * of cornell_box_shortest.py:102-122 x spp launches, kernel sample() of
 * bunny_sdf_glass.py:393-416 / tokyo_ibl.py:403-423 x spp launches.  Asynchronous. */
RTPBR_API int rtpbr_pathtrace(RtpbrContext* ctx, int spp);
/* replaces kernel post_process() (src/postprocessor.py:24-43) / the tonemap tail of render()
 * (cornell_box_shortest.py:124-129).  mode: 0 family A, 1 family B, 2 family C, 3 v3. */
RTPBR_API int rtpbr_post_process(RtpbrContext* ctx, int mode, float exposure, double gamma);  /* gamma in binary64: the
   reference folds 1.0 / camera_gamma in Python before casting to f32 (src/postprocessor.py:32).
   Known divergences of this (f)-row pass, unlike the accumulation buffer, which is bit-exact: modes 0, 1 and 3 raise to the
   gamma with the device's powf (within 2-3e-5 of the reference-source pixels, tests/test_golden.py; mode 2, the src/ order,
   is bit-exact under the fp32 contract), and the NaN pixels cornell_box.py's tonemap produces for 0/0 are written as 0. */

/* replaces kernel denoise(image_pixels, denoise_pixels, threshold) of examples/denoise/denoise_test_1.py:86-118 (the
 * temporal blend + bright-neighbour fill after shadertoy 7tKGzD) as an optional pass after rtpbr_post_process.  The
 * reference filters denoise_pixels in place while neighbouring threads read it, so its own result depends on the thread
 * schedule; this entry point is the deterministic form: neighbours are read from the previous denoise_pixels, the result
 * becomes the new denoise_pixels (double-buffered; zero before the first call, like a fresh Taichi field). */
RTPBR_API int rtpbr_denoise(RtpbrContext* ctx, float threshold);

/* replaces field.to_numpy() / canvas.set_image(field) / ti.tools.imwrite(field) reads */
RTPBR_API int rtpbr_download(RtpbrContext* ctx, int which, void* host, size_t bytes);
/* resume from a saved accumulation buffer (reference has no checkpointing) */
RTPBR_API int rtpbr_upload(RtpbrContext* ctx, int which, const void* host, size_t bytes);
RTPBR_API int rtpbr_sync(RtpbrContext* ctx);
/* page-locked host memory for rtpbr_download / rtpbr_upload / rtpbr_set_envmap buffers (cudaHostAlloc / cudaFreeHost) */
RTPBR_API int rtpbr_alloc_host(size_t bytes, void** out);
RTPBR_API int rtpbr_free_host(void* ptr);
/* benchmark hygiene: evict L2 by writing a 256 MiB scratch buffer on the launch stream */
RTPBR_API int rtpbr_flush_l2(RtpbrContext* ctx);

/* CUDA-event timing on the context's launch stream (bench.py) */
RTPBR_API int rtpbr_timer_start(RtpbrContext* ctx);
RTPBR_API int rtpbr_timer_stop(RtpbrContext* ctx, float* elapsed_ms);   /* synchronises */
/* ms spent in the path-tracing kernels only since the last call (sum of per-launch events) */
RTPBR_API int rtpbr_kernel_time(RtpbrContext* ctx, float* kernel_ms, int* launches);

RTPBR_API int rtpbr_get_counters(RtpbrContext* ctx, uint64_t out[RTPBR_CNT_COUNT]);
RTPBR_API int rtpbr_device_info(RtpbrContext* ctx, int* sm_count, int* cc_major, int* cc_minor, int* blocks_per_sm);

/* Scene-specialised kernels.  Like Taichi, which JIT-compiles the reference's kernels per scene
 * (`ti.static` object loops, src/scene.py:48-51), rtpbr_pathtrace compiles the march loop for the
 * current scene with NVRTC (object constants as immediates; bit-identical results).  Enabled by
 * default (RTPBR_JIT=0 or rtpbr_set_jit(ctx, 0) selects the ahead-of-time kernels; so does any
 * NVRTC failure).  rtpbr_jit_status returns 1 when the specialised kernel is in use and copies a
 * one-line description to buf.  rtpbr_jit_generate / rtpbr_jit_compile_check need no GPU:
 * they return the generated source (its length) / run NVRTC on it (0 on success). */
RTPBR_API int rtpbr_set_jit(RtpbrContext* ctx, int enable);
RTPBR_API int rtpbr_jit_status(RtpbrContext* ctx, char* buf, size_t cap);
RTPBR_API long long rtpbr_jit_generate(const RtpbrConfig* cfg, const RtpbrObject* objects, int n, char* buf, size_t cap);
RTPBR_API int rtpbr_jit_compile_check(const RtpbrConfig* cfg, const RtpbrObject* objects, int n, char* log, size_t cap);

/* multi-GPU: one context per process/GPU; NCCL only at tonemap time */
RTPBR_API int rtpbr_nccl_unique_id(void* id128);                       /* 128 bytes out (rank 0) */
RTPBR_API int rtpbr_nccl_init(RtpbrContext* ctx, const void* id128, int rank, int nranks);
RTPBR_API int rtpbr_reduce_tiles(RtpbrContext* ctx, int root);         /* root < 0: all-reduce */

/* After rtpbr_reduce_tiles the root's (all ranks', for root < 0) image_buffer holds the SUM over ranks: tracing on would
 * count the other ranks' samples twice at the next reduce, so rtpbr_pathtrace fails with RTPBR_ERR_STATE until
 * rtpbr_refresh.  A context sharded with rtpbr_set_shard(nranks > 1) that has no communicator fails too. */

/* Single-process multi-GPU (SURVEY.md 8(b) `pt_create_multi`): n contexts, one per device, column bands of `band`
 * columns interleaved across them (rank = (column / band) mod n), one NCCL communicator per GPU created by the calling
 * thread.  The setters broadcast; rtpbr_multi_pathtrace queues one asynchronous launch per GPU; rtpbr_multi_post_process
 * sums the per-GPU sample sums onto rank 0 (the only collective: NCCL at tonemap time) and tone-maps there;
 * rtpbr_multi_download reads rank 0.  rtpbr_multi_context lends the per-GPU context (timers, counters, jit status). */
typedef struct RtpbrMulti RtpbrMulti;
RTPBR_API int rtpbr_multi_create(const RtpbrConfig* cfg, const int* devices, int n, int band, RtpbrMulti** out);
RTPBR_API int rtpbr_multi_destroy(RtpbrMulti* m);
RTPBR_API int rtpbr_multi_count(RtpbrMulti* m);
RTPBR_API RtpbrContext* rtpbr_multi_context(RtpbrMulti* m, int rank);
RTPBR_API int rtpbr_multi_set_scene(RtpbrMulti* m, const RtpbrObject* objects, int n);
RTPBR_API int rtpbr_multi_set_camera(RtpbrMulti* m, const RtpbrCamera* cam);
RTPBR_API int rtpbr_multi_set_envmap(RtpbrMulti* m, const float* rgb, int w, int h);
RTPBR_API int rtpbr_multi_set_frame(RtpbrMulti* m, int frame);
RTPBR_API int rtpbr_multi_set_sample_base(RtpbrMulti* m, uint32_t sample_base);
RTPBR_API int rtpbr_multi_refresh(RtpbrMulti* m);
RTPBR_API int rtpbr_multi_pathtrace(RtpbrMulti* m, int spp);
RTPBR_API int rtpbr_multi_reduce(RtpbrMulti* m, int root);
RTPBR_API int rtpbr_multi_post_process(RtpbrMulti* m, int mode, float exposure, double gamma);
RTPBR_API int rtpbr_multi_download(RtpbrMulti* m, int which, void* host, size_t bytes);
RTPBR_API int rtpbr_multi_sync(RtpbrMulti* m);

/* device pointer of a buffer (for zero-copy interop, e.g. __cuda_array_interface__) */
RTPBR_API int rtpbr_device_ptr(RtpbrContext* ctx, int which, uint64_t* ptr);

RTPBR_API const char* rtpbr_last_error(void);
RTPBR_API int rtpbr_version(void);
RTPBR_API int rtpbr_device_count(void);      /* usable CUDA devices (0 without a driver) */
RTPBR_API int rtpbr_sizeof_config(void);
RTPBR_API int rtpbr_sizeof_object(void);
RTPBR_API int rtpbr_sizeof_camera(void);

#ifdef __cplusplus
}
#endif
#endif /* RTPBR_H_ */
