// rt_params.h -- kernel parameter block (lives in the constant bank via __grid_constant__).
#pragma once
#if !defined(__CUDACC_RTC__)
#include <cstdint>
#include <vector_types.h>
#endif

namespace rt {

constexpr int kMaxObjects = 16;

// Hot geometry of one SDF object: 16 words, 16-byte aligned so it loads as 4 x 128 bit.
// `m` is Transform.matrix (src/dataclass.py:28), derived on the host by rtpbr_set_scene.
struct alignas(16) DevGeom {
    float px, py, pz;
    float m[9];
    float sx, sy, sz;
    int32_t type;
};

// src/dataclass.py:13-20 Material: 12 words.
struct DevMaterial {
    float albedo[3];
    float emission[3];
    float roughness, metallic, transmission, ior;
    float pad0, pad1;
};

struct DevCamera {
    float origin[3];      // lookfrom
    float llc[3];         // lower_left_corner
    float horizontal[3];
    float vertical[3];
    float x[3], y[3];     // lens basis
    float lens_radius;    // aperture * 0.5
    float inv_w, inv_h;   // SCREEN_PIXEL_SIZE (families B/C: coord * SCREEN_PIXEL_SIZE)
    float fw, fh;         // float(W), float(H) (family A: coord / vec2(resolution))
};

struct KParams {
    int32_t width, height;
    int32_t spp;             // families A/B: samples per pixel in this launch; family C: reference launches
    uint32_t sample_base;    // launch index of the first sample (Philox counter word 1)
    uint32_t seed;
    int32_t max_bounces, max_steps;
    float t_start, hit_eps, t_far;
    float relax_w0, relax_w_reset;
    int32_t relax_guard, relax_reset;
    float normal_h, box_round;
    float visibility_min, visibility_max;
    int32_t bsdf, f0_variant;
    int32_t sky;
    float sky_scale;
    float min_dis, pixel_radius, quality_per_sample;
    float inv_max_bounces;   // float(1.0 / MAX_RAYTRACE), src/pathtracer.py:68
    int32_t black_background;
    int32_t nearest_seed, normal_mode, samples_per_pixel;
    int32_t adaptive;        // ADAPTIVE_SAMPLING (family C): skip pixels with diff_pixels <= noise_threshold
    float noise_threshold;
    int32_t frame;
    float anim_m[9];         // bunny programmatic animation: angle(vec3(0, 0, t)), bunny_sdf_glass.py:214
    float anim_bob;          // 0.1 * sin(t), :215
    // shard: this context renders global columns i with (i / band) % nranks == rank
    int32_t rank, nranks, band;
    int32_t local_cols;      // number of columns owned by this rank
    uint32_t total_work;     // work items (pixels incl. tile padding) on this rank
    int32_t tiles_per_col;   // ceil(H / 8)
    int32_t nobj;
    int32_t resolve_min;     // pool kernel: resolve when this many slots are pending (or lanes would idle)
    int32_t inner_spp;       // bunny_sdf.py / bunny_sdf_v2.py: in-kernel sample loop sharing one RNG stream (simple kernel only)
    int32_t primary_miss;    // 0 x sky, 1 white, 2 x 0 x sky for camera rays that miss
    int32_t bunny_bob;       // 1: the bunny's animation includes the bob
    int32_t count_mlp;       // count_work builds: count evaluations of the neural bunny's MLP (RTPBR_CNT_MLP_EVALS)
    DevCamera cam;
    DevGeom geom[kMaxObjects];
    DevMaterial mat[kMaxObjects];
    // device pointers
    float4* image_buffer;           // (W,H) vec4, j fastest
    float* ray_buffer;              // family C: (W,H,10) AOS Ray, src/fileds.py:7
    const float* rr_prob;           // [max_bounces] Russian-roulette table (families A/B)
    const float* env;               // (env_w, env_h, 3) or nullptr
    int32_t env_w, env_h;
    const float* diff_pixels;       // family C adaptive sampling: (W,H) f32, src/fileds.py:22
    float4* scratch;                // families A/B: per-sample radiance, [total_work][spp] (pool kernel -> k_fold_samples)
    unsigned long long* work_counter;   // persistent-kernel work queue head
    unsigned long long* counters;   // RTPBR_CNT_* (count_work builds)
};

}  // namespace rt
