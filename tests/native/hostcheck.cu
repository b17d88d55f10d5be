// hostcheck.cu -- TEST HARNESS ONLY.  Runs the product's __host__ __device__ path-tracing
// building blocks (raytracingpbr_b200/csrc/rt_integrator.cuh) on the CPU so that their fp32
// operation sequence can be compared bit-for-bit with the independent C oracle and the golden
// fixtures on a machine without a GPU.  Not linked into librtpbr.so and not reachable from the
// product API.
#include <cstring>
#include <vector>

#include "../../include/rtpbr.h"
#include "../../raytracingpbr_b200/csrc/host_setup.h"
// tests/test_jit.py compiles this file a second time per scene with the scene-specialised translation unit that
// jit_codegen.h generates: HC_JIT_PREAMBLE = its #defines, HC_JIT_BODY = its functions.  trace_sample() then marches
// the way the specialised pool kernel does (fast region, t_stop), on the CPU.
#if defined(HC_JIT_PREAMBLE)
#include HC_JIT_PREAMBLE
#endif
#include "../../raytracingpbr_b200/csrc/rt_integrator.cuh"
#if defined(HC_JIT_BODY)
namespace rt {
#include HC_JIT_BODY
}
#endif

using namespace rt;

#define HC_API extern "C" __attribute__((visibility("default")))

#if defined(RT_JIT_SCENE)
template <class VAR>
static int jit_march_check_t(const KParams& P, const float* rays, int nrays, int* out)
{
    int bad = 0;
    for (int k = 0; k < nrays; ++k) {
        MarchState a, b;
        memset(&a, 0, sizeof(a));
        a.ro = V3(rays[6 * k], rays[6 * k + 1], rays[6 * k + 2]);
        a.rd = V3(rays[6 * k + 3], rays[6 * k + 4], rays[6 * k + 5]);
        if (ray_is_irregular(a)) continue;
        march_begin<VAR>(P, a);
        b = a;
        int sa;
        do { sa = march_step<VAR, true>(P, a); } while (sa == MARCH_CONTINUE);      // the generic code, to the end
        const int sb = march_to_end_jit<VAR>(P, b);
        bool same = sa == sb;
        if (same && sa == MARCH_HIT) {
            const vec3 pa = hit_position<VAR>(a), pb = hit_position<VAR>(b);
            same = memcmp(&pa, &pb, sizeof(pa)) == 0;
            if (VAR::MARCHER == MARCH_SRC) same = same && a.idx == b.idx;
        }
        if (!same) ++bad;
        out[0] += sa == MARCH_HIT;
        out[1] += a.steps;
        out[2] += b.steps;
    }
    return bad;
}
// rays: (origin, direction) x nrays.  Returns the number of rays whose specialised march ends differently from the generic
// one (status; hit position bits); out = { hits, generic steps, specialised steps }.
HC_API int hostcheck_jit_march(const RtpbrConfig* cfg, const RtpbrObject* objs, int n, int frame, const float* rays, int nrays, int* out)
{
    KParams P;
    memset(&P, 0, sizeof(P));
    fill_config(P, *cfg);
    fill_objects(P, objs, n);
    fill_frame(P, frame);
    out[0] = out[1] = out[2] = 0;
    bool bunny = false;
    for (int k = 0; k < n; ++k) bunny = bunny || objs[k].type == RTPBR_SHAPE_BUNNY;
    if (cfg->family == RTPBR_FAMILY_A) return jit_march_check_t<Variant<FAMILY_A, 0, SHAPESET_BOX, MARCH_PLAIN, false>>(P, rays, nrays, out);
    if (cfg->family == RTPBR_FAMILY_B && cfg->marcher == RTPBR_MARCH_PLAIN) return jit_march_check_t<Variant<FAMILY_B, 0, SHAPESET_ANALYTIC, MARCH_PLAIN, false>>(P, rays, nrays, out);
    if (cfg->family == RTPBR_FAMILY_B && bunny) return jit_march_check_t<Variant<FAMILY_B, 0, SHAPESET_BUNNY, MARCH_ENHANCED, false>>(P, rays, nrays, out);
    if (cfg->family == RTPBR_FAMILY_B) return jit_march_check_t<Variant<FAMILY_B, 0, SHAPESET_ANALYTIC, MARCH_ENHANCED, false>>(P, rays, nrays, out);
    return jit_march_check_t<Variant<FAMILY_C, 0, SHAPESET_ANALYTIC, MARCH_SRC, false>>(P, rays, nrays, out);
}
#if defined(RT_JIT_STATS)
HC_API unsigned long long* hostcheck_jit_stats(void) { return g_jit_stats; }
#endif
// points x npts: wherever jit_nearest_fast() says ok its result must have the bits of jit_nearest_dist().  Returns the
// number of violations, -1 when the scene has no fast region; *n_ok = points inside the region.
HC_API int hostcheck_jit_fast(const RtpbrConfig* cfg, const RtpbrObject* objs, int n, int frame, const float* pts, int npts, int* n_ok)
{
    *n_ok = 0;
#if defined(RT_JIT_FAST)
    KParams P;
    memset(&P, 0, sizeof(P));
    fill_config(P, *cfg);
    fill_objects(P, objs, n);
    fill_frame(P, frame);
    int bad = 0;
    for (int k = 0; k < npts; ++k) {
        const vec3 p = V3(pts[3 * k], pts[3 * k + 1], pts[3 * k + 2]);
        bool ok;
        const float f = jit_nearest_fast(P, p, ok), d = jit_nearest_dist(P, p);
        bool ok2;
        int i_fast = -1, i_full = -2;
        const float f2 = jit_nearest_fast_idx(P, p, ok2, i_fast), d2 = jit_nearest(P, p, i_full);
        if (ok != ok2) ++bad;
        if (!ok) continue;
        ++*n_ok;
        if (memcmp(&f, &d, 4) != 0) ++bad;
        if (memcmp(&f2, &d2, 4) != 0 || i_fast != i_full) ++bad;      // the argmin form: bits and index of jit_nearest()
    }
    return bad;
#else
    return -1;
#endif
}
#endif

template <class VAR>
static void run(const KParams& P, float4* buf)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (long long w = 0; w < (long long)P.total_work; ++w) {
        int i, j;
        if (!work_to_pixel(P, (uint32_t)w, i, j)) continue;
        const uint32_t pixel = (uint32_t)(i * P.height + j);
        if (VAR::FAMILY == FAMILY_C && P.adaptive && !(P.diff_pixels[pixel] > P.noise_threshold)) continue;
        float4 acc = buf[pixel];
        if (VAR::FAMILY == FAMILY_C) {
            trace_pixel_c<VAR>(P, pixel, i, j, acc, nullptr);
        } else if (VAR::FAMILY == FAMILY_B && P.inner_spp > 0) {
            acc = trace_pixel_inner<VAR>(P, pixel, i, j, P.sample_base + (uint32_t)(P.spp - 1), nullptr);
        } else {
            for (int s = 0; s < P.spp; ++s) {
                vec3 c = trace_sample<VAR>(P, pixel, i, j, P.sample_base + (uint32_t)s, nullptr);
                acc.x += c.x; acc.y += c.y; acc.z += c.z; acc.w += 1.0f;
            }
        }
        buf[pixel] = acc;
    }
}

HC_API int hostcheck_pathtrace_ex(const RtpbrConfig* cfg, const RtpbrCamera* cam, const RtpbrObject* objs, int n,
                                  float* image_buffer, float* ray_buffer, const float* env, int env_w, int env_h, int frame,
                                  int spp, uint32_t sample_base, int rank, int nranks, int band, const float* diff_pixels)
{
    KParams P;
    memset(&P, 0, sizeof(P));
    fill_config(P, *cfg);
    fill_shard(P, rank, nranks, band);
    fill_objects(P, objs, n);
    fill_camera(P, *cfg, *cam);
    fill_frame(P, frame);
    std::vector<float> rr = rr_table(*cfg);
    P.rr_prob = rr.data();
    P.spp = spp;
    P.sample_base = sample_base;
    P.env = env; P.env_w = env_w; P.env_h = env_h;
    P.ray_buffer = ray_buffer;
    P.diff_pixels = diff_pixels;
    if (!diff_pixels) P.adaptive = 0;
    float4* buf = reinterpret_cast<float4*>(image_buffer);
    bool bunny = false;
    for (int k = 0; k < n; ++k) bunny = bunny || objs[k].type == RTPBR_SHAPE_BUNNY;
    if (cfg->family == RTPBR_FAMILY_A) run<Variant<FAMILY_A, 0, SHAPESET_BOX, MARCH_PLAIN, false>>(P, buf);
    else if (cfg->family == RTPBR_FAMILY_B && cfg->marcher == RTPBR_MARCH_PLAIN) run<Variant<FAMILY_B, 0, SHAPESET_ANALYTIC, MARCH_PLAIN, false>>(P, buf);
    else if (cfg->family == RTPBR_FAMILY_B && bunny) run<Variant<FAMILY_B, 0, SHAPESET_BUNNY, MARCH_ENHANCED, false>>(P, buf);
    else if (cfg->family == RTPBR_FAMILY_B) run<Variant<FAMILY_B, 0, SHAPESET_ANALYTIC, MARCH_ENHANCED, false>>(P, buf);
    else if (cfg->family == RTPBR_FAMILY_C && ray_buffer) run<Variant<FAMILY_C, 0, SHAPESET_ANALYTIC, MARCH_SRC, false>>(P, buf);
    else return RTPBR_ERR_UNSUPPORTED;
    return 0;
}

HC_API int hostcheck_pathtrace(const RtpbrConfig* cfg, const RtpbrCamera* cam, const RtpbrObject* objs, int n, float* image_buffer,
                               int spp, uint32_t sample_base, int rank, int nranks, int band)
{
    return hostcheck_pathtrace_ex(cfg, cam, objs, n, image_buffer, nullptr, nullptr, 0, 0, 0, spp, sample_base, rank, nranks, band,
                                  nullptr);
}

HC_API void hostcheck_sincos(float x, float* s, float* c) { sincos_rt(x, *s, *c); }
HC_API float hostcheck_atan2(float y, float x) { return atan2_rt(y, x); }
HC_API float hostcheck_asin(float x) { return asin_rt(x); }
HC_API void hostcheck_euler(const float rot_deg[3], float out9[9]) { euler_matrix_deg(rot_deg, out9); }
// three draws starting at draw n of stream (pixel, launch): out[0..2] by three rng_next calls, out[3..5] by
// rng_next2_ahead followed by rng_next (the order on_hit + begin_bounce use); out[6] = draws consumed by each
HC_API void hostcheck_rng3(uint32_t seed, uint32_t pixel, uint32_t launch, uint32_t n, float out[7])
{
    KParams P;
    memset(&P, 0, sizeof(P));
    P.seed = seed;
    Rng a = rng_make(pixel, launch, n), b = rng_make(pixel, launch, n);
    out[0] = rng_next(P, a); out[1] = rng_next(P, a); out[2] = rng_next(P, a);
    rng_next2_ahead(P, b, out[3], out[4]);
    out[5] = rng_next(P, b);
    out[6] = (a.n == n + 3u && b.n == n + 3u) ? 3.0f : -1.0f;
}
HC_API float hostcheck_sd_bunny(const float p[3]) { return sd_bunny(V3(p[0], p[1], p[2])); }
