// kernels.cu -- sm_100a kernels of the path-tracing hot path.
//
//   k_pathtrace_pool    persistent wavefront kernel with a per-warp path pool (the product path)
//   k_pathtrace_simple  one thread per pixel, run-to-completion (validation / baseline)
//   k_refresh_depth     family C part of refresh(): ray_buffer[i, j].depth = 0
//   k_post_process      tonemap (src/postprocessor.py:24-38 and the example variants)
//
// Compiled with -fmad=false: the fp32 contract (rt_math.cuh) allows only explicit fmaf().
#include <cuda_runtime.h>

#include "kernels.h"
#include "pool_kernel.cuh"

namespace rt {

// ------------------------------------------------------------------------------------------
// Simple kernel: one thread per work item, all samples, every path run to completion.
// ------------------------------------------------------------------------------------------
template <class VAR>
__global__ void __launch_bounds__(kSimpleBlock) k_pathtrace_simple(const __grid_constant__ KParams P)
{
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    int i, j;
    if (w >= P.total_work || !work_to_pixel(P, w, i, j)) return;
    const uint32_t pixel = (uint32_t)(i * P.height + j);
    if (VAR::FAMILY == FAMILY_C && P.adaptive && !(P.diff_pixels[pixel] > P.noise_threshold)) return;   // src/pathtracer.py:97-101
    float4 acc = P.image_buffer[pixel];
    WorkCounters cnt = { 0, 0, 0, 0 };
    if (VAR::FAMILY == FAMILY_C) {
        trace_pixel_c<VAR>(P, pixel, i, j, acc, VAR::COUNT ? &cnt : nullptr);
    } else if (VAR::FAMILY == FAMILY_B && P.inner_spp > 0) {
        // P.spp launches of render(); every launch overwrites the buffer, so the last one is what remains
        acc = trace_pixel_inner<VAR>(P, pixel, i, j, P.sample_base + (uint32_t)(P.spp - 1), VAR::COUNT ? &cnt : nullptr);
    } else {
        for (int s = 0; s < P.spp; ++s) {
            vec3 c = trace_sample<VAR>(P, pixel, i, j, P.sample_base + (uint32_t)s, VAR::COUNT ? &cnt : nullptr);
            acc.x += c.x; acc.y += c.y; acc.z += c.z; acc.w += 1.0f;
        }
    }
    P.image_buffer[pixel] = acc;
    if (VAR::COUNT) {
        atomicAdd(&P.counters[0], cnt.evals);
        atomicAdd(&P.counters[1], cnt.rays);
        atomicAdd(&P.counters[2], cnt.normals);
        atomicAdd(&P.counters[3], cnt.samples);
    }
}

// ------------------------------------------------------------------------------------------
// k_fold_samples: image_buffer[pixel] += scratch[item * spp + s] for s = 0 .. spp-1, IN ORDER
// (shortest:121 `buffer += vec4(ray.color, 1.0)` once per launch).  HBM-bound: 16 B per sample
// read + 32 B per pixel; one thread per pixel item, each thread streams its own contiguous run.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fold_samples(const __grid_constant__ KParams P)
{
    const uint32_t item = blockIdx.x * blockDim.x + threadIdx.x;
    int i, j;
    if (item >= P.total_work || !work_to_pixel(P, item, i, j)) return;
    const uint32_t pixel = (uint32_t)(i * P.height + j);
    float4 acc = P.image_buffer[pixel];
    const float4* src = P.scratch + (size_t)item * (size_t)P.spp;
    for (int s = 0; s < P.spp; ++s) {
        const float4 c = __ldcs(src + s);
        acc.x += c.x; acc.y += c.y; acc.z += c.z; acc.w += c.w;
    }
    P.image_buffer[pixel] = acc;
}

// refresh() of src/renderer.py:12-22: the accumulators are cleared by a memset; the ray buffer keeps
// origin / direction / colour and only has `depth` reset (the reference does not reset the colour).
__global__ void __launch_bounds__(256) k_refresh_depth(float* ray_buffer, int n)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) reinterpret_cast<int*>(ray_buffer)[(size_t)p * 10 + 9] = 0;
}

// ------------------------------------------------------------------------------------------
// Tonemap.  mode 0: cornell_box_shortest.py:124-129 (gamma -> ACES with its truncated
// constants -> clamp); mode 1: cornell_box.py:374-379 / tokyo_ibl.py:434-439 (exposure ->
// ACES -> gamma -> clamp); mode 2: src/postprocessor.py:34-38 (exposure -> gamma -> ACES ->
// clamp); mode 3: cornell_box_v3/postprocessor.py (exposure -> gamma -> ACES -> clamp).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float3 aces_fitted(float3 c, bool truncated)
{
    float3 v;
    if (truncated) {  // cornell_box_shortest.py:126
        v.x = 0.597190f * c.x + 0.35458f * c.y + 0.04823f * c.z;
        v.y = 0.07600f * c.x + 0.90834f * c.y + 0.01566f * c.z;
        v.z = 0.02840f * c.x + 0.13383f * c.y + 0.83777f * c.z;
        v.x = (v.x * (v.x + 0.024578f) - 0.0000905f) / (v.x * (0.983729f * v.x + 0.4329510f) + 0.238081f);
        v.y = (v.y * (v.y + 0.024578f) - 0.0000905f) / (v.y * (0.983729f * v.y + 0.4329510f) + 0.238081f);
        v.z = (v.z * (v.z + 0.024578f) - 0.0000905f) / (v.z * (0.983729f * v.z + 0.4329510f) + 0.238081f);
        float3 o;
        o.x = 1.60475f * v.x + -0.531f * v.y + -0.0736f * v.z;
        o.y = -0.102f * v.x + 1.10813f * v.y + -0.00605f * v.z;
        o.z = -0.00327f * v.x + -0.07276f * v.y + 1.07602f * v.z;
        return o;
    }
    // src/aces.py:5-30
    v.x = 0.59719f * c.x + 0.35458f * c.y + 0.04823f * c.z;
    v.y = 0.07600f * c.x + 0.90834f * c.y + 0.01566f * c.z;
    v.z = 0.02840f * c.x + 0.13383f * c.y + 0.83777f * c.z;
    v.x = (v.x * (v.x + 0.0245786f) - 0.000090537f) / (v.x * (0.983729f * v.x + 0.4329510f) + 0.238081f);
    v.y = (v.y * (v.y + 0.0245786f) - 0.000090537f) / (v.y * (0.983729f * v.y + 0.4329510f) + 0.238081f);
    v.z = (v.z * (v.z + 0.0245786f) - 0.000090537f) / (v.z * (0.983729f * v.z + 0.4329510f) + 0.238081f);
    float3 o;
    o.x = 1.60475f * v.x + -0.53108f * v.y + -0.07367f * v.z;
    o.y = -0.10208f * v.x + 1.10813f * v.y + -0.00605f * v.z;
    o.z = -0.00327f * v.x + -0.07276f * v.y + 1.07602f * v.z;
    return o;
}

__device__ __forceinline__ float3 pow3(float3 c, float e) { return make_float3(powf(c.x, e), powf(c.y, e), powf(c.z, e)); }
__device__ __forceinline__ float3 clamp01(float3 c)
{
    return make_float3(fminf(fmaxf(c.x, 0.f), 1.f), fminf(fmaxf(c.y, 0.f), 1.f), fminf(fmaxf(c.z, 0.f), 1.f));
}

__global__ void __launch_bounds__(256) k_post_process(const float4* __restrict__ image_buffer, float* __restrict__ image_pixels,
                                                      int n, int mode, float exposure, float inv_gamma)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const float4 b = image_buffer[p];
    float3 c = make_float3(b.x / b.w, b.y / b.w, b.z / b.w);  // average(), src/postprocessor.py:12-14
    if (mode == 0) {
        c = clamp01(aces_fitted(pow3(c, inv_gamma), true));
    } else if (mode == 1) {
        c = make_float3(c.x * exposure, c.y * exposure, c.z * exposure);
        c = clamp01(pow3(aces_fitted(c, false), inv_gamma));
    } else {
        c = make_float3(c.x * exposure, c.y * exposure, c.z * exposure);
        c = clamp01(aces_fitted(pow3(c, inv_gamma), false));
    }
    image_pixels[3 * p + 0] = c.x;
    image_pixels[3 * p + 1] = c.y;
    image_pixels[3 * p + 2] = c.z;
}

// kernel post_process() of the src/ package (src/postprocessor.py:24-43, src/aces.py:5-30) under the fp32
// contract -- matrix products as fmaf chains, pow in binary64 rounded once -- because with
// ADAPTIVE_SAMPLING its output feeds back into which pixels pathtrace() samples.
__device__ __forceinline__ float pow_contract(float x, float e) { return (float)pow((double)x, (double)e); }
__global__ void __launch_bounds__(256) k_post_process_src(const float4* __restrict__ image_buffer, float* __restrict__ image_pixels,
                                                          float2* __restrict__ diff_buffer, float* __restrict__ diff_pixels, int n,
                                                          float exposure, float gamma_inv)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const float Min[9] = { 0.59719f, 0.35458f, 0.04823f, 0.07600f, 0.90834f, 0.01566f, 0.02840f, 0.13383f, 0.83777f };
    const float Mout[9] = { 1.60475f, -0.53108f, -0.07367f, -0.10208f, 1.10813f, -0.00605f, -0.00327f, -0.07276f, 1.07602f };
    const float4 b = image_buffer[p];
    const vec3 last = V3(image_pixels[3 * p], image_pixels[3 * p + 1], image_pixels[3 * p + 2]);
    vec3 c = V3(b.x / b.w, b.y / b.w, b.z / b.w);
    c = c * exposure;
    c = V3(pow_contract(c.x, gamma_inv), pow_contract(c.y, gamma_inv), pow_contract(c.z, gamma_inv));
    c = mat_mul(Min, c);
    {
        vec3 a = V3(c.x * (c.x + 0.0245786f) - 0.000090537f, c.y * (c.y + 0.0245786f) - 0.000090537f,
                    c.z * (c.z + 0.0245786f) - 0.000090537f);
        vec3 d = V3(c.x * (0.983729f * c.x + 0.4329510f) + 0.238081f, c.y * (0.983729f * c.y + 0.4329510f) + 0.238081f,
                    c.z * (0.983729f * c.z + 0.4329510f) + 0.238081f);
        c = V3(a.x / d.x, a.y / d.y, a.z / d.z);
    }
    c = mat_mul(Mout, c);
    c = V3(fminf(fmaxf(c.x, 0.f), 1.f), fminf(fmaxf(c.y, 0.f), 1.f), fminf(fmaxf(c.z, 0.f), 1.f));
    image_pixels[3 * p] = c.x; image_pixels[3 * p + 1] = c.y; image_pixels[3 * p + 2] = c.z;
    if (diff_buffer != nullptr) {
        const vec3 dc = V3(fabsf(c.x - last.x), fabsf(c.y - last.y), fabsf(c.z - last.z));
        float2 db = diff_buffer[p];
        db.x += brightness(dc);
        db.y += 1.0f;
        diff_buffer[p] = db;
        diff_pixels[p] = db.x / db.y;
    }
}

// kernel denoise() of examples/denoise/denoise_test_1.py:86-118 as one DETERMINISTIC pass: the reference filters
// pixels_out in place while neighbouring threads read it (a race); here every neighbour comes from the previous output
// and the result goes to a second buffer (capi.cu swaps them).  28 B read + 12 B written per pixel (+ 4 neighbours from L2).
__global__ void __launch_bounds__(256) k_denoise(const float* __restrict__ pixels_in, const float* __restrict__ out_prev,
                                                 float* __restrict__ out_new, int W, int H, float threshold)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= W * H) return;
    const int i = p / H, j = p - i * H;
    const vec3 pixel1 = V3(pixels_in[3 * p], pixels_in[3 * p + 1], pixels_in[3 * p + 2]);
    const vec3 pixel2 = V3(out_prev[3 * p], out_prev[3 * p + 1], out_prev[3 * p + 2]);
    vec3 col = mix3(pixel1, pixel2, 0.2f);
    if (brightness(pixel1) < threshold) {
        const int ip = min(i + 1, W - 1), im = max(i - 1, 0), jp = min(j + 1, H - 1);
        const int q[4] = { ip * H + j, im * H + j, i * H + jp, i * H + jp };      // sur3 repeats j + 1, as written (:95-96)
        vec3 sum = V3(0.0f);
        float counter = 0.0f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const vec3 sur = V3(out_prev[3 * q[k]], out_prev[3 * q[k] + 1], out_prev[3 * q[k] + 2]);
            if (brightness(sur) > threshold) { sum = sum + sur; counter += 1.0f; }
        }
        col = V3(sum.x / counter, sum.y / counter, sum.z / counter);           // 0 / 0 = NaN when no neighbour qualifies (:113)
    }
    out_new[3 * p] = col.x; out_new[3 * p + 1] = col.y; out_new[3 * p + 2] = col.z;
}

// refresh() with ADAPTIVE_SAMPLING: diff_buffer = vec2(1), diff_pixels = 1e32 (src/renderer.py:18-20)
__global__ void __launch_bounds__(256) k_refresh_adaptive(float2* diff_buffer, float* diff_pixels, int n)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    diff_buffer[p] = make_float2(1.0f, 1.0f);
    diff_pixels[p] = 1e32f;
}

// ------------------------------------------------------------------------------------------
// Host-side launchers (called from capi.cu)
// ------------------------------------------------------------------------------------------
template <int NSLOT>
constexpr size_t pool_smem_bytes() { return (size_t)pool_smem_bytes_for(kPoolBlock, NSLOT); }

template <class VAR>
static cudaError_t launch_pool_t(const KParams& P, int grid, cudaStream_t stream)
{
    constexpr size_t smem = pool_smem_bytes<kPoolSlots>();
    // the attribute is per device: remember which devices this instantiation has been configured on
    static unsigned long long configured = 0ull;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 64 || !((configured >> dev) & 1ull)) {
        e = cudaFuncSetAttribute(k_pathtrace_pool<VAR, kPoolSlots>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev < 64) configured |= 1ull << dev;
    }
    k_pathtrace_pool<VAR, kPoolSlots><<<grid, kPoolBlock, smem, stream>>>(P);
    return cudaGetLastError();
}
template <class VAR>
static cudaError_t launch_simple_t(const KParams& P, cudaStream_t stream)
{
    const unsigned grid = (P.total_work + kSimpleBlock - 1) / kSimpleBlock;
    k_pathtrace_simple<VAR><<<grid, kSimpleBlock, 0, stream>>>(P);
    return cudaGetLastError();
}
template <class VAR>
static cudaError_t occupancy_t(int* blocks_per_sm)
{
    constexpr size_t smem = pool_smem_bytes<kPoolSlots>();
    cudaError_t e = cudaFuncSetAttribute(k_pathtrace_pool<VAR, kPoolSlots>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k_pathtrace_pool<VAR, kPoolSlots>, kPoolBlock, smem);
}

// Dispatch on (family, marcher, shape set, object count, count_work).
#define RT_CASE(FN, FAM, NOBJ, SHAPES, MARCH, ...)                                                   \
    do {                                                                                             \
        if (sel.count) return FN<Variant<FAM, NOBJ, SHAPES, MARCH, true>>(__VA_ARGS__);              \
        return FN<Variant<FAM, NOBJ, SHAPES, MARCH, false>>(__VA_ARGS__);                            \
    } while (0)

#define RT_DISPATCH(FN, ...)                                                                         \
    do {                                                                                             \
        if (sel.family == FAMILY_A && sel.marcher == MARCH_PLAIN && !sel.bunny) {                    \
            if (sel.nobj == 8) RT_CASE(FN, FAMILY_A, 8, SHAPESET_BOX, MARCH_PLAIN, __VA_ARGS__);     \
            RT_CASE(FN, FAMILY_A, 0, SHAPESET_BOX, MARCH_PLAIN, __VA_ARGS__);                        \
        }                                                                                            \
        if (sel.family == FAMILY_B && sel.marcher == MARCH_PLAIN && !sel.bunny)                      \
            RT_CASE(FN, FAMILY_B, 0, SHAPESET_ANALYTIC, MARCH_PLAIN, __VA_ARGS__);                   \
        if (sel.family == FAMILY_B && sel.marcher == MARCH_ENHANCED && !sel.bunny)                   \
            RT_CASE(FN, FAMILY_B, 0, SHAPESET_ANALYTIC, MARCH_ENHANCED, __VA_ARGS__);                \
        if (sel.family == FAMILY_B && sel.marcher == MARCH_ENHANCED && sel.bunny)                    \
            RT_CASE(FN, FAMILY_B, 0, SHAPESET_BUNNY, MARCH_ENHANCED, __VA_ARGS__);                   \
        if (sel.family == FAMILY_C && sel.marcher == MARCH_SRC && !sel.bunny)                        \
            RT_CASE(FN, FAMILY_C, 0, SHAPESET_ANALYTIC, MARCH_SRC, __VA_ARGS__);                     \
        return cudaErrorNotSupported;                                                                \
    } while (0)

size_t pool_dynamic_smem() { return pool_smem_bytes<kPoolSlots>(); }

cudaError_t launch_pathtrace_pool(const KernelSelect& sel, const KParams& P, int grid, cudaStream_t stream)
{
    RT_DISPATCH(launch_pool_t, P, grid, stream);
}
cudaError_t launch_pathtrace_simple(const KernelSelect& sel, const KParams& P, cudaStream_t stream)
{
    RT_DISPATCH(launch_simple_t, P, stream);
}
cudaError_t pool_occupancy(const KernelSelect& sel, int* blocks_per_sm)
{
    RT_DISPATCH(occupancy_t, blocks_per_sm);
}
bool kernel_supported(const KernelSelect& sel)
{
    if (sel.family == FAMILY_A) return sel.marcher == MARCH_PLAIN && !sel.bunny;
    if (sel.family == FAMILY_B) return (sel.marcher == MARCH_PLAIN && !sel.bunny) || sel.marcher == MARCH_ENHANCED;
    if (sel.family == FAMILY_C) return sel.marcher == MARCH_SRC && !sel.bunny;
    return false;
}
cudaError_t launch_fold_samples(const KParams& P, cudaStream_t stream)
{
    k_fold_samples<<<(P.total_work + 255) / 256, 256, 0, stream>>>(P);
    return cudaGetLastError();
}
cudaError_t launch_refresh_depth(float* ray_buffer, int n, cudaStream_t stream)
{
    k_refresh_depth<<<(n + 255) / 256, 256, 0, stream>>>(ray_buffer, n);
    return cudaGetLastError();
}
cudaError_t launch_post_process_src(const float4* image_buffer, float* image_pixels, float* diff_buffer, float* diff_pixels, int n,
                                    float exposure, float gamma_inv, cudaStream_t stream)
{
    k_post_process_src<<<(n + 255) / 256, 256, 0, stream>>>(image_buffer, image_pixels, reinterpret_cast<float2*>(diff_buffer),
                                                             diff_pixels, n, exposure, gamma_inv);
    return cudaGetLastError();
}
cudaError_t launch_refresh_adaptive(float* diff_buffer, float* diff_pixels, int n, cudaStream_t stream)
{
    k_refresh_adaptive<<<(n + 255) / 256, 256, 0, stream>>>(reinterpret_cast<float2*>(diff_buffer), diff_pixels, n);
    return cudaGetLastError();
}
cudaError_t launch_denoise(const float* pixels_in, const float* out_prev, float* out_new, int W, int H, float threshold, cudaStream_t stream)
{
    k_denoise<<<(W * H + 255) / 256, 256, 0, stream>>>(pixels_in, out_prev, out_new, W, H, threshold);
    return cudaGetLastError();
}
cudaError_t launch_post_process(const float4* image_buffer, float* image_pixels, int n, int mode, float exposure,
                                float inv_gamma, cudaStream_t stream)
{
    k_post_process<<<(n + 255) / 256, 256, 0, stream>>>(image_buffer, image_pixels, n, mode, exposure, inv_gamma);
    return cudaGetLastError();
}

}  // namespace rt
