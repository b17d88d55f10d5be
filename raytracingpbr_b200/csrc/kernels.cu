// kernels.cu -- sm_100a kernels of the path-tracing hot path.
//
//   k_pathtrace_persistent  persistent-threads wavefront kernel (the product path)
//   k_pathtrace_simple      one thread per pixel, run-to-completion (validation / baseline)
//   k_post_process          tonemap (src/postprocessor.py:24-38 and the example variants)
//
// Compiled with -fmad=false: the fp32 contract (rt_math.cuh) allows only explicit fmaf().
#include <cuda_runtime.h>

#include "kernels.h"
#include "rt_integrator.cuh"

namespace rt {

constexpr unsigned kFull = 0xffffffffu;

// ------------------------------------------------------------------------------------------
// Persistent-threads wavefront kernel.
//
// Every lane owns one path at a time and walks a small state machine:
//     FETCH -> NEWPATH -> MARCH -> (HIT -> MARCH ...) -> DONE -> NEWPATH/FETCH ... -> IDLE
// The warp-wide loop body is ONE sphere-tracing step (scene SDF evaluation) for the lanes in
// MARCH.  Lanes whose ray has hit / left / been terminated wait ("pending") until at least
// resolve_q/32 of the warp's live lanes are pending (or nobody marches); then one *resolve
// round* shades all pending hits together, accumulates finished samples, regenerates camera
// paths in the freed lanes (path regeneration), and pulls new pixels from the global work
// queue with one warp-aggregated atomic.  This bounds SIMT divergence in the march loop (the
// >90% cost, heavy-tailed step counts) to 1 - resolve_q/32 idle lanes and keeps the divergent
// shading code off the hot loop.  Samples of one pixel are traced by one lane in index order,
// so the fp32 accumulation order equals the reference's launch-by-launch `buffer += color`
// (cornell_box_shortest.py:121) and the result is independent of scheduling.
// ------------------------------------------------------------------------------------------
enum : int { M_MARCH = 0, M_HIT = 1, M_DONE = 2, M_FETCH = 3, M_NEWPATH = 4, M_IDLE = 5 };

template <class VAR>
__global__ void __launch_bounds__(kPersistentBlock, kPersistentMinBlocks)
k_pathtrace_persistent(const __grid_constant__ KParams P)
{
    const int lane = threadIdx.x & 31;
    const unsigned lane_lt = (1u << lane) - 1u;

    int mode = M_FETCH;
    PathState st;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t pixel = 0;
    int pi = 0, pj = 0, s = 0;

    unsigned long long c_evals = 0, c_rays = 0, c_normals = 0, c_samples = 0;   // per lane
    unsigned long long c_iters = 0, c_active = 0, c_rounds = 0;                 // per warp

    for (;;) {
        const unsigned m_march = __ballot_sync(kFull, mode == M_MARCH);
        const unsigned m_pend = __ballot_sync(kFull, mode != M_MARCH && mode != M_IDLE);
        if ((m_march | m_pend) == 0u) break;
        const int n_march = __popc(m_march), n_pend = __popc(m_pend);

        if (n_pend > 0 && (n_march == 0 || n_pend * 32 >= (n_pend + n_march) * P.resolve_q)) {
            // ------------------------------------------------------------ resolve round
            if (VAR::COUNT) c_rounds++;
            if (mode == M_HIT) {
                if (VAR::COUNT) c_normals++;
                if (shade<VAR>(P, st) && begin_bounce<VAR>(P, pixel, P.sample_base + (uint32_t)s, st))
                    mode = M_MARCH;
                else
                    mode = M_DONE;
            }
            if (mode == M_DONE) {
                acc.x += st.col.x; acc.y += st.col.y; acc.z += st.col.z; acc.w += 1.0f;
                if (++s == P.spp) {
                    P.image_buffer[pixel] = acc;
                    mode = M_FETCH;
                } else {
                    mode = M_NEWPATH;
                }
            }
            // warp-aggregated pull from the global work queue (tile padding is skipped)
            for (;;) {
                const unsigned m_fetch = __ballot_sync(kFull, mode == M_FETCH);
                if (m_fetch == 0u) break;
                const int leader = __ffs(m_fetch) - 1;
                unsigned base = 0;
                if (lane == leader) base = atomicAdd(P.work_counter, (unsigned)__popc(m_fetch));
                base = __shfl_sync(kFull, base, leader);
                if (mode == M_FETCH) {
                    const unsigned w = base + (unsigned)__popc(m_fetch & lane_lt);
                    if (w >= P.total_work) {
                        mode = M_IDLE;
                    } else if (work_to_pixel(P, w, pi, pj)) {
                        pixel = (uint32_t)(pi * P.height + pj);
                        acc = P.image_buffer[pixel];
                        s = 0;
                        mode = M_NEWPATH;
                    }
                }
            }
            if (mode == M_NEWPATH) {
                if (VAR::COUNT) c_samples++;
                begin_path<VAR>(P, pixel, pi, pj, P.sample_base + (uint32_t)s, st);
                mode = begin_bounce<VAR>(P, pixel, P.sample_base + (uint32_t)s, st) ? M_MARCH : M_DONE;
            }
        }

        // ---------------------------------------------------------------- march step
        if (VAR::COUNT) { c_iters += 32; c_active += (unsigned long long)__popc(__ballot_sync(kFull, mode == M_MARCH)); }
        if (mode == M_MARCH) {
            const int status = march_step<VAR>(P, st);
            if (VAR::COUNT) c_evals++;
            if (status != MARCH_CONTINUE) {
                if (VAR::COUNT) c_rays++;
                if (status == MARCH_HIT) {
                    mode = M_HIT;
                } else {
                    miss(P, st);
                    mode = M_DONE;
                }
            }
        }
    }

    if (VAR::COUNT) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            c_evals += __shfl_xor_sync(kFull, c_evals, o);
            c_rays += __shfl_xor_sync(kFull, c_rays, o);
            c_normals += __shfl_xor_sync(kFull, c_normals, o);
            c_samples += __shfl_xor_sync(kFull, c_samples, o);
        }
        if (lane == 0) {
            atomicAdd(&P.counters[0], c_evals);
            atomicAdd(&P.counters[1], c_rays);
            atomicAdd(&P.counters[2], c_normals);
            atomicAdd(&P.counters[3], c_samples);
            atomicAdd(&P.counters[4], c_iters);
            atomicAdd(&P.counters[5], c_active);
            atomicAdd(&P.counters[6], c_rounds);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Simple kernel: one thread per work item, all spp, every path run to completion.
// ------------------------------------------------------------------------------------------
template <class VAR>
__global__ void __launch_bounds__(kSimpleBlock) k_pathtrace_simple(const __grid_constant__ KParams P)
{
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    int i, j;
    if (w >= P.total_work || !work_to_pixel(P, w, i, j)) return;
    const uint32_t pixel = (uint32_t)(i * P.height + j);
    float4 acc = P.image_buffer[pixel];
    unsigned long long cnt[4] = { 0, 0, 0, 0 };
    for (int s = 0; s < P.spp; ++s) {
        vec3 c = trace_sample<VAR>(P, pixel, i, j, P.sample_base + (uint32_t)s, cnt);
        acc.x += c.x; acc.y += c.y; acc.z += c.z; acc.w += 1.0f;
    }
    P.image_buffer[pixel] = acc;
    if (VAR::COUNT) {
        atomicAdd(&P.counters[0], cnt[0]);
        atomicAdd(&P.counters[1], cnt[1]);
        atomicAdd(&P.counters[2], cnt[2]);
        atomicAdd(&P.counters[3], cnt[3]);
    }
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

// ------------------------------------------------------------------------------------------
// Host-side launchers (called from capi.cu)
// ------------------------------------------------------------------------------------------
template <class VAR>
static cudaError_t launch_persistent_t(const KParams& P, int grid, cudaStream_t stream)
{
    k_pathtrace_persistent<VAR><<<grid, kPersistentBlock, 0, stream>>>(P);
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
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k_pathtrace_persistent<VAR>, kPersistentBlock, 0);
}

// Dispatch on (family, object count, count_work).  Family A scenes are boxes only.
#define RT_DISPATCH(FN, ...)                                                                         \
    do {                                                                                             \
        if (sel.family == FAMILY_A) {                                                                \
            if (sel.nobj == 8) {                                                                     \
                if (sel.count) return FN<Variant<FAMILY_A, 8, true, true>>(__VA_ARGS__);             \
                return FN<Variant<FAMILY_A, 8, true, false>>(__VA_ARGS__);                           \
            }                                                                                        \
            if (sel.count) return FN<Variant<FAMILY_A, 0, true, true>>(__VA_ARGS__);                 \
            return FN<Variant<FAMILY_A, 0, true, false>>(__VA_ARGS__);                               \
        }                                                                                            \
        return cudaErrorNotSupported;                                                                \
    } while (0)

cudaError_t launch_pathtrace_persistent(const KernelSelect& sel, const KParams& P, int grid, cudaStream_t stream)
{
    RT_DISPATCH(launch_persistent_t, P, grid, stream);
}
cudaError_t launch_pathtrace_simple(const KernelSelect& sel, const KParams& P, cudaStream_t stream)
{
    RT_DISPATCH(launch_simple_t, P, stream);
}
cudaError_t persistent_occupancy(const KernelSelect& sel, int* blocks_per_sm)
{
    RT_DISPATCH(occupancy_t, blocks_per_sm);
}
cudaError_t launch_post_process(const float4* image_buffer, float* image_pixels, int n, int mode, float exposure,
                                float inv_gamma, cudaStream_t stream)
{
    k_post_process<<<(n + 255) / 256, 256, 0, stream>>>(image_buffer, image_pixels, n, mode, exposure, inv_gamma);
    return cudaGetLastError();
}

}  // namespace rt
