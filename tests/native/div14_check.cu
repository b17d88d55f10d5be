// Exhaustive device check of div14_2 (rt_integrator.cuh): for EVERY binary32 bit pattern the packed fast path (or its
// fallback) must equal the IEEE division v / 1.4f.  Prints "<values checked> <in fast range> <mismatches>".
#include <cstdio>
#include <cuda_runtime.h>
#include "../../raytracingpbr_b200/csrc/rt_integrator.cuh"

__global__ void k(unsigned long long* out)
{
#if defined(__CUDA_ARCH__)      // div14_2 exists in the sm_100 device pass only
    unsigned long long bad = 0, fast = 0;
    const unsigned stride = gridDim.x * blockDim.x;
    unsigned u = blockIdx.x * blockDim.x + threadIdx.x;
    for (unsigned it = 0; it < (1u << 31) / stride; ++it, u += stride) {
        // pair (u, u | 0x80000000): both signs of the same magnitude
        const float a = __uint_as_float(u), b = __uint_as_float(u | 0x80000000u);
        const float2 q = rt::div14_2(make_float2(a, b));
        const float ra = a / 1.4f, rb = b / 1.4f;
        if (__float_as_uint(q.x) != __float_as_uint(ra) && !(ra != ra && q.x != q.x)) ++bad;
        if (__float_as_uint(q.y) != __float_as_uint(rb) && !(rb != rb && q.y != q.y)) ++bad;
        if (rt::div14_in_range(a)) fast += 2;
    }
    atomicAdd(&out[0], bad);
    atomicAdd(&out[1], fast);
#endif
}

int main()
{
    unsigned long long* d;
    cudaMalloc(&d, 16);
    cudaMemset(d, 0, 16);
    k<<<1024, 256>>>(d);
    unsigned long long h[2] = { ~0ull, 0 };
    if (cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("cuda error\n"); return 2; }
    printf("%llu %llu %llu\n", 1ull << 32, h[1], h[0]);
    return h[0] == 0 ? 0 : 1;
}
