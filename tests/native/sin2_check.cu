// Exhaustive device check of sin2_rt (rt_integrator.cuh): for EVERY binary32 value x with |x| * 2/pi < 2^22 (except -0,
// excluded by its precondition) the packed routine must return the bits of the scalar contract routine sin_rt().
// Prints "<values checked> <mismatches>".
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#include "../../raytracingpbr_b200/csrc/rt_integrator.cuh"

__global__ void k(unsigned long long* out, unsigned limit)
{
#if defined(__CUDA_ARCH__)      // sin2_rt exists in the sm_100 device pass only
    unsigned long long bad = 0, n = 0;
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned u = blockIdx.x * blockDim.x + threadIdx.x; u < limit; u += stride) {
        const float a = __uint_as_float(u), b = u == 0u ? 1.0f : __uint_as_float(u | 0x80000000u);   // both signs (never -0)
        const float2 q = rt::sin2_rt(make_float2(a, b));
        const float ra = rt::sin_rt(a), rb = rt::sin_rt(b);
        if (__float_as_uint(q.x) != __float_as_uint(ra)) ++bad;
        if (__float_as_uint(q.y) != __float_as_uint(rb)) ++bad;
        n += 2;
    }
    atomicAdd(&out[0], bad);
    atomicAdd(&out[1], n);
#endif
}

int main()
{
    unsigned long long* d;
    cudaMalloc(&d, 16);
    cudaMemset(d, 0, 16);
    const float xmax = 4194304.0f * 1.5707963f * 0.999f;       // |x| * 2/pi < 2^22 with a little room
    unsigned limit;
    memcpy(&limit, &xmax, 4);
    k<<<148 * 8, 256>>>(d, limit);
    unsigned long long h[2] = { ~0ull, 0 };
    if (cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("cuda error\n"); return 2; }
    printf("%llu %llu\n", h[1], h[0]);
    return h[0] == 0 ? 0 : 1;
}
