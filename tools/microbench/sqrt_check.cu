// Checks on the device: sqrt_ranged(x) == sqrtf(x), sqrt_ranged(4x) == 2*sqrtf(x), and
// sd_box2_ranged_x2 == 2 * sd_box for random points / extents.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../raytracingpbr_b200/csrc/rt_integrator.cuh"
using namespace rt;

__device__ uint32_t hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

__global__ void k(unsigned long long* bad, int rounds)
{
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    for (int r = 0; r < rounds; ++r) {
        uint32_t h = hash(id * 977u + r * 7919u + 1u);
        // x in [2^-90, 2^40]: random exponent and mantissa
        uint32_t e = 127u - 90u + (hash(h) % 130u);
        float x = __uint_as_float((e << 23) | (h & 0x7fffffu));
        float a = sqrtf(x), b = sqrt_ranged(x), c = sqrt_ranged(4.0f * x);
        if (a != b) atomicAdd(&bad[0], 1ull);
        if (c != 2.0f * a) atomicAdd(&bad[1], 1ull);
        // boxes
        auto rnd = [&](uint32_t s, float lo, float hi) { return lo + (hi - lo) * ((hash(h + s) >> 8) * 0x1p-24f); };
        vec3 pa = V3(rnd(1, -3, 3), rnd(2, -3, 3), rnd(3, -3, 3)), pb = V3(rnd(4, -3, 3), rnd(5, -3, 3), rnd(6, -3, 3));
        if ((h & 7u) == 0u) { pa = pa * 1e-3f; pb = pb * 700.0f; }
        float ax0 = rnd(7, 0.01f, 1.5f), ay0 = rnd(8, 0.01f, 1.5f), az0 = rnd(9, 0.01f, 1.5f);
        if ((h & 3u) == 1u) {   // near-surface points: |p| = extent * (1 + tiny), any sign of tiny
            float sc = exp2f(-(float)(hash(h + 20) % 24u));
            pa.x = ax0 * (1.0f + rnd(21, -1, 1) * sc); if (h & 16u) pa.x = -pa.x;
            if (h & 32u) pa.y = ay0 * (1.0f + rnd(22, -1, 1) * sc);
            if (h & 64u) pa.z = az0 * (1.0f + rnd(23, -1, 1) * sc);
        }
        float ax = rnd(7, 0.01f, 1.5f), ay = rnd(8, 0.01f, 1.5f), az = rnd(9, 0.01f, 1.5f);
        float bx = rnd(10, 0.01f, 1.5f), by = rnd(11, 0.01f, 1.5f), bz = rnd(12, 0.01f, 1.5f);
        float da2, db2;
        sd_box2_ranged_x2(pa, ax, ay, az, pb, bx, by, bz, 0.0f, da2, db2);
        float da = sd_box(pa, ax, ay, az, 0.0f), db = sd_box(pb, bx, by, bz, 0.0f);
        if (da2 != 2.0f * da) { if (atomicAdd(&bad[2], 1ull) < 5) printf("A p=(%a,%a,%a) b=(%a,%a,%a) got %a want %a\n", pa.x, pa.y, pa.z, ax, ay, az, da2, 2.0f * da); }
        if (db2 != 2.0f * db) atomicAdd(&bad[3], 1ull);
    }
}

int main()
{
    unsigned long long* bad; cudaMallocManaged(&bad, 4 * sizeof(*bad)); for (int i = 0; i < 4; ++i) bad[i] = 0;
    k<<<148 * 8, 256>>>(bad, 4096);
    cudaDeviceSynchronize();
    printf("checked %llu values: sqrt_ranged!=sqrtf %llu, sqrt_ranged(4x)!=2sqrt %llu, box2 A mismatches %llu, B %llu (%s)\n",
           148ull * 8 * 256 * 4096, bad[0], bad[1], bad[2], bad[3], cudaGetErrorString(cudaGetLastError()));
    return 0;
}
