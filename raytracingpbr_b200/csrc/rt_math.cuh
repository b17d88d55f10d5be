// rt_math.cuh -- fp32 vector math, transcendental routines and Philox4x32-10 for the
// path-tracing kernels.  All functions are __host__ __device__ so tests/native/hostcheck.cu
// can run the very same inline code on the CPU (test harness only; the product never does).
//
// FP32 CONTRACT.  Device code is compiled with -fmad=false and host
// code with -ffp-contract=off, so the only fused operations are the explicit fmaf() calls:
//   dot(a,b)   = fmaf(a.z,b.z, fmaf(a.y,b.y, a.x*b.x));  M@v = per-row dot
//   length(v)  = sqrtf(dot(v,v)) (IEEE, -prec-sqrt=true);  normalize(v) = v * (1.0f/length(v))
//   mix(a,b,t) = a*(1-t) + b*t;  pow(x, 5.0) = ((x*x)*(x*x))*x
//   every operator the reference itself writes (+ - * /, e.g. origin + t * direction) is ONE
//   IEEE binary32 rounding, left to right, never contracted.
// sin/cos/atan2/asin are the polynomial routines below -- never libdevice/libm, which differ
// bitwise between host and device.
#pragma once
#if defined(__CUDACC_RTC__)
typedef unsigned int uint32_t;
typedef int int32_t;
typedef unsigned long long uint64_t;
typedef unsigned char uint8_t;
typedef unsigned long size_t;
#else
#include <cstdint>
#include <cmath>
#endif

#if defined(__CUDACC__)
#define RT_HD __host__ __device__ __forceinline__
#else
#define RT_HD inline
#endif

namespace rt {

struct vec3 {
    float x, y, z;
};

RT_HD vec3 V3(float x, float y, float z) { vec3 r; r.x = x; r.y = y; r.z = z; return r; }
RT_HD vec3 V3(float s) { return V3(s, s, s); }
RT_HD vec3 operator+(vec3 a, vec3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
RT_HD vec3 operator-(vec3 a, vec3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
RT_HD vec3 operator*(vec3 a, vec3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
RT_HD vec3 operator*(vec3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
RT_HD vec3 operator-(vec3 a) { return V3(-a.x, -a.y, -a.z); }
RT_HD float dot(vec3 a, vec3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
RT_HD float length(vec3 a) { return sqrtf(dot(a, a)); }
RT_HD vec3 normalize(vec3 a) { return a * (1.0f / length(a)); }
RT_HD vec3 cross(vec3 a, vec3 b)
{
    return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
RT_HD vec3 at(vec3 o, vec3 d, float t) { return o + d * t; }  // ray.origin + t * ray.direction
RT_HD vec3 mat_mul(const float* m, vec3 v)
{
    return V3(fmaf(m[2], v.z, fmaf(m[1], v.y, m[0] * v.x)), fmaf(m[5], v.z, fmaf(m[4], v.y, m[3] * v.x)),
              fmaf(m[8], v.z, fmaf(m[7], v.y, m[6] * v.x)));
}
RT_HD float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
RT_HD vec3 mix3(vec3 a, vec3 b, float t) { return V3(mixf(a.x, b.x, t), mixf(a.y, b.y, t), mixf(a.z, b.z, t)); }
RT_HD float brightness(vec3 c) { return dot(c, V3(0.299f, 0.587f, 0.114f)); }  // src/util.py:31-33

constexpr float kPi = 3.14159274101257324f;       // fl32(pi)
constexpr float kHalfPi = 1.57079637050628662f;   // fl32(pi/2)
constexpr float kDegToRad = 0.01745329238474369f; // fl32(pi/180)

// sin and cos by 3-term Cody-Waite reduction (fmaf) + minimax polynomials on [-pi/4, pi/4].
RT_HD void sincos_rt(float x, float& sn, float& cs)
{
    float j = rintf(x * 0.636619746685028076f);
    float r = fmaf(-j, 0x1.921fb6p+0f, x);
    r = fmaf(-j, -0x1.777a5cp-25f, r);
    r = fmaf(-j, -0x1.ee59dap-50f, r);
    int q = (int)j;
    float r2 = r * r;
    float sp = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
    sp = fmaf(sp, r2, -1.6666654611e-1f);
    float s = fmaf(sp * r2, r, r);
    float cp = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
    cp = fmaf(cp, r2, 4.166664568298827e-2f);
    cp = fmaf(cp, r2, -0.5f);
    float c = fmaf(cp, r2, 1.0f);
    if (q & 1) { float t = s; s = c; c = t; }
    if (q & 2) s = -s;
    if ((q + 1) & 2) c = -c;
    sn = s;
    cs = c;
}

RT_HD float atan01_rt(float a)
{
    float s = a * a;
    float p = fmaf(s, 0.00282363896258175373077393f, -0.0159569028764963150024414f);
    p = fmaf(p, s, 0.0425049886107444763183594f);
    p = fmaf(p, s, -0.0748900920152664184570312f);
    p = fmaf(p, s, 0.106347933411598205566406f);
    p = fmaf(p, s, -0.142027363181114196777344f);
    p = fmaf(p, s, 0.199926957488059997558594f);
    p = fmaf(p, s, -0.333331018686294555664062f);
    return fmaf(p * s, a, a);
}
RT_HD float atan2_rt(float y, float x)
{
    float ax = fabsf(x), ay = fabsf(y);
    float mx = ax > ay ? ax : ay, mn = ax < ay ? ax : ay;
    float a = (mx == 0.0f) ? 0.0f : mn / mx;
    float r = atan01_rt(a);
    if (ay > ax) r = 0x1.921fb6p+0f - r;
    if (x < 0.0f) r = kPi - r;
    return (y < 0.0f) ? -r : r;
}
RT_HD float asin_rt(float x)
{
    x = x > 1.0f ? 1.0f : (x < -1.0f ? -1.0f : x);
    return atan2_rt(x, sqrtf((1.0f - x) * (1.0f + x)));
}

// ------------------------------------------------------------------ Philox4x32-10
struct uint4_rt { uint32_t x, y, z, w; };

RT_HD void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo)
{
#if defined(__CUDA_ARCH__)
    lo = a * b;
    hi = __umulhi(a, b);
#else
    uint64_t p = (uint64_t)a * b;
    lo = (uint32_t)p;
    hi = (uint32_t)(p >> 32);
#endif
}

// RT_RESOLVE_OOL (set by jit_codegen.h for the PBR families): out of line.  Their resolve phase draws from up
// to six call sites per bounce, and six inlined copies of the ten rounds only bloat a code path that already
// misses the 32 KB instruction cache (tokyo_ibl: stall_no_instruction 3.0 per issue; +5.5 % with this).  The
// diffuse-only family A has two call sites and is 1 % faster with the rounds inlined.
#if defined(RT_RESOLVE_OOL) && defined(__CUDACC__)
static __host__ __device__ __noinline__
#else
RT_HD
#endif
uint4_rt philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo(0xD2511F53u, c0, hi0, lo0);
        mulhilo(0xCD9E8D57u, c2, hi1, lo1);
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    uint4_rt o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

RT_HD float u01(uint32_t u) { return (float)(u >> 8) * 0x1p-24f; }  // Taichi u32 -> f32 rule

constexpr uint32_t kPhiloxKey1 = 0x52545042u;  // "RTPB"

// RNG CONTRACT: the n-th ti.random() call (n = 0, 1, ...) a pixel makes inside kernel launch
// number L is word (n & 3) of Philox4x32-10(counter = (pixel, L, n >> 2, 0), key = (seed,
// "RTPB")), pixel = i * height + j.  Families A/B trace one sample per launch, so L is the
// sample index.  Nothing else about the path structure enters the RNG, so regeneration and
// compaction cannot change which numbers a sample sees.
RT_HD float pick4(const uint4_rt& o, uint32_t k)
{
    uint32_t u = k == 0u ? o.x : (k == 1u ? o.y : (k == 2u ? o.z : o.w));
    return u01(u);
}
RT_HD float rng_at(uint32_t seed, uint32_t pixel, uint32_t launch, uint32_t n)
{
    return pick4(philox4x32_10(pixel, launch, n >> 2, 0u, seed, kPhiloxKey1), n & 3u);
}
// Draws n, n+1, n+2 of the stream with at most two Philox evaluations.
RT_HD void rng_at3(uint32_t seed, uint32_t pixel, uint32_t launch, uint32_t n, float& a, float& b, float& c)
{
    const uint32_t blk = n >> 2, k = n & 3u;
    uint4_rt o = philox4x32_10(pixel, launch, blk, 0u, seed, kPhiloxKey1);
    if (k <= 1u) {
        a = pick4(o, k); b = pick4(o, k + 1u); c = pick4(o, k + 2u);
    } else {
        uint4_rt q = philox4x32_10(pixel, launch, blk + 1u, 0u, seed, kPhiloxKey1);
        a = pick4(o, k);
        b = k == 2u ? u01(o.w) : u01(q.x);
        c = k == 2u ? u01(q.x) : u01(q.y);
    }
}

}  // namespace rt
