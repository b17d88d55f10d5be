// hostcheck.cu -- TEST HARNESS ONLY.  Runs the product's __host__ __device__ path-tracing
// building blocks (raytracingpbr_b200/csrc/rt_integrator.cuh) on the CPU so that their fp32
// operation sequence can be compared bit-for-bit with the independent C oracle on a machine
// without a GPU.  Not linked into librtpbr.so and not reachable from the product API.
#include <cstring>
#include <vector>

#include "../../include/rtpbr.h"
#include "../../raytracingpbr_b200/csrc/host_setup.h"
#include "../../raytracingpbr_b200/csrc/rt_integrator.cuh"

using namespace rt;

extern "C" __attribute__((visibility("default"))) int hostcheck_pathtrace(const RtpbrConfig* cfg, const RtpbrCamera* cam,
                                                                          const RtpbrObject* objs, int n, float* image_buffer,
                                                                          int spp, uint32_t sample_base, int rank, int nranks,
                                                                          int band)
{
    if (cfg->family != RTPBR_FAMILY_A) return RTPBR_ERR_UNSUPPORTED;
    KParams P;
    memset(&P, 0, sizeof(P));
    fill_config(P, *cfg);
    fill_shard(P, rank, nranks, band);
    fill_objects(P, objs, n);
    fill_camera(P, *cfg, *cam);
    std::vector<float> rr = rr_table(*cfg);
    P.rr_prob = rr.data();
    P.spp = spp;
    P.sample_base = sample_base;
    float4* buf = reinterpret_cast<float4*>(image_buffer);
    typedef Variant<FAMILY_A, 0, true, false> VAR;
#pragma omp parallel for schedule(dynamic, 64)
    for (long long w = 0; w < (long long)P.total_work; ++w) {
        int i, j;
        if (!work_to_pixel(P, (uint32_t)w, i, j)) continue;
        const uint32_t pixel = (uint32_t)(i * P.height + j);
        float4 acc = buf[pixel];
        for (int s = 0; s < spp; ++s) {
            vec3 c = trace_sample<VAR>(P, pixel, i, j, sample_base + (uint32_t)s, nullptr);
            acc.x += c.x; acc.y += c.y; acc.z += c.z; acc.w += 1.0f;
        }
        buf[pixel] = acc;
    }
    return 0;
}

extern "C" __attribute__((visibility("default"))) void hostcheck_sincos(float x, float* s, float* c) { sincos_rt(x, *s, *c); }
extern "C" __attribute__((visibility("default"))) float hostcheck_atan2(float y, float x) { return atan2_rt(y, x); }
extern "C" __attribute__((visibility("default"))) float hostcheck_asin(float x) { return asin_rt(x); }
extern "C" __attribute__((visibility("default"))) void hostcheck_euler(const float rot_deg[3], float out9[9])
{
    euler_matrix_deg(rot_deg, out9);
}
