// host_setup.h -- per-frame host precomputation (camera frame, rotation matrices, Russian
// roulette table).  Host code is compiled with -ffp-contract=off; see the fp32 contract in
// rt_math.cuh.  Shared by capi.cu and tests/native/hostcheck.cu.
#pragma once
#include <cmath>
#include <vector>

#include "../../include/rtpbr.h"
#include "rt_math.cuh"
#include "rt_params.h"

namespace rt {

// cornell_box_shortest.py:34-39 `angle` == src/util.py:36-42 `rotate`: Rz @ Ry @ Rx of
// radians(rotation); replaces kernel update_all_transform (src/scene.py:99-109).
inline void euler_matrix_deg(const float rot_deg[3], float out[9])
{
    float sx, cx, sy, cy, sz, cz;
    sincos_rt(rot_deg[0] * kDegToRad, sx, cx);
    sincos_rt(rot_deg[1] * kDegToRad, sy, cy);
    sincos_rt(rot_deg[2] * kDegToRad, sz, cz);
    const float A[9] = { cz, sz, 0, -sz, cz, 0, 0, 0, 1 };
    const float B[9] = { cy, 0, -sy, 0, 1, 0, sy, 0, cy };
    const float C[9] = { 1, 0, 0, 0, cx, sx, 0, -sx, cx };
    float AB[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            AB[3 * i + j] = fmaf(A[3 * i + 2], B[6 + j], fmaf(A[3 * i + 1], B[3 + j], A[3 * i] * B[j]));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            out[3 * i + j] = fmaf(AB[3 * i + 2], C[6 + j], fmaf(AB[3 * i + 1], C[3 + j], AB[3 * i] * C[j]));
}

inline void fill_objects(KParams& P, const RtpbrObject* objs, int n)
{
    P.nobj = n;
    for (int k = 0; k < n; ++k) {
        const RtpbrObject& o = objs[k];
        DevGeom& g = P.geom[k];
        g.px = o.position[0]; g.py = o.position[1]; g.pz = o.position[2];
        euler_matrix_deg(o.rotation, g.m);
        g.sx = o.scale[0]; g.sy = o.scale[1]; g.sz = o.scale[2];
        g.type = o.type;
        DevMaterial& m = P.mat[k];
        for (int c = 0; c < 3; ++c) { m.albedo[c] = o.albedo[c]; m.emission[c] = o.emission[c]; }
        m.roughness = o.roughness; m.metallic = o.metallic; m.transmission = o.transmission; m.ior = o.ior;
        m.pad0 = m.pad1 = 0.f;
    }
}

// Family A: cornell_box_shortest.py:107-114 (pinhole, square frame, half = tan(radians(vfov)/2)).
// Families B/C: get_ray() of src/camera.py:11-36 (thin lens; frame scaled by focus).
inline void fill_camera(KParams& P, const RtpbrConfig& cfg, const RtpbrCamera& c)
{
    vec3 from = V3(c.lookfrom[0], c.lookfrom[1], c.lookfrom[2]);
    vec3 lookat = V3(c.lookat[0], c.lookat[1], c.lookat[2]);
    vec3 up = V3(c.vup[0], c.vup[1], c.vup[2]);
    vec3 z = normalize(from - lookat);
    vec3 x = normalize(cross(up, z));
    vec3 y = cross(z, x);
    float theta = c.vfov * kDegToRad;
    float half_height = (float)tan((double)(theta * 0.5f));
    DevCamera& d = P.cam;
    vec3 llc, hor, ver;
    if (cfg.family == RTPBR_FAMILY_A) {
        float half = half_height;
        llc = ((from - x * half) - y * half) - z;
        hor = x * (2.0f * half);
        ver = y * (2.0f * half);
        d.lens_radius = 0.0f;
    } else {
        float half_width = c.aspect * half_height;
        vec3 hwfx = x * (half_width * c.focus);
        vec3 hhfy = y * (half_height * c.focus);
        llc = ((from - hwfx) - hhfy) - z * c.focus;
        hor = hwfx * 2.0f;
        ver = hhfy * 2.0f;
        d.lens_radius = c.aperture * 0.5f;
    }
    d.origin[0] = from.x; d.origin[1] = from.y; d.origin[2] = from.z;
    d.llc[0] = llc.x; d.llc[1] = llc.y; d.llc[2] = llc.z;
    d.horizontal[0] = hor.x; d.horizontal[1] = hor.y; d.horizontal[2] = hor.z;
    d.vertical[0] = ver.x; d.vertical[1] = ver.y; d.vertical[2] = ver.z;
    d.x[0] = x.x; d.x[1] = x.y; d.x[2] = x.z;
    d.y[0] = y.x; d.y[1] = y.y; d.y[2] = y.z;
    d.fw = (float)cfg.width; d.fh = (float)cfg.height;
    d.inv_w = (float)(1.0 / (double)cfg.width);   // SCREEN_PIXEL_SIZE = 1.0 / vec2(res), src/config.py:19
    d.inv_h = (float)(1.0 / (double)cfg.height);
}

// cornell_box_shortest.py:84-85: inv_pdf = exp(i / light_quality); p = 1 - 1 / inv_pdf.
// Depends on the bounce index only -> host table.
inline std::vector<float> rr_table(const RtpbrConfig& cfg)
{
    std::vector<float> t((size_t)(cfg.max_bounces > 0 ? cfg.max_bounces : 1));
    for (size_t i = 0; i < t.size(); ++i) {
        float inv_pdf = (float)exp((double)((float)i / cfg.light_quality));
        t[i] = 1.0f - (1.0f / inv_pdf);
    }
    return t;
}

inline int owned_columns(int width, int rank, int nranks, int band)
{
    int n = 0;
    for (int i = 0; i < width; ++i)
        if ((i / band) % nranks == rank) ++n;
    return n;
}

inline void fill_config(KParams& P, const RtpbrConfig& c)
{
    P.width = c.width; P.height = c.height;
    P.seed = c.seed;
    P.max_bounces = c.max_bounces; P.max_steps = c.max_steps;
    P.t_start = c.t_start; P.hit_eps = c.hit_eps; P.t_far = c.t_far;
    P.relax_w0 = c.relax_w0; P.relax_w_reset = c.relax_w_reset;
    P.relax_guard = c.relax_guard; P.relax_reset = c.relax_reset;
    P.normal_h = c.normal_h; P.box_round = c.box_round;
    P.visibility_min = c.visibility_min; P.visibility_max = c.visibility_max;
    P.bsdf = c.bsdf; P.f0_variant = c.f0_variant;
    P.nearest_seed = c.nearest_seed; P.normal_mode = c.normal_mode;
    P.samples_per_pixel = c.samples_per_pixel > 0 ? c.samples_per_pixel : 1;
    P.adaptive = c.family == RTPBR_FAMILY_C && c.adaptive_sampling != 0;
    P.noise_threshold = c.noise_threshold;
    P.inv_max_bounces = (float)(1.0 / (double)c.max_bounces);   // ti.static(1.0 / MAX_RAYTRACE), src/pathtracer.py:68
    P.sky = c.sky; P.sky_scale = c.sky_scale;
    P.min_dis = c.min_dis; P.pixel_radius = c.pixel_radius; P.quality_per_sample = c.quality_per_sample;
    P.black_background = c.black_background;
    P.inner_spp = c.inner_spp > 0 ? c.inner_spp : 0;
    P.primary_miss = c.primary_miss;
    P.bunny_bob = c.bunny_bob != 0;
}

// bunny_sdf_glass.py:213-216: t = pi * float(u_frame) / 120.0; angle(vec3(0, 0, t)); 0.1 * sin(t)
inline void fill_frame(KParams& P, int frame)
{
    P.frame = frame;
    float t = kPi * (float)frame / 120.0f, sn, cs;
    sincos_rt(t, sn, cs);
    // angle(vec3(0, 0, t)) = Rz(t) @ I @ I evaluated with the same matrix products as euler_matrix_deg
    const float A[9] = { cs, sn, 0, -sn, cs, 0, 0, 0, 1 };
    const float I3[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
    float AB[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            AB[3 * i + j] = fmaf(A[3 * i + 2], I3[6 + j], fmaf(A[3 * i + 1], I3[3 + j], A[3 * i] * I3[j]));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            P.anim_m[3 * i + j] = fmaf(AB[3 * i + 2], I3[6 + j], fmaf(AB[3 * i + 1], I3[3 + j], AB[3 * i] * I3[j]));
    P.anim_bob = 0.1f * sn;
}

inline void fill_shard(KParams& P, int rank, int nranks, int band)
{
    P.rank = rank; P.nranks = nranks; P.band = band;
    P.local_cols = owned_columns(P.width, rank, nranks, band);
    P.tiles_per_col = (P.height + 7) / 8;
    P.total_work = (uint32_t)(((P.local_cols + 3) / 4) * P.tiles_per_col) * 32u;
}

}  // namespace rt
