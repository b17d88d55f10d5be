// jit_codegen.h -- generates the scene-specialised translation unit (host-only, plain C++).
//
// The reference is JIT-compiled per scene by Taichi (`ti.static` loops over OBJECTS with the shape
// function resolved at compile time, src/scene.py:48-51).  This is the B200 counterpart: the
// march loop's nearest() is emitted as straight-line CUDA with every object constant as an
// immediate operand.  All rewrites are EXACT for finite inputs (bit-identical to the generic
// code):
//   x - 0            -> x
//   m*x with m = 0   -> dropped (0*x = +-0 and adding +-0 never changes a non-zero sum; the sign of a
//                       zero coordinate is irrelevant because every primitive takes |p| or p^2 ... )
//   m*x with m = +-1 -> +-x; fmaf(+-1, y, acc) -> acc +- y (one rounding either way)
//   fmaf(m, y, 0)    -> m*y (one rounding either way)
//
// Two scene-level analyses ride on top (analyse()):
//   * FAST REGION (walls as planes).  For an axis-aligned box, a point whose two in-plane coordinates lie
//     inside the face (q_a <= 0) and whose normal coordinate lies outside the slab (q_n > 0) has
//     length(max(q, 0)) = sqrt(fl(q_n^2)) = q_n EXACTLY (binary fp32, correctly rounded sqrt) and
//     min(max(q), 0) = 0, so the box distance IS the plane distance q_n -- no square root, no dot product,
//     no clamps.  analyse() finds the boxes for which one axis-aligned region of space guarantees the
//     in-plane condition (the five walls of the Cornell box and the room they enclose) and emits
//     jit_nearest_fast(): inside the region (and with every q_n > 0) it returns the very same bits as
//     jit_nearest_dist(); elsewhere it reports ok = false and the caller falls back to the full code.
//   * SCENE BOUNDS (provable misses).  When every primitive is bounded, the translation unit gets the
//     world-space bounding box of all surfaces (RT_JIT_BBOX); rt_integrator.cuh: ray_t_stop() uses it to
//     cut a march as soon as the ray has left the box for good (the reference marches such rays on to
//     t > MAX_DIS, with the same outcome: a miss).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <utility>
#include <vector>

#include "../../include/rtpbr.h"
#include "host_setup.h"

namespace rt {
namespace jit {

inline std::string flit(float v)
{
    char buf[64];
    if (v == 0.0f) return "0.0f";
    snprintf(buf, sizeof(buf), "(%af)", (double)v);
    return buf;
}

// expression for one row of M @ d under the dot contract fmaf(m2, dz, fmaf(m1, dy, m0 * dx))
inline std::string row_expr(const float m[3], const char* const d[3])
{
    std::string acc;   // empty = exact zero so far
    for (int k = 0; k < 3; ++k) {
        const float c = m[k];
        if (c == 0.0f) continue;
        if (acc.empty()) {
            if (c == 1.0f) acc = d[k];
            else if (c == -1.0f) acc = std::string("(-") + d[k] + ")";
            else acc = "(" + flit(c) + " * " + d[k] + ")";
        } else {
            if (c == 1.0f) acc = "(" + acc + " + " + d[k] + ")";
            else if (c == -1.0f) acc = "(" + acc + " - " + d[k] + ")";
            else acc = "fmaf(" + flit(c) + ", " + d[k] + ", " + acc + ")";
        }
    }
    return acc.empty() ? "0.0f" : acc;
}

struct Source {
    std::string text;
    std::string kernel_name;
    bool fast = false;     // jit_nearest_fast() + drop-out handling are compiled in (RT_JIT_FAST)
    bool tstop = false;    // per-ray t_stop in the march loop (RT_JIT_TSTOP)
    bool relaxed = false;  // over-relaxed marcher (enhanced / src) without the neural bunny: its march state wants > 64 registers
};

// Tuning / test knobs: unset = automatic, "0" = off, anything else = on wherever the analysis permits.
inline int knob(const char* name)
{
    const char* v = getenv(name);
    if (!v || !*v) return -1;
    return atoi(v) != 0 ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------------
// Scene analysis
// ---------------------------------------------------------------------------------------------------
struct Wall {
    int object;      // index into the scene
    int normal;      // local axis with the smallest half-extent (the slab's normal)
};
struct Analysis {
    bool bounded = false;          // every primitive has a finite bounding box
    double lo[3] = { 0, 0, 0 }, hi[3] = { 0, 0, 0 };   // world-space box around every surface (conservative)
    double scale = 1.0;            // max(1, largest |coordinate| of the box)
    bool fast = false;             // jit_nearest_fast() exists
    std::vector<Wall> walls;
    float rc[3] = { 0, 0, 0 }, rh[3] = { 0, 0, 0 };    // fast region: |pos_a - rc_a| <= rh_a
};

inline float round_of(const RtpbrConfig& cfg) { return cfg.family == RTPBR_FAMILY_A ? 0.0f : cfg.box_round; }

inline bool box_is_ranged(const RtpbrObject& o)
{
    // ranged sqrt is exact when every half-extent is a normal number in [2^-26, 2^40]
    bool r = o.type == RTPBR_SHAPE_BOX;
    for (int a = 0; a < 3; ++a) r = r && o.scale[a] >= 0x1p-26f && o.scale[a] <= 0x1p40f;
    return r;
}

inline Analysis analyse(const RtpbrConfig& cfg, const RtpbrObject* objs, int n)
{
    Analysis A;
    const double round_ = (double)round_of(cfg);
    // ---- world bounds of every primitive
    bool bounded = true, any = false;
    double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
    for (int k = 0; k < n && bounded; ++k) {
        const RtpbrObject& o = objs[k];
        double ext[3];   // half-extents of the primitive in its own frame
        switch (o.type) {
        case RTPBR_SHAPE_NONE: continue;                                  // constant distance MAX_DIS: no surface
        case RTPBR_SHAPE_SPHERE: ext[0] = ext[1] = ext[2] = std::fabs((double)o.scale[0]); break;
        case RTPBR_SHAPE_BOX: for (int a = 0; a < 3; ++a) ext[a] = std::fabs((double)o.scale[a]) + round_; break;
        case RTPBR_SHAPE_CYLINDER: ext[0] = ext[2] = std::fabs((double)o.scale[0]); ext[1] = std::fabs((double)o.scale[1]); break;
        case RTPBR_SHAPE_BUNNY: ext[0] = ext[1] = ext[2] = 1.2; break;    // unit ball (MLP inside, |p| - 0.8 outside) + 0.1 bob
        default: bounded = false; continue;                               // cone, plane: unbounded
        }
        float m[9];
        euler_matrix_deg(o.rotation, m);
        for (int a = 0; a < 3; ++a) {
            // p = M (x - c)  =>  x = c + M^T p: |x_a - c_a| <= sum_j |M[j][a]| ext_j (a ball for sphere-like primitives)
            double e = 0.0;
            if (o.type == RTPBR_SHAPE_SPHERE || o.type == RTPBR_SHAPE_BUNNY) e = ext[0];
            else for (int j = 0; j < 3; ++j) e += std::fabs((double)m[3 * j + a]) * ext[j];
            e = e * (1.0 + 1e-5) + 1e-6;
            if (!std::isfinite(e) || !std::isfinite((double)o.position[a])) { bounded = false; break; }
            lo[a] = std::min(lo[a], (double)o.position[a] - e);
            hi[a] = std::max(hi[a], (double)o.position[a] + e);
        }
        any = true;
    }
    if (bounded && any) {
        A.bounded = true;
        double s = 1.0;
        for (int a = 0; a < 3; ++a) { A.lo[a] = lo[a]; A.hi[a] = hi[a]; s = std::max(s, std::max(std::fabs(lo[a]), std::fabs(hi[a]))); }
        A.scale = s;
        if (!(s < 1e6)) A.bounded = false;                               // keeps every constant of ray_t_stop() comfortably in range
    }
    const bool have_box = A.bounded;   // the fast region below starts from the box, whether or not t_stop is used
    // ray_t_stop()'s argument needs a relaxation factor in [0, 2] (a step back never passes the previous evaluation point)
    if (cfg.marcher == RTPBR_MARCH_ENHANCED &&
        !(cfg.relax_w0 >= 0.0f && cfg.relax_w0 <= 2.0f && cfg.relax_w_reset >= 0.0f && cfg.relax_w_reset <= 2.0f))
        A.bounded = false;
    if (!(cfg.hit_eps >= 0.0f && cfg.hit_eps < 0.25f && cfg.t_far > 0.0f && cfg.t_far < 1e30f)) A.bounded = false;
    if (knob("RTPBR_JIT_BBOX") == 0) A.bounded = false;

    // ---- fast region: axis-aligned ("wall") boxes whose in-plane condition one region of space guarantees
    if (!have_box) return A;
    if (knob("RTPBR_JIT_FAST") == 0) return A;
    struct Cand { int object, normal; int world_axis[3]; double area; };
    std::vector<Cand> cands;
    for (int k = 0; k < n; ++k) {
        const RtpbrObject& o = objs[k];
        if (!box_is_ranged(o)) continue;
        float m[9];
        euler_matrix_deg(o.rotation, m);
        Cand c; c.object = k;
        bool ok = true;
        int used = 0;
        for (int r = 0; r < 3 && ok; ++r) {          // every row: one entry exactly +-1, the others tiny
            int big = -1;
            for (int j = 0; j < 3; ++j) {
                const float v = std::fabs(m[3 * r + j]);
                if (v == 1.0f) { if (big >= 0) ok = false; big = j; }
                else if (!(v <= 0x1p-20f)) ok = false;
            }
            if (big < 0 || (used >> big & 1)) ok = false;
            else { used |= 1 << big; c.world_axis[r] = big; }
        }
        if (!ok) continue;
        c.normal = 0;
        for (int a = 1; a < 3; ++a) if (o.scale[a] < o.scale[c.normal]) c.normal = a;
        c.area = 1.0;
        for (int a = 0; a < 3; ++a) if (a != c.normal) c.area *= (double)o.scale[a];
        cands.push_back(c);
    }
    std::stable_sort(cands.begin(), cands.end(), [](const Cand& a, const Cand& b) { return a.area > b.area; });
    double rlo[3], rhi[3];
    for (int a = 0; a < 3; ++a) { rlo[a] = A.lo[a]; rhi[a] = A.hi[a]; }
    auto volume = [](const double* l, const double* h) {
        double v = 1.0;
        for (int a = 0; a < 3; ++a) v *= std::max(0.0, h[a] - l[a]);
        return v;
    };
    // |p_a| <= |x_j - c_j| + (two tiny matrix entries) * |x - c| + rounding: slop covers both generously
    const double slop = 1e-5 * (A.scale + 1.0);
    for (const Cand& c : cands) {
        const RtpbrObject& o = objs[c.object];
        double nlo[3], nhi[3];
        for (int a = 0; a < 3; ++a) { nlo[a] = rlo[a]; nhi[a] = rhi[a]; }
        for (int r = 0; r < 3; ++r) {
            if (r == c.normal) continue;
            const int j = c.world_axis[r];
            nlo[j] = std::max(nlo[j], (double)o.position[j] - (double)o.scale[r] + slop);
            nhi[j] = std::min(nhi[j], (double)o.position[j] + (double)o.scale[r] - slop);
        }
        if (volume(nlo, nhi) >= 0.5 * volume(rlo, rhi) && volume(nlo, nhi) > 0.0) {
            for (int a = 0; a < 3; ++a) { rlo[a] = nlo[a]; rhi[a] = nhi[a]; }
            A.walls.push_back(Wall{ c.object, c.normal });
        }
    }
    if (A.walls.size() < 2) { A.walls.clear(); return A; }
    std::sort(A.walls.begin(), A.walls.end(), [](const Wall& a, const Wall& b) { return a.object < b.object; });
    for (int a = 0; a < 3; ++a) {
        const double c = 0.5 * (rlo[a] + rhi[a]), h = 0.5 * (rhi[a] - rlo[a]);
        A.rc[a] = std::fabs(c) < 1e-9 * A.scale ? 0.0f : (float)c;
        // the test runs in fp32: fl(pos - rc) <= rh; shrink by the rounding of rc, of the difference and of rh
        const double hh = h - std::fabs(c - (double)A.rc[a]) - 4e-7 * (A.scale + std::fabs(c) + h);
        if (!(hh > 0.0)) { A.walls.clear(); return A; }
        A.rh[a] = std::nextafterf((float)hh, 0.0f);
    }
    A.fast = true;
    return A;
}

// ---------------------------------------------------------------------------------------------------
// Emission
// ---------------------------------------------------------------------------------------------------
struct Emitter {
    const RtpbrConfig& cfg;
    const RtpbrObject* objs;
    int n;
    float round_;
    int max_pairs;
    std::string s;

    static std::string idx(int k) { return std::to_string(k); }

    // `const float dx<k> = pos.x - c;` ... for the world axes `need` (bit mask) of object k
    void emit_delta(int k, int need)
    {
        static const char* axes[3] = { "x", "y", "z" };
        for (int a = 0; a < 3; ++a) {
            if (!(need >> a & 1)) continue;
            s += "    const float d" + std::string(axes[a]) + idx(k) + " = ";
            if (objs[k].position[a] == 0.0f) s += std::string("pos.") + axes[a] + ";\n";
            else s += std::string("pos.") + axes[a] + " - " + flit(objs[k].position[a]) + ";\n";
        }
    }
    std::string local_row(int k, const float* m, int row)
    {
        const std::string K = idx(k);
        const std::string dn[3] = { "dx" + K, "dy" + K, "dz" + K };
        const char* d[3] = { dn[0].c_str(), dn[1].c_str(), dn[2].c_str() };
        return row_expr(m + 3 * row, d);
    }
    // object-space point p<k>
    void emit_point(int k)
    {
        float m[9];
        euler_matrix_deg(objs[k].rotation, m);
        s += "    // object " + idx(k) + ", type " + std::to_string(objs[k].type) + "\n";
        emit_delta(k, 7);
        s += "    vec3 p" + idx(k) + " = V3(" + local_row(k, m, 0) + ", " + local_row(k, m, 1) + ", " + local_row(k, m, 2) + ");\n";
    }
    std::string box_args(int j) const { return flit(objs[j].scale[0]) + ", " + flit(objs[j].scale[1]) + ", " + flit(objs[j].scale[2]); }

    // Distances of the objects `list` (in this order): fills dist[k] / doubled[k].  mode: 1 = with argmin, 0 = distance only,
    // -1 = distance without the bunny MLP (two-stage march), 2 = fast variant (same forms as 0).
    void emit_distances(const std::vector<int>& list, int mode, std::vector<std::string>& dist, std::vector<bool>& doubled)
    {
        int pairs_left = max_pairs;
        const bool sphere_pairs = getenv("RTPBR_SPHERE_PAIRS") == nullptr || atoi(getenv("RTPBR_SPHERE_PAIRS")) != 0;   // tuning knob
        int alu_pairs_left = getenv("RTPBR_ALU_CLAMPS") ? atoi(getenv("RTPBR_ALU_CLAMPS")) : 0;         // tuning knob: pairs with FMNMX clamps
        int clamp_pairs_left = getenv("RTPBR_PACK_CLAMPS") ? atoi(getenv("RTPBR_PACK_CLAMPS")) : 0;   // tuning knob, see sd_box2_ranged_x2
        for (size_t q = 0; q < list.size(); ++q) {
            const int k = list[q];
            const int k1 = q + 1 < list.size() ? list[q + 1] : -1;
            const RtpbrObject& o = objs[k];
            const std::string K = idx(k), P_ = "p" + K;
            if (box_is_ranged(o) && k1 >= 0 && box_is_ranged(objs[k1]) && pairs_left > 0) {
                --pairs_left;
                const std::string K1 = idx(k1);
                if (alu_pairs_left-- > 0) {      // clamps on the ALU pipe (FMNMX), plain distances
                    s += "    float sd" + K + ", sd" + K1 + ";\n";
                    s += "    sd_box_ranged_x2(p" + K + ", " + box_args(k) + ", p" + K1 + ", " + box_args(k1) + ", " + flit(round_) +
                         ", sd" + K + ", sd" + K1 + ");\n";
                    dist[k] = "sd" + K;
                    dist[k1] = "sd" + K1;
                    ++q;
                    continue;
                }
                s += "    float sd" + K + ", sd" + K1 + ";   // twice the distances\n";
                s += std::string("    sd_box2_ranged_x2<") + (clamp_pairs_left-- > 0 ? "true" : "false") + ">(p" + K + ", " + box_args(k) + ", p" + K1 +
                     ", " + box_args(k1) + ", " + flit(2.0f * round_) + ", sd" + K + ", sd" + K1 + ");\n";
                dist[k] = "sd" + K;
                dist[k1] = "sd" + K1;
                doubled[k] = doubled[k1] = true;
                ++q;
                continue;
            }
            if (o.type == RTPBR_SHAPE_SPHERE && k1 >= 0 && objs[k1].type == RTPBR_SHAPE_SPHERE && sphere_pairs) {
                const std::string K1 = idx(k1);
                s += "    float ss" + K + ", ss" + K1 + ";\n";
                s += "    sd_sphere_x2(p" + K + ", " + flit(o.scale[0]) + ", p" + K1 + ", " + flit(objs[k1].scale[0]) + ", ss" + K + ", ss" + K1 + ");\n";
                dist[k] = "ss" + K;
                dist[k1] = "ss" + K1;
                ++q;
                continue;
            }
            switch (o.type) {
            case RTPBR_SHAPE_SPHERE: dist[k] = "sd_sphere(" + P_ + ", " + flit(o.scale[0]) + ")"; break;
            case RTPBR_SHAPE_BOX:
                dist[k] = std::string(box_is_ranged(o) ? "sd_box_ranged" : "sd_box") + "(" + P_ + ", " + box_args(k) + ", " + flit(round_) + ")";
                break;
            case RTPBR_SHAPE_CYLINDER: dist[k] = "sd_cylinder(" + P_ + ", " + flit(o.scale[0]) + ", " + flit(o.scale[1]) + ")"; break;
            case RTPBR_SHAPE_CONE:
                dist[k] = "sd_cone(" + P_ + ", " + flit(o.scale[0]) + ", " + flit(o.scale[1]) + ", " + flit(o.scale[2]) + ")";
                break;
            case RTPBR_SHAPE_PLANE: dist[k] = "sd_plane(" + P_ + ", " + flit(o.scale[1]) + ")"; break;
            case RTPBR_SHAPE_BUNNY:
                s += "    " + P_ + " = mat_mul(P.anim_m, " + P_ + ");\n";
                if (cfg.bunny_bob) s += "    " + P_ + " = " + P_ + " + V3(0.0f, 0.0f, P.anim_bob);\n";
                if (mode == -1) {
                    s += "    const float len" + K + " = length(" + P_ + ");\n    float hb" + K + " = len" + K + " - 0.8f;\n";
                    s += "    if (!(len" + K + " > 1.0f)) { need_mlp = true; pb = " + P_ + "; hb" + K + " = rt_inf(); }\n";
                    dist[k] = "hb" + K;
                } else {
                    dist[k] = "sd_bunny(" + P_ + ")";
                }
                break;
            default: dist[k] = flit(cfg.t_far); break;
            }
        }
    }

    // jit_nearest (mode 1), jit_nearest_dist (0), jit_nearest_partial (-1)
    void emit_full(int mode)
    {
        if (mode == 1) s += "RT_HD float jit_nearest(const KParams& P, vec3 pos, int& index)\n{\n";
        else if (mode == 0) s += "RT_HD float jit_nearest_dist(const KParams& P, vec3 pos)\n{\n";
        else s += "RT_HD float jit_nearest_partial(const KParams& P, vec3 pos, bool& need_mlp, vec3& pb)\n{\n    need_mlp = false;\n    pb = V3(0.0f);\n";
        s += "    float best = " + flit(2.0f * cfg.t_far) + ";   // doubled, like every a<k> below\n";
        if (mode == 1) s += "    int idx = 0;\n";
        std::vector<int> all(n);
        for (int k = 0; k < n; ++k) { all[k] = k; emit_point(k); }
        std::vector<std::string> dist(n);
        std::vector<bool> doubled(n, false);   // dist[k] already holds 2 * distance
        emit_distances(all, mode, dist, doubled);
        // min / argmin in object order, strict '<' (first index wins ties), on DOUBLED distances
        // (2x and the final 0.5x are exact, so every comparison and the result are unchanged)
        for (int k = 0; k < n; ++k) {
            const std::string K = idx(k);
            if (doubled[k]) s += "    const float a" + K + " = fabsf(" + dist[k] + ");\n";
            else s += "    const float h" + K + " = " + dist[k] + ";\n    const float a" + K + " = fabsf(h" + K + " + h" + K + ");\n";
            if (k == 0 && cfg.nearest_seed == 0) s += "    best = a0;\n";
            else if (mode == 1) s += "    if (a" + K + " < best) { best = a" + K + "; idx = " + K + "; }\n";
            else s += "    best = fminf(best, a" + K + ");\n";
        }
        if (mode == 1) s += "    index = idx;\n";
        s += "    return 0.5f * best;\n}\n";
    }

    // jit_nearest_fast: walls as planes inside the fast region, everything else as in jit_nearest_dist.
    // with_index: jit_nearest_fast_idx, the argmin as well (same compare chain as jit_nearest: object order, strict '<')
    void emit_fast(const Analysis& A, bool with_index = false)
    {
        static const char* axes[3] = { "x", "y", "z" };
        if (!with_index) {
            s += "// ok: pos lies in the fast region and outside every wall's slab; then the result has the bits of jit_nearest_dist(pos)\n";
            s += "RT_HD float jit_nearest_fast(const KParams& P, vec3 pos, bool& ok)\n{\n";
        } else {
            s += "// ... and this one the bits and the index of jit_nearest(pos, index)\n";
            s += "RT_HD float jit_nearest_fast_idx(const KParams& P, vec3 pos, bool& ok, int& index)\n{\n";
        }
        std::string in;
        for (int a = 0; a < 3; ++a) {
            const std::string e = A.rc[a] == 0.0f ? std::string("fabsf(pos.") + axes[a] + ")"
                                                  : std::string("fabsf(pos.") + axes[a] + " - " + flit(A.rc[a]) + ")";
            in += (a ? " && " : "") + std::string("(") + e + " <= " + flit(A.rh[a]) + ")";
        }
        s += "    const bool in_region = " + in + ";\n";
        std::vector<bool> is_wall(n, false);
        std::string qmin, hmin;
        for (const Wall& w : A.walls) {
            const int k = w.object;
            is_wall[k] = true;
            float m[9];
            euler_matrix_deg(objs[k].rotation, m);
            int need = 0;
            for (int j = 0; j < 3; ++j) if (m[3 * w.normal + j] != 0.0f) need |= 1 << j;
            s += "    // wall " + idx(k) + ": distance = |p_normal| - half-extent (exact here)\n";
            emit_delta(k, need);
            s += "    const float q" + idx(k) + " = fabsf(" + local_row(k, m, w.normal) + ") - " + flit(objs[k].scale[w.normal]) + ";\n";
            qmin = qmin.empty() ? "q" + idx(k) : "fminf(" + qmin + ", q" + idx(k) + ")";
            if (round_ != 0.0f) {
                s += "    const float g" + idx(k) + " = fabsf(q" + idx(k) + " - " + flit(round_) + ");\n";
                hmin = hmin.empty() ? "g" + idx(k) : "fminf(" + hmin + ", g" + idx(k) + ")";
            }
        }
        s += "    const float qmin = " + qmin + ";\n";
        s += "    ok = in_region && qmin > 0.0f;\n";
        if (round_ != 0.0f) s += "    const float wall = " + hmin + ";\n";
        else s += "    const float wall = qmin;   // every q > 0 when ok\n";
        std::vector<int> rest;
        for (int k = 0; k < n; ++k) if (!is_wall[k]) rest.push_back(k);
        std::vector<std::string> dist(n);
        std::vector<bool> doubled(n, false);
        for (int k : rest) emit_point(k);
        emit_distances(rest, 2, dist, doubled);
        if (with_index) {
            // doubled distances of ALL objects in object order: the compare chain of jit_nearest
            s += "    float best = " + flit(2.0f * cfg.t_far) + ";\n    int idx = 0;\n";
            for (int k = 0; k < n; ++k) {
                const std::string K = idx(k);
                if (is_wall[k]) {
                    const std::string g = round_ != 0.0f ? "g" + K : "fabsf(q" + K + ")";
                    s += "    const float a" + K + " = " + g + " + " + g + ";\n";
                } else if (doubled[k]) {
                    s += "    const float a" + K + " = fabsf(" + dist[k] + ");\n";
                } else {
                    s += "    const float h" + K + " = " + dist[k] + ";\n    const float a" + K + " = fabsf(h" + K + " + h" + K + ");\n";
                }
                if (k == 0 && cfg.nearest_seed == 0) s += "    best = a0;\n";
                else s += "    if (a" + K + " < best) { best = a" + K + "; idx = " + K + "; }\n";
            }
            s += "    index = idx;\n    return 0.5f * best;\n}\n";
            return;
        }
        std::string best;
        if (cfg.nearest_seed != 0) best = flit(2.0f * cfg.t_far);
        for (int k : rest) {
            const std::string K = idx(k);
            if (doubled[k]) s += "    const float a" + K + " = fabsf(" + dist[k] + ");\n";
            else s += "    const float h" + K + " = " + dist[k] + ";\n    const float a" + K + " = fabsf(h" + K + " + h" + K + ");\n";
            best = best.empty() ? "a" + K : "fminf(" + best + ", a" + K + ")";
        }
        if (best.empty()) s += "    return wall;\n}\n";
        else s += "    return fminf(wall, 0.5f * " + best + ");   // = 0.5 * min(2 * wall, ...), exactly\n}\n";
    }
};

// family / marcher / shape set must already have been validated (kernel_supported).
inline Source generate(const RtpbrConfig& cfg, const RtpbrObject* objs, int n, int max_pairs = 1 << 30)
{
    int bunnies = 0;
    for (int k = 0; k < n; ++k) bunnies += objs[k].type == RTPBR_SHAPE_BUNNY;
    const bool bunny = bunnies > 0;
    const char* shapeset = cfg.family == RTPBR_FAMILY_A ? "SHAPESET_BOX" : (bunny ? "SHAPESET_BUNNY" : "SHAPESET_ANALYTIC");
    const char* marcher = cfg.marcher == RTPBR_MARCH_PLAIN ? "MARCH_PLAIN" : (cfg.marcher == RTPBR_MARCH_ENHANCED ? "MARCH_ENHANCED" : "MARCH_SRC");
    const char* family = cfg.family == RTPBR_FAMILY_A ? "FAMILY_A" : (cfg.family == RTPBR_FAMILY_B ? "FAMILY_B" : "FAMILY_C");
    Analysis A = analyse(cfg, objs, n);
    if (cfg.family == RTPBR_FAMILY_C) { A.bounded = A.fast = false; A.walls.clear(); }   // src/: the marched origin is persisted in ray_buffer, every step counts
    if (bunnies == 1 && cfg.marcher == RTPBR_MARCH_ENHANCED) { A.fast = false; A.walls.clear(); }   // the two-stage bunny loop has its own cheap stage
    // Where the analyses pay (measured on B200, profiles/r02_sweeps.md):
    //   * the fast region wins 25 % on the diffuse-only family A, whose resolve phase is small; in the PBR families the
    //     full-code steps of new camera paths (taken in the resolve phase, at ~4 lanes per batch) cost more than the
    //     march loop saves (cornell_box.py 1244 -> 1032 Msamples/s), and the over-relaxed marcher steps INTO wall slabs
    //     all the time (cornell_box_v3 1225 -> 752).  Automatic: family A only.
    //   * t_stop costs a register and a slot word per ray and a dozen instructions per bounce: +7 % where most paths
    //     end in the sky (bunny_sdf_glass), +1 % on tokyo_ibl, -5 % in a closed room (cornell_box.py).  Automatic: when
    //     the scene has a sky (or is family A).
    if (A.fast && knob("RTPBR_JIT_FAST") < 0 && cfg.family != RTPBR_FAMILY_A) { A.fast = false; A.walls.clear(); }
    if (A.bounded && knob("RTPBR_JIT_BBOX") < 0 && cfg.family != RTPBR_FAMILY_A && cfg.sky == RTPBR_SKY_BLACK) A.bounded = false;

    Emitter E{ cfg, objs, n, round_of(cfg), max_pairs, std::string() };
    std::string& s = E.s;
    s += "// generated by raytracingpbr_b200/csrc/jit_codegen.h -- scene-specialised nearest()\n";
    s += "#define RT_JIT_SCENE 1\n";
    // march-loop constants as literals (same values as the parameter block => same comparisons)
    s += "#define RT_K_HIT_EPS " + flit(cfg.hit_eps) + "\n#define RT_K_T_FAR " + flit(cfg.t_far) + "\n#define RT_K_MAX_STEPS " +
         std::to_string(cfg.max_steps) + "\n";
    // configuration constants as literals (RT_CFG, rt_integrator.cuh): same values as fill_config() puts into the parameter block
    s += "#define RT_K_CFG 1\n";
    {
        const std::pair<const char*, int> ints[] = {
            { "bsdf", cfg.bsdf }, { "f0_variant", cfg.f0_variant }, { "normal_mode", cfg.normal_mode }, { "sky", cfg.sky },
            { "relax_guard", cfg.relax_guard }, { "relax_reset", cfg.relax_reset }, { "black_background", cfg.black_background },
            { "primary_miss", cfg.primary_miss }, { "bunny_bob", cfg.bunny_bob != 0 }, { "max_bounces", cfg.max_bounces },
            { "samples_per_pixel", cfg.samples_per_pixel > 0 ? cfg.samples_per_pixel : 1 }, { "width", cfg.width }, { "height", cfg.height },
            { "tiles_per_col", (cfg.height + 7) / 8 } };       // (fill_shard: work items are 4 x 8 tiles)
        for (const auto& kv : ints) s += std::string("#define RT_K_") + kv.first + " " + std::to_string(kv.second) + "\n";
        const std::pair<const char*, float> floats[] = {
            { "relax_w_reset", cfg.relax_w_reset }, { "relax_w0", cfg.relax_w0 }, { "t_start", cfg.t_start }, { "normal_h", cfg.normal_h },
            { "min_dis", cfg.min_dis }, { "visibility_min", cfg.visibility_min }, { "sky_scale", cfg.sky_scale }, { "box_round", cfg.box_round },
            { "pixel_radius", cfg.pixel_radius }, { "quality_per_sample", cfg.quality_per_sample },
            { "inv_max_bounces", (float)(1.0 / (double)cfg.max_bounces) }, { "visibility_max", cfg.visibility_max } };
        for (const auto& kv : floats)
            s += std::string("#define RT_K_") + kv.first + " " + (std::isinf(kv.second) ? (kv.second > 0 ? "rt::rt_inf()" : "(-rt::rt_inf())") : flit(kv.second)) + "\n";
    }
    if (A.bounded) {
        // world box around every surface; margin: see ray_t_stop() (rt_integrator.cuh)
        s += "#define RT_JIT_BBOX 1\n";
        static const char* AX[3] = { "X", "Y", "Z" };
        for (int a = 0; a < 3; ++a)
            s += std::string("#define RT_BB_LO_") + AX[a] + " " + flit((float)A.lo[a]) + "\n#define RT_BB_HI_" + AX[a] + " " + flit((float)A.hi[a]) + "\n";
        s += "#define RT_BB_SCALE " + flit((float)A.scale) + "\n";
    }
    if (A.fast) s += "#define RT_JIT_FAST 1\n";
    // per-ray t_stop in the march loop: always without a fast region; with one, the same test runs where rays drop out
    // of the region (slow_march) unless RTPBR_JIT_TSTOP=1 (tuning knob)
    const bool tstop = A.bounded && (!A.fast || knob("RTPBR_JIT_TSTOP") == 1);
    if (tstop && A.fast) s += "#define RT_JIT_TSTOP 1\n";
    // bunny scenes with the enhanced marcher: the march loop runs the cheap part of nearest() and the MLP
    // as separate stages (pool_kernel.cuh), which needs the third form of the function (mode -1); the split keeps
    // ONE (need_mlp, pb) pair, so it is used for scenes with exactly one bunny
    // PBR families without the bunny: Philox and the normal's primitive switch out of line (code size, rt_math.cuh)
    const int ool = knob("RTPBR_RESOLVE_OOL");                    // tuning knob: -1 = automatic
    if (ool == 1 || (ool < 0 && cfg.family != RTPBR_FAMILY_A && !bunny)) s += "#define RT_RESOLVE_OOL 1\n";
    const bool split = bunnies == 1 && cfg.marcher == RTPBR_MARCH_ENHANCED;
    if (split) s += "#define RT_JIT_SPLIT_BUNNY 1\n";
    s += "#include \"pool_kernel.cuh\"\nnamespace rt {\n";
    E.emit_full(1);
    E.emit_full(0);
    if (split) E.emit_full(-1);
    if (A.fast) { E.emit_fast(A); E.emit_fast(A, true); }
    s += "}  // namespace rt\n\n";
    const std::string variant = std::string("rt::Variant<rt::") + family + ", 0, rt::" + shapeset + ", rt::" + marcher + ", false>";
    s += "extern \"C\" __global__ void __launch_bounds__(rt::kPoolBlock, rt::PoolLaunch<" + variant + ">::kMinBlocks)\n"
         "k_pathtrace_pool_jit(const __grid_constant__ rt::KParams P)\n{\n";
    s += "    rt::pool_body<" + variant + ", rt::kPoolSlots>(P);\n}\n";
    Source out;
    out.text = s;
    out.kernel_name = "k_pathtrace_pool_jit";
    out.fast = A.fast;
    out.tstop = tstop;
    out.relaxed = cfg.marcher != RTPBR_MARCH_PLAIN && !bunny;
    return out;
}

}  // namespace jit
}  // namespace rt
