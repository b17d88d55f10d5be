"""Build (here, cross-compiled) and run (on the GPU box) a device-side comparison of the generated
jit_nearest / jit_nearest_dist against the generic nearest<VAR> on the points real rays visit.

    python tools/jit_device_check.py build     # writes tools/microbench/jit_device_check(.cu)
    tools/microbench/jit_device_check          # on a B200
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from raytracingpbr_b200 import _native as N, scenes  # noqa: E402

SRC = r'''
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#include "%(csrc)s/host_setup.h"
#include "%(csrc)s/rt_integrator.cuh"
namespace rt {
%(func)s
}
using namespace rt;
typedef Variant<FAMILY_A, 0, SHAPESET_BOX, MARCH_PLAIN, false> VAR;

__global__ void k(const __grid_constant__ KParams P, unsigned long long* bad, int spp)
{
    const uint32_t pixel = blockIdx.x * blockDim.x + threadIdx.x;
    if (pixel >= (uint32_t)(P.width * P.height)) return;
    const int i = pixel / P.height, j = pixel %% P.height;
    for (int s = 0; s < spp; ++s) {
        Path p;
        begin_path<VAR>(P, pixel, i, j, (uint32_t)s, p);
        while (begin_bounce<VAR>(P, p)) {
            int status;
            do {
                vec3 pos = at(p.m.ro, p.m.rd, p.m.t);
                int i0, i1;
                float a = nearest<VAR>(P, pos, i0);
                float b = jit_nearest(P, pos, i1);
                float c = jit_nearest_dist(P, pos);
                atomicAdd(&bad[3], 1ull);
                // non-finite points belong to irregular rays, which the product marches with the generic code
                if (finite3(pos) && (a != b || i0 != i1 || a != c)) {
                    if (atomicAdd(&bad[0], 1ull) < 8)
                        printf("pixel %%u s %%d pos=(%%a,%%a,%%a) generic=%%a/%%d jit=%%a/%%d dist=%%a\n", pixel, s, pos.x, pos.y, pos.z, a, i0, b, i1, c);
                }
                status = march_step<VAR>(P, p.m);
            } while (status == MARCH_CONTINUE);
            if (status == MARCH_MISS) break;
            if (!on_hit<VAR>(P, p)) break;
        }
    }
}

int main()
{
    RtpbrConfig cfg; RtpbrCamera cam; RtpbrObject objs[16]; int n = 0;
%(setup)s
    KParams P; memset(&P, 0, sizeof(P));
    fill_config(P, cfg); fill_shard(P, 0, 1, 32); fill_objects(P, objs, n); fill_camera(P, cfg, cam); fill_frame(P, 0);
    std::vector<float> rr = rr_table(cfg);
    float* d_rr; cudaMalloc(&d_rr, rr.size() * 4); cudaMemcpy(d_rr, rr.data(), rr.size() * 4, cudaMemcpyHostToDevice);
    P.rr_prob = d_rr;
    unsigned long long* bad; cudaMallocManaged(&bad, 4 * sizeof(*bad)); memset(bad, 0, 4 * sizeof(*bad));
    k<<<(P.width * P.height + 127) / 128, 128>>>(P, bad, %(spp)d);
    cudaDeviceSynchronize();
    printf("device check: %%llu scene evaluations, %%llu mismatches (%%s)\n", bad[3], bad[0], cudaGetErrorString(cudaGetLastError()));
    return bad[0] != 0;
}
'''


def c_array(v):
    return "{" + ", ".join(repr(float(x)) + "f" for x in v) + "}"


def build(size=1024, spp=64, out_dir=None):
    cfg, objs, cam, _ = scenes.cornell_box_shortest(size, size, max_bounces=8, seed=0)
    nat = [o.to_native() for o in objs]
    src = N.jit_source(cfg, nat)
    body = src[src.index("namespace rt {") + len("namespace rt {"):src.index("}  // namespace rt")]
    setup = ["    memset(&cfg, 0, sizeof(cfg)); memset(&cam, 0, sizeof(cam)); memset(objs, 0, sizeof(objs));"]
    for name, _ in N.RtpbrConfig._fields_:
        v = getattr(cfg, name)
        if name in ("visibility_max",):
            setup.append(f"    cfg.{name} = INFINITY;")
        else:
            setup.append(f"    cfg.{name} = {v!r}{'f' if isinstance(v, float) else ''};")
    c = cam.to_native()
    for name in ("lookfrom", "lookat", "vup"):
        for a in range(3):
            setup.append(f"    cam.{name}[{a}] = {float(getattr(c, name)[a])!r}f;")
    for name in ("vfov", "aspect", "aperture", "focus"):
        setup.append(f"    cam.{name} = {float(getattr(c, name))!r}f;")
    for k, o in enumerate(nat):
        setup.append(f"    objs[{k}].type = {o.type};")
        for name in ("position", "rotation", "scale", "albedo", "emission"):
            for a in range(3):
                setup.append(f"    objs[{k}].{name}[{a}] = {float(getattr(o, name)[a])!r}f;")
        for name in ("roughness", "metallic", "transmission", "ior"):
            setup.append(f"    objs[{k}].{name} = {float(getattr(o, name))!r}f;")
    setup.append(f"    n = {len(nat)};")
    out = os.path.join(out_dir or os.path.join(ROOT, "tools", "microbench"), "jit_device_check.cu")
    with open(out, "w") as f:
        f.write(SRC % dict(csrc=os.path.join(ROOT, "raytracingpbr_b200", "csrc"), func=body, setup="\n".join(setup), spp=spp))
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-fmad=false", "-prec-div=true",
                           "-prec-sqrt=true", "-Xcompiler", "-ffp-contract=off", "-o", out[:-3], out], stderr=subprocess.DEVNULL)
    print("built", out[:-3])
    return out[:-3]


if __name__ == "__main__":
    build()
