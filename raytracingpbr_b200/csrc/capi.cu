// capi.cu -- C-ABI (include/rtpbr.h) over the CUDA kernels.  One context = one GPU, one
// stream, all device memory.  No CPU fallback: every compute call needs a CUDA device.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <vector>

#include "../../include/rtpbr.h"
#include "host_setup.h"
#include "jit.h"
#include "jit_codegen.h"
#include "kernels.h"

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg)
{
    g_last_error = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return fail(RTPBR_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));       \
    } while (0)

// ---- NCCL, resolved at run time so the library loads on machines without it -------------
struct UniqueId { char internal[128]; };   // == ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, UniqueId /* by value */, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Reduce)(const void*, void*, size_t, int, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

bool load_nccl(std::string& why)
{
    if (g_nccl.handle) return true;
    const char* names[] = { getenv("RTPBR_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
    void* h = nullptr;
    for (const char* n : names) {
        if (!n || !*n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) { why = "cannot dlopen libnccl.so.2 (set RTPBR_NCCL_LIB)"; return false; }
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
    g_nccl.Reduce = (decltype(g_nccl.Reduce))dlsym(h, "ncclReduce");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
    g_nccl.GroupStart = (decltype(g_nccl.GroupStart))dlsym(h, "ncclGroupStart");
    g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))dlsym(h, "ncclGroupEnd");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.Reduce || !g_nccl.CommDestroy ||
        !g_nccl.GroupStart || !g_nccl.GroupEnd) {
        why = "libnccl is missing required symbols";
        return false;
    }
    g_nccl.handle = h;
    return true;
}

}  // namespace

struct RtpbrContext {
    RtpbrConfig cfg{};
    RtpbrCamera cam{};
    rt::KParams P{};
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> kernel_events;  // pool
    size_t kernel_events_used = 0;
    float4* d_image_buffer = nullptr;
    float* d_image_pixels = nullptr;
    float* d_ray_buffer = nullptr;
    float* d_diff_buffer = nullptr;   // adaptive sampling (family C): vec2 per pixel
    float* d_diff_pixels = nullptr;   // adaptive sampling: f32 per pixel
    float* d_denoise[2] = { nullptr, nullptr };   // denoise_pixels, double-buffered (allocated by the first rtpbr_denoise)
    int denoise_cur = 0;
    float* d_rr = nullptr;
    float* d_env = nullptr;
    unsigned long long* d_work = nullptr;
    float4* d_scratch = nullptr;
    size_t scratch_bytes = 0;
    unsigned long long* d_counters = nullptr;
    void* d_flush = nullptr;
    bool have_scene = false, have_camera = false, has_bunny = false;
    std::vector<RtpbrObject> scene;
    uint32_t sample_base = 0;
    int sm_count = 0, cc_major = 0, cc_minor = 0, blocks_per_sm = 0;
    bool blocks_per_sm_is_jit = false;
    // scene-specialised kernel (NVRTC); falls back to the ahead-of-time variant when unavailable
    bool jit_enabled = true, jit_stale = true;
    std::unique_ptr<rt::jit::Kernel> jit_kernel;
    int jit_blocks_per_sm = 0;
    int jit_block = rt::kPoolBlock, jit_slots = rt::kPoolSlots, jit_min_blocks = rt::kPoolMinBlocks;   // pool geometry of the NVRTC build
    std::string jit_log = "not built yet";
    std::string jit_key;          // generated source + build options of the loaded kernel
    int jit_churn = 0;            // recent NVRTC compiles (decays per pathtrace): animated scenes fall back to the ahead-of-time kernel
    void* nccl_comm = nullptr;
    int nccl_rank = 0, nccl_nranks = 1;
    bool holds_reduced = false;   // image_buffer holds the sum over ranks (after rtpbr_reduce_tiles): refresh before tracing on
    unsigned long long launches = 0;
};

namespace {

size_t npixels(const RtpbrContext* c) { return (size_t)c->cfg.width * (size_t)c->cfg.height; }

rt::KernelSelect select_kernel(const RtpbrContext* c)
{
    rt::KernelSelect s;
    s.family = c->cfg.family;
    s.marcher = c->cfg.marcher;
    s.nobj = c->P.nobj;
    s.bunny = c->has_bunny;
    s.count = c->cfg.count_work != 0;
    return s;
}

int validate_config(const RtpbrConfig& c)
{
    if (c.width < 1 || c.height < 1 || c.width > 65536 || c.height > 65536) return fail(RTPBR_ERR_ARG, "bad resolution");
    if ((uint64_t)c.width * (uint64_t)c.height > (1ull << 30)) return fail(RTPBR_ERR_ARG, "resolution too large");
    if (c.family < RTPBR_FAMILY_A || c.family > RTPBR_FAMILY_C) return fail(RTPBR_ERR_ARG, "bad family");
    if (c.max_bounces < 1 || c.max_bounces > RTPBR_MAX_BOUNCES) return fail(RTPBR_ERR_ARG, "bad max_bounces");
    if (c.max_steps < 1) return fail(RTPBR_ERR_ARG, "bad max_steps");
    if (!(c.light_quality > 0.f)) return fail(RTPBR_ERR_ARG, "bad light_quality");
    if (c.marcher < RTPBR_MARCH_PLAIN || c.marcher > RTPBR_MARCH_SRC) return fail(RTPBR_ERR_ARG, "bad marcher");
    if (c.family == RTPBR_FAMILY_A && (c.marcher != RTPBR_MARCH_PLAIN || c.box_round != 0.f || c.bsdf != 0))
        return fail(RTPBR_ERR_ARG, "family A (cornell_box_shortest.py) is plain marching, sharp boxes, diffuse only");
    if (c.family == RTPBR_FAMILY_B && (c.marcher == RTPBR_MARCH_SRC || c.bsdf != 1))
        return fail(RTPBR_ERR_UNSUPPORTED, "family B uses the plain or enhanced marcher and bsdf 1");
    if (c.family == RTPBR_FAMILY_C && (c.marcher != RTPBR_MARCH_SRC || c.bsdf != 2))
        return fail(RTPBR_ERR_UNSUPPORTED, "family C (src/) uses the src marcher and bsdf 2");
    if (c.family == RTPBR_FAMILY_C && c.samples_per_pixel < 1) return fail(RTPBR_ERR_ARG, "bad samples_per_pixel");
    if (c.sky < RTPBR_SKY_BLACK || c.sky > RTPBR_SKY_GRADIENT) return fail(RTPBR_ERR_ARG, "bad sky");
    if (c.inner_spp < 0 || (c.inner_spp > 0 && c.family != RTPBR_FAMILY_B))
        return fail(RTPBR_ERR_ARG, "inner_spp (in-kernel sample loop of bunny_sdf.py / bunny_sdf_v2.py) belongs to family B");
    if (c.primary_miss < 0 || c.primary_miss > 2) return fail(RTPBR_ERR_ARG, "bad primary_miss");
    return RTPBR_OK;
}

}  // namespace

extern "C" {

const char* rtpbr_last_error(void) { return g_last_error.c_str(); }
int rtpbr_version(void) { return RTPBR_VERSION; }
int rtpbr_device_count(void)
{
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}
int rtpbr_sizeof_config(void) { return (int)sizeof(RtpbrConfig); }
int rtpbr_sizeof_object(void) { return (int)sizeof(RtpbrObject); }
int rtpbr_sizeof_camera(void) { return (int)sizeof(RtpbrCamera); }

int rtpbr_create(const RtpbrConfig* cfg, int device, RtpbrContext** out)
{
    if (!cfg || !out) return fail(RTPBR_ERR_ARG, "null argument");
    *out = nullptr;
    int rc = validate_config(*cfg);
    if (rc != RTPBR_OK) return rc;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(RTPBR_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) +
                                        " (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(RTPBR_ERR_ARG, "bad device index");
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(RTPBR_ERR_CUDA, "librtpbr is built for sm_100a only; device is sm_" + std::to_string(prop.major) +
                                        std::to_string(prop.minor));
    RtpbrContext* c = new (std::nothrow) RtpbrContext();
    if (!c) return fail(RTPBR_ERR_ARG, "out of host memory");
    c->cfg = *cfg;
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->cc_major = prop.major;
    c->cc_minor = prop.minor;
    rt::fill_config(c->P, c->cfg);
    rt::fill_shard(c->P, 0, 1, 32);
    rt::fill_frame(c->P, 0);
    if (const char* j = getenv("RTPBR_JIT")) c->jit_enabled = atoi(j) != 0;
    c->P.resolve_min = 8;
    c->P.count_mlp = cfg->count_work != 0;
    if (const char* q = getenv("RTPBR_RESOLVE_MIN")) {
        int v = atoi(q);
        if (v >= 1 && v <= 32) c->P.resolve_min = v;
    }
#define CREATE_TRY(expr)                                                                            \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            rtpbr_destroy(c);                                                                       \
            return fail(RTPBR_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));       \
        }                                                                                           \
    } while (0)
    CREATE_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CREATE_TRY(cudaEventCreate(&c->ev_start));
    CREATE_TRY(cudaEventCreate(&c->ev_stop));
    const size_t n = npixels(c);
    CREATE_TRY(cudaMalloc(&c->d_image_buffer, n * sizeof(float4)));
    CREATE_TRY(cudaMalloc(&c->d_image_pixels, n * 3 * sizeof(float)));
    CREATE_TRY(cudaMalloc(&c->d_work, sizeof(unsigned long long)));
    CREATE_TRY(cudaMalloc(&c->d_counters, RTPBR_CNT_COUNT * sizeof(unsigned long long)));
    CREATE_TRY(cudaMemsetAsync(c->d_image_buffer, 0, n * sizeof(float4), c->stream));
    CREATE_TRY(cudaMemsetAsync(c->d_image_pixels, 0, n * 3 * sizeof(float), c->stream));
    if (cfg->family == RTPBR_FAMILY_C) {   // ray_buffer = Ray.field(), zero-initialised like a fresh Taichi field
        CREATE_TRY(cudaMalloc(&c->d_ray_buffer, n * 10 * sizeof(float)));
        CREATE_TRY(cudaMemsetAsync(c->d_ray_buffer, 0, n * 10 * sizeof(float), c->stream));
        if (cfg->adaptive_sampling) {   // fresh Taichi fields are zero: nothing is sampled before the first refresh()
            CREATE_TRY(cudaMalloc(&c->d_diff_buffer, n * 2 * sizeof(float)));
            CREATE_TRY(cudaMalloc(&c->d_diff_pixels, n * sizeof(float)));
            CREATE_TRY(cudaMemsetAsync(c->d_diff_buffer, 0, n * 2 * sizeof(float), c->stream));
            CREATE_TRY(cudaMemsetAsync(c->d_diff_pixels, 0, n * sizeof(float), c->stream));
        }
    }
    CREATE_TRY(cudaMemsetAsync(c->d_counters, 0, RTPBR_CNT_COUNT * sizeof(unsigned long long), c->stream));
    std::vector<float> rr = rt::rr_table(c->cfg);
    CREATE_TRY(cudaMalloc(&c->d_rr, rr.size() * sizeof(float)));
    CREATE_TRY(cudaMemcpyAsync(c->d_rr, rr.data(), rr.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CREATE_TRY(cudaStreamSynchronize(c->stream));
#undef CREATE_TRY
    c->P.image_buffer = c->d_image_buffer;
    c->P.ray_buffer = c->d_ray_buffer;
    c->P.diff_pixels = c->d_diff_pixels;
    c->P.rr_prob = c->d_rr;
    c->P.env = nullptr;
    c->P.env_w = c->P.env_h = 0;
    c->P.work_counter = c->d_work;
    c->P.counters = c->d_counters;
    *out = c;
    return RTPBR_OK;
}

int rtpbr_destroy(RtpbrContext* c)
{
    if (!c) return RTPBR_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    c->jit_kernel.reset();
    if (c->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl_comm);
    for (auto& p : c->kernel_events) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    if (c->ev_start) cudaEventDestroy(c->ev_start);
    if (c->ev_stop) cudaEventDestroy(c->ev_stop);
    cudaFree(c->d_image_buffer);
    cudaFree(c->d_image_pixels);
    cudaFree(c->d_ray_buffer);
    cudaFree(c->d_diff_buffer);
    cudaFree(c->d_diff_pixels);
    cudaFree(c->d_denoise[0]);
    cudaFree(c->d_denoise[1]);
    cudaFree(c->d_rr);
    cudaFree(c->d_env);
    cudaFree(c->d_work);
    cudaFree(c->d_scratch);
    cudaFree(c->d_counters);
    cudaFree(c->d_flush);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return RTPBR_OK;
}

int rtpbr_set_scene(RtpbrContext* c, const RtpbrObject* objects, int n)
{
    if (!c || !objects) return fail(RTPBR_ERR_ARG, "null argument");
    if (n < 1 || n > RTPBR_MAX_OBJECTS) return fail(RTPBR_ERR_ARG, "object count must be in 1..RTPBR_MAX_OBJECTS");
    bool bunny = false;
    for (int k = 0; k < n; ++k) {
        if (objects[k].type < RTPBR_SHAPE_NONE || objects[k].type > RTPBR_SHAPE_BUNNY)
            return fail(RTPBR_ERR_ARG, "bad shape type");
        if (c->cfg.family == RTPBR_FAMILY_A && objects[k].type != RTPBR_SHAPE_BOX)
            return fail(RTPBR_ERR_ARG, "family A (cornell_box_shortest.py) scenes are boxes only");
        bunny = bunny || objects[k].type == RTPBR_SHAPE_BUNNY;
    }
    rt::KernelSelect sel = select_kernel(c);
    sel.bunny = bunny;
    if (!rt::kernel_supported(sel))
        return fail(RTPBR_ERR_UNSUPPORTED, "no kernel variant for this family / marcher / shape combination "
                                           "(the neural bunny needs family B with the enhanced marcher)");
    // The specialised kernel bakes the GEOMETRY into its instruction stream (materials stay in the parameter block):
    // re-uploading an unchanged scene, or changing materials only, keeps the loaded kernel.
    bool same_geometry = c->have_scene && (int)c->scene.size() == n;
    for (int k = 0; k < n && same_geometry; ++k) {
        const RtpbrObject& a = c->scene[k];
        const RtpbrObject& b = objects[k];
        same_geometry = a.type == b.type && memcmp(a.position, b.position, sizeof(a.position)) == 0 &&
                        memcmp(a.rotation, b.rotation, sizeof(a.rotation)) == 0 && memcmp(a.scale, b.scale, sizeof(a.scale)) == 0;
    }
    rt::fill_objects(c->P, objects, n);
    c->has_bunny = bunny;
    c->scene.assign(objects, objects + n);
    if (!same_geometry) {
        if (c->have_scene && c->jit_churn < 24) c->jit_churn += 2;   // see ensure_jit
        c->jit_stale = true;
        c->blocks_per_sm = 0;  // kernel variant may change with the object count
    }
    c->have_scene = true;
    return RTPBR_OK;
}

int rtpbr_set_camera(RtpbrContext* c, const RtpbrCamera* cam)
{
    if (!c || !cam) return fail(RTPBR_ERR_ARG, "null argument");
    c->cam = *cam;
    rt::fill_camera(c->P, c->cfg, c->cam);
    c->have_camera = true;
    return RTPBR_OK;
}

int rtpbr_set_envmap(RtpbrContext* c, const float* rgb, int w, int h)
{
    if (!c || !rgb || w < 1 || h < 1) return fail(RTPBR_ERR_ARG, "bad envmap");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (c->d_env) { cudaFree(c->d_env); c->d_env = nullptr; }
    const size_t bytes = (size_t)w * h * 3 * sizeof(float);
    CUDA_TRY(cudaMalloc(&c->d_env, bytes));
    CUDA_TRY(cudaMemcpyAsync(c->d_env, rgb, bytes, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->P.env = c->d_env;
    c->P.env_w = w;
    c->P.env_h = h;
    return RTPBR_OK;
}

int rtpbr_set_frame(RtpbrContext* c, int frame)
{
    if (!c) return fail(RTPBR_ERR_ARG, "null context");
    rt::fill_frame(c->P, frame);
    return RTPBR_OK;
}

int rtpbr_set_sample_base(RtpbrContext* c, uint32_t sample_base)
{
    if (!c) return fail(RTPBR_ERR_ARG, "null context");
    c->sample_base = sample_base;
    return RTPBR_OK;
}

int rtpbr_set_shard(RtpbrContext* c, int rank, int nranks, int band)
{
    if (!c) return fail(RTPBR_ERR_ARG, "null context");
    if (nranks < 1 || rank < 0 || rank >= nranks || band < 1) return fail(RTPBR_ERR_ARG, "bad shard");
    rt::fill_shard(c->P, rank, nranks, band);
    return RTPBR_OK;
}

int rtpbr_refresh(RtpbrContext* c)
{
    if (!c) return fail(RTPBR_ERR_ARG, "null context");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemsetAsync(c->d_image_buffer, 0, npixels(c) * sizeof(float4), c->stream));
    c->holds_reduced = false;
    if (c->d_ray_buffer) CUDA_TRY(rt::launch_refresh_depth(c->d_ray_buffer, (int)npixels(c), c->stream));
    if (c->d_diff_buffer) CUDA_TRY(rt::launch_refresh_adaptive(c->d_diff_buffer, c->d_diff_pixels, (int)npixels(c), c->stream));
    return RTPBR_OK;
}

// Pool geometry and scheduling policy of an NVRTC build.  Environment knobs win; otherwise kernels with a fast region
// (family A: short march step, small resolve phase) run 3 CTAs/SM x 88 slots per warp, regenerate paths (and continue
// rays that dropped out of the region, and end the paths of rays that missed) in batches of their own -- when 28 slots
// wait or 12 lanes idle -- and leave the march loop only when 6 lanes have finished
// (profiles/r02_sweeps.md); everything else keeps
// the ahead-of-time geometry.
struct JitBuild {
    int block, slots, min_blocks;
    std::vector<std::string> defs;
};
static JitBuild jit_build_options(const rt::jit::Source& src)
{
    auto env_int = [](const char* name, int lo, int hi, int dflt) {
        const char* v = getenv(name);
        if (!v || !*v) return dflt;
        const int x = atoi(v);
        return x >= lo && x <= hi ? x : dflt;
    };
    JitBuild b;
    b.slots = env_int("RTPBR_POOL_SLOTS", 32, 128, src.fast ? 88 : rt::kPoolSlots);   // 3 x 75.6 KB: the most that fits 3 CTAs/SM
    if (b.slots % 4 != 0) b.slots = rt::kPoolSlots;
    b.block = env_int("RTPBR_POOL_BLOCK", 32, 1024, rt::kPoolBlock);
    if (b.block % 32 != 0) b.block = rt::kPoolBlock;
    // 3 CTAs/SM (<= 85 registers): the fast-region kernels, and the over-relaxed marchers, whose w / s / d state spills at
    // 64 registers (tokyo_ibl +4.6 %, scene_demo +6 %, cornell_box_v3 +3.9 %, src +0.9 %; plain PBR marcher -1 %: stays at 4)
    b.min_blocks = env_int("RTPBR_POOL_MIN_BLOCKS", 1, 16, (src.fast || src.relaxed) ? 3 : rt::kPoolMinBlocks);
    const int regen_min = env_int("RTPBR_REGEN_MIN", 0, 32, src.fast ? 28 : 0);
    const int regen_idle = env_int("RTPBR_REGEN_IDLE", 1, 32, src.fast ? 12 : 1);
    const int fin_min = env_int("RTPBR_FIN_MIN", 1, 32, src.fast ? 6 : 1);
    b.defs = { "-DRT_POOL_BLOCK=" + std::to_string(b.block), "-DRT_POOL_SLOTS=" + std::to_string(b.slots),
               "-DRT_POOL_MIN_BLOCKS=" + std::to_string(b.min_blocks) };
    if (regen_min > 0) b.defs.push_back("-DRT_REGEN_MIN=" + std::to_string(regen_min));
    if (regen_idle > 1) b.defs.push_back("-DRT_REGEN_IDLE=" + std::to_string(regen_idle));
    if (fin_min > 1) b.defs.push_back("-DRT_FIN_MIN=" + std::to_string(fin_min));
    if (const char* v = getenv("RTPBR_SLOW_FIRST")) {            // tuning knob: see pool_kernel.cuh RT_SLOW_FIRST
        if (atoi(v) == 0) b.defs.push_back("-DRT_SLOW_FIRST=0");
    }
    if (const char* v = getenv("RTPBR_MARCH_UNROLL")) {
        if (atoi(v) == 2) b.defs.push_back("-DRT_MARCH_VOTE_EVERY_2=1");
    }
    if (const char* v = getenv("RTPBR_POOL_MIN_BLOCKS_BUNNY")) {
        const int x = atoi(v);
        if (x >= 1 && x <= 16) b.defs.push_back("-DRT_POOL_MIN_BLOCKS_BUNNY=" + std::to_string(x));
    }
    if (const char* v = getenv("RTPBR_SIN4_INLINE")) {          // tuning knob: the MLP's sine routine inlined at its 12 call sites
        if (atoi(v) != 0) b.defs.push_back("-DRT_SIN4_INLINE=1");
    }
    return b;
}

// Launch sequence of one rtpbr_pathtrace call.
//   families A/B, pool kernel: the spp are processed in chunks sized so that the per-sample scratch
//   buffer ([pixel items][chunk] float4) stays within the scratch budget (RTPBR_SCRATCH_MB, default
//   4096); per chunk: k_pathtrace_pool (work item = one path) then k_fold_samples (ordered sum).
//   family C / simple kernel: one launch.
// Build (or reuse) the scene-specialised kernel.  Any failure disables JIT for this context and
// leaves the ahead-of-time kernel in charge; the reason is kept for rtpbr_jit_status().
static void ensure_jit(RtpbrContext* c)
{
    if (c->jit_churn > 0) c->jit_churn--;
    if (!c->jit_enabled || !c->jit_stale) return;
    if (c->cfg.count_work) {
        c->jit_stale = false;
        c->jit_kernel.reset();
        c->jit_log = "disabled: count_work uses the ahead-of-time counting kernel";
        return;
    }
    // A scene whose geometry changes every few launches (animation through set_scene) would pay ~2 s of NVRTC per
    // frame: while that goes on the ahead-of-time kernel renders (same bits); the specialised one comes back once the
    // geometry has been stable for a few launches.
    if (c->jit_churn > 8) {
        if (c->jit_kernel) { cudaStreamSynchronize(c->stream); c->jit_kernel.reset(); c->jit_key.clear(); }
        c->jit_log = "geometry changes every few launches: ahead-of-time kernel until it settles";
        c->blocks_per_sm = 0;
        return;                                   // stays stale: retried at the next launch
    }
    c->jit_stale = false;
    int max_pairs = 1 << 30;      // RTPBR_JIT_PAIRS: how many box pairs use the packed f32x2 form (tuning knob)
    if (const char* v = getenv("RTPBR_JIT_PAIRS")) max_pairs = atoi(v) < 0 ? 0 : atoi(v);
    const rt::jit::Source src = rt::jit::generate(c->cfg, c->scene.data(), (int)c->scene.size(), max_pairs);
    std::shared_ptr<std::vector<char>> cubin;
    std::string log;
    const JitBuild build = jit_build_options(src);
    c->jit_slots = build.slots;
    c->jit_block = build.block;
    c->jit_min_blocks = build.min_blocks;
    const size_t smem = rt::pool_smem_bytes_for(c->jit_block, c->jit_slots);
    const std::vector<std::string>& defs = build.defs;
    std::string key = src.text;
    for (const std::string& d : defs) key += "\n" + d;
    if (c->jit_kernel && key == c->jit_key) return;          // same translation unit as the loaded kernel
    if (c->jit_kernel) {
        cudaStreamSynchronize(c->stream);                    // launches of the old module may still be queued
        c->jit_kernel.reset();
    }
    c->jit_key.clear();
    if (!rt::jit::compile(src.text, rt::jit::default_include_dir(), cubin, log, defs)) {
        c->jit_enabled = false;
        c->jit_log = "NVRTC failed, using the ahead-of-time kernel: " + log;
        fprintf(stderr, "librtpbr: %s\n", c->jit_log.c_str());
        return;
    }
    std::unique_ptr<rt::jit::Kernel> k(new rt::jit::Kernel());
    std::string err;
    if (!rt::jit::load(*cubin, src.kernel_name.c_str(), smem, *k, err) ||
        !rt::jit::occupancy(*k, c->jit_block, smem, &c->jit_blocks_per_sm, err) || c->jit_blocks_per_sm < 1) {
        c->jit_enabled = false;
        c->jit_log = "loading the specialised kernel failed, using the ahead-of-time kernel: " + err;
        fprintf(stderr, "librtpbr: %s\n", c->jit_log.c_str());
        return;
    }
    c->jit_log = "scene-specialised kernel active (NVRTC " + rt::jit::nvrtc_version() + ", " + std::to_string(k->registers) + " registers, " +
                 std::to_string(c->jit_blocks_per_sm) + " CTAs/SM x " + std::to_string(c->jit_block) + " threads, " +
                 std::to_string(c->jit_slots) + " slots/warp)";
    c->jit_kernel = std::move(k);
    c->jit_key = key;
    if (log != "cached" && c->jit_churn < 24) c->jit_churn += 4;   // an actual NVRTC compile
}

static int launch_pool_chunk(RtpbrContext* c, const rt::KernelSelect& sel, std::pair<cudaEvent_t, cudaEvent_t>& ev,
                             unsigned long long items)
{
    ensure_jit(c);
    if (c->jit_kernel) {
        long long grid = (long long)c->sm_count * c->jit_blocks_per_sm;
        const long long per_cta = (long long)(c->jit_block / 32) * c->jit_slots;
        const long long need = (long long)((items + per_cta - 1) / per_cta);
        if (grid > need) grid = need;
        CUDA_TRY(cudaMemsetAsync(c->d_work, 0, sizeof(unsigned long long), c->stream));
        CUDA_TRY(cudaEventRecord(ev.first, c->stream));
        std::string err;
        if (!rt::jit::launch(*c->jit_kernel, c->P, (int)grid, c->jit_block, rt::pool_smem_bytes_for(c->jit_block, c->jit_slots),
                             c->stream, err))
            return fail(RTPBR_ERR_CUDA, err);
        CUDA_TRY(cudaEventRecord(ev.second, c->stream));
        c->blocks_per_sm = c->jit_blocks_per_sm;
        c->blocks_per_sm_is_jit = true;
        return RTPBR_OK;
    }
    if (c->blocks_per_sm == 0 || c->blocks_per_sm_is_jit) {
        CUDA_TRY(rt::pool_occupancy(sel, &c->blocks_per_sm));
        c->blocks_per_sm_is_jit = false;
        if (c->blocks_per_sm < 1) return fail(RTPBR_ERR_CUDA, "pool kernel does not fit on an SM");
    }
    long long grid = (long long)c->sm_count * c->blocks_per_sm;
    const long long per_cta = (long long)(rt::kPoolBlock / 32) * rt::kPoolSlots;   // slots one CTA keeps in flight
    const long long need = (long long)((items + per_cta - 1) / per_cta);
    if (grid > need) grid = need;
    CUDA_TRY(cudaMemsetAsync(c->d_work, 0, sizeof(unsigned long long), c->stream));
    CUDA_TRY(cudaEventRecord(ev.first, c->stream));
    CUDA_TRY(rt::launch_pathtrace_pool(sel, c->P, (int)grid, c->stream));
    CUDA_TRY(cudaEventRecord(ev.second, c->stream));
    return RTPBR_OK;
}

// Per-launch event pairs for rtpbr_kernel_time().  The pool is bounded: a caller that never collects (an interactive
// render loop) wraps around after kMaxKernelEvents launches and overwrites the oldest pairs.
constexpr size_t kMaxKernelEvents = 4096;
static int next_event_pair(RtpbrContext* c, std::pair<cudaEvent_t, cudaEvent_t>** out)
{
    if (c->kernel_events_used == kMaxKernelEvents) c->kernel_events_used = 0;
    if (c->kernel_events_used == c->kernel_events.size()) {
        cudaEvent_t a, b;
        CUDA_TRY(cudaEventCreate(&a));
        CUDA_TRY(cudaEventCreate(&b));
        c->kernel_events.emplace_back(a, b);
    }
    *out = &c->kernel_events[c->kernel_events_used++];
    return RTPBR_OK;
}

int rtpbr_pathtrace(RtpbrContext* c, int spp)
{
    if (!c) return fail(RTPBR_ERR_ARG, "null context");
    if (spp < 1) return fail(RTPBR_ERR_ARG, "spp must be >= 1");
    if (!c->have_scene || !c->have_camera) return fail(RTPBR_ERR_STATE, "set_scene and set_camera must precede pathtrace");
    if (c->holds_reduced)
        return fail(RTPBR_ERR_STATE, "image_buffer holds the sum over all ranks (rtpbr_reduce_tiles): accumulating this rank's "
                                     "shard on top of it would count the other ranks twice at the next reduce; call rtpbr_refresh first");
    CUDA_TRY(cudaSetDevice(c->device));
    const rt::KernelSelect sel = select_kernel(c);
    if (c->P.total_work == 0) {  // this rank owns no columns
        c->sample_base += (uint32_t)spp;
        return RTPBR_OK;
    }
    std::pair<cudaEvent_t, cudaEvent_t>* ev = nullptr;
    int rc;
    if (c->cfg.kernel == RTPBR_KERNEL_SIMPLE || c->cfg.inner_spp > 0) {   // (the in-kernel sample loop is pixel-granular)
        c->P.spp = spp;
        c->P.sample_base = c->sample_base;
        if ((rc = next_event_pair(c, &ev)) != RTPBR_OK) return rc;
        CUDA_TRY(cudaEventRecord(ev->first, c->stream));
        CUDA_TRY(rt::launch_pathtrace_simple(sel, c->P, c->stream));
        CUDA_TRY(cudaEventRecord(ev->second, c->stream));
        c->launches++;
    } else if (c->cfg.family == RTPBR_FAMILY_C) {
        c->P.spp = spp;
        c->P.sample_base = c->sample_base;
        if ((rc = next_event_pair(c, &ev)) != RTPBR_OK) return rc;
        if ((rc = launch_pool_chunk(c, sel, *ev, c->P.total_work)) != RTPBR_OK) return rc;
        c->launches++;
    } else {
        size_t budget = (size_t)4096 << 20;
        if (const char* e = getenv("RTPBR_SCRATCH_MB")) {
            long long v = atoll(e);
            if (v >= 1) budget = (size_t)v << 20;
        }
        long long chunk = (long long)(budget / ((size_t)c->P.total_work * sizeof(float4)));
        if (chunk < 1) chunk = 1;
        if (chunk > spp) chunk = spp;
        const size_t need_bytes = (size_t)c->P.total_work * (size_t)chunk * sizeof(float4);
        if (need_bytes > c->scratch_bytes) {
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            if (c->d_scratch) { cudaFree(c->d_scratch); c->d_scratch = nullptr; c->scratch_bytes = 0; }
            CUDA_TRY(cudaMalloc(&c->d_scratch, need_bytes));
            c->scratch_bytes = need_bytes;
        }
        c->P.scratch = c->d_scratch;
        for (int done = 0; done < spp; done += (int)chunk) {
            const int n = spp - done < (int)chunk ? spp - done : (int)chunk;
            c->P.spp = n;
            c->P.sample_base = c->sample_base + (uint32_t)done;
            if ((rc = next_event_pair(c, &ev)) != RTPBR_OK) return rc;
            if ((rc = launch_pool_chunk(c, sel, *ev, (unsigned long long)c->P.total_work * (unsigned long long)n)) != RTPBR_OK)
                return rc;
            CUDA_TRY(rt::launch_fold_samples(c->P, c->stream));
            c->launches += 2;
        }
    }
    c->sample_base += (uint32_t)spp;
    return RTPBR_OK;
}

int rtpbr_post_process(RtpbrContext* c, int mode, float exposure, double gamma)
{
    if (!c) return fail(RTPBR_ERR_ARG, "null context");
    if (mode < 0 || mode > 3 || !(gamma > 0.0)) return fail(RTPBR_ERR_ARG, "bad tonemap arguments");
    CUDA_TRY(cudaSetDevice(c->device));
    if (c->cfg.family == RTPBR_FAMILY_C && mode == 2) {
        // src/postprocessor.py:24-43 under the fp32 contract (feeds adaptive sampling)
        CUDA_TRY(rt::launch_post_process_src(c->d_image_buffer, c->d_image_pixels, c->d_diff_buffer, c->d_diff_pixels,
                                             (int)npixels(c), exposure, (float)(1.0 / gamma), c->stream));
        return RTPBR_OK;
    }
    CUDA_TRY(rt::launch_post_process(c->d_image_buffer, c->d_image_pixels, (int)npixels(c), mode, exposure,
                                     (float)(1.0 / gamma), c->stream));
    return RTPBR_OK;
}

int rtpbr_denoise(RtpbrContext* c, float threshold)
{
    if (!c) return fail(RTPBR_ERR_ARG, "null context");
    CUDA_TRY(cudaSetDevice(c->device));
    const size_t bytes = npixels(c) * 3 * sizeof(float);
    if (!c->d_denoise[0]) {
        CUDA_TRY(cudaMalloc(&c->d_denoise[0], bytes));
        if (cudaMalloc(&c->d_denoise[1], bytes) != cudaSuccess) {
            cudaFree(c->d_denoise[0]);
            c->d_denoise[0] = nullptr;
            return fail(RTPBR_ERR_CUDA, "cudaMalloc(denoise_pixels) failed");
        }
        CUDA_TRY(cudaMemsetAsync(c->d_denoise[0], 0, bytes, c->stream));
        c->denoise_cur = 0;
    }
    const int prev = c->denoise_cur, next = 1 - prev;
    CUDA_TRY(rt::launch_denoise(c->d_image_pixels, c->d_denoise[prev], c->d_denoise[next], c->cfg.width, c->cfg.height, threshold, c->stream));
    c->denoise_cur = next;
    return RTPBR_OK;
}

static int buffer_of(RtpbrContext* c, int which, void** ptr, size_t* bytes)
{
    switch (which) {
    case RTPBR_BUF_IMAGE_BUFFER: *ptr = c->d_image_buffer; *bytes = npixels(c) * sizeof(float4); return RTPBR_OK;
    case RTPBR_BUF_IMAGE_PIXELS: *ptr = c->d_image_pixels; *bytes = npixels(c) * 3 * sizeof(float); return RTPBR_OK;
    case RTPBR_BUF_RAY_BUFFER:
        if (!c->d_ray_buffer) return fail(RTPBR_ERR_STATE, "ray_buffer exists in family C only");
        *ptr = c->d_ray_buffer; *bytes = npixels(c) * 10 * sizeof(float); return RTPBR_OK;
    case RTPBR_BUF_DIFF_BUFFER:
        if (!c->d_diff_buffer) return fail(RTPBR_ERR_STATE, "diff_buffer exists with adaptive_sampling only");
        *ptr = c->d_diff_buffer; *bytes = npixels(c) * 2 * sizeof(float); return RTPBR_OK;
    case RTPBR_BUF_DIFF_PIXELS:
        if (!c->d_diff_pixels) return fail(RTPBR_ERR_STATE, "diff_pixels exists with adaptive_sampling only");
        *ptr = c->d_diff_pixels; *bytes = npixels(c) * sizeof(float); return RTPBR_OK;
    case RTPBR_BUF_DENOISE_PIXELS:
        if (!c->d_denoise[0]) return fail(RTPBR_ERR_STATE, "denoise_pixels exists after the first rtpbr_denoise");
        *ptr = c->d_denoise[c->denoise_cur]; *bytes = npixels(c) * 3 * sizeof(float); return RTPBR_OK;
    default: return fail(RTPBR_ERR_ARG, "unknown buffer");
    }
}

int rtpbr_download(RtpbrContext* c, int which, void* host, size_t bytes)
{
    if (!c || !host) return fail(RTPBR_ERR_ARG, "null argument");
    void* d; size_t n;
    int rc = buffer_of(c, which, &d, &n);
    if (rc != RTPBR_OK) return rc;
    if (bytes != n) return fail(RTPBR_ERR_ARG, "size mismatch: expected " + std::to_string(n) + " bytes");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpyAsync(host, d, n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return RTPBR_OK;
}

int rtpbr_upload(RtpbrContext* c, int which, const void* host, size_t bytes)
{
    if (!c || !host) return fail(RTPBR_ERR_ARG, "null argument");
    void* d; size_t n;
    int rc = buffer_of(c, which, &d, &n);
    if (rc != RTPBR_OK) return rc;
    if (bytes != n) return fail(RTPBR_ERR_ARG, "size mismatch: expected " + std::to_string(n) + " bytes");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpyAsync(d, host, n, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return RTPBR_OK;
}

int rtpbr_sync(RtpbrContext* c)
{
    if (!c) return fail(RTPBR_ERR_ARG, "null context");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return RTPBR_OK;
}

int rtpbr_alloc_host(size_t bytes, void** out)
{
    if (!out || bytes == 0) return fail(RTPBR_ERR_ARG, "bad argument");
    *out = nullptr;
    CUDA_TRY(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return RTPBR_OK;
}

int rtpbr_free_host(void* ptr)
{
    if (ptr) CUDA_TRY(cudaFreeHost(ptr));
    return RTPBR_OK;
}

int rtpbr_flush_l2(RtpbrContext* c)
{
    if (!c) return fail(RTPBR_ERR_ARG, "null context");
    CUDA_TRY(cudaSetDevice(c->device));
    const size_t bytes = (size_t)256 << 20;   // > 126 MB L2
    if (!c->d_flush) CUDA_TRY(cudaMalloc(&c->d_flush, bytes));
    CUDA_TRY(cudaMemsetAsync(c->d_flush, 0, bytes, c->stream));
    return RTPBR_OK;
}

int rtpbr_timer_start(RtpbrContext* c)
{
    if (!c) return fail(RTPBR_ERR_ARG, "null context");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaEventRecord(c->ev_start, c->stream));
    return RTPBR_OK;
}

int rtpbr_timer_stop(RtpbrContext* c, float* elapsed_ms)
{
    if (!c || !elapsed_ms) return fail(RTPBR_ERR_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaEventRecord(c->ev_stop, c->stream));
    CUDA_TRY(cudaEventSynchronize(c->ev_stop));
    CUDA_TRY(cudaEventElapsedTime(elapsed_ms, c->ev_start, c->ev_stop));
    return RTPBR_OK;
}

int rtpbr_kernel_time(RtpbrContext* c, float* kernel_ms, int* launches)
{
    if (!c || !kernel_ms) return fail(RTPBR_ERR_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    float total = 0.f;
    for (size_t k = 0; k < c->kernel_events_used; ++k) {
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, c->kernel_events[k].first, c->kernel_events[k].second));
        total += ms;
    }
    *kernel_ms = total;
    if (launches) *launches = (int)c->kernel_events_used;
    c->kernel_events_used = 0;
    return RTPBR_OK;
}

int rtpbr_get_counters(RtpbrContext* c, uint64_t out[RTPBR_CNT_COUNT])
{
    if (!c || !out) return fail(RTPBR_ERR_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    unsigned long long h[RTPBR_CNT_COUNT];
    CUDA_TRY(cudaMemcpyAsync(h, c->d_counters, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < RTPBR_CNT_COUNT; ++k) out[k] = h[k];
    out[RTPBR_CNT_LAUNCHES] = c->launches;
    return RTPBR_OK;
}

int rtpbr_device_info(RtpbrContext* c, int* sm_count, int* cc_major, int* cc_minor, int* blocks_per_sm)
{
    if (!c) return fail(RTPBR_ERR_ARG, "null context");
    if (sm_count) *sm_count = c->sm_count;
    if (cc_major) *cc_major = c->cc_major;
    if (cc_minor) *cc_minor = c->cc_minor;
    if (blocks_per_sm) *blocks_per_sm = c->blocks_per_sm;
    return RTPBR_OK;
}

int rtpbr_device_ptr(RtpbrContext* c, int which, uint64_t* ptr)
{
    if (!c || !ptr) return fail(RTPBR_ERR_ARG, "null argument");
    void* d; size_t n;
    int rc = buffer_of(c, which, &d, &n);
    if (rc != RTPBR_OK) return rc;
    *ptr = (uint64_t)(uintptr_t)d;
    return RTPBR_OK;
}

// ------------------------------------------------------------------------------ JIT
int rtpbr_set_jit(RtpbrContext* c, int enable)
{
    if (!c) return fail(RTPBR_ERR_ARG, "null context");
    c->jit_enabled = enable != 0;
    c->jit_stale = true;
    if (!c->jit_enabled) {
        if (c->jit_kernel) cudaStreamSynchronize(c->stream);
        c->jit_kernel.reset();
        c->jit_key.clear();
        c->jit_log = "disabled by rtpbr_set_jit";
        c->blocks_per_sm = 0;
    }
    return RTPBR_OK;
}

int rtpbr_jit_status(RtpbrContext* c, char* buf, size_t cap)
{
    if (!c) return fail(RTPBR_ERR_ARG, "null context");
    if (buf && cap > 0) {
        strncpy(buf, c->jit_log.c_str(), cap - 1);
        buf[cap - 1] = '\0';
    }
    return c->jit_kernel ? 1 : 0;
}

static int jit_validate(const RtpbrConfig* cfg, const RtpbrObject* objects, int n)
{
    if (!cfg || !objects) return fail(RTPBR_ERR_ARG, "null argument");
    if (n < 1 || n > RTPBR_MAX_OBJECTS) return fail(RTPBR_ERR_ARG, "object count must be in 1..RTPBR_MAX_OBJECTS");
    return validate_config(*cfg);
}

long long rtpbr_jit_generate(const RtpbrConfig* cfg, const RtpbrObject* objects, int n, char* buf, size_t cap)
{
    int rc = jit_validate(cfg, objects, n);
    if (rc != RTPBR_OK) return rc;
    const rt::jit::Source src = rt::jit::generate(*cfg, objects, n);
    if (buf && cap > 0) {
        strncpy(buf, src.text.c_str(), cap - 1);
        buf[cap - 1] = '\0';
    }
    return (long long)src.text.size();
}

int rtpbr_jit_compile_check(const RtpbrConfig* cfg, const RtpbrObject* objects, int n, char* log, size_t cap)
{
    int rc = jit_validate(cfg, objects, n);
    if (rc != RTPBR_OK) return rc;
    const rt::jit::Source src = rt::jit::generate(*cfg, objects, n);
    std::shared_ptr<std::vector<char>> cubin;
    std::string l;
    const bool ok = rt::jit::compile(src.text, rt::jit::default_include_dir(), cubin, l, jit_build_options(src).defs);   // the options rtpbr_pathtrace uses
    if (log && cap > 0) {
        strncpy(log, l.c_str(), cap - 1);
        log[cap - 1] = '\0';
    }
    if (!ok) return fail(RTPBR_ERR_UNSUPPORTED, l);
    return RTPBR_OK;
}

// ------------------------------------------------------------------------------ NCCL
int rtpbr_nccl_unique_id(void* id128)
{
    if (!id128) return fail(RTPBR_ERR_ARG, "null argument");
    std::string why;
    if (!load_nccl(why)) return fail(RTPBR_ERR_NCCL, why);
    int r = g_nccl.GetUniqueId(id128);
    if (r != 0) return fail(RTPBR_ERR_NCCL, std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r));
    return RTPBR_OK;
}

int rtpbr_nccl_init(RtpbrContext* c, const void* id128, int rank, int nranks)
{
    if (!c || !id128) return fail(RTPBR_ERR_ARG, "null argument");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(RTPBR_ERR_ARG, "bad rank");
    std::string why;
    if (!load_nccl(why)) return fail(RTPBR_ERR_NCCL, why);
    CUDA_TRY(cudaSetDevice(c->device));
    UniqueId id;
    memcpy(&id, id128, sizeof(id));
    int r = g_nccl.CommInitRank(&c->nccl_comm, nranks, id, rank);
    if (r != 0) return fail(RTPBR_ERR_NCCL, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r));
    c->nccl_rank = rank;
    c->nccl_nranks = nranks;
    return RTPBR_OK;
}

// Sum of the per-rank accumulation buffers.  Every pixel is non-zero on exactly one rank
// (the others hold +0.0f), so the fp32 sum is exact and equals the single-GPU buffer bit for bit.
int rtpbr_reduce_tiles(RtpbrContext* c, int root)
{
    if (!c) return fail(RTPBR_ERR_ARG, "null context");
    if (!c->nccl_comm) {
        if (c->P.nranks == 1) return RTPBR_OK;                 // one rank owns every column: nothing to reduce
        return fail(RTPBR_ERR_STATE, "this context renders one of " + std::to_string(c->P.nranks) +
                                         " shards (rtpbr_set_shard) but rtpbr_nccl_init has not been called");
    }
    if (c->P.nranks != c->nccl_nranks || c->P.rank != c->nccl_rank)
        return fail(RTPBR_ERR_STATE, "rtpbr_set_shard and rtpbr_nccl_init disagree about rank / number of ranks");
    CUDA_TRY(cudaSetDevice(c->device));
    const size_t count = npixels(c) * 4;
    int r;
    if (root < 0)
        r = g_nccl.AllReduce(c->d_image_buffer, c->d_image_buffer, count, /*ncclFloat32*/ 7, /*ncclSum*/ 0, c->nccl_comm,
                             c->stream);
    else
        r = g_nccl.Reduce(c->d_image_buffer, c->d_image_buffer, count, 7, 0, root, c->nccl_comm, c->stream);
    if (r != 0) return fail(RTPBR_ERR_NCCL, std::string("nccl reduce: ") + g_nccl.GetErrorString(r));
    if (root < 0 || root == c->nccl_rank) c->holds_reduced = true;
    return RTPBR_OK;
}

// ------------------------------------------------------------------------------ single-process multi-GPU
}  // extern "C"

struct RtpbrMulti {
    std::vector<RtpbrContext*> ctx;
    int band = 4;
};

extern "C" {

int rtpbr_multi_create(const RtpbrConfig* cfg, const int* devices, int n, int band, RtpbrMulti** out)
{
    if (!cfg || !devices || !out || n < 1 || n > 64 || band < 1) return fail(RTPBR_ERR_ARG, "bad argument");
    *out = nullptr;
    for (int a = 0; a < n; ++a)
        for (int b = a + 1; b < n; ++b)
            if (devices[a] == devices[b]) return fail(RTPBR_ERR_ARG, "the same device appears twice");
    std::unique_ptr<RtpbrMulti> m(new (std::nothrow) RtpbrMulti());
    if (!m) return fail(RTPBR_ERR_ARG, "out of host memory");
    m->band = band;
    auto destroy_all = [&]() { for (RtpbrContext* c : m->ctx) rtpbr_destroy(c); m->ctx.clear(); };
    for (int r = 0; r < n; ++r) {
        RtpbrContext* c = nullptr;
        int rc = rtpbr_create(cfg, devices[r], &c);
        if (rc == RTPBR_OK) rc = rtpbr_set_shard(c, r, n, band);
        if (rc != RTPBR_OK) { if (c) rtpbr_destroy(c); destroy_all(); return rc; }
        m->ctx.push_back(c);
    }
    if (n > 1) {
        // one communicator per GPU, created by this one thread: ncclCommInitRank inside a group (= ncclCommInitAll)
        std::string why;
        if (!load_nccl(why)) { destroy_all(); return fail(RTPBR_ERR_NCCL, why); }
        UniqueId id;
        int r = g_nccl.GetUniqueId(&id);
        if (r != 0) { destroy_all(); return fail(RTPBR_ERR_NCCL, std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r)); }
        g_nccl.GroupStart();
        for (int k = 0; k < n && r == 0; ++k) {
            cudaSetDevice(m->ctx[k]->device);
            r = g_nccl.CommInitRank(&m->ctx[k]->nccl_comm, n, id, k);
            m->ctx[k]->nccl_rank = k;
            m->ctx[k]->nccl_nranks = n;
        }
        const int e = g_nccl.GroupEnd();
        if (r == 0) r = e;
        if (r != 0) { destroy_all(); return fail(RTPBR_ERR_NCCL, std::string("ncclCommInitRank (group): ") + g_nccl.GetErrorString(r)); }
    }
    *out = m.release();
    return RTPBR_OK;
}

int rtpbr_multi_destroy(RtpbrMulti* m)
{
    if (!m) return RTPBR_OK;
    for (RtpbrContext* c : m->ctx) rtpbr_destroy(c);
    delete m;
    return RTPBR_OK;
}

int rtpbr_multi_count(RtpbrMulti* m) { return m ? (int)m->ctx.size() : 0; }

RtpbrContext* rtpbr_multi_context(RtpbrMulti* m, int rank)
{
    if (!m || rank < 0 || rank >= (int)m->ctx.size()) { fail(RTPBR_ERR_ARG, "bad rank"); return nullptr; }
    return m->ctx[rank];
}

#define MULTI_EACH(call)                                                     \
    do {                                                                     \
        if (!m) return fail(RTPBR_ERR_ARG, "null argument");                 \
        for (RtpbrContext* c : m->ctx) {                                     \
            const int rc__ = (call);                                         \
            if (rc__ != RTPBR_OK) return rc__;                               \
        }                                                                    \
        return RTPBR_OK;                                                     \
    } while (0)

int rtpbr_multi_set_scene(RtpbrMulti* m, const RtpbrObject* objects, int n) { MULTI_EACH(rtpbr_set_scene(c, objects, n)); }
int rtpbr_multi_set_camera(RtpbrMulti* m, const RtpbrCamera* cam) { MULTI_EACH(rtpbr_set_camera(c, cam)); }
int rtpbr_multi_set_envmap(RtpbrMulti* m, const float* rgb, int w, int h) { MULTI_EACH(rtpbr_set_envmap(c, rgb, w, h)); }
int rtpbr_multi_set_frame(RtpbrMulti* m, int frame) { MULTI_EACH(rtpbr_set_frame(c, frame)); }
int rtpbr_multi_set_sample_base(RtpbrMulti* m, uint32_t base) { MULTI_EACH(rtpbr_set_sample_base(c, base)); }
int rtpbr_multi_refresh(RtpbrMulti* m) { MULTI_EACH(rtpbr_refresh(c)); }
// launches are asynchronous: the loop queues one kernel per GPU and they run side by side
int rtpbr_multi_pathtrace(RtpbrMulti* m, int spp) { MULTI_EACH(rtpbr_pathtrace(c, spp)); }
int rtpbr_multi_sync(RtpbrMulti* m) { MULTI_EACH(rtpbr_sync(c)); }
#undef MULTI_EACH

int rtpbr_multi_reduce(RtpbrMulti* m, int root)
{
    if (!m) return fail(RTPBR_ERR_ARG, "null argument");
    if (root >= (int)m->ctx.size()) return fail(RTPBR_ERR_ARG, "bad root");
    if (m->ctx.size() == 1) return RTPBR_OK;
    g_nccl.GroupStart();
    int rc = RTPBR_OK;
    for (RtpbrContext* c : m->ctx) {
        rc = rtpbr_reduce_tiles(c, root);
        if (rc != RTPBR_OK) break;
    }
    const int e = g_nccl.GroupEnd();
    if (rc != RTPBR_OK) return rc;
    if (e != 0) return fail(RTPBR_ERR_NCCL, std::string("ncclGroupEnd: ") + g_nccl.GetErrorString(e));
    return RTPBR_OK;
}

int rtpbr_multi_post_process(RtpbrMulti* m, int mode, float exposure, double gamma)
{
    if (!m) return fail(RTPBR_ERR_ARG, "null argument");
    if (m->ctx[0]->cfg.family == RTPBR_FAMILY_C && m->ctx.size() > 1 && !m->ctx[0]->holds_reduced)
        return fail(RTPBR_ERR_UNSUPPORTED, "family C keeps per-pixel ray state between launches: reduce explicitly "
                                           "(rtpbr_multi_reduce) when the frame is finished, then post_process");
    if (m->ctx.size() > 1 && !m->ctx[0]->holds_reduced) {
        const int rc = rtpbr_multi_reduce(m, 0);
        if (rc != RTPBR_OK) return rc;
    }
    return rtpbr_post_process(m->ctx[0], mode, exposure, gamma);
}

int rtpbr_multi_download(RtpbrMulti* m, int which, void* host, size_t bytes)
{
    if (!m) return fail(RTPBR_ERR_ARG, "null argument");
    return rtpbr_download(m->ctx[0], which, host, bytes);
}

}  // extern "C"
