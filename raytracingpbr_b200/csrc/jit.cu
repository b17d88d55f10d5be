// jit.cu -- NVRTC + CUDA driver API, both resolved with dlopen so that librtpbr.so loads on
// machines without them (the ahead-of-time kernels remain available).
#include "jit.h"

#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>

namespace rt {
namespace jit {
namespace {

struct Nvrtc {
    void* h = nullptr;
    int (*CreateProgram)(void**, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    int (*CompileProgram)(void*, int, const char* const*) = nullptr;
    int (*GetCUBINSize)(void*, size_t*) = nullptr;
    int (*GetCUBIN)(void*, char*) = nullptr;
    int (*GetProgramLogSize)(void*, size_t*) = nullptr;
    int (*GetProgramLog)(void*, char*) = nullptr;
    int (*DestroyProgram)(void**) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
struct Driver {
    void* h = nullptr;
    int (*ModuleLoadData)(void**, const void*) = nullptr;
    int (*ModuleUnload)(void*) = nullptr;
    int (*ModuleGetFunction)(void**, void*, const char*) = nullptr;
    int (*FuncSetAttribute)(void*, int, int) = nullptr;
    int (*FuncGetAttribute)(int*, int, void*) = nullptr;
    int (*OccupancyMaxActiveBlocksPerMultiprocessor)(int*, void*, int, size_t) = nullptr;
    int (*LaunchKernel)(void*, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void*, void**, void**) = nullptr;
    int (*GetErrorString)(int, const char**) = nullptr;
};
Nvrtc g_nvrtc;
int g_nvrtc_major = 0, g_nvrtc_minor = 0;
Driver g_drv;
std::mutex g_mu;
std::unordered_map<std::string, std::shared_ptr<std::vector<char>>> g_cache;

template <class F>
bool sym(void* h, const char* name, F& f)
{
    f = reinterpret_cast<F>(dlsym(h, name));
    return f != nullptr;
}

bool load_nvrtc(std::string& why)
{
    if (g_nvrtc.h) return true;
    // The toolkit's own NVRTC first, by absolute path.  A bare soname resolves to whatever copy the process has
    // already mapped -- with PyTorch imported that is the NVRTC 12.8 bundled in its wheels, whose code for the
    // specialised march loop is 5.5 % slower than 12.9's (measured: 70.5 vs 66.8 ms per C1 launch; that, not NCCL,
    // was the per-rank slowdown of the multi-GPU bench).
#if defined(RTPBR_NVRTC_DIR)
    const char* built_with = RTPBR_NVRTC_DIR "/libnvrtc.so.12";
#else
    const char* built_with = nullptr;
#endif
    const char* names[] = { getenv("RTPBR_NVRTC_LIB"), built_with, "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so.12", "libnvrtc.so" };
    void* h = nullptr;
    for (const char* n : names) {
        if (!n || !*n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (h) break;
    }
    if (!h) { why = "cannot dlopen libnvrtc.so.12 (set RTPBR_NVRTC_LIB)"; return false; }
    Nvrtc n;
    bool ok = sym(h, "nvrtcCreateProgram", n.CreateProgram) && sym(h, "nvrtcCompileProgram", n.CompileProgram) &&
              sym(h, "nvrtcGetCUBINSize", n.GetCUBINSize) && sym(h, "nvrtcGetCUBIN", n.GetCUBIN) &&
              sym(h, "nvrtcGetProgramLogSize", n.GetProgramLogSize) && sym(h, "nvrtcGetProgramLog", n.GetProgramLog) &&
              sym(h, "nvrtcDestroyProgram", n.DestroyProgram) && sym(h, "nvrtcGetErrorString", n.GetErrorString);
    if (!ok) { why = "libnvrtc is missing required symbols"; return false; }
    n.h = h;
    g_nvrtc = n;
    int (*version)(int*, int*) = nullptr;
    if (sym(h, "nvrtcVersion", version)) version(&g_nvrtc_major, &g_nvrtc_minor);
    return true;
}

bool load_driver(std::string& why)
{
    if (g_drv.h) return true;
    void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!h) { why = "cannot dlopen libcuda.so.1"; return false; }
    Driver d;
    bool ok = sym(h, "cuModuleLoadData", d.ModuleLoadData) && sym(h, "cuModuleUnload", d.ModuleUnload) &&
              sym(h, "cuModuleGetFunction", d.ModuleGetFunction) && sym(h, "cuFuncSetAttribute", d.FuncSetAttribute) &&
              sym(h, "cuFuncGetAttribute", d.FuncGetAttribute) &&
              sym(h, "cuOccupancyMaxActiveBlocksPerMultiprocessor", d.OccupancyMaxActiveBlocksPerMultiprocessor) &&
              sym(h, "cuLaunchKernel", d.LaunchKernel) && sym(h, "cuGetErrorString", d.GetErrorString);
    if (!ok) { why = "libcuda is missing required symbols"; return false; }
    d.h = h;
    g_drv = d;
    return true;
}

std::string drv_err(const char* what, int rc)
{
    const char* msg = nullptr;
    if (g_drv.GetErrorString) g_drv.GetErrorString(rc, &msg);
    return std::string(what) + ": " + (msg ? msg : "unknown driver error");
}

}  // namespace

Kernel::~Kernel()
{
    if (module && g_drv.ModuleUnload) g_drv.ModuleUnload(module);
}

std::string nvrtc_version()
{
    return g_nvrtc.h ? std::to_string(g_nvrtc_major) + "." + std::to_string(g_nvrtc_minor) : std::string("not loaded");
}

std::string default_include_dir()
{
    Dl_info info;
    if (dladdr(reinterpret_cast<void*>(&default_include_dir), &info) && info.dli_fname) {
        std::string p = info.dli_fname;
        size_t k = p.rfind('/');
        if (k != std::string::npos) return p.substr(0, k) + "/csrc";
    }
    return "csrc";
}

bool compile(const std::string& source, const std::string& include_dir, std::shared_ptr<std::vector<char>>& cubin,
             std::string& log, const std::vector<std::string>& defines)
{
    std::lock_guard<std::mutex> lock(g_mu);
    std::string key = source;
    for (const std::string& d : defines) key += "\n//" + d;
    auto it = g_cache.find(key);
    if (it != g_cache.end()) { cubin = it->second; log = "cached"; return true; }
    if (!load_nvrtc(log)) return false;
    void* prog = nullptr;
    int rc = g_nvrtc.CreateProgram(&prog, source.c_str(), "rtpbr_scene.cu", 0, nullptr, nullptr);
    if (rc != 0) { log = std::string("nvrtcCreateProgram: ") + g_nvrtc.GetErrorString(rc); return false; }
    const std::string inc = "-I" + include_dir;
    std::vector<const char*> opts = { "--gpu-architecture=sm_100a", "-std=c++17", "--fmad=false", "--prec-div=true",
                                      "--prec-sqrt=true", "-lineinfo", inc.c_str() };
    for (const std::string& d : defines) opts.push_back(d.c_str());
    rc = g_nvrtc.CompileProgram(prog, (int)opts.size(), opts.data());
    size_t n = 0;
    g_nvrtc.GetProgramLogSize(prog, &n);
    std::string plog(n, '\0');
    if (n > 1) g_nvrtc.GetProgramLog(prog, &plog[0]);
    if (rc != 0) {
        log = std::string("nvrtcCompileProgram: ") + g_nvrtc.GetErrorString(rc) + "\n" + plog;
        g_nvrtc.DestroyProgram(&prog);
        return false;
    }
    size_t sz = 0;
    rc = g_nvrtc.GetCUBINSize(prog, &sz);
    auto out = std::make_shared<std::vector<char>>(sz);
    if (rc == 0) rc = g_nvrtc.GetCUBIN(prog, out->data());
    g_nvrtc.DestroyProgram(&prog);
    if (rc != 0 || sz == 0) { log = "nvrtcGetCUBIN failed"; return false; }
    if (const char* dump = getenv("RTPBR_JIT_DUMP")) {   // debugging aid: <dump>.cu / <dump>.cubin for nvdisasm
        if (FILE* f = fopen((std::string(dump) + ".cu").c_str(), "w")) { fwrite(source.data(), 1, source.size(), f); fclose(f); }
        if (FILE* f = fopen((std::string(dump) + ".cubin").c_str(), "wb")) { fwrite(out->data(), 1, out->size(), f); fclose(f); }
    }
    // bounded: an animated scene produces a new translation unit per frame
    if (g_cache.size() >= 64) g_cache.erase(g_cache.begin());
    g_cache[key] = out;
    cubin = out;
    log = plog;
    return true;
}

bool load(const std::vector<char>& cubin, const char* kernel_name, size_t dynamic_smem, Kernel& out, std::string& err)
{
    if (!load_driver(err)) return false;
    int rc = g_drv.ModuleLoadData(&out.module, cubin.data());
    if (rc != 0) { err = drv_err("cuModuleLoadData", rc); return false; }
    rc = g_drv.ModuleGetFunction(&out.function, out.module, kernel_name);
    if (rc != 0) { err = drv_err("cuModuleGetFunction", rc); return false; }
    rc = g_drv.FuncSetAttribute(out.function, /*CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES*/ 8, (int)dynamic_smem);
    if (rc != 0) { err = drv_err("cuFuncSetAttribute", rc); return false; }
    g_drv.FuncGetAttribute(&out.registers, /*CU_FUNC_ATTRIBUTE_NUM_REGS*/ 4, out.function);
    return true;
}

bool occupancy(const Kernel& k, int block, size_t dynamic_smem, int* blocks_per_sm, std::string& err)
{
    int rc = g_drv.OccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k.function, block, dynamic_smem);
    if (rc != 0) { err = drv_err("cuOccupancyMaxActiveBlocksPerMultiprocessor", rc); return false; }
    return true;
}

bool launch(const Kernel& k, const KParams& P, int grid, int block, size_t dynamic_smem, cudaStream_t stream, std::string& err)
{
    void* args[] = { const_cast<KParams*>(&P) };
    int rc = g_drv.LaunchKernel(k.function, (unsigned)grid, 1, 1, (unsigned)block, 1, 1, (unsigned)dynamic_smem, stream, args, nullptr);
    if (rc != 0) { err = drv_err("cuLaunchKernel", rc); return false; }
    return true;
}

}  // namespace jit
}  // namespace rt
