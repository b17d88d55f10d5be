// jit.h -- NVRTC compilation and driver-API launch of scene-specialised pool kernels.
#pragma once
#include <cuda_runtime.h>

#include <memory>
#include <string>
#include <vector>

#include "rt_params.h"

namespace rt {
namespace jit {

struct Kernel {
    void* module = nullptr;      // CUmodule
    void* function = nullptr;    // CUfunction
    int registers = 0;
    std::string source_hash;
    ~Kernel();
};

// Compile `source` to an sm_100a CUBIN with NVRTC (works without a GPU).  `include_dir` must
// contain pool_kernel.cuh and friends.  Results are cached in-process by source text.
bool compile(const std::string& source, const std::string& include_dir, std::shared_ptr<std::vector<char>>& cubin,
             std::string& log, const std::vector<std::string>& defines = {});
// Load a CUBIN into the current (primary) context and look the kernel up.
bool load(const std::vector<char>& cubin, const char* kernel_name, size_t dynamic_smem, Kernel& out, std::string& err);
bool occupancy(const Kernel& k, int block, size_t dynamic_smem, int* blocks_per_sm, std::string& err);
bool launch(const Kernel& k, const KParams& P, int grid, int block, size_t dynamic_smem, cudaStream_t stream, std::string& err);
// Directory of this shared library + "/csrc".
std::string default_include_dir();
// "12.9" once NVRTC has been loaded (which copy gets loaded matters: see load_nvrtc in jit.cu)
std::string nvrtc_version();

}  // namespace jit
}  // namespace rt
