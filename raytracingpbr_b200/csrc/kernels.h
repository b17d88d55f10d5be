// kernels.h -- host-callable launchers of kernels.cu
#pragma once
#include <cuda_runtime.h>

#include "kernels_config.h"
#include "rt_params.h"

namespace rt {


struct KernelSelect {
    int family;
    int marcher;
    int nobj;
    bool bunny;
    bool count;
};

bool kernel_supported(const KernelSelect& sel);
size_t pool_dynamic_smem();
cudaError_t launch_pathtrace_pool(const KernelSelect& sel, const KParams& P, int grid, cudaStream_t stream);
cudaError_t launch_pathtrace_simple(const KernelSelect& sel, const KParams& P, cudaStream_t stream);
cudaError_t pool_occupancy(const KernelSelect& sel, int* blocks_per_sm);
cudaError_t launch_fold_samples(const KParams& P, cudaStream_t stream);
cudaError_t launch_refresh_depth(float* ray_buffer, int n, cudaStream_t stream);
cudaError_t launch_post_process_src(const float4* image_buffer, float* image_pixels, float* diff_buffer, float* diff_pixels, int n,
                                    float exposure, float gamma_inv, cudaStream_t stream);
cudaError_t launch_refresh_adaptive(float* diff_buffer, float* diff_pixels, int n, cudaStream_t stream);
cudaError_t launch_denoise(const float* pixels_in, const float* out_prev, float* out_new, int W, int H, float threshold, cudaStream_t stream);
cudaError_t launch_post_process(const float4* image_buffer, float* image_pixels, int n, int mode, float exposure,
                                float inv_gamma, cudaStream_t stream);

}  // namespace rt
