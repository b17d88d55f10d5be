// kernels_config.h -- launch constants shared by host code, AOT kernels and NVRTC translation units.
// The RT_POOL_* macros let the scene-specialised (NVRTC) build use another pool geometry than the
// ahead-of-time default (RTPBR_POOL_SLOTS / RTPBR_POOL_BLOCK / RTPBR_POOL_MIN_BLOCKS, capi.cu).
#pragma once
#ifndef RT_POOL_BLOCK
#define RT_POOL_BLOCK 256
#endif
#ifndef RT_POOL_MIN_BLOCKS
#define RT_POOL_MIN_BLOCKS 4
#endif
#ifndef RT_POOL_SLOTS
#define RT_POOL_SLOTS 64
#endif
namespace rt {
constexpr int kPoolBlock = RT_POOL_BLOCK;            // threads per CTA (8 warps)
constexpr int kPoolMinBlocks = RT_POOL_MIN_BLOCKS;   // <= 64 registers/thread -> 32 warps per SM
constexpr int kPoolSlots = RT_POOL_SLOTS;            // path slots per warp: 32 marching + 32 ready / pending
constexpr int kPoolSlotWords = 26;                   // words of state per slot (pool_kernel.cuh F_COUNT)
constexpr int kPoolStacks = 3;                       // per-warp slot stacks: ready, pending, fresh (one byte per slot each)
constexpr int kPoolQueueWords = 8;                   // per-warp work-queue chunk state (pool_kernel.cuh WQ_*)
constexpr int kSimpleBlock = 128;
// dynamic shared memory of one pool CTA (host side only; NVRTC rejects unannotated functions)
#if !defined(__CUDACC_RTC__)
constexpr unsigned long pool_smem_bytes_for(int block, int slots)
{
    return (unsigned long)(block / 32) * (unsigned long)(kPoolSlotWords * slots + kPoolStacks * (slots / 4) + kPoolQueueWords) * 4ul;
}
#endif
}  // namespace rt
