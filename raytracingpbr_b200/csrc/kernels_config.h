// kernels_config.h -- launch constants shared by host code, AOT kernels and NVRTC translation units.
#pragma once
namespace rt {
constexpr int kPoolBlock = 256;        // 8 warps per CTA
constexpr int kPoolMinBlocks = 4;      // <= 64 registers/thread -> 32 warps per SM
constexpr int kPoolSlots = 64;         // path slots per warp: 32 marching + 32 ready / pending
constexpr int kSimpleBlock = 128;
}  // namespace rt
