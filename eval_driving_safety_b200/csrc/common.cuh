// Shared helpers for libb2attack.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/b2attack.h"

namespace b2 {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

void set_error(const char* fmt, ...);

// Kernel-variant switches (A/B measurements, tests of both variants): value from b2_set_flag() if set, else from the
// environment variable B2_<NAME>, else `dflt`.
enum Flag { kFlagConvDcPair = 0, kFlagConv2dHalo = 1, kFlagConvG2Pair = 2, kFlagDepthHeadX4 = 3, kFlagRoiBwdWarp = 4, kNumFlags };
int flag_value(Flag f, const char* env_name, int dflt);

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

#define B2_REQUIRE(cond, ...)                \
    do {                                     \
        if (!(cond)) {                       \
            b2::set_error(__VA_ARGS__);      \
            return B2_ERR_BAD_ARG;           \
        }                                    \
    } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute of a kernel: remember what has been
// granted per (call site, device) instead of once per process, so a process that drives a second GPU opts in
// there too.  ``granted`` is the call site's own static table (zero-initialised).  Racing host threads at worst
// repeat the (idempotent) driver call.
constexpr int kMaxDevices = 64;
struct SmemOptIn { int granted[kMaxDevices]; };

template <typename Kernel>
inline cudaError_t ensure_dynamic_smem(SmemOptIn& tab, Kernel kernel, int bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= kMaxDevices) return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (bytes <= tab.granted[dev]) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) tab.granted[dev] = bytes;
    return e;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Grid for a grid-stride streaming kernel: whole waves of the 148 SMs.
inline int stream_grid(int64_t work_items, int per_block, int max_waves_blocks = kNumSMs * 16) {
    int64_t need = (work_items + per_block - 1) / per_block;
    if (need < 1) need = 1;
    if (need > max_waves_blocks) need = max_waves_blocks;
    return (int)need;
}

// What the epilogue of the stride-1 / transposed kernels can do besides writing the tile.
struct EpiFusion {
    const float* addend;      // out = acc + addend (nullptr: off)
    float* stat_partial;      // per-CTA partial sums [grid][2][Cout] (nullptr: off)
    int stat_mode;            // 1: GroupNorm FORWARD statistics of the written output v: (sum v, sum v^2)
                              // 2: GroupNorm BACKWARD sums of the norm that produced this conv's input, v being
                              //    the gradient w.r.t. that norm's output: (sum v*x, sum v)
                              // 3: as 2 with that norm's ReLU mask recomputed: v counts where fmaf(x, scale_c, shift_c) > 0
    const float* gn_x;        // modes 2, 3: the norm's input x, same shape as out
    const float* gn_coef;     // mode 3: the norm's forward scale[Cout], shift[Cout]
};

__device__ __forceinline__ float4 ldg_stream(const float4* p) { return __ldcs(p); }
__device__ __forceinline__ void stg_stream(float4* p, float4 v) { __stcs(p, v); }

}  // namespace b2
