// Subsystem (2): frustum -> voxel lifting by bilinear / trilinear grid sampling,
// forward and DETERMINISTIC backward.  Replaces ATen grid_sampler_{2d,3d} and
// their atomicAdd backward, reached through F.grid_sample inside
// StereoNet.forward (attack/DSGN/pgd_attack.py:308, 336).
//
// Semantics = F.grid_sample(mode='bilinear', padding_mode='zeros'), both
// align_corners conventions (ATen grid_sampler_unnormalize formulas).
//
// Layout: channels-last.  A sub-warp of C/4 lanes owns one output voxel, every
// lane a float4 of channels, so each corner read and the output write are one
// contiguous C*4-byte segment (256 B for the 64-channel PSV feature).
//
// Backward: the sampling grid is a fixed function of the calibration, so a CSR
// plan "input cell -> sorted list of (output voxel, weight)" is built once and
// reused by every PGD iteration; the backward is then a pure gather with a fixed
// summation order -> bitwise reproducible, no atomics on floats.
#include "common.cuh"

namespace b2 {

__device__ __forceinline__ float unnorm(float g, int size, int align) {
    return align ? ((g + 1.f) / 2.f) * (float)(size - 1) : ((g + 1.f) * (float)size - 1.f) / 2.f;
}

struct Corners3 {
    int x0, y0, z0;
    float wx1, wy1, wz1;  // weight of the +1 corner along each axis
};

__device__ __forceinline__ Corners3 corners3(const float* g, int D, int H, int W, int align) {
    float ix = unnorm(g[0], W, align), iy = unnorm(g[1], H, align), iz = unnorm(g[2], D, align);
    float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
    Corners3 c;
    // clamp far-out coordinates before the int cast (NaN/inf/huge -> all corners OOB)
    fx = fminf(fmaxf(fx, -2.f), (float)W + 1.f); if (!(ix == ix)) fx = -2.f;
    fy = fminf(fmaxf(fy, -2.f), (float)H + 1.f); if (!(iy == iy)) fy = -2.f;
    fz = fminf(fmaxf(fz, -2.f), (float)D + 1.f); if (!(iz == iz)) fz = -2.f;
    c.x0 = (int)fx; c.y0 = (int)fy; c.z0 = (int)fz;
    c.wx1 = ix - floorf(ix); c.wy1 = iy - floorf(iy); c.wz1 = iz - floorf(iz);
    return c;
}

__device__ __forceinline__ void fma4(float4& a, float w, float4 v) {
    a.x += w * v.x; a.y += w * v.y; a.z += w * v.z; a.w += w * v.w;
}

// LPV = lanes per voxel (= C/4, power of two <= 32).  One-shot blocks (no grid-stride loop): voxels
// outside the frustum issue no loads, so per-voxel cost varies and the hardware block scheduler
// balances it.  All eight corner loads are issued before the first FMA (8 x 16 B in flight per lane).
template <int LPV>
__global__ void __launch_bounds__(256, 4)   // min-blocks hint: without it ptxas squeezes to 32 registers and serialises the loads
grid_sample3d_fwd_kernel(const float4* __restrict__ in, const float* __restrict__ grid,
                         float* __restrict__ out, int N, int D, int H, int W, int64_t nvox_per_n,
                         int out_cstride, int out_coff, int align) {
    const int64_t nvox = (int64_t)N * nvox_per_n;
    const int lane = threadIdx.x % LPV;
    const int64_t v = (int64_t)blockIdx.x * (blockDim.x / LPV) + threadIdx.x / LPV;
    if (v >= nvox) return;
    const int n = (int)(v / nvox_per_n);
    const float* g = grid + v * 3;
    float gg[3] = {__ldg(g), __ldg(g + 1), __ldg(g + 2)};
    Corners3 c = corners3(gg, D, H, W, align);
    const float4* base = in + (int64_t)n * D * H * W * LPV + lane;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    // voxel entirely outside the feature volume: nothing to read
    const bool any_in = c.z0 >= -1 && c.z0 < D && c.y0 >= -1 && c.y0 < H && c.x0 >= -1 && c.x0 < W;
    if (any_in) {
        // Unpredicated loads from CLAMPED (always valid) addresses, out-of-range corners get weight 0:
        // predicated loads made ptxas serialise the eight reads into ~6 dependent round trips.
        float4 val[8];
        float wgt[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            int dz = k >> 2, dy = (k >> 1) & 1, dx = k & 1;
            int z = c.z0 + dz, y = c.y0 + dy, x = c.x0 + dx;
            bool ok = z >= 0 && z < D && y >= 0 && y < H && x >= 0 && x < W;
            float w = (dx ? c.wx1 : 1.f - c.wx1) * (dy ? c.wy1 : 1.f - c.wy1) * (dz ? c.wz1 : 1.f - c.wz1);
            wgt[k] = ok ? w : 0.f;
            z = min(max(z, 0), D - 1); y = min(max(y, 0), H - 1); x = min(max(x, 0), W - 1);
            val[k] = __ldg(base + (uint32_t)(((z * H + y) * W + x) * LPV));   // < 2^31 float4 per sample (checked by the launcher)
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) fma4(acc, wgt[k], val[k]);
    }
    *reinterpret_cast<float4*>(out + v * out_cstride + out_coff + lane * 4) = acc;
}

template <int LPV>
__global__ void __launch_bounds__(256)
grid_sample2d_fwd_kernel(const float4* __restrict__ in, const float* __restrict__ grid,
                         float* __restrict__ out, int N, int H, int W, int64_t nvox_per_n,
                         int out_cstride, int out_coff, int align) {
    const int64_t nvox = (int64_t)N * nvox_per_n;
    const int lane = threadIdx.x % LPV;
    const int64_t vstride = (int64_t)gridDim.x * (blockDim.x / LPV);
    for (int64_t v = (int64_t)blockIdx.x * (blockDim.x / LPV) + threadIdx.x / LPV; v < nvox; v += vstride) {
        int n = (int)(v / nvox_per_n);
        const float* g = grid + v * 2;
        float gg[3] = {__ldg(g), __ldg(g + 1), -1.f};
        Corners3 c = corners3(gg, 1, H, W, align);
        const float4* base = in + (int64_t)n * H * W * LPV + lane;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int dy = k >> 1, dx = k & 1;
            int y = c.y0 + dy, x = c.x0 + dx;
            float w = (dx ? c.wx1 : 1.f - c.wx1) * (dy ? c.wy1 : 1.f - c.wy1);
            if (y >= 0 && y < H && x >= 0 && x < W)
                fma4(acc, w, __ldg(base + ((int64_t)y * W + x) * LPV));
        }
        *reinterpret_cast<float4*>(out + v * out_cstride + out_coff + lane * 4) = acc;
    }
}

// =============================================================================================
// Fused lifting forward (DSGN's frustum -> voxel step: the 64-channel PSV feature sampled trilinearly AND the
// 32-channel image feature sampled bilinearly at the same (u, v), written side by side into one [.., 96]
// voxel row).  One kernel instead of two: the grid is read once, every output row is written once as three
// full 128 B lines.
// What bounds this op (measured, round 2): NOT the memory system.  In the generic kernels above 16 lanes own a
// voxel and EVERY lane recomputes the corner indices, validity masks, weights and addresses (~190 of its ~270
// instructions) for its one float4 per corner: 88 M warp instructions for the PSV sample = ~160 us of issue
// time at IPC 2, the measured 190 us.  Re-ordering the voxels for L1 reuse (blocks of consecutive Z: the
// corner rows of a column drift by less than a pixel per step) changed nothing -- 0.304-0.312 ms for every
// block shape, tools/bench_lift.py.  Here 8 lanes own a voxel and each carries 2 float4 per PSV corner (lane l ->
// float4 l and l+8: one load instruction of the 8 lanes covers one whole 128 B line): the index arithmetic is
// amortised over twice the data without raising the number of cache lines an instruction touches (a 4-lane /
// 4-float4 variant, 64 B per voxel per instruction, doubled the L1 tag work per byte and needed 128 registers:
// ncu l1tex 66 %, 21 % occupancy, 0.239 ms).
// =============================================================================================
// Occupancy is what this latency-bound gather wants: measured 0.198 ms with 3 resident blocks per SM (80 registers, all
// loads of a batch in flight), 0.165 with 4 (63), 0.161 with 5 (48), **0.154 with 6** (40 registers, 8 bytes spilled),
// 0.157 / 0.158 with 7 / 8; streaming stores or issuing the image-feature loads first change nothing.
__global__ void __launch_bounds__(256, 6)
lift_fwd_kernel(const float4* __restrict__ psv, const float4* __restrict__ img, const float* __restrict__ grid,
                float* __restrict__ out, int N, int D, int H, int W, int Hi, int Wi, int64_t nvox_per_n, int align) {
    constexpr int LPV = 8, R3 = 16, R2 = 8, CO = 96;        // lanes per voxel, float4 per PSV row / image row
    const int lane = threadIdx.x % LPV;
    const int64_t v = (int64_t)blockIdx.x * (256 / LPV) + threadIdx.x / LPV;
    if (v >= (int64_t)N * nvox_per_n) return;
    const int n = (int)(v / nvox_per_n);
    const float* g = grid + v * 3;
    float gg[3] = {__ldg(g), __ldg(g + 1), __ldg(g + 2)};
    const Corners3 c = corners3(gg, D, H, W, align);
    float4* orow = reinterpret_cast<float4*>(out + v * CO) + lane;
    // the image feature at the same (u, v); its size may differ from the PSV's, so its corners are recomputed
    float4 v2[4];
    float w2[4];
    auto img_loads = [&]() {
        float g2[3] = {gg[0], gg[1], -1.f};
        const Corners3 c2 = corners3(g2, 1, Hi, Wi, align);
        const float4* base2 = img + (int64_t)n * Hi * Wi * R2 + lane;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int dy = k >> 1, dx = k & 1;
            int yy = c2.y0 + dy, xx = c2.x0 + dx;
            const bool ok = yy >= 0 && yy < Hi && xx >= 0 && xx < Wi;
            const float w = (dx ? c2.wx1 : 1.f - c2.wx1) * (dy ? c2.wy1 : 1.f - c2.wy1);
            w2[k] = ok ? w : 0.f;
            yy = min(max(yy, 0), Hi - 1); xx = min(max(xx, 0), Wi - 1);
            v2[k] = __ldg(base2 + (uint32_t)((yy * Wi + xx) * R2));
        }
    };
    float4 acc[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
    const bool any_in = c.z0 >= -1 && c.z0 < D && c.y0 >= -1 && c.y0 < H && c.x0 >= -1 && c.x0 < W;
    if (any_in) {
        const float4* base = psv + (int64_t)n * D * H * W * R3 + lane;
#pragma unroll
        for (int half = 0; half < 2; ++half) {               // two batches of 4 corners: 8 loads in flight per lane
            float4 val[4][2];
            float wgt[4];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const int k = half * 4 + kk;
                const int dz = k >> 2, dy = (k >> 1) & 1, dx = k & 1;
                int zz = c.z0 + dz, yy = c.y0 + dy, xx = c.x0 + dx;
                const bool ok = zz >= 0 && zz < D && yy >= 0 && yy < H && xx >= 0 && xx < W;
                const float w = (dx ? c.wx1 : 1.f - c.wx1) * (dy ? c.wy1 : 1.f - c.wy1) * (dz ? c.wz1 : 1.f - c.wz1);
                wgt[kk] = ok ? w : 0.f;
                zz = min(max(zz, 0), D - 1); yy = min(max(yy, 0), H - 1); xx = min(max(xx, 0), W - 1);
                const float4* row = base + (uint32_t)(((zz * H + yy) * W + xx) * R3);
                val[kk][0] = __ldg(row);
                val[kk][1] = __ldg(row + LPV);
            }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) { fma4(acc[0], wgt[kk], val[kk][0]); fma4(acc[1], wgt[kk], val[kk][1]); }
        }
    }
    orow[0] = acc[0];
    orow[LPV] = acc[1];
    img_loads();
    float4 a2 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 4; ++k) fma4(a2, w2[k], v2[k]);
    orow[R3] = a2;
}

// ---------------------------------------------------------------------------
// CSR plan.  FILL == 0: count entries per input cell; FILL == 1: write them.
// One thread per output voxel.  Integer atomics only (counts are exact; the
// slot order inside a row is fixed afterwards by plan_sort).
// ---------------------------------------------------------------------------
template <int FILL>
__global__ void __launch_bounds__(256)
grid_plan_kernel(const float* __restrict__ grid, int32_t* __restrict__ counts_or_cursor,
                 const int32_t* __restrict__ row_ptr, int2* __restrict__ entries, int ndim, int N,
                 int D, int H, int W, int64_t nvox_per_n, int align) {
    const int64_t nvox = (int64_t)N * nvox_per_n;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvox;
         v += (int64_t)gridDim.x * blockDim.x) {
        int n = (int)(v / nvox_per_n);
        const float* g = grid + v * ndim;
        float gg[3] = {__ldg(g), __ldg(g + 1), ndim == 3 ? __ldg(g + 2) : -1.f};
        Corners3 c = corners3(gg, D, H, W, align);
        if (ndim == 2) { c.z0 = 0; c.wz1 = 0.f; }
        const int nk = ndim == 3 ? 8 : 4;
        for (int k = 0; k < nk; ++k) {
            int dz = k >> 2, dy = (k >> 1) & 1, dx = k & 1;
            int z = c.z0 + dz, y = c.y0 + dy, x = c.x0 + dx;
            if (z < 0 || z >= D || y < 0 || y >= H || x < 0 || x >= W) continue;
            float w = (dx ? c.wx1 : 1.f - c.wx1) * (dy ? c.wy1 : 1.f - c.wy1);
            if (ndim == 3) w *= (dz ? c.wz1 : 1.f - c.wz1);
            int64_t cell = (((int64_t)n * D + z) * H + y) * W + x;
            if (FILL) {
                int slot = atomicAdd(counts_or_cursor + cell, 1);
                entries[(int64_t)row_ptr[cell] + slot] = make_int2((int)v, __float_as_int(w));
            } else {
                atomicAdd(counts_or_cursor + cell, 1);
            }
        }
    }
}

// Sort each CSR row by (voxel id, then weight bits) -- insertion sort, rows are short.
__global__ void __launch_bounds__(256)
grid_plan_sort_kernel(const int32_t* __restrict__ row_ptr, int2* __restrict__ entries, int64_t ncell) {
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < ncell;
         c += (int64_t)gridDim.x * blockDim.x) {
        int b = row_ptr[c], e = row_ptr[c + 1];
        for (int i = b + 1; i < e; ++i) {
            int2 key = entries[i];
            int j = i - 1;
            while (j >= b) {
                int2 o = entries[j];
                if (o.x < key.x || (o.x == key.x && o.y <= key.y)) break;
                entries[j + 1] = o;
                --j;
            }
            entries[j + 1] = key;
        }
    }
}

// Gather backward: a group of LPV*SPLIT lanes owns one input cell (a lane = one float4 of
// channels).  SPLIT = 1: sub-warp per cell (short rows, e.g. ~6 entries per PSV cell).
// SPLIT = 32/LPV: a whole warp per cell whose sub-groups take every SPLIT-th CSR entry (an
// image-feature pixel receives from ~150 voxels along its ray); the sub-group partial sums are
// combined by a fixed shuffle tree.  Either way the summation order depends only on the plan.
// F4 = float4 per lane: with F4 = 1 every lane of a cell repeats the entry load and the 64-bit address
// arithmetic for a single float4 -- the gather is then instruction-issue bound like the forward (see
// lift_fwd_kernel); F4 = 4 (64 channels on 4 lanes, lane l -> float4 l, l+4, l+8, l+12) amortises them.
template <int LPV, int SPLIT, int F4>
__global__ void __launch_bounds__(256)
grid_sample_bwd_kernel(const float* __restrict__ gout, const int32_t* __restrict__ row_ptr,
                       const int2* __restrict__ entries, float4* __restrict__ gin, int64_t ncell,
                       int gout_cstride, int gout_coff) {
    constexpr int GS = LPV * SPLIT;                 // lanes per cell
    const int cl = threadIdx.x % LPV, sub = (threadIdx.x / LPV) % SPLIT;
    const int64_t cstride = (int64_t)gridDim.x * (blockDim.x / GS);
    // the loop bound is warp-uniform (first cell of the warp) because of the shuffles below
    for (int64_t c0 = (int64_t)blockIdx.x * (blockDim.x / GS) + (threadIdx.x / 32) * (32 / GS); c0 < ncell; c0 += cstride) {
        const int64_t c = c0 + (threadIdx.x & 31) / GS;
        const bool valid = c < ncell;
        const int b = valid ? __ldg(row_ptr + c) : 0, e = valid ? __ldg(row_ptr + c + 1) : 0;
        float4 acc[F4];
#pragma unroll
        for (int q = 0; q < F4; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        int i = b + sub;
        // two independent chains per trip (a 4-way unroll was measured slower in round 1: occupancy hides this
        // two-level entry -> row latency better than more ILP)
        for (; i + SPLIT < e; i += 2 * SPLIT) {
            const int2 e0 = __ldg(entries + i), e1 = __ldg(entries + i + SPLIT);
            const float4* r0 = reinterpret_cast<const float4*>(gout + (int64_t)e0.x * gout_cstride + gout_coff) + cl;
            const float4* r1 = reinterpret_cast<const float4*>(gout + (int64_t)e1.x * gout_cstride + gout_coff) + cl;
            float4 g0[F4], g1[F4];
#pragma unroll
            for (int q = 0; q < F4; ++q) { g0[q] = __ldg(r0 + LPV * q); g1[q] = __ldg(r1 + LPV * q); }
#pragma unroll
            for (int q = 0; q < F4; ++q) { fma4(acc[q], __int_as_float(e0.y), g0[q]); fma4(acc[q], __int_as_float(e1.y), g1[q]); }
        }
        if (i < e) {
            const int2 e0 = __ldg(entries + i);
            const float4* r0 = reinterpret_cast<const float4*>(gout + (int64_t)e0.x * gout_cstride + gout_coff) + cl;
#pragma unroll
            for (int q = 0; q < F4; ++q) fma4(acc[q], __int_as_float(e0.y), __ldg(r0 + LPV * q));
        }
#pragma unroll
        for (int off = GS / 2; off >= LPV; off >>= 1) {
#pragma unroll
            for (int q = 0; q < F4; ++q) {
                acc[q].x += __shfl_down_sync(0xffffffffu, acc[q].x, off, GS);
                acc[q].y += __shfl_down_sync(0xffffffffu, acc[q].y, off, GS);
                acc[q].z += __shfl_down_sync(0xffffffffu, acc[q].z, off, GS);
                acc[q].w += __shfl_down_sync(0xffffffffu, acc[q].w, off, GS);
            }
        }
        if (valid && sub == 0) {
#pragma unroll
            for (int q = 0; q < F4; ++q) gin[c * (LPV * F4) + cl + LPV * q] = acc[q];
        }
    }
}

}  // namespace b2

using namespace b2;

#define B2_LPV_SWITCH(C, ...)                                                         \
    switch ((C) / 4) {                                                                \
        case 1: { constexpr int LPV = 1; __VA_ARGS__; } break;                               \
        case 2: { constexpr int LPV = 2; __VA_ARGS__; } break;                               \
        case 4: { constexpr int LPV = 4; __VA_ARGS__; } break;                               \
        case 8: { constexpr int LPV = 8; __VA_ARGS__; } break;                               \
        case 16: { constexpr int LPV = 16; __VA_ARGS__; } break;                             \
        case 32: { constexpr int LPV = 32; __VA_ARGS__; } break;                             \
        default:                                                                      \
            b2::set_error("grid_sample: C must be 4,8,16,32,64 or 128 (got %d)", (C)); \
            return B2_ERR_UNSUPPORTED;                                                \
    }

extern "C" int b2_grid_sample3d_fwd(const float* in, const float* grid, float* out, int N, int C,
                                    int D, int H, int W, int64_t nvox_per_n, int out_cstride,
                                    int out_coff, int align_corners, void* stream) {
    B2_REQUIRE(in && grid && out, "grid_sample3d_fwd: null pointer");
    B2_REQUIRE(C % 4 == 0 && out_cstride % 4 == 0 && out_coff % 4 == 0 && aligned16(in) && aligned16(out),
               "grid_sample3d_fwd: channel counts/offsets must be multiples of 4 and pointers 16B aligned");
    if ((int64_t)N * nvox_per_n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t nvox = (int64_t)N * nvox_per_n;
    B2_LPV_SWITCH(C, {
        const int64_t nblk = (nvox + 256 / LPV - 1) / (256 / LPV);
        B2_REQUIRE(nblk < ((int64_t)1 << 31), "grid_sample3d_fwd: too many voxels");
        B2_REQUIRE((int64_t)D * H * W * LPV < ((int64_t)1 << 31), "grid_sample3d_fwd: input sample has more than 2^31 float4");
        grid_sample3d_fwd_kernel<LPV><<<(unsigned)nblk, 256, 0, st>>>((const float4*)in, grid, out, N, D, H, W,
                                                             nvox_per_n, out_cstride, out_coff, align_corners);
    });
    return check_launch("grid_sample3d_fwd");
}

extern "C" int b2_grid_sample2d_fwd(const float* in, const float* grid, float* out, int N, int C,
                                    int H, int W, int64_t nvox_per_n, int out_cstride, int out_coff,
                                    int align_corners, void* stream) {
    B2_REQUIRE(in && grid && out, "grid_sample2d_fwd: null pointer");
    B2_REQUIRE(C % 4 == 0 && out_cstride % 4 == 0 && out_coff % 4 == 0 && aligned16(in) && aligned16(out),
               "grid_sample2d_fwd: channel counts/offsets must be multiples of 4 and pointers 16B aligned");
    if ((int64_t)N * nvox_per_n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t nvox = (int64_t)N * nvox_per_n;
    B2_LPV_SWITCH(C, {
        int grid_x = stream_grid(nvox, 256 / LPV, kNumSMs * 16);
        grid_sample2d_fwd_kernel<LPV><<<grid_x, 256, 0, st>>>((const float4*)in, grid, out, N, H, W,
                                                             nvox_per_n, out_cstride, out_coff, align_corners);
    });
    return check_launch("grid_sample2d_fwd");
}

static int plan_args_ok(int ndim, int N, int D, int H, int W, int64_t nvox_per_n) {
    if (!(ndim == 2 || ndim == 3)) { set_error("grid_plan: ndim must be 2 or 3"); return B2_ERR_BAD_ARG; }
    if (ndim == 2 && D != 1) { set_error("grid_plan: ndim 2 needs D == 1"); return B2_ERR_BAD_ARG; }
    if ((int64_t)N * nvox_per_n >= (int64_t)1 << 31 || (int64_t)N * D * H * W >= (int64_t)1 << 31) {
        set_error("grid_plan: more than 2^31 voxels or cells"); return B2_ERR_UNSUPPORTED;
    }
    return 0;
}

extern "C" int b2_grid_plan_count(const float* grid, int32_t* counts, int ndim, int N, int D, int H,
                                  int W, int64_t nvox_per_n, int align_corners, void* stream) {
    B2_REQUIRE(grid && counts, "grid_plan_count: null pointer");
    if (int e = plan_args_ok(ndim, N, D, H, W, nvox_per_n)) return e;
    int64_t nvox = (int64_t)N * nvox_per_n;
    if (nvox == 0) return 0;
    grid_plan_kernel<0><<<stream_grid(nvox, 256), 256, 0, (cudaStream_t)stream>>>(
        grid, counts, nullptr, nullptr, ndim, N, D, H, W, nvox_per_n, align_corners);
    return check_launch("grid_plan_count");
}

extern "C" int b2_grid_plan_fill(const float* grid, const int32_t* row_ptr, int32_t* cursor,
                                 void* entries, int ndim, int N, int D, int H, int W,
                                 int64_t nvox_per_n, int align_corners, void* stream) {
    B2_REQUIRE(grid && row_ptr && cursor && entries, "grid_plan_fill: null pointer");
    if (int e = plan_args_ok(ndim, N, D, H, W, nvox_per_n)) return e;
    int64_t nvox = (int64_t)N * nvox_per_n;
    if (nvox == 0) return 0;
    grid_plan_kernel<1><<<stream_grid(nvox, 256), 256, 0, (cudaStream_t)stream>>>(
        grid, cursor, row_ptr, (int2*)entries, ndim, N, D, H, W, nvox_per_n, align_corners);
    return check_launch("grid_plan_fill");
}

extern "C" int b2_grid_plan_sort(const int32_t* row_ptr, void* entries, int64_t ncell, void* stream) {
    B2_REQUIRE(row_ptr && entries, "grid_plan_sort: null pointer");
    if (ncell == 0) return 0;
    grid_plan_sort_kernel<<<stream_grid(ncell, 256), 256, 0, (cudaStream_t)stream>>>(row_ptr, (int2*)entries, ncell);
    return check_launch("grid_plan_sort");
}

extern "C" int b2_grid_sample_bwd(const float* gout, const int32_t* row_ptr, const void* entries,
                                  float* gin, int64_t ncell, int C, int gout_cstride, int gout_coff,
                                  int long_rows, void* stream) {
    B2_REQUIRE(gout && row_ptr && entries && gin, "grid_sample_bwd: null pointer");
    B2_REQUIRE(C % 4 == 0 && gout_cstride % 4 == 0 && gout_coff % 4 == 0 && aligned16(gout) && aligned16(gin),
               "grid_sample_bwd: channel counts/offsets must be multiples of 4 and pointers 16B aligned");
    if (ncell == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    static int wide = -1;
    if (wide < 0) { const char* e = getenv("B2_GS_BWD_WIDE"); wide = (e && e[0] == '0') ? 0 : 1; }
    // Wide lanes (several float4 per lane, the entry load and address arithmetic amortised) pay on the long rows of the
    // image-feature lift (0.051 -> 0.046 ms) but not on the ~6-entry rows of the PSV lift (0.188 -> 0.198 ms: the 4x
    // fewer threads per cell hide the entry -> row latency worse), tools/bench_lift.py -- so only the former uses them.
    if (wide && C == 32 && long_rows) {
        int grid_x = stream_grid(ncell, 256 / 32, kNumSMs * 32);
        grid_sample_bwd_kernel<4, 8, 2><<<grid_x, 256, 0, st>>>(gout, row_ptr, (const int2*)entries, (float4*)gin, ncell,
                                                               gout_cstride, gout_coff);
        return check_launch("grid_sample_bwd");
    }
    B2_LPV_SWITCH(C, {
        if (long_rows) {
            int grid_x = stream_grid(ncell, 256 / 32, kNumSMs * 32);
            grid_sample_bwd_kernel<LPV, 32 / LPV, 1><<<grid_x, 256, 0, st>>>(gout, row_ptr, (const int2*)entries,
                                                                            (float4*)gin, ncell, gout_cstride, gout_coff);
        } else {
            int grid_x = stream_grid(ncell, 256 / LPV, kNumSMs * 16);
            grid_sample_bwd_kernel<LPV, 1, 1><<<grid_x, 256, 0, st>>>(gout, row_ptr, (const int2*)entries, (float4*)gin,
                                                                     ncell, gout_cstride, gout_coff);
        }
    });
    return check_launch("grid_sample_bwd");
}

extern "C" int b2_lift_fwd(const float* psv, const float* img, const float* grid, float* out, int N, int C3, int C2,
                           int D, int H, int W, int Hi, int Wi, int64_t nvox_per_n, int align_corners, void* stream) {
    B2_REQUIRE(psv && img && grid && out, "lift_fwd: null pointer");
    B2_REQUIRE(C3 == 64 && C2 == 32, "lift_fwd: the fused kernel serves the DSGN widths (64 + 32 channels); got %d + %d "
               "(use b2_grid_sample3d_fwd / b2_grid_sample2d_fwd)", C3, C2);
    B2_REQUIRE(aligned16(psv) && aligned16(img) && aligned16(out), "lift_fwd: pointers must be 16-byte aligned");
    B2_REQUIRE(N >= 0 && D > 0 && H > 0 && W > 0 && Hi > 0 && Wi > 0 && nvox_per_n >= 0, "lift_fwd: bad dims");
    B2_REQUIRE((int64_t)D * H * W * 16 < ((int64_t)1 << 31), "lift_fwd: input sample has more than 2^31 float4");
    const int64_t nvox = (int64_t)N * nvox_per_n;
    if (nvox == 0) return 0;
    const int64_t nblk = (nvox + 31) / 32;
    B2_REQUIRE(nblk < ((int64_t)1 << 31), "lift_fwd: too many voxels");
    lift_fwd_kernel<<<(unsigned)nblk, 256, 0, (cudaStream_t)stream>>>(
        (const float4*)psv, (const float4*)img, grid, out, N, D, H, W, Hi, Wi, nvox_per_n, align_corners);
    return check_launch("lift_fwd");
}
