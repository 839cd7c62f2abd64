// Subsystem (1): plane-sweep stereo cost volume, forward and gather-form backward.
// Replaces upstream dsgn._C build_cost_volume_{forward,backward} (reached from
// attack/DSGN/pgd_attack.py:308 / :336).  Semantics pinned in oracle/dsgn_ref.py
// build_cost_volume():
//   s = shifts[n,d]; s0 = floor(s); f = s - s0; x0 = w - s0; valid = x0 >= 0
//   cost[n, 0:C , d,h,w] = L[n,:,h,w] * valid
//   cost[n, C:2C, d,h,w] = ((1-f)*R[n,:,h,x0] + f*R[n,:,h,x0-1]*[x0-1>=0]) * valid
//
// Roofline: HBM write-bound forward (2*3.83 MB in, 368 MB out per KITTI pair),
// HBM read-bound backward.  The features (7.7 MB) stay L2 resident across the D
// planes.  No atomics anywhere: the backward sums over planes in a fixed order.
#include "common.cuh"

namespace b2 {

constexpr int kCvThreads = 256;

__device__ __forceinline__ float4 lerp_valid(float4 r0, float4 r1, float f, float v1) {
    float omf = __fsub_rn(1.f, f);
    float4 o;
    o.x = __fadd_rn(__fmul_rn(omf, r0.x), __fmul_rn(f, __fmul_rn(r1.x, v1)));
    o.y = __fadd_rn(__fmul_rn(omf, r0.y), __fmul_rn(f, __fmul_rn(r1.y, v1)));
    o.z = __fadd_rn(__fmul_rn(omf, r0.z), __fmul_rn(f, __fmul_rn(r1.z, v1)));
    o.w = __fadd_rn(__fmul_rn(omf, r0.w), __fmul_rn(f, __fmul_rn(r1.w, v1)));
    return o;
}

// Channels-last forward.  One thread owns one float4 of one (n,h,w) feature row and walks the D
// planes: the left value is loaded once and written D times, the right value is re-gathered per
// plane from the L2-resident features; per-plane (integer shift, fraction) pairs sit in smem, so
// the inner loop is address arithmetic-free.  For a fixed plane consecutive threads write
// consecutive 16 B -> every store instruction is a fully coalesced 512 B per warp.
constexpr int kCvMaxD = 256;

__global__ void __launch_bounds__(kCvThreads)
cost_volume_fwd_cl(const float4* __restrict__ left, const float4* __restrict__ right,
                   const float* __restrict__ shifts, float4* __restrict__ cost,
                   int N, int C4, int D, int H, int W) {
    __shared__ int s_s0[kCvMaxD];
    __shared__ float s_f[kCvMaxD];
    const int n = blockIdx.y;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float s = fmaxf(__ldg(shifts + n * D + d), 0.f), s0 = floorf(s);
        s_s0[d] = (int)s0;
        s_f[d] = __fsub_rn(s, s0);
    }
    __syncthreads();
    const int Q = 2 * C4;
    const int64_t per_n = (int64_t)H * W * Q;          // float4 per plane
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int q = (int)(i % Q);
        const int64_t p = i / Q;
        const int w = (int)(p % W);
        const int64_t row = (int64_t)n * H * W + (p - w);          // feature row index of (n,h,0)
        float4* out = cost + (int64_t)n * D * per_n + i;
        // blockIdx.z splits the planes into gridDim.z segments: more, shorter work items -> the
        // last wave of blocks is a small fraction of the kernel (tail effect)
        const int dseg = (D + gridDim.z - 1) / gridDim.z;
        const int d_lo = blockIdx.z * dseg, d_hi = min(D, d_lo + dseg);
        if (q < C4) {
            const float4 l = __ldg(left + (row + w) * C4 + q);
#pragma unroll 4
            for (int d = d_lo; d < d_hi; ++d) out[(int64_t)d * per_n] = (w - s_s0[d] >= 0) ? l : zero;
        } else {
            const float4* rrow = right + row * C4 + (q - C4);
#pragma unroll 4
            for (int d = d_lo; d < d_hi; ++d) {
                const int x0 = w - s_s0[d];
                float4 o = zero;
                if (x0 >= 0) {
                    const float4 r0 = __ldg(rrow + (int64_t)min(x0, W - 1) * C4);
                    const int x1 = x0 - 1;
                    const float4 r1 = __ldg(rrow + (int64_t)min(max(x1, 0), W - 1) * C4);
                    o = lerp_valid(r0, r1, s_f[d], x1 >= 0 ? 1.f : 0.f);
                }
                out[(int64_t)d * per_n] = o;
            }
        }
    }
}

// NCDHW forward (the upstream op's layout): one thread per output element.
__global__ void __launch_bounds__(kCvThreads)
cost_volume_fwd_ncdhw(const float* __restrict__ left, const float* __restrict__ right,
                      const float* __restrict__ shifts, float* __restrict__ cost,
                      int N, int C, int D, int H, int W) {
    const int64_t total = (int64_t)N * 2 * C * D * H * W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int w = (int)(i % W);
        int h = (int)((i / W) % H);
        int d = (int)((i / ((int64_t)W * H)) % D);
        int c2 = (int)((i / ((int64_t)W * H * D)) % (2 * C));
        int n = (int)(i / ((int64_t)W * H * D * 2 * C));
        float s = fmaxf(__ldg(shifts + n * D + d), 0.f);
        float s0 = floorf(s);
        float f = __fsub_rn(s, s0);
        int x0 = w - (int)s0;
        float o = 0.f;
        if (x0 >= 0) {
            if (c2 < C) {
                o = __ldg(left + (((int64_t)n * C + c2) * H + h) * W + w);
            } else {
                const float* rp = right + (((int64_t)n * C + (c2 - C)) * H + h) * W;
                float r0 = __ldg(rp + min(x0, W - 1));
                int x1 = x0 - 1;
                float v1 = x1 >= 0 ? 1.f : 0.f;
                float r1 = __ldg(rp + min(max(x1, 0), W - 1));
                o = __fadd_rn(__fmul_rn(__fsub_rn(1.f, f), r0), __fmul_rn(f, __fmul_rn(r1, v1)));
            }
        }
        __stcs(cost + i, o);
    }
}

// Channels-last backward, gather form.  One thread per float4 of (n,h,w, 2C):
//   q <  C4: gL[n,h,w]  = sum_d [w - s0_d >= 0] G[n,d,h,w,q]
//   q >= C4: gR[n,h,x]  = sum_d (1-f_d) G[n,d,h,x+s0_d, C+c]      (x+s0_d   < W)
//                              +   f_d  G[n,d,h,x+s0_d+1, C+c]    (x+s0_d+1 < W)
// Fixed d order -> bitwise deterministic.
__global__ void __launch_bounds__(kCvThreads)
cost_volume_bwd_cl(const float4* __restrict__ gcost, const float* __restrict__ shifts,
                   float4* __restrict__ gleft, float4* __restrict__ gright,
                   int N, int C4, int D, int H, int W) {
    __shared__ int s_s0[kCvMaxD];
    __shared__ float s_f[kCvMaxD];
    const int n = blockIdx.y;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float s = fmaxf(__ldg(shifts + n * D + d), 0.f), s0 = floorf(s);
        s_s0[d] = (int)s0;
        s_f[d] = s - s0;
    }
    __syncthreads();
    const int Q = 2 * C4;
    const int64_t per_n = (int64_t)H * W * Q;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int q = (int)(i % Q);
        const int64_t p = i / Q;
        const int w = (int)(p % W);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* gbase = gcost + (int64_t)n * D * per_n + i;     // (d = 0, h, w, q)
        if (q < C4) {
            // unconditional loads + select: the 8 loads of an unrolled trip are all in flight at once
#pragma unroll 8
            for (int d = 0; d < D; ++d) {
                const float4 g = __ldg(gbase + (int64_t)d * per_n);
                const float m = (w - s_s0[d] >= 0) ? 1.f : 0.f;
                acc.x += m * g.x; acc.y += m * g.y; acc.z += m * g.z; acc.w += m * g.w;
            }
            gleft[((int64_t)n * H * W + p) * C4 + q] = acc;
        } else {
#pragma unroll 8
            for (int d = 0; d < D; ++d) {
                const int sa = s_s0[d];
                const float f = s_f[d];
                // columns w+s0 and w+s0+1, clamped into the row; out-of-row terms get weight 0
                const int c0 = min(w + sa, W - 1) - w, c1 = min(w + sa + 1, W - 1) - w;
                const float4* gp = gbase + (int64_t)d * per_n;
                const float4 g0 = __ldg(gp + (int64_t)c0 * Q);
                const float4 g1 = __ldg(gp + (int64_t)c1 * Q);
                const float a = (w + sa < W) ? 1.f - f : 0.f, bb = (w + sa + 1 < W) ? f : 0.f;
                acc.x += a * g0.x; acc.y += a * g0.y; acc.z += a * g0.z; acc.w += a * g0.w;
                acc.x += bb * g1.x; acc.y += bb * g1.y; acc.z += bb * g1.z; acc.w += bb * g1.w;
            }
            gright[((int64_t)n * H * W + p) * C4 + (q - C4)] = acc;
        }
    }
}

// NCDHW backward, gather form; one thread per feature-gradient element.
__global__ void __launch_bounds__(kCvThreads)
cost_volume_bwd_ncdhw(const float* __restrict__ gcost, const float* __restrict__ shifts,
                      float* __restrict__ gleft, float* __restrict__ gright,
                      int N, int C, int D, int H, int W) {
    const int64_t total = (int64_t)N * 2 * C * H * W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int w = (int)(i % W);
        int h = (int)((i / W) % H);
        int c2 = (int)((i / ((int64_t)W * H)) % (2 * C));
        int n = (int)(i / ((int64_t)W * H * 2 * C));
        const float* gb = gcost + (((int64_t)n * 2 * C + c2) * D) * H * W + (int64_t)h * W;
        const float* sh = shifts + n * D;
        float acc = 0.f;
        if (c2 < C) {
            for (int d = 0; d < D; ++d) {
                int s0 = (int)floorf(fmaxf(__ldg(sh + d), 0.f));
                if (w - s0 >= 0) acc += __ldcs(gb + (int64_t)d * H * W + w);
            }
            gleft[(((int64_t)n * C + c2) * H + h) * W + w] = acc;
        } else {
            for (int d = 0; d < D; ++d) {
                float s = fmaxf(__ldg(sh + d), 0.f);
                float s0f = floorf(s);
                float f = s - s0f;
                int wa = w + (int)s0f;
                if (wa < W) acc += (1.f - f) * __ldcs(gb + (int64_t)d * H * W + wa);
                if (wa + 1 < W) acc += f * __ldcs(gb + (int64_t)d * H * W + wa + 1);
            }
            gright[(((int64_t)n * C + (c2 - C)) * H + h) * W + w] = acc;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Row-staged variants (default for the channels-last layout).  The plane-strided walk of the
// kernels above has every warp touching 48 planes 7.7 MB apart and re-gathers the right features
// through L2 for every plane (measured 3.4 / 2.4 TB/s).  Here a block owns ONE feature row (n, h)
// and a segment of planes:
//   fwd: the right-feature row (W*C floats) is staged in smem once; each plane of the segment is
//        then produced as one contiguous W*2C*4-byte stream (left value from L1, right from smem).
//   bwd: each plane's right-half gradient row is staged in smem (coalesced, read once) and
//        gathered twice from there; per-segment partial sums go to a scratch buffer and a second
//        tiny kernel adds the segments in a fixed order (deterministic).
// ---------------------------------------------------------------------------------------------
constexpr int kCvRowThreads = 512;
constexpr int kCvMaxW32 = 12;                           // per-thread w positions: W <= 32 * 12

// Thread (q, wl): q = threadIdx.x % (2*C4) is the float4 column inside a voxel row (left half
// q < C4, right half otherwise), wl = threadIdx.x / (2*C4) the w lane; w = wl + k * WL.  All index
// math is compile-time strength-reduced (no integer division in the hot loops: at ~6 TB/s the
// budget is only ~50 instructions per 16 B).
template <int C4>
__global__ void __launch_bounds__(kCvRowThreads, 2)
cost_volume_fwd_row(const float4* __restrict__ left, const float4* __restrict__ right,
                    const float* __restrict__ shifts, float4* __restrict__ cost,
                    int D, int H, int W, int dseg) {
    constexpr int Q = 2 * C4, WL = kCvRowThreads / Q;
    extern __shared__ float4 s_r[];                     // [W][C4] right-feature row
    __shared__ int s_s0[kCvMaxD];
    __shared__ float s_f[kCvMaxD];
    const int h = blockIdx.x, n = blockIdx.y;
    const int d_lo = blockIdx.z * dseg, d_hi = min(D, d_lo + dseg);
    const int64_t row = ((int64_t)n * H + h) * W;
    for (int i = threadIdx.x; i < W * C4; i += kCvRowThreads) s_r[i] = __ldg(right + row * C4 + i);
    for (int d = d_lo + threadIdx.x; d < d_hi; d += kCvRowThreads) {
        float s = fmaxf(__ldg(shifts + n * D + d), 0.f), s0 = floorf(s);
        s_s0[d] = (int)s0;
        s_f[d] = __fsub_rn(s, s0);
    }
    const int q = threadIdx.x % Q, wl = threadIdx.x / Q;
    const bool is_left = q < C4;
    float4 lv[kCvMaxW32];                               // left values of this thread's w positions
#pragma unroll
    for (int k = 0; k < kCvMaxW32; ++k) {
        const int w = wl + k * WL;
        lv[k] = (is_left && w < W) ? __ldg(left + (row + w) * C4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int d = d_lo; d < d_hi; ++d) {
        const int s0 = s_s0[d];
        const float f = s_f[d];
        float4* out = cost + ((((int64_t)n * D + d) * H + h) * W) * Q + threadIdx.x;
        if (is_left) {
#pragma unroll
            for (int k = 0; k < kCvMaxW32; ++k) {
                const int w = wl + k * WL;
                if (w < W) stg_stream(out + k * kCvRowThreads, (w - s0 >= 0) ? lv[k] : zero);
            }
        } else {
            const float4* sr = s_r + (q - C4);
#pragma unroll
            for (int k = 0; k < kCvMaxW32; ++k) {
                const int w = wl + k * WL;
                if (w < W) {
                    const int x0 = w - s0;
                    float4 o = zero;
                    if (x0 >= 0) {
                        const float4 r0 = sr[min(x0, W - 1) * C4];
                        const float4 r1 = sr[max(x0 - 1, 0) * C4];
                        o = lerp_valid(r0, r1, f, x0 >= 1 ? 1.f : 0.f);
                    }
                    stg_stream(out + k * kCvRowThreads, o);
                }
            }
        }
    }
}

// partial[seg][n][h][w][2C] (left half then right half of every voxel row)
//
// Round 2: software-pipelined with cp.async.  The first version loaded a plane's row into registers, used it, and only
// then asked for the next plane: one HBM round trip per plane exposed, 120 registers, one block per SM (ncu: DRAM 56 %).
// Now the WHOLE row (both halves, 2C x W floats = 80 KB at KITTI size) of plane d+1 streams into the second smem buffer
// with cp.async (L1-bypassing .cg, no registers) while plane d is consumed from the first; the left half accumulates
// straight from smem, the right half is gathered from it as before.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int C4>
__global__ void __launch_bounds__(kCvRowThreads)
cost_volume_bwd_row(const float4* __restrict__ gcost, const float* __restrict__ shifts,
                    float4* __restrict__ partial, int N, int D, int H, int W, int dseg) {
    constexpr int Q = 2 * C4, WL = kCvRowThreads / Q;
    extern __shared__ float4 s_g[];                     // [2][W][Q]: whole gradient rows of two planes
    __shared__ int s_s0[kCvMaxD];
    __shared__ float s_f[kCvMaxD];
    const int h = blockIdx.x, n = blockIdx.y, seg = blockIdx.z;
    const int d_lo = seg * dseg, d_hi = min(D, d_lo + dseg);
    for (int d = d_lo + threadIdx.x; d < d_hi; d += kCvRowThreads) {
        float s = fmaxf(__ldg(shifts + n * D + d), 0.f), s0 = floorf(s);
        s_s0[d] = (int)s0;
        s_f[d] = s - s0;
    }
    const int q = threadIdx.x % Q, wl = threadIdx.x / Q;
    const bool is_left = q < C4;
    const int row4 = W * Q;                              // float4 per row
    auto prefetch = [&](int d, int buf) {
        const float4* g = gcost + ((((int64_t)n * D + d) * H + h) * W) * Q;
        float4* dst = s_g + buf * row4;
        for (int i = threadIdx.x; i < row4; i += kCvRowThreads) cp_async16(dst + i, g + i);
        cp_async_commit();
    };
    float4 acc[kCvMaxW32];
#pragma unroll
    for (int k = 0; k < kCvMaxW32; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (d_lo < d_hi) prefetch(d_lo, 0);
    for (int d = d_lo; d < d_hi; ++d) {
        const int buf = (d - d_lo) & 1;
        cp_async_wait_all();
        __syncthreads();                                // plane d has landed; everybody is done with the other buffer
        if (d + 1 < d_hi) prefetch(d + 1, buf ^ 1);
        const int s0 = s_s0[d];
        const float f = s_f[d];
        const float4* sg = s_g + buf * row4;
        if (is_left) {
#pragma unroll
            for (int k = 0; k < kCvMaxW32; ++k) {
                const int w = wl + k * WL;
                if (w < W && w - s0 >= 0) {
                    const float4 v = sg[w * Q + q];
                    acc[k].x += v.x; acc[k].y += v.y; acc[k].z += v.z; acc[k].w += v.w;
                }
            }
        } else {
            const float a = 1.f - f;
#pragma unroll
            for (int k = 0; k < kCvMaxW32; ++k) {
                const int wa = wl + k * WL + s0;        // gR[x] += (1-f) G[x+s0] + f G[x+s0+1]
                if (wa < W) {
                    const float4 t = sg[wa * Q + q];
                    acc[k].x += a * t.x; acc[k].y += a * t.y; acc[k].z += a * t.z; acc[k].w += a * t.w;
                }
                if (wa + 1 < W) {
                    const float4 t = sg[(wa + 1) * Q + q];
                    acc[k].x += f * t.x; acc[k].y += f * t.y; acc[k].z += f * t.z; acc[k].w += f * t.w;
                }
            }
        }
    }
    float4* pout = partial + ((((int64_t)seg * N + n) * H + h) * W) * Q + threadIdx.x;
#pragma unroll
    for (int k = 0; k < kCvMaxW32; ++k)
        if (wl + k * WL < W) pout[k * kCvRowThreads] = acc[k];
}

__global__ void __launch_bounds__(256)
cost_volume_bwd_combine(const float4* __restrict__ partial, float4* __restrict__ gleft, float4* __restrict__ gright,
                        int64_t nrows, int C4, int nseg) {
    const int Q = 2 * C4;
    const int64_t total = nrows * Q;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        float4 a = partial[i];
        for (int s = 1; s < nseg; ++s) {
            const float4 v = partial[(int64_t)s * total + i];
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        const int q = (int)(i % Q);
        const int64_t r = i / Q;
        if (q < C4) gleft[r * C4 + q] = a; else gright[r * C4 + (q - C4)] = a;
    }
}

constexpr int kCvBwdSegs = 3;

}  // namespace b2

using namespace b2;

extern "C" int b2_cost_volume_fwd(const float* left, const float* right, const float* shifts,
                                  float* cost, int N, int C, int D, int H, int W, int layout,
                                  void* stream) {
    B2_REQUIRE(left && right && shifts && cost, "cost_volume_fwd: null pointer");
    B2_REQUIRE(N >= 0 && C > 0 && D > 0 && H > 0 && W > 0, "cost_volume_fwd: bad dims");
    if (N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t total = (int64_t)N * 2 * C * D * H * W;
    if (layout == 1) {
        B2_REQUIRE(C % 4 == 0, "cost_volume_fwd: channels-last layout needs C %% 4 == 0 (C=%d)", C);
        B2_REQUIRE(aligned16(left) && aligned16(right) && aligned16(cost), "cost_volume_fwd: pointers must be 16B aligned");
        B2_REQUIRE(D <= kCvMaxD, "cost_volume_fwd: at most %d planes", kCvMaxD);
        const int row_bytes = W * C * 4;
        if (C == 32 && W <= 32 * kCvMaxW32 && H <= 65535 && N <= 65535) {      // row-staged kernel (DSGN: 32 ch)
            static SmemOptIn optin;
            cudaError_t ea = ensure_dynamic_smem(optin, cost_volume_fwd_row<8>, 64 * 1024);
            if (ea != cudaSuccess) { set_error("cost_volume_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(ea)); return (int)ea; }
            const int dseg = D >= 16 ? 4 : D;
            cost_volume_fwd_row<8><<<dim3(H, N, (D + dseg - 1) / dseg), kCvRowThreads, row_bytes, st>>>(
                (const float4*)left, (const float4*)right, shifts, (float4*)cost, D, H, W, dseg);
            return check_launch("cost_volume_fwd(row)");
        }
        int grid = stream_grid((int64_t)H * W * (C / 2), kCvThreads, kNumSMs * 16);
        cost_volume_fwd_cl<<<dim3(grid, N, D >= 16 ? 4 : 1), kCvThreads, 0, st>>>((const float4*)left, (const float4*)right, shifts,
                                                                 (float4*)cost, N, C / 4, D, H, W);
    } else if (layout == 0) {
        int grid = stream_grid(total, kCvThreads, kNumSMs * 32);
        cost_volume_fwd_ncdhw<<<grid, kCvThreads, 0, st>>>(left, right, shifts, cost, N, C, D, H, W);
    } else {
        B2_REQUIRE(false, "cost_volume_fwd: unknown layout %d", layout);
    }
    return check_launch("cost_volume_fwd");
}

extern "C" int64_t b2_cost_volume_bwd_workspace_bytes(int N, int C, int H, int W) {
    return (int64_t)kCvBwdSegs * N * H * W * 2 * C * (int64_t)sizeof(float);
}

extern "C" int b2_cost_volume_bwd(const float* gcost, const float* shifts, float* gleft,
                                  float* gright, int N, int C, int D, int H, int W, int layout,
                                  void* workspace, void* stream) {
    B2_REQUIRE(gcost && shifts && gleft && gright, "cost_volume_bwd: null pointer");
    B2_REQUIRE(N >= 0 && C > 0 && D > 0 && H > 0 && W > 0, "cost_volume_bwd: bad dims");
    if (N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t total = (int64_t)N * 2 * C * H * W;
    if (layout == 1) {
        B2_REQUIRE(C % 4 == 0, "cost_volume_bwd: channels-last layout needs C %% 4 == 0 (C=%d)", C);
        B2_REQUIRE(aligned16(gcost) && aligned16(gleft) && aligned16(gright), "cost_volume_bwd: pointers must be 16B aligned");
        B2_REQUIRE(D <= kCvMaxD, "cost_volume_bwd: at most %d planes", kCvMaxD);
        const int row_bytes = W * C * 4;
        if (workspace && C == 32 && W <= 32 * kCvMaxW32 && H <= 65535 && N <= 65535 && aligned16(workspace)) {
            static SmemOptIn optin;
            cudaError_t ea = ensure_dynamic_smem(optin, cost_volume_bwd_row<8>, 200 * 1024);
            if (ea != cudaSuccess) { set_error("cost_volume_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(ea)); return (int)ea; }
            const int nseg = D >= 2 * kCvBwdSegs ? kCvBwdSegs : 1;
            const int dseg = (D + nseg - 1) / nseg;
            cost_volume_bwd_row<8><<<dim3(H, N, nseg), kCvRowThreads, 4 * row_bytes, st>>>(
                (const float4*)gcost, shifts, (float4*)workspace, N, D, H, W, dseg);
            const int64_t nrows = (int64_t)N * H * W;
            cost_volume_bwd_combine<<<stream_grid(nrows * (C / 2), 256, kNumSMs * 8), 256, 0, st>>>(
                (const float4*)workspace, (float4*)gleft, (float4*)gright, nrows, C / 4, nseg);
            return check_launch("cost_volume_bwd(row)");
        }
        int grid = stream_grid((int64_t)H * W * (C / 2), kCvThreads, kNumSMs * 16);
        cost_volume_bwd_cl<<<dim3(grid, N), kCvThreads, 0, st>>>((const float4*)gcost, shifts, (float4*)gleft,
                                                                 (float4*)gright, N, C / 4, D, H, W);
    } else if (layout == 0) {
        int grid = stream_grid(total, kCvThreads, kNumSMs * 32);
        cost_volume_bwd_ncdhw<<<grid, kCvThreads, 0, st>>>(gcost, shifts, gleft, gright, N, C, D, H, W);
    } else {
        B2_REQUIRE(false, "cost_volume_bwd: unknown layout %d", layout);
    }
    return check_launch("cost_volume_bwd");
}
