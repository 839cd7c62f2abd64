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

// Channels-last forward.  One thread per float4 of the output row; consecutive
// threads write consecutive 16 B -> fully coalesced 368 MB stream.
__global__ void __launch_bounds__(kCvThreads)
cost_volume_fwd_cl(const float4* __restrict__ left, const float4* __restrict__ right,
                   const float* __restrict__ shifts, float4* __restrict__ cost,
                   int N, int C4, int D, int H, int W) {
    const int Q = 2 * C4;  // float4 per output voxel
    const int64_t total = (int64_t)N * D * H * W * Q;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int q = (int)(i % Q);
        int64_t v = i / Q;
        int w = (int)(v % W);
        int h = (int)((v / W) % H);
        int d = (int)((v / ((int64_t)W * H)) % D);
        int n = (int)(v / ((int64_t)W * H * D));
        float s = __ldg(shifts + n * D + d);
        float s0 = floorf(s);
        float f = __fsub_rn(s, s0);
        int x0 = w - (int)s0;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (x0 >= 0) {
            int64_t row = ((int64_t)n * H + h) * W;
            if (q < C4) {
                o = __ldg(left + (row + w) * C4 + q);
            } else {
                int x0c = min(x0, W - 1);
                float4 r0 = __ldg(right + (row + x0c) * C4 + (q - C4));
                int x1 = x0 - 1;
                float v1 = x1 >= 0 ? 1.f : 0.f;
                int x1c = min(max(x1, 0), W - 1);
                float4 r1 = __ldg(right + (row + x1c) * C4 + (q - C4));
                o = lerp_valid(r0, r1, f, v1);
            }
        }
        stg_stream(cost + i, o);
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
        float s = __ldg(shifts + n * D + d);
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
    const int Q = 2 * C4;
    const int64_t total = (int64_t)N * H * W * Q;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int q = (int)(i % Q);
        int64_t p = i / Q;
        int w = (int)(p % W);
        int h = (int)((p / W) % H);
        int n = (int)(p / ((int64_t)W * H));
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const int64_t plane = (int64_t)H * W * Q;
        const float4* gbase = gcost + (int64_t)n * D * plane + (int64_t)h * W * Q + q;
        const float* sh = shifts + n * D;
        if (q < C4) {
#pragma unroll 4
            for (int d = 0; d < D; ++d) {
                int s0 = (int)floorf(__ldg(sh + d));
                if (w - s0 >= 0) {
                    float4 g = ldg_stream(gbase + d * plane + (int64_t)w * Q);
                    acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
                }
            }
            gleft[p * C4 + q] = acc;
        } else {
#pragma unroll 4
            for (int d = 0; d < D; ++d) {
                float s = __ldg(sh + d);
                float s0f = floorf(s);
                float f = s - s0f;
                int wa = w + (int)s0f;
                if (wa < W) {
                    float4 g = ldg_stream(gbase + d * plane + (int64_t)wa * Q);
                    float a = 1.f - f;
                    acc.x += a * g.x; acc.y += a * g.y; acc.z += a * g.z; acc.w += a * g.w;
                }
                if (wa + 1 < W) {
                    float4 g = ldg_stream(gbase + d * plane + (int64_t)(wa + 1) * Q);
                    acc.x += f * g.x; acc.y += f * g.y; acc.z += f * g.z; acc.w += f * g.w;
                }
            }
            gright[p * C4 + (q - C4)] = acc;
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
                int s0 = (int)floorf(__ldg(sh + d));
                if (w - s0 >= 0) acc += __ldcs(gb + (int64_t)d * H * W + w);
            }
            gleft[(((int64_t)n * C + c2) * H + h) * W + w] = acc;
        } else {
            for (int d = 0; d < D; ++d) {
                float s = __ldg(sh + d);
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
        int grid = stream_grid(total / 4, kCvThreads, kNumSMs * 32);
        cost_volume_fwd_cl<<<grid, kCvThreads, 0, st>>>((const float4*)left, (const float4*)right, shifts,
                                                       (float4*)cost, N, C / 4, D, H, W);
    } else if (layout == 0) {
        int grid = stream_grid(total, kCvThreads, kNumSMs * 32);
        cost_volume_fwd_ncdhw<<<grid, kCvThreads, 0, st>>>(left, right, shifts, cost, N, C, D, H, W);
    } else {
        B2_REQUIRE(false, "cost_volume_fwd: unknown layout %d", layout);
    }
    return check_launch("cost_volume_fwd");
}

extern "C" int b2_cost_volume_bwd(const float* gcost, const float* shifts, float* gleft,
                                  float* gright, int N, int C, int D, int H, int W, int layout,
                                  void* stream) {
    B2_REQUIRE(gcost && shifts && gleft && gright, "cost_volume_bwd: null pointer");
    B2_REQUIRE(N >= 0 && C > 0 && D > 0 && H > 0 && W > 0, "cost_volume_bwd: bad dims");
    if (N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t total = (int64_t)N * 2 * C * H * W;
    if (layout == 1) {
        B2_REQUIRE(C % 4 == 0, "cost_volume_bwd: channels-last layout needs C %% 4 == 0 (C=%d)", C);
        B2_REQUIRE(aligned16(gcost) && aligned16(gleft) && aligned16(gright), "cost_volume_bwd: pointers must be 16B aligned");
        int grid = stream_grid(total / 4, kCvThreads, kNumSMs * 32);
        cost_volume_bwd_cl<<<grid, kCvThreads, 0, st>>>((const float4*)gcost, shifts, (float4*)gleft,
                                                       (float4*)gright, N, C / 4, D, H, W);
    } else if (layout == 0) {
        int grid = stream_grid(total, kCvThreads, kNumSMs * 32);
        cost_volume_bwd_ncdhw<<<grid, kCvThreads, 0, st>>>(gcost, shifts, gleft, gright, N, C, D, H, W);
    } else {
        B2_REQUIRE(false, "cost_volume_bwd: unknown layout %d", layout);
    }
    return check_launch("cost_volume_bwd");
}
