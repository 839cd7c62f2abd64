// Voxel -> BEV hand-off of the detection branch: AvgPool3d over the height axis Y followed by
// folding (C, Y/p) into the channel axis (upstream StereoNet; F.avg_pool3d + permute + reshape in
// the restatement, reached from attack/DSGN/pgd_attack.py:308/:336).  The stock sequence costs a
// pool kernel, two layout copies and -- in backward -- an uncoalesced 299 MB permute copy
// (1.4 ms measured); this is one streaming pass each way.
//   in  v   [N, Z, Y, X, C]      (channels-last 3-D volume)
//   out bev [N, Z, X, C*YY]      (channels-last 2-D map, channel = c*YY + yy, YY = Y/p)
#include "common.cuh"

namespace b2 {

__global__ void __launch_bounds__(256)
bev_pool_fwd_kernel(const float4* __restrict__ v, float* __restrict__ bev, int N, int Z, int Y, int X, int C4,
                    int p) {
    const int YY = Y / p;
    const int64_t total = (int64_t)N * Z * YY * X * C4;
    const float inv = 1.f / (float)p;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        const int x = (int)((i / C4) % X);
        const int yy = (int)((i / ((int64_t)C4 * X)) % YY);
        const int64_t nz = i / ((int64_t)C4 * X * YY);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k < p; ++k) {
            float4 t = ldg_stream(v + ((nz * Y + yy * p + k) * X + x) * C4 + c4);
            acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
        float* o = bev + (nz * X + x) * (int64_t)(C4 * 4 * YY) + (int64_t)(c4 * 4) * YY + yy;
        o[0] = acc.x * inv; o[YY] = acc.y * inv; o[2 * YY] = acc.z * inv; o[3 * YY] = acc.w * inv;
    }
}

__global__ void __launch_bounds__(256)
bev_pool_bwd_kernel(const float* __restrict__ gbev, float4* __restrict__ gv, int N, int Z, int Y, int X, int C4,
                    int p) {
    const int YY = Y / p;
    const int64_t total = (int64_t)N * Z * Y * X * C4;
    const float inv = 1.f / (float)p;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        const int x = (int)((i / C4) % X);
        const int y = (int)((i / ((int64_t)C4 * X)) % Y);
        const int64_t nz = i / ((int64_t)C4 * X * Y);
        const int yy = y / p;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (yy < YY) {
            const float* s = gbev + (nz * X + x) * (int64_t)(C4 * 4 * YY) + (int64_t)(c4 * 4) * YY + yy;
            g = make_float4(__ldg(s) * inv, __ldg(s + YY) * inv, __ldg(s + 2 * YY) * inv, __ldg(s + 3 * YY) * inv);
        }
        stg_stream(gv + i, g);
    }
}

}  // namespace b2

using namespace b2;

extern "C" int b2_bev_pool_fwd(const float* v, float* bev, int N, int C, int Z, int Y, int X, int p, void* stream) {
    B2_REQUIRE(v && bev, "bev_pool_fwd: null pointer");
    B2_REQUIRE(C % 4 == 0 && p >= 1 && Y >= p, "bev_pool_fwd: need C %% 4 == 0 and 1 <= p <= Y");
    int64_t total = (int64_t)N * Z * (Y / p) * X * (C / 4);
    if (total == 0) return 0;
    bev_pool_fwd_kernel<<<stream_grid(total, 256, kNumSMs * 16), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)v, bev, N, Z, Y, X, C / 4, p);
    return check_launch("bev_pool_fwd");
}

extern "C" int b2_bev_pool_bwd(const float* gbev, float* gv, int N, int C, int Z, int Y, int X, int p, void* stream) {
    B2_REQUIRE(gbev && gv, "bev_pool_bwd: null pointer");
    B2_REQUIRE(C % 4 == 0 && p >= 1 && Y >= p, "bev_pool_bwd: need C %% 4 == 0 and 1 <= p <= Y");
    int64_t total = (int64_t)N * Z * Y * X * (C / 4);
    if (total == 0) return 0;
    bev_pool_bwd_kernel<<<stream_grid(total, 256, kNumSMs * 16), 256, 0, (cudaStream_t)stream>>>(
        gbev, (float4*)gv, N, Z, Y, X, C / 4, p);
    return check_launch("bev_pool_bwd");
}
