// Voxel -> BEV hand-off of the detection branch: AvgPool3d over the height axis Y followed by
// folding (C, Y/p) into the channel axis (upstream StereoNet; F.avg_pool3d + permute + reshape in
// the restatement, reached from attack/DSGN/pgd_attack.py:308/:336).  The stock sequence costs a
// pool kernel, two layout copies and -- in backward -- an uncoalesced 299 MB permute copy
// (1.4 ms measured); this is one streaming pass each way.
//   in  v   [N, Z, Y, X, C]      (channels-last 3-D volume)
//   out bev [N, Z, X, C*YY]      (channels-last 2-D map, channel = c*YY + yy, YY = Y/p)
#include "common.cuh"

namespace b2 {

// Thread mapping: blockIdx.z = (n, z) slab, blockIdx.y = output (fwd: yy, bwd: y) row, threads run
// over the (x, float4-of-channels) pairs of that row -> 32-bit index math only (the first version
// spent ~200 instructions per float4 on 64-bit div/mod and ran at 28 % of the HBM peak in backward).
__global__ void __launch_bounds__(256)
bev_pool_fwd_kernel(const float4* __restrict__ v, float* __restrict__ bev, int Y, int X, int C4, int p) {
    const int YY = Y / p;
    const int yy = blockIdx.y;
    const int64_t nz = blockIdx.z;
    const float inv = 1.f / (float)p;
    const int row = X * C4;
    const float4* vin = v + (nz * Y + (int64_t)yy * p) * row;
    float* orow = bev + nz * (int64_t)row * 4 * YY + yy;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < row; j += gridDim.x * blockDim.x) {
        const int x = j / C4, c4 = j - x * C4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k < p; ++k) {
            float4 t = ldg_stream(vin + (int64_t)k * row + j);
            acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
        float* o = orow + ((int64_t)x * C4 * 4 + c4 * 4) * YY;
        o[0] = acc.x * inv; o[YY] = acc.y * inv; o[2 * YY] = acc.z * inv; o[3 * YY] = acc.w * inv;
    }
}

// backward: one block row per pooling WINDOW yy: the strided gradient values are gathered once and
// written to the p voxel rows of the window (and zeros to the rows past the last whole window).
__global__ void __launch_bounds__(256)
bev_pool_bwd_kernel(const float* __restrict__ gbev, float4* __restrict__ gv, int Y, int X, int C4, int p) {
    const int YY = Y / p;
    const int yy = blockIdx.y;
    const int64_t nz = blockIdx.z;
    const float inv = 1.f / (float)p;
    const int row = X * C4;
    float4* obase = gv + (nz * Y + (int64_t)yy * p) * (int64_t)row;
    const float* srow = gbev + nz * (int64_t)row * 4 * YY + yy;
    const int tail = (yy == YY - 1) ? Y - YY * p : 0;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < row; j += gridDim.x * blockDim.x) {
        const int x = j / C4, c4 = j - x * C4;
        const float* s = srow + ((int64_t)x * C4 * 4 + c4 * 4) * YY;
        const float4 g = make_float4(__ldg(s) * inv, __ldg(s + YY) * inv, __ldg(s + 2 * YY) * inv, __ldg(s + 3 * YY) * inv);
        for (int k = 0; k < p; ++k) stg_stream(obase + (int64_t)k * row + j, g);
        for (int k = 0; k < tail; ++k) stg_stream(obase + (int64_t)(p + k) * row + j, make_float4(0.f, 0.f, 0.f, 0.f));
    }
}

}  // namespace b2

using namespace b2;

extern "C" int b2_bev_pool_fwd(const float* v, float* bev, int N, int C, int Z, int Y, int X, int p, void* stream) {
    B2_REQUIRE(v && bev, "bev_pool_fwd: null pointer");
    B2_REQUIRE(C % 4 == 0 && p >= 1 && Y >= p, "bev_pool_fwd: need C %% 4 == 0 and 1 <= p <= Y");
    int64_t total = (int64_t)N * Z * (Y / p) * X * (C / 4);
    if (total == 0) return 0;
    B2_REQUIRE((int64_t)N * Z <= 65535 && Y / p <= 65535 && (int64_t)X * (C / 4) < (1 << 30), "bev_pool_fwd: dims too large");
    const int row = X * (C / 4);
    bev_pool_fwd_kernel<<<dim3((row + 1023) / 1024, Y / p, N * Z), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)v, bev, Y, X, C / 4, p);
    return check_launch("bev_pool_fwd");
}

extern "C" int b2_bev_pool_bwd(const float* gbev, float* gv, int N, int C, int Z, int Y, int X, int p, void* stream) {
    B2_REQUIRE(gbev && gv, "bev_pool_bwd: null pointer");
    B2_REQUIRE(C % 4 == 0 && p >= 1 && Y >= p, "bev_pool_bwd: need C %% 4 == 0 and 1 <= p <= Y");
    int64_t total = (int64_t)N * Z * Y * X * (C / 4);
    if (total == 0) return 0;
    B2_REQUIRE((int64_t)N * Z <= 65535 && Y <= 65535 && (int64_t)X * (C / 4) < (1 << 30), "bev_pool_bwd: dims too large");
    const int row = X * (C / 4);
    bev_pool_bwd_kernel<<<dim3((row + 1023) / 1024, Y / p, N * Z), 256, 0, (cudaStream_t)stream>>>(
        gbev, (float4*)gv, Y, X, C / 4, p);
    return check_launch("bev_pool_bwd");
}
