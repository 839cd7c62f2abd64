// Subsystem (3), verification mode: fp32 SIMT implicit-GEMM 3x3x3 convolution
// (impl = 1 of b2_conv3d) plus the bandwidth-bound Cout == 1 head.  The tensor-core
// path lives in conv3d_tcgen05.cu; this kernel is the exact-fp32 cross-check for
// it and covers every (Cin, Cout) % 4 == 0, both modes, stride 1|2.
//
// Gather definition shared by both implementations (see b2attack.h):
//   CONV  : out[o] = sum_k in[o*stride + k - 1] . wp[k]
//   DECONV: out[o] = sum_k [(o+1-k) even, in range] in[(o+1-k)/2] . wp[k]
#include "common.cuh"

namespace b2 {

constexpr int BM = 64, BN = 64, BK = 16, kConvThreads = 256;

struct ConvGeom {
    int N, Cin, Cout, Di, Hi, Wi, Do, Ho, Wo, stride, mode;
};

// input coordinate along one axis for output coordinate o and tap k; -1 if none
__device__ __forceinline__ int in_coord(int o, int k, int stride, int mode, int isz) {
    if (mode == 0) {
        int i = o * stride + k - 1;
        return (i >= 0 && i < isz) ? i : -1;
    }
    int t = o + 1 - k;
    if (t < 0 || (t & 1)) return -1;
    t >>= 1;
    return t < isz ? t : -1;
}

__global__ void __launch_bounds__(kConvThreads)
conv3d_simt_kernel(const float* __restrict__ in, const float* __restrict__ wp,
                   float* __restrict__ out, ConvGeom g) {
    __shared__ float As[2][BK][BM + 4];
    __shared__ float Bs[2][BK][BN + 4];
    const int t = threadIdx.x;
    const int64_t M = (int64_t)g.N * g.Do * g.Ho * g.Wo;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    // loader role: row lr (voxel for A, cout for B), k-quad lq
    const int lr = t >> 2, lq = t & 3;
    int64_t mv = m0 + lr;
    bool mvalid = mv < M;
    int ow = 0, oh = 0, od = 0, nb = 0;
    if (mvalid) {
        ow = (int)(mv % g.Wo); oh = (int)((mv / g.Wo) % g.Ho);
        od = (int)((mv / ((int64_t)g.Wo * g.Ho)) % g.Do); nb = (int)(mv / ((int64_t)g.Wo * g.Ho * g.Do));
    }
    const int co_l = n0 + lr;
    const bool bvalid = co_l < g.Cout;

    // compute role: 4x4 micro tile
    const int ty = t >> 4, tx = t & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int kchunks = (g.Cin + BK - 1) / BK;
    const int steps = 27 * kchunks;

    auto load = [&](int step, float4& a, float4& b) {
        int tap = step / kchunks, c0 = (step % kchunks) * BK + lq * 4;
        int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
        a = make_float4(0.f, 0.f, 0.f, 0.f);
        b = a;
        if (c0 < g.Cin) {
            if (mvalid) {
                int id = in_coord(od, kd, g.stride, g.mode, g.Di);
                int ih = in_coord(oh, kh, g.stride, g.mode, g.Hi);
                int iw = in_coord(ow, kw, g.stride, g.mode, g.Wi);
                if (id >= 0 && ih >= 0 && iw >= 0)
                    a = __ldg(reinterpret_cast<const float4*>(
                        in + ((((int64_t)nb * g.Di + id) * g.Hi + ih) * g.Wi + iw) * g.Cin + c0));
            }
            if (bvalid)
                b = __ldg(reinterpret_cast<const float4*>(wp + ((int64_t)tap * g.Cout + co_l) * g.Cin + c0));
        }
    };
    auto stash = [&](int buf, float4 a, float4 b) {
        As[buf][lq * 4 + 0][lr] = a.x; As[buf][lq * 4 + 1][lr] = a.y;
        As[buf][lq * 4 + 2][lr] = a.z; As[buf][lq * 4 + 3][lr] = a.w;
        Bs[buf][lq * 4 + 0][lr] = b.x; Bs[buf][lq * 4 + 1][lr] = b.y;
        Bs[buf][lq * 4 + 2][lr] = b.z; Bs[buf][lq * 4 + 3][lr] = b.w;
    };

    float4 ra, rb;
    load(0, ra, rb);
    stash(0, ra, rb);
    __syncthreads();
    for (int s = 0; s < steps; ++s) {
        int buf = s & 1;
        if (s + 1 < steps) load(s + 1, ra, rb);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float4 a = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (s + 1 < steps) stash(buf ^ 1, ra, rb);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t m = m0 + ty * 4 + i;
        int co = n0 + tx * 4;
        if (m < M && co < g.Cout)
            *reinterpret_cast<float4*>(out + m * g.Cout + co) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
}

// ---------------------------------------------------------------------------
// Cout == 1 head.  LPV = Cin/4 lanes per voxel.
// ---------------------------------------------------------------------------
template <int LPV>
__global__ void __launch_bounds__(256)
conv3d_c1_fwd_kernel(const float4* __restrict__ in, const float4* __restrict__ w1,
                     float* __restrict__ out, int N, int D, int H, int W) {
    __shared__ float4 ws[27 * LPV];
    for (int i = threadIdx.x; i < 27 * LPV; i += blockDim.x) ws[i] = w1[i];
    __syncthreads();
    const int lane = threadIdx.x % LPV;
    const int64_t nvox = (int64_t)N * D * H * W;
    const int64_t vstride = (int64_t)gridDim.x * (blockDim.x / LPV);
    // all lanes of a warp iterate the same number of times (shuffles below)
    const int64_t vbase0 = (int64_t)blockIdx.x * (blockDim.x / LPV);
    for (int64_t vb = vbase0; vb < nvox; vb += vstride) {
        int64_t v = vb + threadIdx.x / LPV;
        bool ok = v < nvox;
        float acc = 0.f;
        if (ok) {
            int w = (int)(v % W), h = (int)((v / W) % H), d = (int)((v / ((int64_t)W * H)) % D);
            int n = (int)(v / ((int64_t)W * H * D));
            const float4* base = in + (int64_t)n * D * H * W * LPV + lane;
#pragma unroll
            for (int tap = 0; tap < 27; ++tap) {
                int id = d + tap / 9 - 1, ih = h + (tap / 3) % 3 - 1, iw = w + tap % 3 - 1;
                if (id >= 0 && id < D && ih >= 0 && ih < H && iw >= 0 && iw < W) {
                    float4 x = __ldg(base + (((int64_t)id * H + ih) * W + iw) * LPV);
                    float4 k = ws[tap * LPV + lane];
                    acc += x.x * k.x + x.y * k.y + x.z * k.z + x.w * k.w;
                }
            }
        }
#pragma unroll
        for (int o = LPV / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (ok && lane == 0) out[v] = acc;
    }
}

template <int LPV>
__global__ void __launch_bounds__(256)
conv3d_c1_dgrad_kernel(const float* __restrict__ gout, const float4* __restrict__ w1,
                       float4* __restrict__ gin, int N, int D, int H, int W) {
    __shared__ float4 ws[27 * LPV];
    for (int i = threadIdx.x; i < 27 * LPV; i += blockDim.x) ws[i] = w1[i];
    __syncthreads();
    const int lane = threadIdx.x % LPV;
    const int64_t nvox = (int64_t)N * D * H * W;
    const int64_t vstride = (int64_t)gridDim.x * (blockDim.x / LPV);
    for (int64_t v = (int64_t)blockIdx.x * (blockDim.x / LPV) + threadIdx.x / LPV; v < nvox; v += vstride) {
        int w = (int)(v % W), h = (int)((v / W) % H), d = (int)((v / ((int64_t)W * H)) % D);
        int n = (int)(v / ((int64_t)W * H * D));
        const float* gb = gout + (int64_t)n * D * H * W;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int tap = 0; tap < 27; ++tap) {
            // gin[i] = sum_k gout[i - k + 1] * w[k]
            int od = d - (tap / 9) + 1, oh = h - ((tap / 3) % 3) + 1, ow = w - (tap % 3) + 1;
            if (od >= 0 && od < D && oh >= 0 && oh < H && ow >= 0 && ow < W) {
                float g = __ldg(gb + ((int64_t)od * H + oh) * W + ow);
                float4 k = ws[tap * LPV + lane];
                acc.x += g * k.x; acc.y += g * k.y; acc.z += g * k.z; acc.w += g * k.w;
            }
        }
        stg_stream(gin + v * LPV + lane, acc);
    }
}

int conv3d_simt_launch(const float* in, const float* wp, float* out, int N, int Cin, int Cout, int Di,
                       int Hi, int Wi, int Do, int Ho, int Wo, int stride, int mode, cudaStream_t st) {
    if (Cin % 4 != 0 || Cout % 4 != 0) {
        set_error("conv3d(simt): Cin and Cout must be multiples of 4 (got %d, %d)", Cin, Cout);
        return B2_ERR_UNSUPPORTED;
    }
    ConvGeom g{N, Cin, Cout, Di, Hi, Wi, Do, Ho, Wo, stride, mode};
    int64_t M = (int64_t)N * Do * Ho * Wo;
    dim3 grid((unsigned)((M + BM - 1) / BM), (Cout + BN - 1) / BN);
    conv3d_simt_kernel<<<grid, kConvThreads, 0, st>>>(in, wp, out, g);
    return check_launch("conv3d(simt)");
}

}  // namespace b2

using namespace b2;

#define B2_C1_SWITCH(Cin, ...)                                                     \
    switch ((Cin) / 4) {                                                           \
        case 4: { constexpr int LPV = 4; __VA_ARGS__; } break;                            \
        case 8: { constexpr int LPV = 8; __VA_ARGS__; } break;                            \
        case 16: { constexpr int LPV = 16; __VA_ARGS__; } break;                          \
        case 32: { constexpr int LPV = 32; __VA_ARGS__; } break;                          \
        default:                                                                   \
            b2::set_error("conv3d_c1: Cin must be 16, 32, 64 or 128 (got %d)", (Cin)); \
            return B2_ERR_UNSUPPORTED;                                             \
    }

extern "C" int b2_conv3d_c1_fwd(const float* in, const float* w1, float* out, int N, int Cin, int D,
                                int H, int W, void* stream) {
    B2_REQUIRE(in && w1 && out, "conv3d_c1_fwd: null pointer");
    int64_t nvox = (int64_t)N * D * H * W;
    if (nvox == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    B2_C1_SWITCH(Cin, {
        int grid = stream_grid(nvox, 256 / LPV, kNumSMs * 16);
        conv3d_c1_fwd_kernel<LPV><<<grid, 256, 0, st>>>((const float4*)in, (const float4*)w1, out, N, D, H, W);
    });
    return check_launch("conv3d_c1_fwd");
}

extern "C" int b2_conv3d_c1_dgrad(const float* gout, const float* w1, float* gin, int N, int Cin,
                                  int D, int H, int W, void* stream) {
    B2_REQUIRE(gout && w1 && gin, "conv3d_c1_dgrad: null pointer");
    int64_t nvox = (int64_t)N * D * H * W;
    if (nvox == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    B2_C1_SWITCH(Cin, {
        int grid = stream_grid(nvox, 256 / LPV, kNumSMs * 16);
        conv3d_c1_dgrad_kernel<LPV><<<grid, 256, 0, st>>>(gout, (const float4*)w1, (float4*)gin, N, D, H, W);
    });
    return check_launch("conv3d_c1_dgrad");
}
