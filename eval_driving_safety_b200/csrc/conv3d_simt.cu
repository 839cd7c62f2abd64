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
// Cout == 1 head (classif1's last layer) and its data gradient.  Bandwidth-bound by
// construction (368 MB in / 5.75 MB out per KITTI pair); the work is arranged so that the
// 64-channel volume is streamed exactly once, fully coalesced:
//   fwd  pass A: one thread per voxel of a 128-voxel tile (32 KB contiguous, staged in smem)
//                computes the 27 tap products P[t][v] = x[v] . w[t]   (tap-major scratch)
//        pass B: out[o] = sum_t P[t][o + off_t]  (27 coalesced reads per output voxel)
//   dgrad      : one thread per voxel gathers its 27 g values, forms the 64-channel row in
//                registers/smem and the block stores the 32 KB tile coalesced.
// Cin is a template parameter (C4 = Cin/4 float4 per voxel row).
// ---------------------------------------------------------------------------
constexpr int kC1Tile = 128;

template <int C4>
__global__ void __launch_bounds__(kC1Tile)
conv3d_c1_partial_kernel(const float4* __restrict__ in, const float4* __restrict__ w1, float* __restrict__ P,
                         int64_t nvox) {
    __shared__ float4 xs[kC1Tile][C4 + 1];
    __shared__ float4 ws[27 * C4];
    for (int i = threadIdx.x; i < 27 * C4; i += kC1Tile) ws[i] = w1[i];
    const int64_t v0 = (int64_t)blockIdx.x * kC1Tile;
    for (int i = threadIdx.x; i < kC1Tile * C4; i += kC1Tile) {
        int64_t gi = v0 * C4 + i;
        xs[i / C4][i % C4] = gi < nvox * C4 ? ldg_stream(in + gi) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    float acc[27];
#pragma unroll
    for (int t = 0; t < 27; ++t) acc[t] = 0.f;
#pragma unroll 2
    for (int c = 0; c < C4; ++c) {
        const float4 x = xs[threadIdx.x][c];
#pragma unroll
        for (int t = 0; t < 27; ++t) {
            const float4 k = ws[t * C4 + c];
            acc[t] = fmaf(x.x, k.x, fmaf(x.y, k.y, fmaf(x.z, k.z, fmaf(x.w, k.w, acc[t]))));
        }
    }
    const int64_t v = v0 + threadIdx.x;
    if (v < nvox) {
#pragma unroll
        for (int t = 0; t < 27; ++t) P[(int64_t)t * nvox + v] = acc[t];
    }
}

__global__ void __launch_bounds__(256, 6)
conv3d_c1_gather_kernel(const float* __restrict__ P, float* __restrict__ out, int N, int D, int H, int W) {
    const int64_t nvox = (int64_t)N * D * H * W;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += (int64_t)gridDim.x * blockDim.x) {
        const int w = (int)(v % W), h = (int)((v / W) % H), d = (int)((v / ((int64_t)W * H)) % D);
        float acc = 0.f;
#pragma unroll
        for (int t = 0; t < 27; ++t) {
            const int dd = t / 9 - 1, dh = (t / 3) % 3 - 1, dw = t % 3 - 1;
            const int id = d + dd, ih = h + dh, iw = w + dw;
            if (id >= 0 && id < D && ih >= 0 && ih < H && iw >= 0 && iw < W)
                acc += __ldg(P + (int64_t)t * nvox + v + ((int64_t)dd * H + dh) * W + dw);
        }
        out[v] = acc;
    }
}

template <int C4>
__global__ void __launch_bounds__(kC1Tile)
conv3d_c1_dgrad_kernel(const float* __restrict__ gout, const float4* __restrict__ w1, float4* __restrict__ gin,
                       int N, int D, int H, int W) {
    __shared__ float4 os[kC1Tile][C4 + 1];
    __shared__ float4 ws[27 * C4];
    for (int i = threadIdx.x; i < 27 * C4; i += kC1Tile) ws[i] = w1[i];
    const int64_t nvox = (int64_t)N * D * H * W;
    const int64_t v0 = (int64_t)blockIdx.x * kC1Tile;
    const int64_t v = v0 + threadIdx.x;
    float g[27];
#pragma unroll
    for (int t = 0; t < 27; ++t) g[t] = 0.f;
    if (v < nvox) {
        const int w = (int)(v % W), h = (int)((v / W) % H), d = (int)((v / ((int64_t)W * H)) % D);
#pragma unroll
        for (int t = 0; t < 27; ++t) {
            // gin[i] = sum_k gout[i - k + 1] * w[k]
            const int dd = 1 - t / 9, dh = 1 - (t / 3) % 3, dw = 1 - t % 3;
            const int od = d + dd, oh = h + dh, ow = w + dw;
            if (od >= 0 && od < D && oh >= 0 && oh < H && ow >= 0 && ow < W)
                g[t] = __ldg(gout + v + ((int64_t)dd * H + dh) * W + dw);
        }
    }
    __syncthreads();
#pragma unroll 2
    for (int c = 0; c < C4; ++c) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int t = 0; t < 27; ++t) {
            const float4 k = ws[t * C4 + c];
            a.x = fmaf(g[t], k.x, a.x); a.y = fmaf(g[t], k.y, a.y); a.z = fmaf(g[t], k.z, a.z); a.w = fmaf(g[t], k.w, a.w);
        }
        os[threadIdx.x][c] = a;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kC1Tile * C4; i += kC1Tile) {
        int64_t gi = v0 * C4 + i;
        if (gi < nvox * C4) stg_stream(gin + gi, os[i / C4][i % C4]);
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
        case 4: { constexpr int C4 = 4; __VA_ARGS__; } break;                      \
        case 8: { constexpr int C4 = 8; __VA_ARGS__; } break;                      \
        case 16: { constexpr int C4 = 16; __VA_ARGS__; } break;                    \
        default:                                                                   \
            b2::set_error("conv3d_c1: Cin must be 16, 32 or 64 (got %d)", (Cin)); \
            return B2_ERR_UNSUPPORTED;                                             \
    }

extern "C" int64_t b2_conv3d_c1_workspace_bytes(int N, int D, int H, int W) {
    return (int64_t)27 * N * D * H * W * (int64_t)sizeof(float);
}

extern "C" int b2_conv3d_c1_fwd(const float* in, const float* w1, float* out, int N, int Cin, int D,
                                int H, int W, void* workspace, void* stream) {
    B2_REQUIRE(in && w1 && out && workspace, "conv3d_c1_fwd: null pointer");
    int64_t nvox = (int64_t)N * D * H * W;
    if (nvox == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned tiles = (unsigned)((nvox + kC1Tile - 1) / kC1Tile);
    B2_C1_SWITCH(Cin, {
        conv3d_c1_partial_kernel<C4><<<tiles, kC1Tile, 0, st>>>((const float4*)in, (const float4*)w1,
                                                               (float*)workspace, nvox);
    });
    conv3d_c1_gather_kernel<<<stream_grid(nvox, 256, kNumSMs * 16), 256, 0, st>>>((const float*)workspace, out, N, D, H, W);
    return check_launch("conv3d_c1_fwd");
}

extern "C" int b2_conv3d_c1_dgrad(const float* gout, const float* w1, float* gin, int N, int Cin,
                                  int D, int H, int W, void* stream) {
    B2_REQUIRE(gout && w1 && gin, "conv3d_c1_dgrad: null pointer");
    int64_t nvox = (int64_t)N * D * H * W;
    if (nvox == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned tiles = (unsigned)((nvox + kC1Tile - 1) / kC1Tile);
    B2_C1_SWITCH(Cin, {
        conv3d_c1_dgrad_kernel<C4><<<tiles, kC1Tile, 0, st>>>(gout, (const float4*)w1, (float4*)gin, N, D, H, W);
    });
    return check_launch("conv3d_c1_dgrad");
}
