// Depth head of the PSV branch, fused: trilinear upsample of the 1-channel cost volume to
// (maxdisp, H, W) -> softmax over depth -> expectation over the plane depths, forward and a
// deterministic backward.  Replaces F.interpolate(cost1, [maxdisp,H,W], 'trilinear') + F.softmax +
// (prob*z).sum (upstream StereoNet, consumed at attack/DSGN/pgd_attack.py:310-317): the stock
// sequence materialises the 368 MB [192,384,1248] tensor three times and its upsample backward
// uses atomics (SURVEY 8f "next" row 1).  Here nothing larger than the cost volume itself
// (5.75 MB) and one [D,H,W] scratch (92 MB) ever exists.
//
// Semantics = ATen upsample_trilinear3d(align_corners=False): src = in/out*(dst+.5)-.5 clamped at 0.
#include "common.cuh"

namespace b2 {

struct Lerp { int i0, i1; float l0, l1; };

__device__ __forceinline__ Lerp lerp_src(int dst, float scale, int in_size) {
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
    Lerp r;
    r.i0 = (int)src;
    if (r.i0 > in_size - 1) r.i0 = in_size - 1;
    r.i1 = r.i0 + (r.i0 < in_size - 1 ? 1 : 0);
    r.l1 = src - (float)r.i0;
    r.l0 = 1.f - r.l1;
    return r;
}

struct DhGeom { int N, D, Hc, Wc, H, W, J; float z0, dz, sd, sh, sw; };

__device__ __forceinline__ float dh_col(const float* __restrict__ cn, int d, const Lerp& ly, const Lerp& lx,
                                        int Hc, int Wc) {
    const float* p = cn + (int64_t)d * Hc * Wc;
    // ATen order: depth outermost, then h, then w
    return ly.l0 * (lx.l0 * __ldg(p + ly.i0 * Wc + lx.i0) + lx.l1 * __ldg(p + ly.i0 * Wc + lx.i1)) +
           ly.l1 * (lx.l0 * __ldg(p + ly.i1 * Wc + lx.i0) + lx.l1 * __ldg(p + ly.i1 * Wc + lx.i1));
}

// One thread per output pixel; online softmax over the J interpolated planes.
// MODE 0: write depth.  MODE 1: also write G[n,d,y,x] = d(loss)/d(column value c[d]).
template <int MODE>
__global__ void __launch_bounds__(256)
depth_head_pixel_kernel(const float* __restrict__ cost, const float* __restrict__ gdepth,
                        float* __restrict__ depth, float* __restrict__ G, DhGeom g,
                        float2* __restrict__ sm_out, const float2* __restrict__ sm_in,
                        const float* __restrict__ depth_in) {
    const int64_t total = (int64_t)g.N * g.H * g.W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % g.W), y = (int)((i / g.W) % g.H), n = (int)(i / ((int64_t)g.W * g.H));
        const Lerp ly = lerp_src(y, g.sh, g.Hc), lx = lerp_src(x, g.sw, g.Wc);
        const float* cn = cost + (int64_t)n * g.D * g.Hc * g.Wc;
        // pass A: running max / sum / weighted sum (skipped in the backward when the forward saved them)
        float m = -INFINITY, s = 0.f, t = 0.f;
        int cur = -1; float c0 = 0.f, c1 = 0.f;
        const bool saved = MODE == 1 && sm_in != nullptr;
        if (saved) { const float2 ms = __ldg(sm_in + i); m = ms.x; s = ms.y; }
        for (int j = 0; !saved && j < g.J; ++j) {
            const Lerp ld = lerp_src(j, g.sd, g.D);
            if (ld.i0 != cur) {
                c0 = (ld.i0 == cur + 1 && cur >= 0) ? c1 : dh_col(cn, ld.i0, ly, lx, g.Hc, g.Wc);
                c1 = ld.i1 == ld.i0 ? c0 : dh_col(cn, ld.i1, ly, lx, g.Hc, g.Wc);
                cur = ld.i0;
            }
            const float v = ld.l0 * c0 + ld.l1 * c1;
            const float z = __fadd_rn(g.z0, __fmul_rn((float)j + 0.5f, g.dz));
            if (v > m) { const float r = __expf(m - v); s *= r; t *= r; m = v; }
            const float e = __expf(v - m);
            s += e; t += e * z;
        }
        const float dep = saved ? __ldg(depth_in + i) : t / s;
        if (MODE == 0) {
            depth[i] = dep;
            if (sm_out) sm_out[i] = make_float2(m, s);
            continue;
        }
        // pass B: g_up[j] = g * p_j * (z_j - depth); accumulate onto the two source planes
        const float go = __ldg(gdepth + i) / s;
        float* Gp = G + (int64_t)n * g.D * g.H * g.W + (int64_t)y * g.W + x;
        const int64_t plane = (int64_t)g.H * g.W;
        cur = -1;
        float a0 = 0.f, a1 = 0.f;           // gradient accumulators of planes cur and cur+1
        int written = 0;                    // planes [0, written) already stored
        for (int j = 0; j < g.J; ++j) {
            const Lerp ld = lerp_src(j, g.sd, g.D);
            if (ld.i0 != cur) {
                if (cur >= 0) {
                    // planes cur .. ld.i0-1 are complete
                    Gp[(int64_t)cur * plane] = a0; written = cur + 1;
                    if (ld.i0 == cur + 1) { a0 = a1; a1 = 0.f; }
                    else {
                        Gp[(int64_t)(cur + 1) * plane] = a1; written = cur + 2;
                        for (int d = cur + 2; d < ld.i0; ++d) { Gp[(int64_t)d * plane] = 0.f; written = d + 1; }
                        a0 = 0.f; a1 = 0.f;
                    }
                } else {
                    for (int d = 0; d < ld.i0; ++d) { Gp[(int64_t)d * plane] = 0.f; written = d + 1; }
                }
                c0 = dh_col(cn, ld.i0, ly, lx, g.Hc, g.Wc);
                c1 = ld.i1 == ld.i0 ? c0 : dh_col(cn, ld.i1, ly, lx, g.Hc, g.Wc);
                cur = ld.i0;
            }
            const float v = ld.l0 * c0 + ld.l1 * c1;
            const float z = __fadd_rn(g.z0, __fmul_rn((float)j + 0.5f, g.dz));
            const float gu = go * __expf(v - m) * (z - dep);
            a0 += ld.l0 * gu;
            if (ld.i1 != ld.i0) a1 += ld.l1 * gu; else a0 += ld.l1 * gu;
        }
        if (cur >= 0) {
            Gp[(int64_t)cur * plane] = a0; written = cur + 1;
            if (cur + 1 < g.D) { Gp[(int64_t)(cur + 1) * plane] = a1; written = cur + 2; }
        }
        for (int d = written; d < g.D; ++d) Gp[(int64_t)d * plane] = 0.f;
    }
}

// Gather: one thread per cost-volume cell; its footprint is the output pixels whose bilinear
// stencil touches (h, w).  Fixed (y, x) order -> deterministic.
__global__ void __launch_bounds__(256)
depth_head_gather_kernel(const float* __restrict__ G, float* __restrict__ gcost, DhGeom g) {
    const int64_t total = (int64_t)g.N * g.D * g.Hc * g.Wc;
    const int ry = (int)ceilf(1.f / g.sh), rx = (int)ceilf(1.f / g.sw);      // output pixels per input cell
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int w = (int)(i % g.Wc), h = (int)((i / g.Wc) % g.Hc);
        const int64_t nd = i / ((int64_t)g.Wc * g.Hc);
        const float* Gp = G + nd * g.H * g.W;
        const int y_lo = max(0, (h - 1) * ry - 1), y_hi = min(g.H - 1, (h + 2) * ry);
        const int x_lo = max(0, (w - 1) * rx - 1), x_hi = min(g.W - 1, (w + 2) * rx);
        float acc = 0.f;
        for (int y = y_lo; y <= y_hi; ++y) {
            const Lerp ly = lerp_src(y, g.sh, g.Hc);
            const float wy = (ly.i0 == h ? ly.l0 : 0.f) + (ly.i1 == h ? ly.l1 : 0.f);
            if (ly.i0 != h && ly.i1 != h) continue;
            for (int x = x_lo; x <= x_hi; ++x) {
                const Lerp lx = lerp_src(x, g.sw, g.Wc);
                if (lx.i0 != w && lx.i1 != w) continue;
                const float wx = (lx.i0 == w ? lx.l0 : 0.f) + (lx.i1 == w ? lx.l1 : 0.f);
                acc += wy * wx * __ldg(Gp + (int64_t)y * g.W + x);
            }
        }
        gcost[i] = acc;
    }
}


// ---------------------------------------------------------------------------------------------
// x4 specialisation (the KITTI configuration: 48x96x312 -> 192x384x1248).  With an integer ratio of 4 the depth
// lerp is periodic: fine planes 4k+2 .. 4k+5 blend coarse planes (k, k+1) with l1 = 1/8, 3/8, 5/8, 7/8, planes 0, 1
// are coarse plane 0 and the last two blend plane D-1 with itself -- the per-plane source-index arithmetic, its
// branches and the per-plane online-softmax rescale (the kernels above are instruction-bound on exactly those) become
// four unrolled FMAs + one exp per plane.  A lerp never exceeds its end points, so the running maximum is taken over
// the COARSE column (one conditional rescale per coarse plane); plane depths come from a shared-memory table.
// Same operation order per plane as the generic kernel (ATen's l0*c0 + l1*c1, ascending planes).
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256, 4)
depth_head_pixel_x4_kernel(const float* __restrict__ cost, const float* __restrict__ gdepth,
                           float* __restrict__ depth, float* __restrict__ G, DhGeom g,
                           float2* __restrict__ sm_out, const float2* __restrict__ sm_in,
                           const float* __restrict__ depth_in) {
    extern __shared__ float zt[];                               // [J] plane depths
    for (int j = threadIdx.x; j < g.J; j += blockDim.x) zt[j] = __fadd_rn(g.z0, __fmul_rn((float)j + 0.5f, g.dz));
    __syncthreads();
    const int64_t total = (int64_t)g.N * g.H * g.W;
    const int D = g.D;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % g.W), y = (int)((i / g.W) % g.H), n = (int)(i / ((int64_t)g.W * g.H));
        const Lerp ly = lerp_src(y, g.sh, g.Hc), lx = lerp_src(x, g.sw, g.Wc);
        const float* cn = cost + (int64_t)n * D * g.Hc * g.Wc;
        float m, s, dep;
        const bool saved = MODE == 1 && sm_in != nullptr;
        if (saved) {
            const float2 ms = __ldg(sm_in + i); m = ms.x; s = ms.y; dep = __ldg(depth_in + i);
        } else {
            float c0 = dh_col(cn, 0, ly, lx, g.Hc, g.Wc), t;
            m = c0; s = 2.f; t = zt[0] + zt[1];                  // planes 0, 1: v = c0, exp(0) = 1
            for (int k = 0; k + 1 < D; ++k) {
                const float c1 = dh_col(cn, k + 1, ly, lx, g.Hc, g.Wc);
                if (c1 > m) { const float r = __expf(m - c1); s *= r; t *= r; m = c1; }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float l1 = 0.125f + 0.25f * q, l0 = 1.f - l1;
                    const float e = __expf(l0 * c0 + l1 * c1 - m);
                    s += e; t += e * zt[4 * k + 2 + q];
                }
                c0 = c1;
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {                       // last two planes: both sources are plane D-1
                const float l1 = 0.125f + 0.25f * q, l0 = 1.f - l1;
                const float e = __expf(l0 * c0 + l1 * c0 - m);
                s += e; t += e * zt[4 * D - 2 + q];
            }
            dep = t / s;
        }
        if (MODE == 0) {
            depth[i] = dep;
            if (sm_out) sm_out[i] = make_float2(m, s);
            continue;
        }
        // g_up[j] = g * p_j * (z_j - depth), accumulated onto its two source planes in ascending plane order
        const float go = __ldg(gdepth + i) / s;
        float* Gp = G + (int64_t)n * D * g.H * g.W + (int64_t)y * g.W + x;
        const int64_t plane = (int64_t)g.H * g.W;
        float c0 = dh_col(cn, 0, ly, lx, g.Hc, g.Wc);
        float a0 = 0.f;
        {
            const float e = go * __expf(c0 - m);
            a0 += e * (zt[0] - dep);
            a0 += e * (zt[1] - dep);
        }
        for (int k = 0; k + 1 < D; ++k) {
            const float c1 = dh_col(cn, k + 1, ly, lx, g.Hc, g.Wc);
            float a1 = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float l1 = 0.125f + 0.25f * q, l0 = 1.f - l1;
                const float gu = go * __expf(l0 * c0 + l1 * c1 - m) * (zt[4 * k + 2 + q] - dep);
                a0 += l0 * gu; a1 += l1 * gu;
            }
            Gp[(int64_t)k * plane] = a0;
            a0 = a1; c0 = c1;
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const float l1 = 0.125f + 0.25f * q, l0 = 1.f - l1;
            const float gu = go * __expf(l0 * c0 + l1 * c0 - m) * (zt[4 * D - 2 + q] - dep);
            a0 += l0 * gu; a0 += l1 * gu;
        }
        Gp[(int64_t)(D - 1) * plane] = a0;
    }
}

// x4 gather: the output rows whose bilinear stencil touches coarse row h are exactly 4h-2 .. 4h+5 (clipped): an 8 x 8
// window with the weights hoisted, instead of a 14 x 14 scan that recomputes source indices per pixel.
__global__ void __launch_bounds__(256)
depth_head_gather_x4_kernel(const float* __restrict__ G, float* __restrict__ gcost, DhGeom g) {
    const int64_t total = (int64_t)g.N * g.D * g.Hc * g.Wc;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int w = (int)(i % g.Wc), h = (int)((i / g.Wc) % g.Hc);
        const int64_t nd = i / ((int64_t)g.Wc * g.Hc);
        const float* Gp = G + nd * g.H * g.W;
        float wx[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int x = 4 * w - 2 + u;
            float v = 0.f;
            if (x >= 0 && x < g.W) {
                const Lerp lx = lerp_src(x, g.sw, g.Wc);
                v = (lx.i0 == w ? lx.l0 : 0.f) + (lx.i1 == w ? lx.l1 : 0.f);
            }
            wx[u] = v;
        }
        float acc = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int y = 4 * h - 2 + r;
            if (y < 0 || y >= g.H) continue;
            const Lerp ly = lerp_src(y, g.sh, g.Hc);
            const float wy = (ly.i0 == h ? ly.l0 : 0.f) + (ly.i1 == h ? ly.l1 : 0.f);
            const float* row = Gp + (int64_t)y * g.W;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int x = 4 * w - 2 + u;
                if (x >= 0 && x < g.W) acc += wy * wx[u] * __ldg(row + x);
            }
        }
        gcost[i] = acc;
    }
}

static bool dh_is_x4(int D, int Hc, int Wc, int H, int W, int J) {
    return flag_value(kFlagDepthHeadX4, "B2_DEPTH_HEAD_X4", 1) && J == 4 * D && H == 4 * Hc && W == 4 * Wc && D >= 2;
}

static DhGeom dh_geom(int N, int D, int Hc, int Wc, int H, int W, int J, float z0, float dz) {
    DhGeom g{N, D, Hc, Wc, H, W, J, z0, dz, (float)D / (float)J, (float)Hc / (float)H, (float)Wc / (float)W};
    return g;
}

}  // namespace b2

using namespace b2;

extern "C" int b2_depth_head_fwd(const float* cost, float* depth, float* sm_stats, int N, int D, int Hc, int Wc,
                                 int H, int W, int J, float z0, float dz, void* stream) {
    B2_REQUIRE(cost && depth, "depth_head_fwd: null pointer");
    B2_REQUIRE(D >= 1 && Hc >= 1 && Wc >= 1 && H >= Hc && W >= Wc && J >= D, "depth_head_fwd: upsampling only");
    B2_REQUIRE(H % Hc == 0 && W % Wc == 0 && J % D == 0, "depth_head_fwd: integer upsampling ratios only (%dx%dx%d -> %dx%dx%d)", D, Hc, Wc, J, H, W);
    int64_t total = (int64_t)N * H * W;
    if (total == 0) return 0;
    if (dh_is_x4(D, Hc, Wc, H, W, J))
        depth_head_pixel_x4_kernel<0><<<stream_grid(total, 256, kNumSMs * 8), 256, J * sizeof(float), (cudaStream_t)stream>>>(
            cost, nullptr, depth, nullptr, dh_geom(N, D, Hc, Wc, H, W, J, z0, dz), (float2*)sm_stats, nullptr, nullptr);
    else
        depth_head_pixel_kernel<0><<<stream_grid(total, 256, kNumSMs * 16), 256, 0, (cudaStream_t)stream>>>(
            cost, nullptr, depth, nullptr, dh_geom(N, D, Hc, Wc, H, W, J, z0, dz), (float2*)sm_stats, nullptr, nullptr);
    return check_launch("depth_head_fwd");
}

extern "C" int64_t b2_depth_head_workspace_bytes(int N, int D, int H, int W) {
    return (int64_t)N * D * H * W * (int64_t)sizeof(float);
}

extern "C" int b2_depth_head_bwd(const float* cost, const float* gdepth, float* gcost, const float* sm_stats,
                                 const float* depth, int N, int D, int Hc, int Wc, int H, int W, int J, float z0,
                                 float dz, void* workspace, void* stream) {
    B2_REQUIRE(cost && gdepth && gcost && workspace, "depth_head_bwd: null pointer");
    B2_REQUIRE(!sm_stats == !depth, "depth_head_bwd: sm_stats and depth (both saved by the forward) go together");
    B2_REQUIRE(D >= 1 && Hc >= 1 && Wc >= 1 && H >= Hc && W >= Wc && J >= D, "depth_head_bwd: upsampling only");
    B2_REQUIRE(H % Hc == 0 && W % Wc == 0 && J % D == 0, "depth_head_bwd: integer upsampling ratios only (the gather window of the "
               "backward is derived from them); got %dx%dx%d -> %dx%dx%d", D, Hc, Wc, J, H, W);
    int64_t total = (int64_t)N * H * W;
    if (total == 0) return 0;
    DhGeom g = dh_geom(N, D, Hc, Wc, H, W, J, z0, dz);
    cudaStream_t st = (cudaStream_t)stream;
    int64_t cells = (int64_t)N * D * Hc * Wc;
    if (dh_is_x4(D, Hc, Wc, H, W, J)) {
        depth_head_pixel_x4_kernel<1><<<stream_grid(total, 256, kNumSMs * 8), 256, J * sizeof(float), st>>>(
            cost, gdepth, nullptr, (float*)workspace, g, nullptr, (const float2*)sm_stats, depth);
        depth_head_gather_x4_kernel<<<stream_grid(cells, 256, kNumSMs * 16), 256, 0, st>>>((const float*)workspace, gcost, g);
    } else {
        depth_head_pixel_kernel<1><<<stream_grid(total, 256, kNumSMs * 16), 256, 0, st>>>(
            cost, gdepth, nullptr, (float*)workspace, g, nullptr, (const float2*)sm_stats, depth);
        depth_head_gather_kernel<<<stream_grid(cells, 256, kNumSMs * 16), 256, 0, st>>>((const float*)workspace, gcost, g);
    }
    return check_launch("depth_head_bwd");
}
