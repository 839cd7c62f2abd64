// Subsystem (4): fused perturbation update, patch blend and patch update.
// Replaces attack/DSGN/pgd_attack.py:339-354 (+196-207), attack/DSGN/patch_attack.py:
// 369-376 and 416-430, attack/Stereo-RCNN/pgd_attack.py:177-217 -- see b2attack.h.
//
// Roofline: pure HBM streaming (3 reads + 1 write per element, 16 B/element).
// The arithmetic uses explicit round-to-nearest intrinsics so that no FMA
// contraction happens: results are bit-identical to the reference's sequence of
// separate fp32 ATen ops.
#include "common.cuh"

namespace b2 {

struct PgdSets {
    const float* x[4];
    const float* g[4];
    const float* c[4];
    float* o[4];
};

struct ChanParams {
    float mean[4], std_[4], lo[4], hi[4];
};

__device__ __forceinline__ float sgnf(float g) { return (float)((g > 0.f) - (g < 0.f)); }

__device__ __forceinline__ float pgd_elem(float x, float g, float c, float alpha, float eps,
                                          bool denorm, float mean, float sd, float lo, float hi) {
    float v = denorm ? __fadd_rn(__fmul_rn(x, sd), mean) : x;
    float adv = __fadd_rn(v, __fmul_rn(alpha, sgnf(g)));
    float eta = fminf(fmaxf(__fsub_rn(adv, c), -eps), eps);
    float o = fminf(fmaxf(__fadd_rn(c, eta), lo), hi);
    return denorm ? __fdiv_rn(__fsub_rn(o, mean), sd) : o;
}

constexpr int kPgdThreads = 256;
constexpr int kPgdUnroll = 4;

// One launch over every image of every set.  blockIdx.y = (set, image, channel) plane, so the
// channel constants are block-uniform and no thread ever divides; blockIdx.x tiles the plane in
// chunks of kPgdThreads*kPgdUnroll float4 (loads of a chunk are all issued before the first use).
__global__ void __launch_bounds__(kPgdThreads)
pgd_update_vec4(PgdSets s, int planes_per_set, int C, int hw4, float alpha, float eps, int denorm, ChanParams cp) {
    const int plane = blockIdx.y;
    const int si = plane / planes_per_set, pl = plane - si * planes_per_set;
    const int ch = pl % C;
    const float m = cp.mean[ch], sd = cp.std_[ch], lo = cp.lo[ch], hi = cp.hi[ch];
    const int64_t off = (int64_t)pl * hw4;
    const float4* x = reinterpret_cast<const float4*>(s.x[si]) + off;
    const float4* g = reinterpret_cast<const float4*>(s.g[si]) + off;
    const float4* c = reinterpret_cast<const float4*>(s.c[si]) + off;
    float4* o = reinterpret_cast<float4*>(s.o[si]) + off;
    const int tile = kPgdThreads * kPgdUnroll;
    for (int base = blockIdx.x * tile; base < hw4; base += gridDim.x * tile) {
        float4 xv[kPgdUnroll], gv[kPgdUnroll], cv[kPgdUnroll];
#pragma unroll
        for (int u = 0; u < kPgdUnroll; ++u) {
            const int i = base + u * kPgdThreads + threadIdx.x;
            if (i < hw4) { xv[u] = ldg_stream(x + i); gv[u] = ldg_stream(g + i); cv[u] = ldg_stream(c + i); }
        }
#pragma unroll
        for (int u = 0; u < kPgdUnroll; ++u) {
            const int i = base + u * kPgdThreads + threadIdx.x;
            if (i >= hw4) continue;
            float4 r;
            r.x = pgd_elem(xv[u].x, gv[u].x, cv[u].x, alpha, eps, denorm, m, sd, lo, hi);
            r.y = pgd_elem(xv[u].y, gv[u].y, cv[u].y, alpha, eps, denorm, m, sd, lo, hi);
            r.z = pgd_elem(xv[u].z, gv[u].z, cv[u].z, alpha, eps, denorm, m, sd, lo, hi);
            r.w = pgd_elem(xv[u].w, gv[u].w, cv[u].w, alpha, eps, denorm, m, sd, lo, hi);
            stg_stream(o + i, r);
        }
    }
}

// Scalar path for hw % 4 != 0 or unaligned pointers.
__global__ void __launch_bounds__(kPgdThreads)
pgd_update_scalar(PgdSets s, int n_sets, int64_t per_set, int64_t hw, int C, float alpha, float eps,
                  int denorm, ChanParams cp) {
    const int64_t total = per_set * n_sets;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int si = (int)(i / per_set);
        int64_t li = i - (int64_t)si * per_set;
        int ch = (int)((li / hw) % C);
        s.o[si][li] = pgd_elem(s.x[si][li], s.g[si][li], s.c[si][li], alpha, eps, denorm,
                               cp.mean[ch], cp.std_[ch], cp.lo[ch], cp.hi[ch]);
    }
}

static int fill_chan(ChanParams& cp, int C, int denorm, const float* mean, const float* std_,
                     const float* lo, const float* hi) {
    for (int c = 0; c < 4; ++c) {
        cp.mean[c] = 0.f; cp.std_[c] = 1.f; cp.lo[c] = 0.f; cp.hi[c] = 1.f;
    }
    for (int c = 0; c < C; ++c) {
        if (denorm) {
            if (!mean || !std_) { set_error("pgd_update: denorm needs mean/std"); return B2_ERR_BAD_ARG; }
            cp.mean[c] = mean[c]; cp.std_[c] = std_[c];
        }
        if (lo) cp.lo[c] = lo[c];
        if (hi) cp.hi[c] = hi[c];
    }
    return 0;
}

// ---------------------------------------------------------------------------
// L2 variant: three launches, fixed-order reductions (deterministic).
// ws layout: double part_g[n_img][kL2Blocks], double part_e[n_img][kL2Blocks]
// ---------------------------------------------------------------------------
constexpr int kL2Blocks = 128;
constexpr int kL2Threads = 256;

__device__ __forceinline__ double block_sum(double v) {
    __shared__ double sh[kL2Threads / 32];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < kL2Threads / 32; ++i) t += sh[i];
    return t;  // valid on thread 0
}

__global__ void __launch_bounds__(kL2Threads)
l2_gnorm_partials(const float* g, int64_t per_img, double* part_g) {
    int img = blockIdx.y;
    const float* gp = g + (int64_t)img * per_img;
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kL2Threads + threadIdx.x; i < per_img;
         i += (int64_t)kL2Blocks * kL2Threads) {
        float v = gp[i];
        acc += (double)v * (double)v;
    }
    double t = block_sum(acc);
    if (threadIdx.x == 0) part_g[img * kL2Blocks + blockIdx.x] = t;
}

__device__ __forceinline__ float reduce_partials_to_norm(const double* part, int img) {
    // every thread sums the same kL2Blocks values in the same order
    double t = 0.0;
    for (int i = 0; i < kL2Blocks; ++i) t += part[img * kL2Blocks + i];
    return fmaxf((float)sqrt(t), 1e-12f);
}

// eta (unprojected) is written to `out`; partial sums of eta^2 to part_e.
__global__ void __launch_bounds__(kL2Threads)
l2_step(const float* x, const float* g, const float* c, float* out, int64_t per_img, int64_t hw,
        int C, float alpha, int denorm, ChanParams cp, const double* part_g, double* part_e) {
    int img = blockIdx.y;
    float gnorm = reduce_partials_to_norm(part_g, img);
    float scale = alpha / gnorm;
    int64_t off = (int64_t)img * per_img;
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kL2Threads + threadIdx.x; i < per_img;
         i += (int64_t)kL2Blocks * kL2Threads) {
        int ch = (int)((i / hw) % C);
        float xv = x[off + i];
        float v = denorm ? __fadd_rn(__fmul_rn(xv, cp.std_[ch]), cp.mean[ch]) : xv;
        float adv = __fadd_rn(v, __fmul_rn(scale, g[off + i]));
        float eta = __fsub_rn(adv, c[off + i]);
        out[off + i] = eta;
        acc += (double)eta * (double)eta;
    }
    double t = block_sum(acc);
    if (threadIdx.x == 0) part_e[img * kL2Blocks + blockIdx.x] = t;
}

__global__ void __launch_bounds__(kL2Threads)
l2_project(const float* c, float* out, int64_t per_img, int64_t hw, int C, float eps, int denorm,
           ChanParams cp, const double* part_e) {
    int img = blockIdx.y;
    float enorm = reduce_partials_to_norm(part_e, img);
    float factor = fminf(eps / enorm, 1.0f);
    int64_t off = (int64_t)img * per_img;
    for (int64_t i = (int64_t)blockIdx.x * kL2Threads + threadIdx.x; i < per_img;
         i += (int64_t)kL2Blocks * kL2Threads) {
        int ch = (int)((i / hw) % C);
        float eta = __fmul_rn(out[off + i], factor);
        float o = fminf(fmaxf(__fadd_rn(c[off + i], eta), cp.lo[ch]), cp.hi[ch]);
        out[off + i] = denorm ? __fdiv_rn(__fsub_rn(o, cp.mean[ch]), cp.std_[ch]) : o;
    }
}

// ---------------------------------------------------------------------------
// Patch kernels.  One thread per (image, channel, box pixel).
// ---------------------------------------------------------------------------
struct Centers { int cy[8]; int cx[8]; };

// centers_dev != nullptr: the centres are read from device memory ([n_img][2] = (cy, cx)) instead of the
// launch arguments, so a captured CUDA graph can be replayed for images with other patch positions
__global__ void patch_apply_kernel(float* img, const float* patch, int n_img, int C, int H, int W,
                                   Centers ctr, int radius, const int* __restrict__ centers_dev) {
    if (centers_dev)
        for (int n = 0; n < n_img && n < 8; ++n) { ctr.cy[n] = centers_dev[2 * n]; ctr.cx[n] = centers_dev[2 * n + 1]; }
    const int dim = 2 * radius + 1;
    const int64_t total = (int64_t)n_img * C * dim * dim;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int px = (int)(i % dim);
        int py = (int)((i / dim) % dim);
        int c = (int)((i / ((int64_t)dim * dim)) % C);
        int n = (int)(i / ((int64_t)dim * dim * C));
        int y = ctr.cy[n] - radius + py, x = ctr.cx[n] - radius + px;
        if (y < 0 || y >= H || x < 0 || x >= W) continue;
        int dy = py - radius, dx = px - radius;
        float m = (dy * dy + dx * dx <= radius * radius) ? 1.f : 0.f;
        float* p = img + (((int64_t)n * C + c) * H + y) * W + x;
        float pv = patch[((int64_t)c * dim + py) * dim + px];
        // (1 - m) * img + m * pad, same op order as the reference
        *p = __fadd_rn(__fmul_rn(__fsub_rn(1.f, m), *p), __fmul_rn(m, pv));
    }
}

__global__ void patch_update_kernel(float* patch, const float* gL, const float* gR, int C, int H,
                                    int W, int cyL, int cxL, int cyR, int cxR, int radius,
                                    float half_alpha, float eps, int has_range, ChanParams cp,
                                    float* delta_out, const int* __restrict__ centers_dev) {
    if (centers_dev) {                  // (cyL, cxL, cyR, cxR) from device memory: graph-replayable
        cyL = centers_dev[0]; cxL = centers_dev[1]; cyR = centers_dev[2]; cxR = centers_dev[3];
        if (cyL - radius < 0 || cyL + radius >= H || cxL - radius < 0 || cxL + radius >= W ||
            cyR - radius < 0 || cyR + radius >= H || cxR - radius < 0 || cxR + radius >= W) {
            // a box that leaves the frame cannot be validated on the host here: contribute a zero step
            const int tot = C * (2 * radius + 1) * (2 * radius + 1);
            if (delta_out)
                for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += gridDim.x * blockDim.x) delta_out[i] = 0.f;
            return;
        }
    }
    const int dim = 2 * radius + 1;
    const int total = C * dim * dim;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int px = i % dim, py = (i / dim) % dim, c = i / (dim * dim);
        int64_t oL = ((int64_t)c * H + (cyL - radius + py)) * W + (cxL - radius + px);
        int64_t oR = ((int64_t)c * H + (cyR - radius + py)) * W + (cxR - radius + px);
        float d = __fmul_rn(half_alpha, __fadd_rn(gL[oL], gR[oR]));
        d = fminf(fmaxf(d, -eps), eps);
        if (delta_out) { delta_out[i] = d; continue; }
        float v = __fsub_rn(patch[i], d);
        if (has_range) v = fminf(fmaxf(v, cp.lo[c]), cp.hi[c]);
        patch[i] = v;
    }
}

// patch -= delta (the all-reduced clipped step), then the optional per-channel range clamp
__global__ void patch_axpy_kernel(float* patch, const float* delta, int C, int per_c, int has_range, ChanParams cp) {
    const int total = C * per_c;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        float v = __fsub_rn(patch[i], delta[i]);
        if (has_range) { const int c = i / per_c; v = fminf(fmaxf(v, cp.lo[c]), cp.hi[c]); }
        patch[i] = v;
    }
}

}  // namespace b2

using namespace b2;

extern "C" int b2_pgd_update(const float* const* x, const float* const* g, const float* const* clean,
                             float* const* out, int n_sets, int n_img, int C, int64_t hw,
                             float alpha, float eps, int denorm, const float* mean,
                             const float* std_, const float* lo, const float* hi, void* stream) {
    B2_REQUIRE(n_sets >= 1 && n_sets <= 4, "pgd_update: n_sets must be 1..4 (got %d)", n_sets);
    B2_REQUIRE(C >= 1 && C <= 4, "pgd_update: C must be 1..4 (got %d)", C);
    B2_REQUIRE(n_img >= 0 && hw >= 0, "pgd_update: negative size");
    ChanParams cp;
    if (int e = fill_chan(cp, C, denorm, mean, std_, lo, hi)) return e;
    PgdSets s{};
    bool vec = (hw % 4) == 0;
    for (int i = 0; i < n_sets; ++i) {
        B2_REQUIRE(x[i] && g[i] && clean[i] && out[i], "pgd_update: null pointer in set %d", i);
        s.x[i] = x[i]; s.g[i] = g[i]; s.c[i] = clean[i]; s.o[i] = out[i];
        vec = vec && aligned16(x[i]) && aligned16(g[i]) && aligned16(clean[i]) && aligned16(out[i]);
    }
    int64_t per_set = (int64_t)n_img * C * hw;
    if (per_set == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (vec && hw / 4 < (int64_t)1 << 30 && (int64_t)n_sets * n_img * C <= 65535) {
        const int hw4 = (int)(hw / 4);
        const int planes = n_img * C;
        int gx = (hw4 + kPgdThreads * kPgdUnroll - 1) / (kPgdThreads * kPgdUnroll);
        const int cap = (kNumSMs * 16 + planes * n_sets - 1) / (planes * n_sets);     // keep ~16 blocks per SM overall
        if (gx > cap) gx = cap < 1 ? 1 : cap;
        pgd_update_vec4<<<dim3(gx, planes * n_sets), kPgdThreads, 0, st>>>(s, planes, C, hw4, alpha, eps, denorm, cp);
    } else {
        int grid = stream_grid(per_set * n_sets, kPgdThreads);
        pgd_update_scalar<<<grid, kPgdThreads, 0, st>>>(s, n_sets, per_set, hw, C, alpha, eps, denorm, cp);
    }
    return check_launch("pgd_update");
}

extern "C" int64_t b2_pgd_update_l2_workspace_bytes(int n_img) {
    return (int64_t)2 * n_img * kL2Blocks * sizeof(double);
}

extern "C" int b2_pgd_update_l2(const float* x, const float* g, const float* clean, float* out,
                                int n_img, int C, int64_t hw, float alpha, float eps, int denorm,
                                const float* mean, const float* std_, const float* lo,
                                const float* hi, void* workspace, void* stream) {
    B2_REQUIRE(C >= 1 && C <= 4, "pgd_update_l2: C must be 1..4 (got %d)", C);
    B2_REQUIRE(x && g && clean && out && workspace, "pgd_update_l2: null pointer");
    ChanParams cp;
    if (int e = fill_chan(cp, C, denorm, mean, std_, lo, hi)) return e;
    if (n_img == 0 || hw == 0) return 0;
    int64_t per_img = (int64_t)C * hw;
    double* part_g = (double*)workspace;
    double* part_e = part_g + (int64_t)n_img * kL2Blocks;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(kL2Blocks, n_img);
    l2_gnorm_partials<<<grid, kL2Threads, 0, st>>>(g, per_img, part_g);
    l2_step<<<grid, kL2Threads, 0, st>>>(x, g, clean, out, per_img, hw, C, alpha, denorm, cp, part_g, part_e);
    l2_project<<<grid, kL2Threads, 0, st>>>(clean, out, per_img, hw, C, eps, denorm, cp, part_e);
    return check_launch("pgd_update_l2");
}

extern "C" int b2_patch_apply(float* img, const float* patch, int n_img, int C, int H, int W,
                              const int* centers, int radius, void* stream) {
    B2_REQUIRE(img && patch && centers, "patch_apply: null pointer");
    B2_REQUIRE(n_img >= 1 && n_img <= 8, "patch_apply: n_img must be 1..8 (got %d)", n_img);
    B2_REQUIRE(radius >= 0, "patch_apply: negative radius");
    Centers ctr{};
    for (int i = 0; i < n_img; ++i) { ctr.cy[i] = centers[2 * i]; ctr.cx[i] = centers[2 * i + 1]; }
    int dim = 2 * radius + 1;
    int64_t total = (int64_t)n_img * C * dim * dim;
    int grid = stream_grid(total, 256);
    patch_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img, patch, n_img, C, H, W, ctr, radius, nullptr);
    return check_launch("patch_apply");
}

extern "C" int b2_patch_apply_dev(float* img, const float* patch, int n_img, int C, int H, int W,
                                  const int* centers_dev, int radius, void* stream) {
    B2_REQUIRE(img && patch && centers_dev, "patch_apply_dev: null pointer");
    B2_REQUIRE(n_img >= 1 && n_img <= 8, "patch_apply_dev: n_img must be 1..8 (got %d)", n_img);
    B2_REQUIRE(radius >= 0, "patch_apply_dev: negative radius");
    Centers ctr{};
    int dim = 2 * radius + 1;
    int64_t total = (int64_t)n_img * C * dim * dim;
    int grid = stream_grid(total, 256);
    patch_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img, patch, n_img, C, H, W, ctr, radius, centers_dev);
    return check_launch("patch_apply_dev");
}

extern "C" int b2_patch_update(float* patch, const float* gL, const float* gR, int C, int H, int W,
                               int cyL, int cxL, int cyR, int cxR, int radius, float alpha,
                               float eps, const float* lo, const float* hi, float* delta_out,
                               void* stream) {
    B2_REQUIRE(patch && gL && gR, "patch_update: null pointer");
    B2_REQUIRE(C >= 1 && C <= 4, "patch_update: C must be 1..4");
    B2_REQUIRE(cyL - radius >= 0 && cyL + radius < H && cxL - radius >= 0 && cxL + radius < W &&
               cyR - radius >= 0 && cyR + radius < H && cxR - radius >= 0 && cxR + radius < W,
               "patch_update: patch box leaves the %dx%d frame", H, W);
    ChanParams cp;
    if (int e = fill_chan(cp, C, 0, nullptr, nullptr, lo, hi)) return e;
    int dim = 2 * radius + 1;
    int total = C * dim * dim;
    // 0.5 * alpha is a Python double product in the reference, then cast to fp32
    float half_alpha = (float)(0.5 * (double)alpha);
    patch_update_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        patch, gL, gR, C, H, W, cyL, cxL, cyR, cxR, radius, half_alpha, eps, (lo && hi) ? 1 : 0, cp,
        delta_out, nullptr);
    return check_launch("patch_update");
}

extern "C" int b2_patch_update_dev(float* patch, const float* gL, const float* gR, int C, int H, int W,
                                   const int* centers_dev, int radius, float alpha, float eps, const float* lo,
                                   const float* hi, float* delta_out, void* stream) {
    B2_REQUIRE(patch && gL && gR && centers_dev, "patch_update_dev: null pointer");
    B2_REQUIRE(C >= 1 && C <= 4, "patch_update_dev: C must be 1..4");
    B2_REQUIRE(radius >= 0 && 2 * radius + 1 <= H && 2 * radius + 1 <= W, "patch_update_dev: patch larger than the frame");
    ChanParams cp;
    if (int e = fill_chan(cp, C, 0, nullptr, nullptr, lo, hi)) return e;
    int dim = 2 * radius + 1;
    int total = C * dim * dim;
    float half_alpha = (float)(0.5 * (double)alpha);
    patch_update_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        patch, gL, gR, C, H, W, 0, 0, 0, 0, radius, half_alpha, eps, (lo && hi) ? 1 : 0, cp, delta_out, centers_dev);
    return check_launch("patch_update_dev");
}

extern "C" int b2_patch_axpy(float* patch, const float* delta, int C, int dim, const float* lo, const float* hi,
                             void* stream) {
    B2_REQUIRE(patch && delta, "patch_axpy: null pointer");
    B2_REQUIRE(C >= 1 && C <= 4 && dim >= 1, "patch_axpy: C must be 1..4, dim >= 1");
    ChanParams cp;
    if (int e = fill_chan(cp, C, 0, nullptr, nullptr, lo, hi)) return e;
    const int total = C * dim * dim;
    patch_axpy_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(patch, delta, C, dim * dim, (lo && hi) ? 1 : 0, cp);
    return check_launch("patch_axpy");
}
