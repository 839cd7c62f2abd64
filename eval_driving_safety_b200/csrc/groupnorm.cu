// GroupNorm (+ residual) (+ ReLU) on channels-last 3-D volumes, forward and data
// gradient: the norm/activation that follows every 3-D conv of the hourglass
// stacks (upstream convbn_3d; reached from attack/DSGN/pgd_attack.py:308/:336).
//
//   y = act( (x - mean_g) * rstd_g * gamma_c + beta_c (+ res) )
//
// Roofline: HBM streaming.  fwd = stats pass (read x) + apply pass (read x [,res],
// write y); bwd = stats pass (read gy, x, y) + apply pass (read gy, x, y, write gx).
// All reductions are two-stage with a fixed order (per-block partials, then a
// serial fp64 combine) -> deterministic.  gamma/beta are frozen in an attack, so no
// parameter gradients are produced.
#include "common.cuh"

namespace b2 {

constexpr int kGnBlocks = kNumSMs * 2;   // partial blocks per sample
constexpr int kGnMaxC = 256;

struct GnLayout {
    int lpr;      // lanes (float4) per row = C/4
    int rpb;      // rows per block step
    int threads;  // lpr * rpb
};

// partial blocks per sample: at least ~256 rows each, at most two (512-thread blocks) per SM
static int gn_nblocks(int64_t S) {
    int64_t nb = (S + 255) / 256;
    if (nb < 1) nb = 1;
    if (nb > kGnBlocks) nb = kGnBlocks;
    return (int)nb;
}

static GnLayout gn_layout(int C) {
    GnLayout l;
    l.lpr = C / 4;
    l.rpb = 512 / l.lpr;
    if (l.rpb < 1) l.rpb = 1;
    l.threads = l.lpr * l.rpb;
    return l;
}

// ws layout (floats): partial[N][kGnBlocks][2][C]  then  coef[N][3][C]
__host__ __device__ inline int64_t gn_partial_floats(int N, int C) { return (int64_t)N * kGnBlocks * 2 * C; }

// MODE 0: sums of (x, x^2) per channel.  MODE 1: sums of (gz*x, gz) per channel,
// gz = gy * mask; relu == 1: mask = saved output y > 0; relu == 2: mask recomputed from x with the
// forward's own scale/shift (fcoef = stats tail, identical fmaf) so y need not be read (or kept).
// (the backward-sums variant holds up to 12 float4 loads in flight: 85 registers = ONE 512-thread block per SM; capped at
// 64 registers it runs two and spills 128 bytes -- measured slightly slower in situ, 7.80 vs 7.61 ms per two pair-iterations)
template <int MODE>
__global__ void gn_partials_kernel(const float4* __restrict__ a, const float4* __restrict__ x,
                                   const float4* __restrict__ y, float* __restrict__ partial,
                                   int C, int64_t S, int lpr, int rpb, int relu, const float* __restrict__ fstats, int G) {
    extern __shared__ float4 sh[];  // [2][rpb][lpr] (+ [2][C] floats of forward coefficients)
    const int n = blockIdx.y;
    const int lane = threadIdx.x % lpr, r = threadIdx.x / lpr;
    float4 fs = make_float4(0.f, 0.f, 0.f, 0.f), fb = fs;     // this lane's 4 channels: scale, shift
    if (MODE == 1 && relu == 2) {
        const float* fc = fstats + (int64_t)n * (2 * G + 2 * C) + 2 * G;      // forward scale / shift of sample n
        fs = make_float4(fc[lane * 4], fc[lane * 4 + 1], fc[lane * 4 + 2], fc[lane * 4 + 3]);
        fb = make_float4(fc[C + lane * 4], fc[C + lane * 4 + 1], fc[C + lane * 4 + 2], fc[C + lane * 4 + 3]);
    }
    const int64_t rows_per_block = (S + gridDim.x - 1) / gridDim.x;
    const int64_t s_begin = (int64_t)blockIdx.x * rows_per_block;
    const int64_t s_end = min(S, s_begin + rows_per_block);
    const int64_t base = (int64_t)n * S * lpr;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f), q = p;
    auto accum = [&](float4 g, float4 v, float4 o) {
        if (MODE == 0) {
            p.x += v.x; p.y += v.y; p.z += v.z; p.w += v.w;
            q.x += v.x * v.x; q.y += v.y * v.y; q.z += v.z * v.z; q.w += v.w * v.w;
        } else {
            if (relu == 2) { o.x = fmaf(v.x, fs.x, fb.x); o.y = fmaf(v.y, fs.y, fb.y); o.z = fmaf(v.z, fs.z, fb.z); o.w = fmaf(v.w, fs.w, fb.w); }
            if (relu) {
                g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f;
                g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f;
            }
            p.x += g.x * v.x; p.y += g.y * v.y; p.z += g.z * v.z; p.w += g.w * v.w;
            q.x += g.x; q.y += g.y; q.z += g.z; q.w += g.w;
        }
    };
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    int64_t s = s_begin + r;
    // 4 rows per trip: all loads are issued before the first use (memory-level parallelism)
    for (; s + 3 * (int64_t)rpb < s_end; s += 4 * (int64_t)rpb) {
        float4 vv[4], gg[4], oo[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            int64_t i = base + (s + (int64_t)u * rpb) * lpr + lane;
            vv[u] = ldg_stream(x + i);
            gg[u] = MODE == 1 ? ldg_stream(a + i) : zero4;
            oo[u] = (MODE == 1 && relu == 1) ? ldg_stream(y + i) : zero4;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) accum(gg[u], vv[u], oo[u]);
    }
    for (; s < s_end; s += rpb) {
        int64_t i = base + s * lpr + lane;
        float4 v = ldg_stream(x + i);
        float4 g = MODE == 1 ? ldg_stream(a + i) : zero4;
        float4 o = (MODE == 1 && relu == 1) ? ldg_stream(y + i) : zero4;
        accum(g, v, o);
    }
    sh[r * lpr + lane] = p;
    sh[(rpb + r) * lpr + lane] = q;
    __syncthreads();
    if (r == 0) {
        float4 sp = make_float4(0.f, 0.f, 0.f, 0.f), sq = sp;
        for (int k = 0; k < rpb; ++k) {
            float4 u = sh[k * lpr + lane], w = sh[(rpb + k) * lpr + lane];
            sp.x += u.x; sp.y += u.y; sp.z += u.z; sp.w += u.w;
            sq.x += w.x; sq.y += w.y; sq.z += w.z; sq.w += w.w;
        }
        float* dst = partial + (((int64_t)n * gridDim.x + blockIdx.x) * 2) * C;
        *reinterpret_cast<float4*>(dst + lane * 4) = sp;
        *reinterpret_cast<float4*>(dst + C + lane * 4) = sq;
    }
}

// Finalize: one block per (group, sample).  Thread t adds the partials k = t, t+128, ... of the
// group's channels (fp64), then a fixed-shape tree over the 128 threads combines them -> the
// summation order is a function of (nblocks, C, G) only, i.e. bitwise reproducible, and the
// serial tail of a single-block finalize (11 us measured) is gone.
constexpr int kGnFinThreads = 128;
constexpr int kGnMaxCpg = 8;

__device__ __forceinline__ void gn_group_sums(const float* __restrict__ partial, int n, int nblocks, int C, int c0,
                                              int cpg, const float* __restrict__ wgt /* per-channel weight or null */,
                                              double& S1, double& S2) {
    __shared__ double sh[2][kGnFinThreads];
    double a = 0.0, b = 0.0;
    for (int k = threadIdx.x; k < nblocks; k += kGnFinThreads) {
        const float* p = partial + (((int64_t)n * nblocks + k) * 2) * C + c0;
        for (int j = 0; j < cpg; ++j) {
            const double w = wgt ? (double)__ldg(wgt + c0 + j) : 1.0;
            a += w * (double)__ldg(p + j);
            b += w * (double)__ldg(p + C + j);
        }
    }
    sh[0][threadIdx.x] = a; sh[1][threadIdx.x] = b;
    __syncthreads();
    for (int off = kGnFinThreads / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
            sh[0][threadIdx.x] += sh[0][threadIdx.x + off];
            sh[1][threadIdx.x] += sh[1][threadIdx.x + off];
        }
        __syncthreads();
    }
    S1 = sh[0][0]; S2 = sh[1][0];
}

// fwd: stats[n] = [(mean, rstd) x G][scale[C]][shift[C]]; coef[n][0][c] = scale, coef[n][1][c] = shift.
__global__ void __launch_bounds__(kGnFinThreads)
gn_finalize_fwd(const float* __restrict__ partial, const float* __restrict__ gamma,
                const float* __restrict__ beta, float* __restrict__ stats,
                float* __restrict__ coef, int C, int64_t S, int G, float eps, int nblocks) {
    const int g = blockIdx.x, n = blockIdx.y, cpg = C / G;
    double sum, sq;
    gn_group_sums(partial, n, nblocks, C, g * cpg, cpg, nullptr, sum, sq);
    const double m = (double)cpg * (double)S;
    const double mean = sum / m;
    double var = sq / m - mean * mean;
    if (var < 0.0) var = 0.0;
    const double rstd = 1.0 / sqrt(var + (double)eps);
    float* st = stats + (int64_t)n * (2 * G + 2 * C);
    if (threadIdx.x == 0) { st[g * 2 + 0] = (float)mean; st[g * 2 + 1] = (float)rstd; }
    if (threadIdx.x < cpg) {
        const int c = g * cpg + threadIdx.x;
        const double scale = (double)gamma[c] * rstd;
        const float fsc = (float)scale, fsh = (float)((double)beta[c] - mean * scale);
        coef[((int64_t)n * 3 + 0) * C + c] = fsc;
        coef[((int64_t)n * 3 + 1) * C + c] = fsh;
        st[2 * G + c] = fsc;
        st[2 * G + C + c] = fsh;
    }
}

// bwd: coef[n][0][c] = gamma_c*rstd_g, coef[n][1][c] = c2_g, coef[n][2][c] = c3_g with
//   ds = sum_c gamma_c * sum(gz*x), db = sum_c gamma_c * sum(gz)
//   c2 = (db*mean - ds) * rstd^3 / m ; c3 = -c2*mean - db*rstd/m ; gx = coef0*gz + c2*x + c3
__global__ void __launch_bounds__(kGnFinThreads)
gn_finalize_bwd(const float* __restrict__ partial, const float* __restrict__ gamma,
                const float* __restrict__ stats, float* __restrict__ coef, int C,
                int64_t S, int G, int nblocks) {
    const int g = blockIdx.x, n = blockIdx.y, cpg = C / G;
    double ds, db;
    gn_group_sums(partial, n, nblocks, C, g * cpg, cpg, gamma, ds, db);
    const float* st = stats + (int64_t)n * (2 * G + 2 * C);
    const double mean = st[g * 2 + 0], rstd = st[g * 2 + 1];
    const double m = (double)cpg * (double)S;
    const double c2 = (db * mean - ds) * rstd * rstd * rstd / m;
    const double c3 = -c2 * mean - db * rstd / m;
    if (threadIdx.x < cpg) {
        const int c = g * cpg + threadIdx.x;
        coef[((int64_t)n * 3 + 0) * C + c] = (float)((double)gamma[c] * rstd);
        coef[((int64_t)n * 3 + 1) * C + c] = (float)c2;
        coef[((int64_t)n * 3 + 2) * C + c] = (float)c3;
    }
}

__global__ void __launch_bounds__(256)
gn_apply_fwd(const float4* __restrict__ x, const float4* __restrict__ res, const float* __restrict__ coef,
             float4* __restrict__ y, int C, int64_t S, int relu) {
    extern __shared__ float shc[];  // [2][C]
    const int n = blockIdx.y, lpr = C / 4;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) shc[i] = coef[(int64_t)n * 3 * C + i];
    __syncthreads();
    const int64_t total = S * lpr, base = (int64_t)n * total;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int q = (int)(i % lpr) * 4;
        float4 v = ldg_stream(x + base + i);
        float4 o;
        o.x = fmaf(v.x, shc[q], shc[C + q]);
        o.y = fmaf(v.y, shc[q + 1], shc[C + q + 1]);
        o.z = fmaf(v.z, shc[q + 2], shc[C + q + 2]);
        o.w = fmaf(v.w, shc[q + 3], shc[C + q + 3]);
        if (res) {
            float4 r = ldg_stream(res + base + i);
            o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        y[base + i] = o;
    }
}

__global__ void __launch_bounds__(256)
gn_apply_bwd(const float4* gy, const float4* __restrict__ x, const float4* __restrict__ y,
             const float* __restrict__ coef, float4* __restrict__ gx, float4* gres, int C,
             int64_t S, int relu, const float* __restrict__ fstats, int G) {
    extern __shared__ float shc[];  // [3][C] backward coefficients (+ [2][C] forward scale / shift)
    const int n = blockIdx.y, lpr = C / 4;
    for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) shc[i] = coef[(int64_t)n * 3 * C + i];
    if (relu == 2)
        for (int i = threadIdx.x; i < 2 * C; i += blockDim.x)
            shc[3 * C + i] = fstats[(int64_t)n * (2 * G + 2 * C) + 2 * G + i];
    __syncthreads();
    const int64_t total = S * lpr, base = (int64_t)n * total;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int q = (int)(i % lpr) * 4;
        float4 g = ldg_stream(gy + base + i);
        float4 v = ldg_stream(x + base + i);
        if (relu) {
            float4 o;
            if (relu == 2) {
                const float* fs = shc + 3 * C;
                o.x = fmaf(v.x, fs[q], fs[C + q]); o.y = fmaf(v.y, fs[q + 1], fs[C + q + 1]);
                o.z = fmaf(v.z, fs[q + 2], fs[C + q + 2]); o.w = fmaf(v.w, fs[q + 3], fs[C + q + 3]);
            } else {
                o = ldg_stream(y + base + i);
            }
            g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f;
            g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f;
        }
        float4 d;
        d.x = shc[q] * g.x + shc[C + q] * v.x + shc[2 * C + q];
        d.y = shc[q + 1] * g.y + shc[C + q + 1] * v.y + shc[2 * C + q + 1];
        d.z = shc[q + 2] * g.z + shc[C + q + 2] * v.z + shc[2 * C + q + 2];
        d.w = shc[q + 3] * g.w + shc[C + q + 3] * v.w + shc[2 * C + q + 3];
        if (gres) gres[base + i] = g;   // may alias gy: read above, same index
        gx[base + i] = d;
    }
}

}  // namespace b2

using namespace b2;

extern "C" int64_t b2_groupnorm_workspace_bytes(int N, int C) {
    return (gn_partial_floats(N, C) + (int64_t)N * 3 * C) * (int64_t)sizeof(float);
}

static int gn_check(const char* who, int N, int C, int64_t S, int G) {
    if (C % 4 != 0 || C > kGnMaxC || C < 4) { set_error("%s: C must be a multiple of 4 in [4,%d] (got %d)", who, kGnMaxC, C); return B2_ERR_UNSUPPORTED; }
    if (G < 1 || C % G != 0) { set_error("%s: G must divide C (C=%d, G=%d)", who, C, G); return B2_ERR_BAD_ARG; }
    if (C / G > kGnFinThreads) { set_error("%s: at most %d channels per group", who, kGnFinThreads); return B2_ERR_UNSUPPORTED; }
    if (N < 0 || S < 0) { set_error("%s: negative size", who); return B2_ERR_BAD_ARG; }
    return 0;
}

static int groupnorm_fwd_impl(const float* x, const float* res, const float* gamma, const float* beta,
                              float* y, float* stats, int N, int C, int64_t S, int G, float eps,
                              int relu, const float* ext_partial, int ext_rows, void* workspace, void* stream) {
    B2_REQUIRE(x && gamma && beta && y && stats && workspace, "groupnorm_fwd: null pointer");
    if (int e = gn_check("groupnorm_fwd", N, C, S, G)) return e;
    B2_REQUIRE(aligned16(x) && aligned16(y) && (!res || aligned16(res)), "groupnorm_fwd: pointers must be 16B aligned");
    B2_REQUIRE(!ext_partial || ext_rows >= 1, "groupnorm_fwd: external partial sums need >= 1 row per sample");
    if (N == 0 || S == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    GnLayout l = gn_layout(C);
    float* partial = (float*)workspace;
    float* coef = partial + gn_partial_floats(N, C);
    if (ext_partial) {
        // statistics pass already done by the producer (conv epilogue): table [N][ext_rows][2][C]
        gn_finalize_fwd<<<dim3(G, N), kGnFinThreads, 0, st>>>(ext_partial, gamma, beta, stats, coef, C, S, G, eps, ext_rows);
    } else {
        int nblocks = gn_nblocks(S);
        gn_partials_kernel<0><<<dim3(nblocks, N), l.threads, 2 * l.rpb * l.lpr * sizeof(float4), st>>>(
            nullptr, (const float4*)x, nullptr, partial, C, S, l.lpr, l.rpb, 0, nullptr, G);
        gn_finalize_fwd<<<dim3(G, N), kGnFinThreads, 0, st>>>(partial, gamma, beta, stats, coef, C, S, G, eps, nblocks);
    }
    int gx = stream_grid(S * l.lpr, 256 * 4, kNumSMs * 8);
    gn_apply_fwd<<<dim3(gx, N), 256, 2 * C * sizeof(float), st>>>((const float4*)x, (const float4*)res, coef,
                                                                  (float4*)y, C, S, relu);
    return check_launch("groupnorm_fwd");
}

extern "C" int b2_groupnorm_fwd(const float* x, const float* res, const float* gamma, const float* beta,
                                float* y, float* stats, int N, int C, int64_t S, int G, float eps,
                                int relu, void* workspace, void* stream) {
    return groupnorm_fwd_impl(x, res, gamma, beta, y, stats, N, C, S, G, eps, relu, nullptr, 0, workspace, stream);
}

extern "C" int b2_groupnorm_fwd_ext(const float* x, const float* res, const float* gamma, const float* beta,
                                    float* y, float* stats, int N, int C, int64_t S, int G, float eps,
                                    int relu, const float* ext_partial, int ext_rows, void* workspace,
                                    void* stream) {
    B2_REQUIRE(ext_partial, "groupnorm_fwd_ext: null partial-sum table");
    return groupnorm_fwd_impl(x, res, gamma, beta, y, stats, N, C, S, G, eps, relu, ext_partial, ext_rows, workspace, stream);
}

static int groupnorm_bwd_impl(const float* gy, const float* x, const float* y, const float* gamma,
                              const float* stats, float* gx, float* gres, int N, int C, int64_t S,
                              int G, int relu, const float* ext_partial, int ext_rows, void* workspace, void* stream) {
    B2_REQUIRE(gy && x && gamma && stats && gx && workspace, "groupnorm_bwd: null pointer");
    B2_REQUIRE(relu != 1 || y, "groupnorm_bwd: relu == 1 needs the saved output y (relu == 2 recomputes the mask)");
    B2_REQUIRE(relu >= 0 && relu <= 2, "groupnorm_bwd: relu must be 0, 1 or 2");
    if (int e = gn_check("groupnorm_bwd", N, C, S, G)) return e;
    B2_REQUIRE(aligned16(gy) && aligned16(x) && aligned16(gx), "groupnorm_bwd: pointers must be 16B aligned");
    B2_REQUIRE(!ext_partial || (ext_rows >= 1 && relu != 1),
               "groupnorm_bwd: external partial sums need >= 1 row per sample and relu in {0, 2}");
    if (N == 0 || S == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    GnLayout l = gn_layout(C);
    float* partial = (float*)workspace;
    float* coef = partial + gn_partial_floats(N, C);
    if (ext_partial) {
        // (sum gz*x, sum gz) per channel already added up by the kernel that produced gy
        gn_finalize_bwd<<<dim3(G, N), kGnFinThreads, 0, st>>>(ext_partial, gamma, stats, coef, C, S, G, ext_rows);
    } else {
        int nblocks = gn_nblocks(S);
        gn_partials_kernel<1><<<dim3(nblocks, N), l.threads, 2 * l.rpb * l.lpr * sizeof(float4), st>>>(
            (const float4*)gy, (const float4*)x, (const float4*)y, partial, C, S, l.lpr, l.rpb, relu,
            stats, G);
        gn_finalize_bwd<<<dim3(G, N), kGnFinThreads, 0, st>>>(partial, gamma, stats, coef, C, S, G, nblocks);
    }
    int gxd = stream_grid(S * l.lpr, 256 * 4, kNumSMs * 8);
    gn_apply_bwd<<<dim3(gxd, N), 256, 5 * C * sizeof(float), st>>>((const float4*)gy, (const float4*)x,
                                                                   (const float4*)y, coef, (float4*)gx,
                                                                   (float4*)gres, C, S, relu, stats, G);
    return check_launch("groupnorm_bwd");
}

extern "C" int b2_groupnorm_bwd(const float* gy, const float* x, const float* y, const float* gamma,
                                const float* stats, float* gx, float* gres, int N, int C, int64_t S,
                                int G, int relu, void* workspace, void* stream) {
    return groupnorm_bwd_impl(gy, x, y, gamma, stats, gx, gres, N, C, S, G, relu, nullptr, 0, workspace, stream);
}

extern "C" int b2_groupnorm_bwd_ext(const float* gy, const float* x, const float* y, const float* gamma,
                                    const float* stats, float* gx, float* gres, int N, int C, int64_t S,
                                    int G, int relu, const float* ext_partial, int ext_rows, void* workspace,
                                    void* stream) {
    B2_REQUIRE(ext_partial, "groupnorm_bwd_ext: null partial-sum table");
    return groupnorm_bwd_impl(gy, x, y, gamma, stats, gx, gres, N, C, S, G, relu, ext_partial, ext_rows, workspace, stream);
}
