// 2-D convolutions of the feature extractor and the BEV head (the layers of upstream
// dsgn.models feature_extraction / bev_conv / bbox_* reached from attack/DSGN/pgd_attack.py:308
// forward and :336 backward) as an implicit GEMM on the 5th-gen tensor cores, with the
// error-compensated 3xTF32 operand split done INSIDE the kernel.
//
//   GEMM view : M = output pixels (tile = 16 h x 8 w = 128 rows), N = Cout (one tile, <= 256),
//               K = taps x Cin.  One tap x 32-channel chunk per pipeline stage.
//   A operand : {32 ch, 8 w, 16 h, 1 n} TMA box of the channels-last input, shifted by the tap
//               offset (dilation 1|2); out-of-range = zero fill = the conv padding.  Stride-2
//               convs use a tensor map with traversal stride 2; transposed stride-2 convs (= the
//               data gradient of a stride-2 conv) are split into 4 output-parity classes.
//   3xTF32    : the reference computes these layers in fp32.  kind::tf32 reads only the upper
//               19 bits of an operand, so a plain TF32 product carries ~2^-11 relative error per
//               factor -- measured to flip ~1 % of the attack's gradient signs through the 56-conv
//               extractor.  With split = 1 four extra warps rewrite every landed A tile as
//               hi = x & ~0x1fff (exactly representable) and lo = x - hi (exact in fp32) into a
//               second smem tile, the weights arrive pre-split (w_hi, w_lo), and the MMA warp
//               issues  [acc0 | acc1] += hi * [w_hi ; w_lo]  (one MMA of N = 2 Cout: the two weight tiles
//               sit back to back in smem, so x_hi is read once for both products -- the stage is
//               smem-bandwidth bound) and  acc0 += lo * w_hi;  the epilogue adds acc0 + acc1.
//               Products are exact, the dropped lo*w_lo term is ~2^-22, accumulation is fp32.
//               No extra HBM traffic or launches.  (split = 2 additionally rewrites the A tile as its
//               truncated value in place -- bit-identical results, i.e. the tensor core truncates; test
//               test_conv2d_tensor_core_truncates_tf32_operands.)
//   Roles     : warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue (tcgen05.ld -> +bias
//               +addend -> st.global rows), warps 6-9 operand split.  Two TMEM accumulators.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "tcgen05.cuh"

namespace b2 {

constexpr int kC2Threads = 320;
constexpr int kC2TileH = 16, kC2TileW = 8;
constexpr int kC2K = 32;                       // floats per K chunk = one 128B swizzle row
constexpr int kC2ABytes = 128 * 128;
constexpr int kC2MaxStages = 8;

struct C2Params {
    int N, Cin, Cout;
    int nt, n_tiles;           // N tile width (<= 256, % 16 == 0), number of N tiles (Cout = nt * n_tiles)
    int Hi, Wi, Ho, Wo;
    int Ht, Wt;                // tile-grid extents (conv: output dims; transposed: input dims)
    int ks, dil;
    int mode;                  // 0 conv stride 1, 1 conv stride 2, 2 transposed stride 2
    int split;                 // 3xTF32
    int stack;                 // split: x_hi meets [w_hi ; w_lo] in ONE MMA of N = 2 nt (two accumulator halves)
    int rewrite_hi;            // split: also rewrite the landed A tile as its TF32-truncated value (not needed if the
                               // tensor core ignores the low 13 mantissa bits itself; kept for A/B verification)
    int tiles_w, tiles_h;
    int kchunks, stages, stage_bytes, b_bytes, tmem_cols;
    int out_cstride;           // floats between consecutive output pixels (>= Cout)
    long long total_tiles;
};

struct C2Tile { int nti, cls, n, h0, w0; };

__device__ __forceinline__ C2Tile c2_decode(const C2Params& p, long long t) {
    C2Tile c;
    c.w0 = (int)(t % p.tiles_w) * kC2TileW; t /= p.tiles_w;
    c.h0 = (int)(t % p.tiles_h) * kC2TileH; t /= p.tiles_h;
    c.n = (int)(t % p.N); t /= p.N;
    const int ncls = p.mode == 2 ? 4 : 1;
    c.cls = (int)(t % ncls); t /= ncls;
    c.nti = (int)t;
    return c;
}
__device__ __forceinline__ int c2_taps(const C2Params& p, int cls) {
    if (p.mode != 2) return p.ks * p.ks;
    if (p.ks == 1) return cls == 0 ? 1 : 0;
    return (1 + (cls & 1)) * (1 + ((cls >> 1) & 1));
}
// transposed conv, one axis: output parity par, j-th contributing tap -> (kernel index, input shift)
__device__ __forceinline__ void c2_deconv_axis(int par, int j, int& k, int& shift) {
    if (par == 0) { k = 1; shift = 0; }
    else if (j == 0) { k = 0; shift = 1; }
    else { k = 2; shift = 0; }
}
// tap_i-th tap of a tile: A-box origin (ah, aw) and weight tap index kh*ks+kw
__device__ __forceinline__ void c2_tap(const C2Params& p, const C2Tile& tc, int tap_i, int& ah, int& aw, int& tap) {
    if (p.mode == 2) {
        if (p.ks == 1) { ah = tc.h0; aw = tc.w0; tap = 0; return; }
        const int pw = tc.cls & 1, ph = (tc.cls >> 1) & 1;
        const int nw = 1 + pw;
        int kw, kh, sw, sh;
        c2_deconv_axis(pw, tap_i % nw, kw, sw);
        c2_deconv_axis(ph, tap_i / nw, kh, sh);
        ah = tc.h0 + sh; aw = tc.w0 + sw; tap = kh * 3 + kw;
        return;
    }
    const int kh = tap_i / p.ks, kw = tap_i % p.ks, half = p.ks / 2;
    const int s = p.mode == 1 ? 2 : 1;
    ah = s * tc.h0 + (kh - half) * p.dil;
    aw = s * tc.w0 + (kw - half) * p.dil;
    tap = tap_i;
}

// ---------------------------------------------------------------- epilogue: bias, addend, GroupNorm sums, store
// What rides along while an accumulator tile is drained (all optional):
//   bias / addend : out = conv + bias + addend (addend = the other consumers' gradient of a forked tensor)
//   stat_mode 1   : per-channel (sum, sum of squares) of the OUTPUT = the statistics pass of the GroupNorm that follows
//   stat_mode 2/3 : the launch writes the gradient gy w.r.t. the output of a GroupNorm (+ReLU) whose input was gn_x:
//                   per-channel (sum gz*x, sum gz), gz = gy (2) or gy * [fma(x, scale, shift) > 0] (3: the ReLU mask
//                   recomputed with the forward's own scale / shift) = the statistics pass of that norm's backward.
// Sums are kept per CTA and per sample in shared memory (one row per epilogue warp) and written as row blockIdx.x of
// the table [N][gridDim.x][2][Cout] when the CTA's tile sequence leaves a (channel tile, sample) -- fixed order, no
// atomics: bitwise reproducible for a given grid.
constexpr int kC2StatMaxNt = 128;

struct C2Epi {
    const float* bias;
    const float* addend;
    int stat_mode;
    float* stat_partial;
    const float* gn_x;
    const float* gn_coef;        // sample 0: scale[Cout], shift[Cout]; sample n at + n * gn_coef_stride
    int gn_coef_stride;
};

// v[0..15] = s, v[16..31] = q of this lane's pixel for 16 channels; on return v[0] of lane l is the warp total of
// s[l] (l < 16) or q[l - 16] (l >= 16).  Recursive halving, 31 shuffles, fixed order.
__device__ __forceinline__ void c2_transpose_sum32(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = upper ? v[i] : v[i + off];
            const float keep = upper ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
}

// grp = epilogue group (0: warps 2-5; 1: warps 6-9, which drain every other tile when no operand split keeps them
// busy); each group owns table row blockIdx.x * ngroups + grp and named barrier 1 + grp.
__device__ __forceinline__ void c2_stat_flush(const C2Params& p, const C2Epi& e, int key, float (*tot)[2][kC2StatMaxNt],
                                              int lane_grp, int lane, int grp, int ngroups) {
    asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");   // the group's four warps walk the same tile sequence
    if (lane_grp == 0) {
        const int n = key % p.N, nti = key / p.N;
        float* row = e.stat_partial + (((long long)n * gridDim.x + blockIdx.x) * ngroups + grp) * 2 * p.Cout + nti * p.nt;
        for (int c = lane; c < p.nt; c += 32) {
            row[c] = ((tot[0][0][c] + tot[1][0][c]) + tot[2][0][c]) + tot[3][0][c];
            row[p.Cout + c] = ((tot[0][1][c] + tot[1][1][c]) + tot[2][1][c]) + tot[3][1][c];
        }
    }
    asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
    for (int c = lane; c < 2 * kC2StatMaxNt; c += 32) (&tot[lane_grp][0][0])[c] = 0.f;
    __syncwarp();
}

// One accumulator tile: TMEM -> registers (+ second half of a stacked accumulator) -> +bias +addend -> sums -> st.global.
// off = element offset of this lane's output pixel (channel tile included); ok = the pixel exists.
__device__ __forceinline__ void c2_epilogue_tile(const C2Params& p, const C2Epi& e, const C2Tile& tc, float* __restrict__ out,
                                                 uint32_t taddr, bool active, bool ok, long long off, int lane,
                                                 float (*mytot)[kC2StatMaxNt]) {
    float* optr = out + off;
    const float* aptr = e.addend ? e.addend + off : nullptr;
    const float* xptr = e.stat_mode >= 2 ? e.gn_x + off : nullptr;
    const float* bptr = e.bias ? e.bias + tc.nti * p.nt : nullptr;
    const float* cptr = e.stat_mode == 3 ? e.gn_coef + (long long)tc.n * e.gn_coef_stride + tc.nti * p.nt : nullptr;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c0 = 0; c0 < p.nt; c0 += 16) {
        float4 av[4], xv[4];
        // operand rows requested before the accumulator is waited for
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            av[q] = (aptr && ok) ? __ldg(reinterpret_cast<const float4*>(aptr + c0) + q) : z4;
            xv[q] = (xptr && ok) ? __ldg(reinterpret_cast<const float4*>(xptr + c0) + q) : z4;
        }
        uint32_t r[16];
        if (active) {
            tmem_ld16(taddr + c0, r);
            if (p.stack) {                                    // + the x_hi*w_lo half of the accumulator
                uint32_t r2[16];
                tmem_ld16(taddr + p.nt + c0, r2);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
            } else {
                tmem_ld_wait();
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) r[j] = 0u;
        }
        float v[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float4 t = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                                   __uint_as_float(r[4 * q + 3]));
            if (bptr) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(bptr + c0) + q);
                t.x += b.x; t.y += b.y; t.z += b.z; t.w += b.w;
            }
            if (aptr) { t.x += av[q].x; t.y += av[q].y; t.z += av[q].z; t.w += av[q].w; }
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
            if (ok) *reinterpret_cast<float4*>(optr + c0 + 4 * q) = t;
        }
        if (e.stat_mode) {
            float sq[32];
            if (e.stat_mode == 1) {
#pragma unroll
                for (int j = 0; j < 16; ++j) { const float u = ok ? v[j] : 0.f; sq[j] = u; sq[16 + j] = u * u; }
            } else {
                const float* x = reinterpret_cast<const float*>(xv);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float g = ok ? v[j] : 0.f;
                    if (cptr && !(fmaf(x[j], __ldg(cptr + c0 + j), __ldg(cptr + p.Cout + c0 + j)) > 0.f)) g = 0.f;
                    sq[j] = g * x[j]; sq[16 + j] = g;
                }
            }
            c2_transpose_sum32(sq, lane);
            mytot[lane >> 4][c0 + (lane & 15)] += sq[0];
        }
    }
}

__global__ void __launch_bounds__(kC2Threads, 1)
conv2d_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      float* __restrict__ out, const C2Epi epi, const C2Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[3 * kC2MaxStages + 4];
    __shared__ uint32_t tmem_base_slot;
    __shared__ float stat_tot[2][4][2][kC2StatMaxNt];      // [epilogue group][warp][sum | second sum][channel]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* const smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };                       // TMA landed
    auto empty_bar = [&](int s) { return bar0 + 8u * (kC2MaxStages + s); };      // MMAs retired
    auto ready_bar = [&](int s) { return bar0 + 8u * (2 * kC2MaxStages + s); };  // operand split done
    auto tfull_bar = [&](int a) { return bar0 + 8u * (3 * kC2MaxStages + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (3 * kC2MaxStages + 2 + a); };
    const uint32_t a_region = p.split ? 2u * kC2ABytes : (uint32_t)kC2ABytes;
    const int taps_total = p.ks * p.ks;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); mbar_init(ready_bar(s), 128);
        }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&tmem_base_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            const uint32_t tx = (uint32_t)kC2ABytes + (uint32_t)(p.split ? 2 : 1) * (uint32_t)p.b_bytes;
            for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                const C2Tile tc = c2_decode(p, t);
                const int ntaps = c2_taps(p, tc.cls);
                for (int tap_i = 0; tap_i < ntaps; ++tap_i) {
                    int ah, aw, tap;
                    c2_tap(p, tc, tap_i, ah, aw, tap);
                    for (int kc = 0; kc < p.kchunks; ++kc) {
                        mbar_wait(empty_bar(stage), phase ^ 1);
                        mbar_expect_tx(full_bar(stage), tx);
                        const uint32_t sa = smem_base + (uint32_t)stage * p.stage_bytes;
                        tma_load_4d(sa, &map_a, full_bar(stage), kc * kC2K, aw, ah, tc.n);
                        tma_load_2d(sa + a_region, &map_b, full_bar(stage), kc * kC2K, tap * p.Cout + tc.nti * p.nt);
                        if (p.split)
                            tma_load_2d(sa + a_region + (uint32_t)p.b_bytes, &map_b, full_bar(stage), kc * kC2K,
                                        (taps_total + tap) * p.Cout + tc.nti * p.nt);
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            // split: accumulator = 2 nt columns [x_hi*w_hi + x_lo*w_hi | x_hi*w_lo]: the weight tiles w_hi, w_lo sit
            // back to back in smem, so ONE MMA of N = 2 nt multiplies x_hi with both (A is read from smem once
            // instead of twice -- the stage is smem-bandwidth bound), a second MMA of N = nt adds x_lo*w_hi
            const uint32_t idesc = umma_idesc_tf32(128, p.nt), idesc2 = umma_idesc_tf32(128, 2 * p.nt);
            const int acc_cols = p.stack ? 2 * p.nt : p.nt;
            int stage = 0; uint32_t phase = 0;
            long long it = 0;
            for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                const C2Tile tc = c2_decode(p, t);
                const int nsteps = c2_taps(p, tc.cls) * p.kchunks;
                if (nsteps == 0) continue;                        // all-zero parity class: the epilogue writes it
                const int acc = (int)(it & 1);
                mbar_wait(tempty_bar(acc), (uint32_t)(((it >> 1) & 1) ^ 1));
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * acc_cols);
                for (int s = 0; s < nsteps; ++s) {
                    mbar_wait(p.split ? ready_bar(stage) : full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + (uint32_t)stage * p.stage_bytes;
                    const uint64_t a_hi = umma_desc_sw128(sa), b_hi = umma_desc_sw128(sa + a_region);
                    if (p.split && p.stack) {
                        const uint64_t a_lo = umma_desc_sw128(sa + kC2ABytes);
#pragma unroll
                        for (int k = 0; k < kC2K / 8; ++k) umma_tf32(tmem_d, a_hi + 2 * k, b_hi + 2 * k, idesc2, (s | k) != 0);
#pragma unroll
                        for (int k = 0; k < kC2K / 8; ++k) umma_tf32(tmem_d, a_lo + 2 * k, b_hi + 2 * k, idesc, 1u);
                    } else if (p.split) {                         // one accumulator, small terms first
                        const uint64_t a_lo = umma_desc_sw128(sa + kC2ABytes),
                                       b_lo = umma_desc_sw128(sa + a_region + (uint32_t)p.b_bytes);
#pragma unroll
                        for (int k = 0; k < kC2K / 8; ++k) umma_tf32(tmem_d, a_lo + 2 * k, b_hi + 2 * k, idesc, (s | k) != 0);
#pragma unroll
                        for (int k = 0; k < kC2K / 8; ++k) umma_tf32(tmem_d, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
#pragma unroll
                        for (int k = 0; k < kC2K / 8; ++k) umma_tf32(tmem_d, a_hi + 2 * k, b_hi + 2 * k, idesc, 1u);
                    } else {
#pragma unroll
                        for (int k = 0; k < kC2K / 8; ++k) umma_tf32(tmem_d, a_hi + 2 * k, b_hi + 2 * k, idesc, (s | k) != 0);
                    }
                    umma_commit(empty_bar(stage));
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                umma_commit(tfull_bar(acc));
                ++it;
            }
        }
        __syncwarp();
    } else if (warp < 6 || !p.split) {
        // ===================== epilogue (one output pixel row per thread) =====================
        // warps 2-5; without the operand split warps 6-9 are a second group: group g drains accumulator g
        const int grp = warp >= 6 ? 1 : 0, ngroups = p.split ? 1 : 2;
        const int lane_grp = warp & 3;                            // TMEM lanes [32*lane_grp, +32)
        const int m = lane_grp * 32 + lane;
        const int hl = m / kC2TileW, wl = m % kC2TileW;
        const bool stats = epi.stat_mode != 0;
        float (*tot)[2][kC2StatMaxNt] = stat_tot[grp];
        if (stats) {
            for (int c = lane; c < 2 * kC2StatMaxNt; c += 32) (&tot[lane_grp][0][0])[c] = 0.f;
            __syncwarp();
        }
        int cur_key = 0;
        long long it = 0, ti = 0;
        for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++ti) {
            const C2Tile tc = c2_decode(p, t);
            if (stats)
                for (const int key = tc.nti * p.N + tc.n; cur_key < key; ++cur_key)
                    c2_stat_flush(p, epi, cur_key, tot, lane_grp, lane, grp, ngroups);
            const bool active = c2_taps(p, tc.cls) != 0;
            const int acc = (int)(it & 1);
            const bool mine = ngroups == 1 || (int)((active ? it : ti) & 1) == grp;
            if (mine) {
                const int h = tc.h0 + hl, w = tc.w0 + wl;
                const bool ok = h < p.Ht && w < p.Wt;
                int oh = h, ow = w;
                if (p.mode == 2) { oh = 2 * h + ((tc.cls >> 1) & 1); ow = 2 * w + (tc.cls & 1); }
                const long long off = (((long long)tc.n * p.Ho + oh) * p.Wo + ow) * p.out_cstride + tc.nti * p.nt;
                if (active) {
                    mbar_wait(tfull_bar(acc), (uint32_t)((it >> 1) & 1));
                    tc_fence_after();
                }
                const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(acc * (p.stack ? 2 * p.nt : p.nt));
                c2_epilogue_tile(p, epi, tc, out, taddr, active, ok, off, lane, tot[lane_grp]);
                if (active) {
                    tc_fence_before();
                    mbar_arrive(tempty_bar(acc));
                }
            }
            if (active) ++it;
        }
        if (stats)
            for (; cur_key < p.n_tiles * p.N; ++cur_key) c2_stat_flush(p, epi, cur_key, tot, lane_grp, lane, grp, ngroups);
    } else {
        // ===================== operand split (4 warps): A tile -> (hi in place, lo) =====================
        const int tid = threadIdx.x - 192;
        int stage = 0; uint32_t phase = 0;
        for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            const C2Tile tc = c2_decode(p, t);
            const int nsteps = c2_taps(p, tc.cls) * p.kchunks;
            for (int s = 0; s < nsteps; ++s) {
                mbar_wait(full_bar(stage), phase);
                float4* A = reinterpret_cast<float4*>(smem_gen + (size_t)stage * p.stage_bytes);
                float4* L = A + kC2ABytes / 16;
#pragma unroll
                for (int i = 0; i < kC2ABytes / 16 / 128; ++i) {
                    const int idx = i * 128 + tid;                 // elementwise: any mapping works, this one is conflict-free
                    const float4 v = A[idx];
                    float4 hi, lo;
                    hi.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); lo.x = v.x - hi.x;
                    hi.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); lo.y = v.y - hi.y;
                    hi.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); lo.z = v.z - hi.z;
                    hi.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); lo.w = v.w - hi.w;
                    if (p.rewrite_hi) A[idx] = hi;
                    L[idx] = lo;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor-core reads
                mbar_arrive(ready_bar(stage));
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

// =============================================================================================
// Stride-1 3x3 fast path ("halo" kernel): the bulk of the extractor.  The kernel above reloads -- and, with the
// split, re-splits -- the 16 KB A box for every tap.  Here, per (kw, 32-channel chunk) "generation", ONE
// {32 ch, 8 w, 16 + 2 dil rows} halo box is loaded and split once; the three kh taps read that smem tile at row
// offsets 0 / dil / 2 dil (= whole 1024 B swizzle atoms, so only the descriptor's start address moves).  The
// weight tiles (w_hi, w_lo pairs) stream through their own ring.  Per tap the smem traffic of the A side drops
// from 16 KB (TMA) + 32 KB (split) to a third of 18 KB + 36 KB; the kernel is smem-bandwidth bound, so that is
// what buys time.
// =============================================================================================
constexpr int kC2HaloMaxNA = 4;            // A ring slots: 2 with the split (hi + lo tiles, 40 KB a slot), 4 without --
                                           // a plain-TF32 generation retires in ~0.5 us, less than one TMA round trip
constexpr int kC2HaloMaxNW = 12;           // weight ring slots

struct C2HaloParams {
    C2Params c;
    int a_rows, a_bytes, a_slot, w_slot, nw, na;
};

__global__ void __launch_bounds__(kC2Threads, 1)
conv2d_halo_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                           float* __restrict__ out, const C2Epi epi, const C2HaloParams hp) {
    const C2Params& p = hp.c;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[3 * kC2HaloMaxNA + 2 * kC2HaloMaxNW + 4];
    __shared__ uint32_t tmem_base_slot;
    __shared__ float stat_tot[2][4][2][kC2StatMaxNt];      // [epilogue group][warp][sum | second sum][channel]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* const smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t a_base = smem_base, w_base = smem_base + (uint32_t)(hp.na * hp.a_slot);
    const uint32_t bar0 = smem_u32(bars);
    auto fullA = [&](int s) { return bar0 + 8u * s; };
    auto emptyA = [&](int s) { return bar0 + 8u * (kC2HaloMaxNA + s); };
    auto readyA = [&](int s) { return bar0 + 8u * (2 * kC2HaloMaxNA + s); };
    auto fullW = [&](int s) { return bar0 + 8u * (3 * kC2HaloMaxNA + s); };
    auto emptyW = [&](int s) { return bar0 + 8u * (3 * kC2HaloMaxNA + kC2HaloMaxNW + s); };
    auto tfull_bar = [&](int a) { return bar0 + 8u * (3 * kC2HaloMaxNA + 2 * kC2HaloMaxNW + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (3 * kC2HaloMaxNA + 2 * kC2HaloMaxNW + 2 + a); };

    if (threadIdx.x == 0) {
        for (int s = 0; s < hp.na; ++s) { mbar_init(fullA(s), 1); mbar_init(emptyA(s), 1); mbar_init(readyA(s), 128); }
        for (int s = 0; s < hp.nw; ++s) { mbar_init(fullW(s), 1); mbar_init(emptyW(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&tmem_base_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    const int ngen = 3 * p.kchunks;            // (kw, chunk) generations per tile

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t aslot = 0, aphase = 0, wslot = 0, wphase = 0;
            const uint32_t wtx = (uint32_t)hp.w_slot;
            for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                const C2Tile tc = c2_decode(p, t);
                for (int g = 0; g < ngen; ++g) {
                    const int kw = g / p.kchunks, kc = g % p.kchunks;
                    mbar_wait(emptyA(aslot), aphase ^ 1);
                    mbar_expect_tx(fullA(aslot), (uint32_t)hp.a_bytes);
                    tma_load_4d(a_base + aslot * (uint32_t)hp.a_slot, &map_a, fullA(aslot), kc * kC2K,
                                tc.w0 + (kw - 1) * p.dil, tc.h0 - p.dil, tc.n);
                    if (++aslot == (uint32_t)hp.na) { aslot = 0; aphase ^= 1; }
                    for (int kh = 0; kh < 3; ++kh) {
                        const int tap = kh * 3 + kw;
                        mbar_wait(emptyW(wslot), wphase ^ 1);
                        mbar_expect_tx(fullW(wslot), wtx);
                        const uint32_t wa = w_base + wslot * (uint32_t)hp.w_slot;
                        tma_load_2d(wa, &map_b, fullW(wslot), kc * kC2K, tap * p.Cout + tc.nti * p.nt);
                        if (p.split)
                            tma_load_2d(wa + (uint32_t)p.b_bytes, &map_b, fullW(wslot), kc * kC2K,
                                        (9 + tap) * p.Cout + tc.nti * p.nt);
                        if (++wslot == (uint32_t)hp.nw) { wslot = 0; wphase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(128, p.nt), idesc2 = umma_idesc_tf32(128, 2 * p.nt);
            const int acc_cols = p.stack ? 2 * p.nt : p.nt;
            uint32_t aslot = 0, aphase = 0, wslot = 0, wphase = 0;
            long long it = 0;
            for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
                const int acc = (int)(it & 1);
                mbar_wait(tempty_bar(acc), (uint32_t)(((it >> 1) & 1) ^ 1));
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * acc_cols);
                for (int g = 0; g < ngen; ++g) {
                    mbar_wait(p.split ? readyA(aslot) : fullA(aslot), aphase);
                    tc_fence_after();
                    const uint32_t sa = a_base + aslot * (uint32_t)hp.a_slot;
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) {
                        mbar_wait(fullW(wslot), wphase);
                        tc_fence_after();
                        const uint32_t off = (uint32_t)(kh * p.dil) * 1024u;          // kh * dil rows of 8 pixels
                        const uint64_t a_hi = umma_desc_sw128(sa + off);
                        const uint64_t b_hi = umma_desc_sw128(w_base + wslot * (uint32_t)hp.w_slot);
                        const uint32_t first = (uint32_t)(g | kh);
                        if (p.split && p.stack) {
                            const uint64_t a_lo = umma_desc_sw128(sa + (uint32_t)hp.a_bytes + off);
#pragma unroll
                            for (int k = 0; k < kC2K / 8; ++k) umma_tf32(tmem_d, a_hi + 2 * k, b_hi + 2 * k, idesc2, (first | (uint32_t)k) != 0);
#pragma unroll
                            for (int k = 0; k < kC2K / 8; ++k) umma_tf32(tmem_d, a_lo + 2 * k, b_hi + 2 * k, idesc, 1u);
                        } else if (p.split) {
                            const uint64_t a_lo = umma_desc_sw128(sa + (uint32_t)hp.a_bytes + off),
                                           b_lo = umma_desc_sw128(w_base + wslot * (uint32_t)hp.w_slot + (uint32_t)p.b_bytes);
#pragma unroll
                            for (int k = 0; k < kC2K / 8; ++k) umma_tf32(tmem_d, a_lo + 2 * k, b_hi + 2 * k, idesc, (first | (uint32_t)k) != 0);
#pragma unroll
                            for (int k = 0; k < kC2K / 8; ++k) umma_tf32(tmem_d, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
#pragma unroll
                            for (int k = 0; k < kC2K / 8; ++k) umma_tf32(tmem_d, a_hi + 2 * k, b_hi + 2 * k, idesc, 1u);
                        } else {
#pragma unroll
                            for (int k = 0; k < kC2K / 8; ++k) umma_tf32(tmem_d, a_hi + 2 * k, b_hi + 2 * k, idesc, (first | (uint32_t)k) != 0);
                        }
                        umma_commit(emptyW(wslot));
                        if (++wslot == (uint32_t)hp.nw) { wslot = 0; wphase ^= 1; }
                    }
                    umma_commit(emptyA(aslot));
                    if (++aslot == (uint32_t)hp.na) { aslot = 0; aphase ^= 1; }
                }
                umma_commit(tfull_bar(acc));
            }
        }
        __syncwarp();
    } else if (warp < 6 || !p.split) {
        // ===================== epilogue =====================
        // warps 2-5; without the operand split warps 6-9 are a second group: group g drains accumulator g
        const int grp = warp >= 6 ? 1 : 0, ngroups = p.split ? 1 : 2;
        const int lane_grp = warp & 3;
        const int m = lane_grp * 32 + lane;
        const int hl = m / kC2TileW, wl = m % kC2TileW;
        const bool stats = epi.stat_mode != 0;
        float (*tot)[2][kC2StatMaxNt] = stat_tot[grp];
        if (stats) {
            for (int c = lane; c < 2 * kC2StatMaxNt; c += 32) (&tot[lane_grp][0][0])[c] = 0.f;
            __syncwarp();
        }
        int cur_key = 0;
        long long it = 0;
        for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
            const C2Tile tc = c2_decode(p, t);
            if (stats)
                for (const int key = tc.nti * p.N + tc.n; cur_key < key; ++cur_key)
                    c2_stat_flush(p, epi, cur_key, tot, lane_grp, lane, grp, ngroups);
            const int acc = (int)(it & 1);
            if (ngroups == 2 && acc != grp) continue;
            const int h = tc.h0 + hl, w = tc.w0 + wl;
            const bool ok = h < p.Ht && w < p.Wt;
            const long long off = (((long long)tc.n * p.Ho + h) * p.Wo + w) * p.out_cstride + tc.nti * p.nt;
            mbar_wait(tfull_bar(acc), (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(acc * (p.stack ? 2 * p.nt : p.nt));
            c2_epilogue_tile(p, epi, tc, out, taddr, true, ok, off, lane, tot[lane_grp]);
            tc_fence_before();
            mbar_arrive(tempty_bar(acc));
        }
        if (stats)
            for (; cur_key < p.n_tiles * p.N; ++cur_key) c2_stat_flush(p, epi, cur_key, tot, lane_grp, lane, grp, ngroups);
    } else {
        // ===================== operand split: the whole halo tile, once per generation =====================
        const int tid = threadIdx.x - 192;
        const int n4 = hp.a_bytes / 16;
        uint32_t aslot = 0, aphase = 0;
        for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            for (int g = 0; g < ngen; ++g) {
                mbar_wait(fullA(aslot), aphase);
                float4* A = reinterpret_cast<float4*>(smem_gen + (size_t)aslot * hp.a_slot);
                float4* L = A + n4;
                for (int idx = tid; idx < n4; idx += 128) {
                    const float4 v = A[idx];
                    float4 hi, lo;
                    hi.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); lo.x = v.x - hi.x;
                    hi.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); lo.y = v.y - hi.y;
                    hi.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); lo.z = v.z - hi.z;
                    hi.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); lo.w = v.w - hi.w;
                    if (p.rewrite_hi) A[idx] = hi;
                    L[idx] = lo;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(readyA(aslot));
                if (++aslot == (uint32_t)hp.na) { aslot = 0; aphase ^= 1; }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

// stat_rows (if given) receives the rows per sample of the statistics table [N][rows][2][Cout] this launch
// configuration writes (0: the kernel serving it has no statistics epilogue); query = compute that only.
int conv2d_tcgen05_launch(const float* in, const float* wp, const C2Epi& epi, float* out,
                          int N, int Cin, int Cout, int Hi, int Wi, int ks, int stride, int dil, int mode, int split,
                          cudaStream_t st, int* stat_rows, bool query) {
    // N tile: <= 256 columns per MMA; with the split an accumulator is 2 nt columns wide and double buffered
    const int nt_max = (split == 1 || split == 2) ? 128 : 256;
    int n_tiles = 0;
    for (int t = 1; t <= 16 && !n_tiles; ++t)
        if (Cout % t == 0 && (Cout / t) % 16 == 0 && Cout / t <= nt_max) n_tiles = t;
    if (Cin % 32 != 0 || Cout < 16 || n_tiles == 0) {
        set_error("conv2d(tcgen05): needs Cin %% 32 == 0 and Cout splitting into equal tiles of 16..256 channels "
                  "(multiples of 16); got %d -> %d", Cin, Cout);
        return B2_ERR_UNSUPPORTED;
    }
    C2Params p{};
    p.n_tiles = n_tiles; p.nt = Cout / n_tiles;
    p.N = N; p.Cin = Cin; p.Cout = Cout; p.Hi = Hi; p.Wi = Wi; p.ks = ks; p.dil = dil; p.split = split ? 1 : 0;
    if (mode == 1) {
        p.mode = 2; p.Ho = 2 * Hi; p.Wo = 2 * Wi; p.Ht = Hi; p.Wt = Wi;
    } else {
        p.mode = stride == 2 ? 1 : 0;
        p.Ho = (Hi - 1) / stride + 1; p.Wo = (Wi - 1) / stride + 1; p.Ht = p.Ho; p.Wt = p.Wo;
    }
    p.out_cstride = Cout;
    p.tiles_w = (p.Wt + kC2TileW - 1) / kC2TileW;
    p.tiles_h = (p.Ht + kC2TileH - 1) / kC2TileH;
    p.kchunks = Cin / kC2K;
    p.b_bytes = p.nt * 128;
    p.stage_bytes = (p.split ? 2 : 1) * (kC2ABytes + p.b_bytes);
    p.stages = (208 * 1024) / p.stage_bytes;
    if (p.stages > kC2MaxStages) p.stages = kC2MaxStages;
    if (p.stages < 2) { set_error("conv2d(tcgen05): stage of %d bytes does not fit twice", p.stage_bytes); return B2_ERR_UNSUPPORTED; }
    p.tmem_cols = 32;
    p.stack = (split == 1 || split == 2) ? 1 : 0;
    p.rewrite_hi = split >= 2 ? 1 : 0;
    while (p.tmem_cols < 2 * (p.stack ? 2 : 1) * p.nt) p.tmem_cols *= 2;
    p.total_tiles = (long long)p.n_tiles * (p.mode == 2 ? 4 : 1) * N * p.tiles_h * p.tiles_w;

    const int grid = (int)(p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs);
    const bool stats_ok = p.mode != 2 && p.nt <= kC2StatMaxNt;
    if (stat_rows) *stat_rows = stats_ok ? grid * (p.split ? 1 : 2) : 0;   // one row per CTA and epilogue group
    if (query) return 0;
    if (epi.stat_mode && !stats_ok) {
        set_error("conv2d(tcgen05): no statistics epilogue for this shape (transposed, or channel tile of %d > %d)", p.nt, kC2StatMaxNt);
        return B2_ERR_UNSUPPORTED;
    }
    EncodeTiledFn encode = get_encode();
    if (!encode) { set_error("conv2d(tcgen05): cuTensorMapEncodeTiled not available from the driver"); return B2_ERR_DRIVER; }

    const bool halo = flag_value(kFlagConv2dHalo, "B2_CONV2D_HALO", 1) && p.mode == 0 && ks == 3;
    C2HaloParams hp{};
    if (halo) {
        hp.a_rows = kC2TileH + 2 * dil;
        hp.a_bytes = hp.a_rows * kC2TileW * 128;
        hp.a_slot = (p.split ? 2 : 1) * hp.a_bytes;
        hp.w_slot = (p.split ? 2 : 1) * p.b_bytes;
        hp.na = p.split ? 2 : kC2HaloMaxNA;
        hp.nw = (208 * 1024 - hp.na * hp.a_slot) / hp.w_slot;
        if (hp.nw > kC2HaloMaxNW) hp.nw = kC2HaloMaxNW;
        if (hp.nw < 2) { set_error("conv2d(tcgen05,halo): weight tiles of %d bytes do not fit", hp.w_slot); return B2_ERR_UNSUPPORTED; }
    }

    CUtensorMap map_a, map_b;
    {
        cuuint64_t gdim[4] = {(cuuint64_t)Cin, (cuuint64_t)Wi, (cuuint64_t)Hi, (cuuint64_t)N};
        cuuint64_t gstr[3] = {(cuuint64_t)Cin * 4, (cuuint64_t)Wi * Cin * 4, (cuuint64_t)Hi * Wi * Cin * 4};
        const cuuint32_t s = (p.mode == 1) ? 2 : 1;
        cuuint32_t box[4] = {(cuuint32_t)kC2K, (cuuint32_t)kC2TileW * s, (cuuint32_t)(halo ? hp.a_rows : kC2TileH * s), 1};
        cuuint32_t estr[4] = {1, s, s, 1};
        CUresult r = encode(&map_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)in, gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("conv2d(tcgen05): cuTensorMapEncodeTiled(A) failed: %d", (int)r); return B2_ERR_DRIVER; }
    }
    {
        cuuint64_t gdim[2] = {(cuuint64_t)Cin, (cuuint64_t)(p.split ? 2 : 1) * ks * ks * Cout};
        cuuint64_t gstr[1] = {(cuuint64_t)Cin * 4};
        cuuint32_t box[2] = {(cuuint32_t)kC2K, (cuuint32_t)p.nt};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&map_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)wp, gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("conv2d(tcgen05): cuTensorMapEncodeTiled(B) failed: %d", (int)r); return B2_ERR_DRIVER; }
    }
    if (halo) {
        hp.c = p;
        const int smem_h = hp.na * hp.a_slot + hp.nw * hp.w_slot + 1024;
        static SmemOptIn optin_h;
        cudaError_t eh = ensure_dynamic_smem(optin_h, conv2d_halo_tcgen05_kernel, 209 * 1024 + 1024);
        if (eh != cudaSuccess) { set_error("conv2d(tcgen05,halo): cudaFuncSetAttribute: %s", cudaGetErrorString(eh)); return (int)eh; }
        conv2d_halo_tcgen05_kernel<<<grid, kC2Threads, smem_h, st>>>(map_a, map_b, out, epi, hp);
        return check_launch("conv2d(tcgen05,halo)");
    }
    const int smem = p.stages * p.stage_bytes + 1024;
    static SmemOptIn optin;
    cudaError_t e = ensure_dynamic_smem(optin, conv2d_tcgen05_kernel, 209 * 1024 + 1024);
    if (e != cudaSuccess) { set_error("conv2d(tcgen05): cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    conv2d_tcgen05_kernel<<<grid, kC2Threads, smem, st>>>(map_a, map_b, out, epi, p);
    return check_launch("conv2d(tcgen05)");
}

// =============================================================================================
// First layer of the extractor: Conv2d(3 -> Cout, k3, stride 2, pad 1) on the NCHW image and its
// data gradient back to the NCHW pixels (the tensor the PGD update kernel reads).  K = 27 is no
// tensor-core shape and the layer is 0.4 GFLOP: exact fp32 SIMT, channels-last on the feature side.
// =============================================================================================
constexpr int kF1MaxC = 64;

__global__ void __launch_bounds__(256)
conv2d_first_fwd_kernel(const float* __restrict__ img, const float* __restrict__ w, float* __restrict__ out,
                        int N, int Cout, int H, int W, int Ho, int Wo) {
    __shared__ float ws[27 * kF1MaxC];                        // [ci*9 + kh*3 + kw][co]
    for (int i = threadIdx.x; i < 27 * Cout; i += blockDim.x) {
        const int co = i % Cout, k = i / Cout;
        ws[k * Cout + co] = w[co * 27 + k];
    }
    __syncthreads();
    const int groups = Cout / 8;
    const long long total = (long long)N * Ho * Wo * groups;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int cg = (int)(idx % groups);
        long long pix = idx / groups;
        const int ow = (int)(pix % Wo), oh = (int)((pix / Wo) % Ho), n = (int)(pix / ((long long)Wo * Ho));
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        for (int ci = 0; ci < 3; ++ci) {
            const float* plane = img + ((long long)n * 3 + ci) * H * W;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const int ih = 2 * oh + kh - 1;
                if (ih < 0 || ih >= H) continue;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int iw = 2 * ow + kw - 1;
                    if (iw < 0 || iw >= W) continue;
                    const float v = __ldg(plane + (long long)ih * W + iw);
                    const float* wr = ws + (ci * 9 + kh * 3 + kw) * Cout + cg * 8;
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] = fmaf(v, wr[j], acc[j]);
                }
            }
        }
        float4* o = reinterpret_cast<float4*>(out + pix * Cout + cg * 8);
        o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
}

__global__ void __launch_bounds__(256)
conv2d_first_dgrad_kernel(const float* __restrict__ gout, const float* __restrict__ w, float* __restrict__ gimg,
                          int N, int Cout, int H, int W, int Ho, int Wo) {
    __shared__ __align__(16) float ws[27 * kF1MaxC];          // [(kh*3 + kw)*3 + ci][co]
    for (int i = threadIdx.x; i < 27 * Cout; i += blockDim.x) {
        const int co = i % Cout, r = i / Cout;
        const int ci = r % 3, k = r / 3;
        ws[r * Cout + co] = w[co * 27 + ci * 9 + k];
    }
    __syncthreads();
    const long long total = (long long)N * H * W;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int iw = (int)(idx % W), ih = (int)((idx / W) % H), n = (int)(idx / ((long long)W * H));
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int th = ih + 1 - kh;
            if (th < 0 || (th & 1)) continue;
            const int oh = th >> 1;
            if (oh >= Ho) continue;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int tw = iw + 1 - kw;
                if (tw < 0 || (tw & 1)) continue;
                const int ow = tw >> 1;
                if (ow >= Wo) continue;
                const float4* g = reinterpret_cast<const float4*>(gout + (((long long)n * Ho + oh) * Wo + ow) * Cout);
                const float4* w0 = reinterpret_cast<const float4*>(ws + ((kh * 3 + kw) * 3 + 0) * Cout);
                const float4* w1 = reinterpret_cast<const float4*>(ws + ((kh * 3 + kw) * 3 + 1) * Cout);
                const float4* w2 = reinterpret_cast<const float4*>(ws + ((kh * 3 + kw) * 3 + 2) * Cout);
                for (int c = 0; c < Cout / 4; ++c) {
                    const float4 gv = __ldg(g + c);
                    const float4 x0 = w0[c], x1 = w1[c], x2 = w2[c];
                    a0 = fmaf(gv.x, x0.x, a0); a0 = fmaf(gv.y, x0.y, a0); a0 = fmaf(gv.z, x0.z, a0); a0 = fmaf(gv.w, x0.w, a0);
                    a1 = fmaf(gv.x, x1.x, a1); a1 = fmaf(gv.y, x1.y, a1); a1 = fmaf(gv.z, x1.z, a1); a1 = fmaf(gv.w, x1.w, a1);
                    a2 = fmaf(gv.x, x2.x, a2); a2 = fmaf(gv.y, x2.y, a2); a2 = fmaf(gv.z, x2.z, a2); a2 = fmaf(gv.w, x2.w, a2);
                }
            }
        }
        const long long plane = (long long)H * W, base = (long long)n * 3 * plane + (long long)ih * W + iw;
        gimg[base] = a0;
        gimg[base + plane] = a1;
        gimg[base + 2 * plane] = a2;
    }
}

}  // namespace b2

static int conv2d_check(const float* in, const float* wp, const float* bias, const float* addend, float* out, int N,
                        int Cin, int Cout, int Hi, int Wi, int ksize, int stride, int dilation, int mode) {
    B2_REQUIRE(in && wp && out, "conv2d: null pointer");
    B2_REQUIRE(N >= 0 && Cin > 0 && Cout > 0 && Hi > 0 && Wi > 0, "conv2d: bad dims");
    B2_REQUIRE(ksize == 1 || ksize == 3, "conv2d: kernel size must be 1 or 3 (got %d)", ksize);
    B2_REQUIRE(mode == 0 || mode == 1, "conv2d: mode must be 0 (CONV) or 1 (DECONV)");
    B2_REQUIRE((mode == 0 && (stride == 1 || stride == 2)) || (mode == 1 && stride == 2),
               "conv2d: unsupported stride %d for mode %d", stride, mode);
    B2_REQUIRE(dilation == 1 || (dilation == 2 && stride == 1 && mode == 0), "conv2d: dilation %d unsupported here", dilation);
    B2_REQUIRE(b2::aligned16(in) && b2::aligned16(out) && b2::aligned16(wp) && (!bias || b2::aligned16(bias)) &&
               (!addend || b2::aligned16(addend)), "conv2d: pointers must be 16-byte aligned");
    return 0;
}

extern "C" int b2_conv2d(const float* in, const float* wp, const float* bias, const float* addend, float* out,
                         int N, int Cin, int Cout, int Hi, int Wi, int ksize, int stride, int dilation, int mode,
                         int split, void* stream) {
    if (int e = conv2d_check(in, wp, bias, addend, out, N, Cin, Cout, Hi, Wi, ksize, stride, dilation, mode)) return e;
    if (N == 0) return 0;
    b2::C2Epi epi{};
    epi.bias = bias; epi.addend = addend;
    return b2::conv2d_tcgen05_launch(in, wp, epi, out, N, Cin, Cout, Hi, Wi, ksize, stride, dilation, mode, split,
                                     (cudaStream_t)stream, nullptr, false);
}

extern "C" int b2_conv2d_fused(const float* in, const float* wp, const float* bias, const float* addend, float* out,
                               int N, int Cin, int Cout, int Hi, int Wi, int ksize, int stride, int dilation, int mode,
                               int split, int stat_mode, float* stat_partial, const float* gn_x, const float* gn_coef,
                               int gn_coef_stride, void* stream) {
    if (int e = conv2d_check(in, wp, bias, addend, out, N, Cin, Cout, Hi, Wi, ksize, stride, dilation, mode)) return e;
    B2_REQUIRE(stat_mode >= 0 && stat_mode <= 3, "conv2d_fused: stat_mode must be 0..3");
    B2_REQUIRE(stat_mode == 0 || stat_partial, "conv2d_fused: stat_mode %d needs the statistics table", stat_mode);
    B2_REQUIRE(stat_mode < 2 || (gn_x && b2::aligned16(gn_x)), "conv2d_fused: stat_mode %d needs gn_x (16-byte aligned)", stat_mode);
    B2_REQUIRE(stat_mode != 3 || (gn_coef && gn_coef_stride >= 2 * Cout), "conv2d_fused: stat_mode 3 needs gn_coef with a per-sample stride >= 2*Cout");
    if (N == 0) return 0;
    b2::C2Epi epi{};
    epi.bias = bias; epi.addend = addend; epi.stat_mode = stat_mode; epi.stat_partial = stat_partial;
    epi.gn_x = gn_x; epi.gn_coef = gn_coef; epi.gn_coef_stride = gn_coef_stride;
    return b2::conv2d_tcgen05_launch(in, wp, epi, out, N, Cin, Cout, Hi, Wi, ksize, stride, dilation, mode, split,
                                     (cudaStream_t)stream, nullptr, false);
}

extern "C" int b2_conv2d_stat_rows(int N, int Cin, int Cout, int Hi, int Wi, int ksize, int stride, int dilation, int mode,
                                   int split, int* rows) {
    B2_REQUIRE(rows, "conv2d_stat_rows: null pointer");
    *rows = 0;
    if (N <= 0 || Cin <= 0 || Cout <= 0 || Hi <= 0 || Wi <= 0 || (ksize != 1 && ksize != 3) || (mode != 0 && mode != 1) ||
        (stride != 1 && stride != 2) || (mode == 1 && stride != 2))
        return 0;
    b2::C2Epi epi{};
    int e = b2::conv2d_tcgen05_launch(nullptr, nullptr, epi, nullptr, N, Cin, Cout, Hi, Wi, ksize, stride, dilation, mode,
                                      split, nullptr, rows, true);
    if (e) { *rows = 0; }
    return 0;
}

extern "C" int b2_conv2d_first_fwd(const float* img, const float* w, float* out, int N, int Cout, int H, int W,
                                   void* stream) {
    B2_REQUIRE(img && w && out, "conv2d_first_fwd: null pointer");
    B2_REQUIRE(N >= 0 && H > 0 && W > 0 && Cout % 8 == 0 && Cout > 0 && Cout <= b2::kF1MaxC,
               "conv2d_first_fwd: Cout must be a multiple of 8, <= %d", b2::kF1MaxC);
    B2_REQUIRE(b2::aligned16(out), "conv2d_first_fwd: out must be 16-byte aligned");
    if (N == 0) return 0;
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const int grid = b2::stream_grid((int64_t)N * Ho * Wo * (Cout / 8), 256);
    b2::conv2d_first_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img, w, out, N, Cout, H, W, Ho, Wo);
    return b2::check_launch("conv2d_first_fwd");
}

extern "C" int b2_conv2d_first_dgrad(const float* gout, const float* w, float* gimg, int N, int Cout, int H, int W,
                                     void* stream) {
    B2_REQUIRE(gout && w && gimg, "conv2d_first_dgrad: null pointer");
    B2_REQUIRE(N >= 0 && H > 0 && W > 0 && Cout % 4 == 0 && Cout > 0 && Cout <= b2::kF1MaxC,
               "conv2d_first_dgrad: Cout must be a multiple of 4, <= %d", b2::kF1MaxC);
    B2_REQUIRE(b2::aligned16(gout), "conv2d_first_dgrad: gout must be 16-byte aligned");
    if (N == 0) return 0;
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const int grid = b2::stream_grid((int64_t)N * H * W, 256);
    b2::conv2d_first_dgrad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(gout, w, gimg, N, Cout, H, W, Ho, Wo);
    return b2::check_launch("conv2d_first_dgrad");
}
