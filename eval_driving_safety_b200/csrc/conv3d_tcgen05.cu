// Subsystem (3): 3x3x3 convolutions of the hourglass stacks as an implicit GEMM on
// the 5th-gen tensor cores (impl = 0 of b2_conv3d).  Replaces cuDNN Conv3d /
// ConvTranspose3d forward + data gradient reached from
// attack/DSGN/pgd_attack.py:308 / :336.
//
//   GEMM view : M = output voxels (tile = 16 h x 8 w = 128 rows), N = Cout (<= 256, one
//               tile), K = taps x Cin.  D[M,N] += A[M,K] * B[N,K]^T, TF32 in, fp32 acc.
//   A operand : never materialised.  For every tap the rows of the A tile are the
//               SAME 16x8 voxel box shifted by the tap offset, so one tiled TMA load
//               of a {32 ch, 8 w, 16 h, 1 d, 1 n} box (128B swizzle, zero fill outside
//               the volume = the conv padding) lands a K-major UMMA tile in smem.
//               Stride-2 convs use a second tensor map with traversal strides 2;
//               transposed convs are split into the 8 output-parity classes, each a
//               unit-stride gather with 1..8 taps.
//   B operand : packed weights wp[27][Cout][Cin] seen as a 2-D K-major matrix
//               {Cin, 27*Cout}; TMA box {32, Cout}.
//   Pipeline  : warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (one thread),
//               warps 2-5 = epilogue (tcgen05.ld -> st.global, one voxel row of
//               Cout*4 contiguous bytes per thread).  smem ring of full/empty
//               mbarriers; two TMEM accumulators so the epilogue of tile i overlaps
//               the MMAs of tile i+1.  Persistent: one CTA per SM, static tile striding.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "tcgen05.cuh"

namespace b2 {

constexpr int kTcThreads = 192;
constexpr int kTileH = 16, kTileW = 8;          // 128 voxel rows per tile
constexpr int kKChunk = 32;                     // floats per K block = one 128B swizzle row
constexpr int kABytes = 128 * 128;              // A stage: 128 rows x 128 B
constexpr int kMaxStages = 8;

struct TcParams {
    int N, Cin, Cout;
    int Dt, Ht, Wt;            // extents of the tile grid space in ROLE order (planes, rows, w)
                               // (conv: output dims; deconv: input dims)
    int Do, Ho, Wo;            // output dims
    int mode;                  // 0 conv s1, 1 conv s2, 2 deconv s2
    int swap;                  // 0: tile rows run along H, planes along D; 1: rows along D, planes along H
    int tiles_w, tiles_h;
    int kchunks;               // Cin / 32
    int stages;
    int stage_bytes;           // kABytes + Cout*128
    int tmem_cols;             // power of two >= 2*Cout
    long long total_tiles;
};


// ---------------------------------------------------------------- GroupNorm statistics in the epilogue
// v[i] (i = 0..31) holds this lane's value for channel i; on return v[0] of lane l is the sum over
// the warp's 32 lanes of channel l (recursive halving: 16+8+4+2+1 = 31 shuffles, fixed order).
__device__ __forceinline__ void warp_transpose_sum32(float (&v)[32], int lane) {
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

constexpr int kEpiMaxC = 128;              // widest norm whose scale / shift the epilogue keeps in smem (stat_mode 3)
constexpr int kEpiStageBytes = 1024;       // per epilogue warp: hand-over of the statistics totals at kernel end

// Row access of the epilogue: lane = voxel row (how tcgen05.ld delivers a tile), 32 channels = 128 B
// contiguous per lane.  (A warp-transposed variant through smem -- 8 lanes per 128 B segment, 4 full
// lines per instruction -- was measured SLOWER: plain launches unchanged, fused-operand launches
// +50 %; the epilogue is bound by the dependent-load latency of each step, not by LSU wavefronts.)
// ptr_of(r) -> pointer to the 32-channel segment of warp row r (0..31), nullptr when the row is outside the volume
template <typename PtrFn>
__device__ __forceinline__ void epi_store32(float4*, int lane, const uint32_t (&rr)[32], PtrFn ptr_of) {
    float* gp = ptr_of(lane);
    if (!gp) return;
#pragma unroll
    for (int q = 0; q < 8; ++q)
        *reinterpret_cast<float4*>(gp + 4 * q) = make_float4(__uint_as_float(rr[4 * q]), __uint_as_float(rr[4 * q + 1]),
                                                             __uint_as_float(rr[4 * q + 2]), __uint_as_float(rr[4 * q + 3]));
}

template <typename PtrFn>
__device__ __forceinline__ void epi_load32(float4*, int lane, float (&xv)[32], PtrFn ptr_of) {
    const float* gp = ptr_of(lane);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float4 t = gp ? __ldg(reinterpret_cast<const float4*>(gp) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        xv[4 * q] = t.x; xv[4 * q + 1] = t.y; xv[4 * q + 2] = t.z; xv[4 * q + 3] = t.w;
    }
}

// One 32-channel chunk of one tile row set: fused addend, statistics, store.
// off_of(r) -> element offset of warp row r's segment (negative when the row is outside); ok = this lane's row is inside.
template <typename OffFn>
__device__ __forceinline__ void epi_chunk32(const EpiFusion& ef, float* __restrict__ out, uint32_t (&rr)[32], float4* stage,
                                            int lane, bool ok, OffFn off_of, float (&s0)[32], float (&s1)[32], int cbase,
                                            const float* coef_sh) {
    const long long myoff = off_of(lane);
    float av[32], xv[32];
    // both operand rows are requested up front (16 independent 16-byte loads in flight per lane)
    if (ef.addend) epi_load32(stage, lane, av, [&](int) { return myoff < 0 ? (const float*)nullptr : ef.addend + myoff; });
    if (ef.stat_mode >= 2) epi_load32(stage, lane, xv, [&](int) { return myoff < 0 ? (const float*)nullptr : ef.gn_x + myoff; });
    if (ef.addend) {
#pragma unroll
        for (int i = 0; i < 32; ++i) rr[i] = __float_as_uint(__uint_as_float(rr[i]) + av[i]);
    }
    if (ef.stat_mode == 1) {
        if (ok) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float v = __uint_as_float(rr[i]);
                s0[i] += v;
                s1[i] = fmaf(v, v, s1[i]);
            }
        }
    } else if (ef.stat_mode >= 2) {
        if (ok) {
            const bool mask = ef.stat_mode == 3;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                float v = __uint_as_float(rr[i]);
                if (mask && !(fmaf(xv[i], coef_sh[cbase + i], coef_sh[kEpiMaxC + cbase + i]) > 0.f)) v = 0.f;
                s0[i] = fmaf(v, xv[i], s0[i]);
                s1[i] += v;
            }
        }
    }
    epi_store32(stage, lane, rr, [&](int r) { const long long o = off_of(r); return o < 0 ? (float*)nullptr : out + o; });
}

// 16-channel tail (Nt = 16 or 48): direct per-row access, addend only (no statistics)
__device__ __forceinline__ void epi_tail16(const EpiFusion& ef, float* __restrict__ out, uint32_t (&rr)[16], bool ok, long long off) {
    if (!ok) return;
    if (ef.addend) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(ef.addend + off) + q);
            rr[4 * q] = __float_as_uint(__uint_as_float(rr[4 * q]) + a.x);
            rr[4 * q + 1] = __float_as_uint(__uint_as_float(rr[4 * q + 1]) + a.y);
            rr[4 * q + 2] = __float_as_uint(__uint_as_float(rr[4 * q + 2]) + a.z);
            rr[4 * q + 3] = __float_as_uint(__uint_as_float(rr[4 * q + 3]) + a.w);
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
        *reinterpret_cast<float4*>(out + off + 4 * q) =
            make_float4(__uint_as_float(rr[4 * q]), __uint_as_float(rr[4 * q + 1]), __uint_as_float(rr[4 * q + 2]),
                        __uint_as_float(rr[4 * q + 3]));
}

// Per-lane running totals of one epilogue warp: slot = n_tile * 2 + (32-channel chunk), lane = channel
// inside the chunk.  Static indexing only (stays in registers).
struct StatTotals {
    float s[4], q[4];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < 4; ++i) { s[i] = 0.f; q[i] = 0.f; }
    }
    __device__ __forceinline__ void add(int slot, float ds, float dq) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i == slot) { s[i] += ds; q[i] += dq; }
    }
};

// After the CTA-wide barrier: one warp adds the four epilogue warps' totals in a fixed order and writes
// this CTA's row of the partial-sum table [gridDim.x][2][Cout] (sum, sum of squares per channel).
__device__ __forceinline__ void stat_store_row(const float* shp, float* __restrict__ stat_partial, int Cout,
                                               int nt, int n_tiles, int lane) {
    const float (*sh)[8][32] = reinterpret_cast<const float (*)[8][32]>(shp);
    float* row = stat_partial + (long long)blockIdx.x * 2 * Cout;
    for (int nti = 0; nti < n_tiles; ++nti)
        for (int ch = 0; ch < nt / 32; ++ch) {
            const int slot = nti * 2 + ch;
            const float ts = ((sh[0][slot][lane] + sh[1][slot][lane]) + sh[2][slot][lane]) + sh[3][slot][lane];
            const float tq = ((sh[0][4 + slot][lane] + sh[1][4 + slot][lane]) + sh[2][4 + slot][lane]) + sh[3][4 + slot][lane];
            row[nti * nt + ch * 32 + lane] = ts;
            row[Cout + nti * nt + ch * 32 + lane] = tq;
        }
}

// ---------------------------------------------------------------- tile / step decode
struct TileCoord { int cls, n, d, h0, w0; };

__device__ __forceinline__ TileCoord decode_tile(const TcParams& p, long long t) {
    TileCoord c;
    c.w0 = (int)(t % p.tiles_w) * kTileW; t /= p.tiles_w;
    c.h0 = (int)(t % p.tiles_h) * kTileH; t /= p.tiles_h;
    c.d = (int)(t % p.Dt); t /= p.Dt;
    c.n = (int)(t % p.N); t /= p.N;
    c.cls = (int)t;
    return c;
}
__device__ __forceinline__ int tile_taps(const TcParams& p, int cls) {
    if (p.mode != 2) return 27;
    return (1 + (cls & 1)) * (1 + ((cls >> 1) & 1)) * (1 + ((cls >> 2) & 1));
}
// deconv: parity p along one axis, j-th tap of that axis -> (k, input shift)
__device__ __forceinline__ void deconv_axis(int par, int j, int& k, int& shift) {
    if (par == 0) { k = 1; shift = 0; }
    else if (j == 0) { k = 0; shift = 1; }
    else { k = 2; shift = 0; }
}

__global__ void __launch_bounds__(kTcThreads, 1)
conv3d_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      float* __restrict__ out, const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 4];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;     // 128B swizzle needs 1024 B alignment
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (kMaxStages + s); };
    auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * kMaxStages + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * kMaxStages + 2 + a); };

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {   // whole warp: allocate TMEM (2 accumulators of Cout fp32 columns)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&tmem_base_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        // ===================== TMA producer (one thread) =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            const uint32_t tx = (uint32_t)p.stage_bytes;
            for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                TileCoord tc = decode_tile(p, t);
                const int ntaps = tile_taps(p, tc.cls);
                const int pw = tc.cls & 1, ph = (tc.cls >> 1) & 1, pd = (tc.cls >> 2) & 1;
                for (int tap_i = 0; tap_i < ntaps; ++tap_i) {
                    int aw, ah, ad, tap;
                    // kd / kh below are the taps along the PLANE / ROW roles
                    int kw, kh, kd;
                    if (p.mode == 0) {
                        kd = tap_i / 9; kh = (tap_i / 3) % 3; kw = tap_i % 3;
                        aw = tc.w0 + kw - 1; ah = tc.h0 + kh - 1; ad = tc.d + kd - 1;
                    } else if (p.mode == 1) {
                        kd = tap_i / 9; kh = (tap_i / 3) % 3; kw = tap_i % 3;
                        aw = 2 * tc.w0 + kw - 1; ah = 2 * tc.h0 + kh - 1; ad = 2 * tc.d + kd - 1;
                    } else {
                        int nw = 1 + pw, nh = 1 + ph;
                        int jw = tap_i % nw, jh = (tap_i / nw) % nh, jd = tap_i / (nw * nh);
                        int sw, sh, sd;
                        deconv_axis(pw, jw, kw, sw); deconv_axis(ph, jh, kh, sh); deconv_axis(pd, jd, kd, sd);
                        aw = tc.w0 + sw; ah = tc.h0 + sh; ad = tc.d + sd;
                    }
                    tap = p.swap ? (kh * 3 + kd) * 3 + kw : (kd * 3 + kh) * 3 + kw;   // weights are [kD][kH][kW]
                    for (int kc = 0; kc < p.kchunks; ++kc) {
                        mbar_wait(empty_bar(stage), phase ^ 1);
                        mbar_expect_tx(full_bar(stage), tx);
                        uint32_t sa = smem_base + (uint32_t)stage * p.stage_bytes;
                        tma_load_5d(sa, &map_a, full_bar(stage), kc * kKChunk, aw, ah, ad, tc.n);
                        tma_load_2d(sa + kABytes, &map_b, full_bar(stage), kc * kKChunk, tap * p.Cout);
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(128, p.Cout);
            int stage = 0; uint32_t phase = 0;
            long long it = 0;
            for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
                const int acc = (int)(it & 1);
                const uint32_t acc_phase = (uint32_t)((it >> 1) & 1);
                mbar_wait(tempty_bar(acc), acc_phase ^ 1);        // epilogue drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * p.Cout);
                TileCoord tc = decode_tile(p, t);
                const int nsteps = tile_taps(p, tc.cls) * p.kchunks;
                for (int s = 0; s < nsteps; ++s) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    uint32_t sa = smem_base + (uint32_t)stage * p.stage_bytes;
                    uint64_t adesc = umma_desc_sw128(sa), bdesc = umma_desc_sw128(sa + kABytes);
#pragma unroll
                    for (int k = 0; k < kKChunk / 8; ++k) {
                        // +32 B per K step inside the 128B swizzle row: +2 in the (addr >> 4) field
                        umma_tf32(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (s | k) != 0);
                    }
                    umma_commit(empty_bar(stage));                // frees the smem slot when the MMAs retire
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                umma_commit(tfull_bar(acc));                      // accumulator ready for the epilogue
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue (4 warps, one voxel row per thread) =====================
        const int lane_grp = warp & 3;                            // TMEM lanes [32*lane_grp, +32)
        const int m = lane_grp * 32 + lane;
        const int hl = m / kTileW, wl = m % kTileW;
        long long it = 0;
        for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
            const int acc = (int)(it & 1);
            const uint32_t acc_phase = (uint32_t)((it >> 1) & 1);
            TileCoord tc = decode_tile(p, t);
            const int h = tc.h0 + hl, w = tc.w0 + wl;
            const bool ok = h < p.Ht && w < p.Wt;
            int op = tc.d, orow = h, ow = w;                       // plane / row / w coordinates of the output
            if (p.mode == 2) {
                op = 2 * tc.d + ((tc.cls >> 2) & 1); orow = 2 * h + ((tc.cls >> 1) & 1); ow = 2 * w + (tc.cls & 1);
            }
            const int od = p.swap ? orow : op, oh = p.swap ? op : orow;
            const long long vox = (((long long)tc.n * p.Do + od) * p.Ho + oh) * p.Wo + ow;
            float* optr = out + vox * p.Cout;
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(acc * p.Cout);
            for (int c0 = 0; c0 < p.Cout; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(taddr + c0, r);
                tmem_ld_wait();
                if (ok) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 v = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                               __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
                        *reinterpret_cast<float4*>(optr + c0 + 4 * j) = v;
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(tempty_bar(acc));
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
// Stride-1 fast path ("super tile").  The generic kernel above reloads the A box and the weight
// tile for every tap: 1.33 MB of L2->SM traffic per 128-voxel tile, measured L2-bound at
// 290 TFLOP/s (profiles/r1_conv3d_tcgen05_ncu.txt).  Here one CTA owns kS1Planes = 4 consecutive
// depth planes of a 16h x 8w box (4 accumulators of Nt columns, double buffered = all of TMEM
// for Nt = 64) and walks (kw, K-chunk) "generations":
//   * A: per generation the 6 input planes d0-1 .. d0+4 are loaded ONCE each as an
//     {32 ch, 8 w, 18 h} halo box (18 KB).  The three kh taps read the same smem tile at row
//     offsets 0 / 8 / 16 (= +0 / +1024 / +2048 B, whole 128B-swizzle atoms), the three kd taps
//     feed three different accumulators -> each A byte is used by up to 9 MMA groups.
//   * B: the 9 (kd, kh) weight tiles of the generation are loaded once (ring of slots with their
//     own full/empty barriers) and reused by all 4 planes.
//   L2->SM traffic: 1080 KB per 512 voxels (4.9x less).  Cout > 64 is split in two N tiles.
// =============================================================================================
constexpr int kS1Planes = 4;
constexpr int kS1ARows = (kTileH + 2) * kTileW;        // 144 rows
constexpr int kS1ABytes = kS1ARows * 128;              // 18432 B = 18 swizzle atoms
constexpr int kS1NA = 4;                               // A ring slots
constexpr int kS1MaxNB = 18;                           // B ring slots (>= 9)

struct S1Params {
    int N, Cin, Cout, D, H, W;
    int P, R, swap;            // planes / rows extents in role order; swap = rows along D, planes along H
    int nt, n_tiles;           // N tile width, number of N tiles
    int tiles_w, tiles_h, dblocks;
    int kchunks;
    int nb;                    // B ring slots
    int b_bytes;               // nt * 128
    int tmem_cols;
    long long total_tiles;
};


struct S1Tile { int nti, n, d0, h0, w0; };

__device__ __forceinline__ S1Tile s1_decode(const S1Params& p, long long t) {
    S1Tile c;
    c.w0 = (int)(t % p.tiles_w) * kTileW; t /= p.tiles_w;
    c.h0 = (int)(t % p.tiles_h) * kTileH; t /= p.tiles_h;
    c.d0 = (int)(t % p.dblocks) * kS1Planes; t /= p.dblocks;
    c.n = (int)(t % p.N); t /= p.N;
    c.nti = (int)t;
    return c;
}

__global__ void __launch_bounds__(kTcThreads, 1)
conv3d_s1_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                         float* __restrict__ out, const S1Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * kS1NA + 2 * kS1MaxNB + 4];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_base = smem_base, b_base = smem_base + kS1NA * kS1ABytes;
    const uint32_t bar0 = smem_u32(bars);
    auto fullA = [&](int s) { return bar0 + 8u * s; };
    auto emptyA = [&](int s) { return bar0 + 8u * (kS1NA + s); };
    auto fullB = [&](int s) { return bar0 + 8u * (2 * kS1NA + s); };
    auto emptyB = [&](int s) { return bar0 + 8u * (2 * kS1NA + kS1MaxNB + s); };
    auto tfull = [&](int a) { return bar0 + 8u * (2 * kS1NA + 2 * kS1MaxNB + a); };
    auto tempty = [&](int a) { return bar0 + 8u * (2 * kS1NA + 2 * kS1MaxNB + 2 + a); };

    if (threadIdx.x == 0) {
        for (int s = 0; s < kS1NA; ++s) { mbar_init(fullA(s), 1); mbar_init(emptyA(s), 1); }
        for (int s = 0; s < p.nb; ++s) { mbar_init(fullB(s), 1); mbar_init(emptyB(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), 128); }
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
    const int ngen = 3 * p.kchunks;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t a_ord = 0, b_ord = 0;
            for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                S1Tile tc = s1_decode(p, t);
                for (int g = 0; g < ngen; ++g) {
                    const int kw = g / p.kchunks, kc = g % p.kchunks;
                    for (int pr = 0; pr < kS1Planes + 2; ++pr) {
                        if (pr < 3) {
                            for (int kh = 0; kh < 3; ++kh, ++b_ord) {
                                const int slot = b_ord % p.nb;
                                mbar_wait(emptyB(slot), ((b_ord / p.nb) & 1) ^ 1);
                                mbar_expect_tx(fullB(slot), (uint32_t)p.b_bytes);
                                const int tap = p.swap ? (kh * 3 + pr) * 3 + kw : (pr * 3 + kh) * 3 + kw;
                                tma_load_2d(b_base + (uint32_t)slot * p.b_bytes, &map_b, fullB(slot), kc * kKChunk,
                                            tap * p.Cout + tc.nti * p.nt);
                            }
                        }
                        const int slot = a_ord % kS1NA;
                        mbar_wait(emptyA(slot), ((a_ord / kS1NA) & 1) ^ 1);
                        mbar_expect_tx(fullA(slot), (uint32_t)kS1ABytes);
                        tma_load_5d(a_base + (uint32_t)slot * kS1ABytes, &map_a, fullA(slot), kc * kKChunk,
                                    tc.w0 + kw - 1, tc.h0 - 1, tc.d0 - 1 + pr, tc.n);
                        ++a_ord;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // One thread issues 864 MMAs per super tile; each MMA occupies the tensor pipe for only
        // 32 cycles, so the scalar code between MMAs must stay tiny: the plane / kd / kh loops are
        // fully unrolled (accumulator index, first/last-use tests are compile-time) and the ring
        // slots of the 9 weight tiles of a generation are resolved once per generation.
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(128, p.nt);
            uint32_t aslot = 0, aphase = 0;            // A ring position
            uint32_t bslot0 = 0, bphase0 = 0;          // B ring position of the generation's first tap
            long long it = 0;
            for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
                const int accbuf = (int)(it & 1);
                mbar_wait(tempty(accbuf), (uint32_t)(((it >> 1) & 1) ^ 1));
                tc_fence_after();
                const uint32_t tmem_acc0 = tmem_base + (uint32_t)(accbuf * kS1Planes * p.nt);
                for (int g = 0; g < ngen; ++g) {
                    uint64_t bdesc[9];
                    uint32_t bfull[9], bempty[9], bpar[9];
                    {
                        uint32_t sl = bslot0, ph = bphase0;
#pragma unroll
                        for (int q = 0; q < 9; ++q) {
                            bdesc[q] = umma_desc_sw128(b_base + sl * (uint32_t)p.b_bytes);
                            bfull[q] = fullB(sl); bempty[q] = emptyB(sl); bpar[q] = ph;
                            if (++sl == (uint32_t)p.nb) { sl = 0; ph ^= 1; }
                        }
                        bslot0 = sl; bphase0 = ph;
                    }
#pragma unroll
                    for (int pr = 0; pr < kS1Planes + 2; ++pr) {
                        mbar_wait(fullA(aslot), aphase);
                        tc_fence_after();
                        const uint64_t adesc0 = umma_desc_sw128(a_base + aslot * (uint32_t)kS1ABytes);
#pragma unroll
                        for (int kd = 0; kd < 3; ++kd) {
                            const int j = pr - kd;                 // accumulator = output plane d0 + j
                            if (j < 0 || j >= kS1Planes) continue;
                            const uint32_t tmem_d = tmem_acc0 + (uint32_t)(j * p.nt);
#pragma unroll
                            for (int kh = 0; kh < 3; ++kh) {
                                const int q = kd * 3 + kh;
                                if (j == 0) {                      // first use of this weight tile
                                    mbar_wait(bfull[q], bpar[q]);
                                    tc_fence_after();
                                }
                                const uint64_t adesc = adesc0 + (uint64_t)(kh * (1024 >> 4));
                                const uint32_t later = (uint32_t)g | (uint32_t)q;   // 0 only for the very first MMA group of acc j
#pragma unroll
                                for (int k = 0; k < kKChunk / 8; ++k)
                                    umma_tf32(tmem_d, adesc + 2 * k, bdesc[q] + 2 * k, idesc, (later | (uint32_t)k) != 0);
                                if (j == kS1Planes - 1) umma_commit(bempty[q]);   // last use of the tile
                            }
                        }
                        umma_commit(emptyA(aslot));
                        if (++aslot == kS1NA) { aslot = 0; aphase ^= 1; }
                    }
                }
                umma_commit(tfull(accbuf));
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue =====================
        const int lane_grp = warp & 3;
        const int m = lane_grp * 32 + lane;
        const int hl = m / kTileW, wl = m % kTileW;
        long long it = 0;
        for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
            const int accbuf = (int)(it & 1);
            S1Tile tc = s1_decode(p, t);
            const int r = tc.h0 + hl, w = tc.w0 + wl;
            const bool ok_hw = r < p.R && w < p.W;
            mbar_wait(tfull(accbuf), (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            for (int j = 0; j < kS1Planes; ++j) {
                const int pl = tc.d0 + j;
                const bool ok = ok_hw && pl < p.P;
                const int d = p.swap ? r : pl, h = p.swap ? pl : r;
                float* orow = out + ((((long long)tc.n * p.D + d) * p.H + h) * p.W + w) * p.Cout + tc.nti * p.nt;
                const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) +
                                       (uint32_t)((accbuf * kS1Planes + j) * p.nt);
                int c0 = 0;
                for (; c0 + 32 <= p.nt; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld32(taddr + c0, r);
                    tmem_ld_wait();
                    if (ok) {
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            *reinterpret_cast<float4*>(orow + c0 + 4 * q) =
                                make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                            __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
                    }
                }
                if (c0 < p.nt) {                                   // 16-column remainder (Nt = 48)
                    uint32_t r[16];
                    tmem_ld16(taddr + c0, r);
                    tmem_ld_wait();
                    if (ok) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            *reinterpret_cast<float4*>(orow + c0 + 4 * q) =
                                make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                            __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(tempty(accbuf));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

// ---------------------------------------------------------------- host side


// =============================================================================================
// Stride-1 fast path, "N-stacked" variant (default).  For one input plane pr and one (kh, kw, chunk),
// the taps kd = 2,1,0 feed the accumulators of output planes pr-2, pr-1, pr -- which sit in
// ADJACENT TMEM column ranges.  With the three weight tiles [W(kd=2); W(kd=1); W(kd=0)] stored
// contiguously in smem, ONE tcgen05.mma with N = 3*Nt (192 for Nt = 64) covers all three planes:
// the A tile is read from smem once instead of three times.  Operand smem traffic per generation
// drops from 864 KB to 583 KB (the super-tile kernel above is smem-read bound: ~192 B/clk needed at
// N = 64 vs 128 B/clk available), tensor-pipe bound goes from ~56 % to ~77 %.
// Every MMA accumulates (the epilogue zeroes an accumulator with tcgen05.st after draining it),
// because one MMA may touch a fresh and a partially summed plane at the same time.
// =============================================================================================

__global__ void __launch_bounds__(kTcThreads, 1)
conv3d_s1n_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                          float* __restrict__ out, const S1Params p, const EpiFusion ef) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * kS1NA + 2 * kS1MaxNB + 4];
    __shared__ uint32_t tmem_base_slot;
    __shared__ float coef_sh[2 * kEpiMaxC];      // stat_mode 3: forward scale / shift of the preceding norm
    float* __restrict__ stat_partial = ef.stat_partial;
    if (ef.stat_mode == 3)
        for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) {
            coef_sh[i] = ef.gn_coef[i];
            coef_sh[kEpiMaxC + i] = ef.gn_coef[p.Cout + i];
        }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_base = smem_base, b_base = smem_base + kS1NA * kS1ABytes;
    // epilogue staging (4 warps x 2 KB) behind the rings; reused for the statistics hand-over at the end
    float* const stage_all = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) + kS1NA * kS1ABytes +
                                                      p.nb * p.b_bytes);
    const uint32_t bar0 = smem_u32(bars);
    auto fullA = [&](int s) { return bar0 + 8u * s; };
    auto emptyA = [&](int s) { return bar0 + 8u * (kS1NA + s); };
    auto fullB = [&](int s) { return bar0 + 8u * (2 * kS1NA + s); };
    auto emptyB = [&](int s) { return bar0 + 8u * (2 * kS1NA + kS1MaxNB + s); };
    auto tfull = [&](int a) { return bar0 + 8u * (2 * kS1NA + 2 * kS1MaxNB + a); };
    auto tempty = [&](int a) { return bar0 + 8u * (2 * kS1NA + 2 * kS1MaxNB + 2 + a); };

    if (threadIdx.x == 0) {
        for (int s = 0; s < kS1NA; ++s) { mbar_init(fullA(s), 1); mbar_init(emptyA(s), 1); }
        for (int s = 0; s < p.nb; ++s) { mbar_init(fullB(s), 1); mbar_init(emptyB(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), 128); }
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
    const int ngen = 3 * p.kchunks;
    const int gper = p.nb / 9;                  // generations resident in the B ring (1 or 2)

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t aslot = 0, aphase = 0, gord = 0;
            for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                S1Tile tc = s1_decode(p, t);
                for (int g = 0; g < ngen; ++g, ++gord) {
                    const int kw = g / p.kchunks, kc = g % p.kchunks;
                    const uint32_t bbase = (gord % gper) * 9, bphase = (gord / gper) & 1;
                    for (int pr = 0; pr < kS1Planes + 2; ++pr) {
                        if (pr < 3) {                       // weight tiles of tap kd = pr, stored at position 2-kd
                            for (int kh = 0; kh < 3; ++kh) {
                                const uint32_t slot = bbase + kh * 3 + (2 - pr);
                                mbar_wait(emptyB(slot), bphase ^ 1);
                                mbar_expect_tx(fullB(slot), (uint32_t)p.b_bytes);
                                const int tap = p.swap ? (kh * 3 + pr) * 3 + kw : (pr * 3 + kh) * 3 + kw;
                                tma_load_2d(b_base + slot * (uint32_t)p.b_bytes, &map_b, fullB(slot), kc * kKChunk,
                                            tap * p.Cout + tc.nti * p.nt);
                            }
                        }
                        mbar_wait(emptyA(aslot), aphase ^ 1);
                        mbar_expect_tx(fullA(aslot), (uint32_t)kS1ABytes);
                        tma_load_5d(a_base + aslot * (uint32_t)kS1ABytes, &map_a, fullA(aslot), kc * kKChunk,
                                    tc.w0 + kw - 1, tc.h0 - 1, tc.d0 - 1 + pr, tc.n);
                        if (++aslot == kS1NA) { aslot = 0; aphase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc1 = umma_idesc_tf32(128, p.nt), idesc2 = umma_idesc_tf32(128, 2 * p.nt),
                           idesc3 = umma_idesc_tf32(128, 3 * p.nt);
            uint32_t aslot = 0, aphase = 0, gord = 0;
            long long it = 0;
            for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
                const int accbuf = (int)(it & 1);
                mbar_wait(tempty(accbuf), (uint32_t)((it >> 1) & 1));     // drained AND zeroed by the epilogue
                tc_fence_after();
                const uint32_t tmem_acc0 = tmem_base + (uint32_t)(accbuf * kS1Planes * p.nt);
                for (int g = 0; g < ngen; ++g, ++gord) {
                    const uint32_t bbase = (gord % gper) * 9, bphase = (gord / gper) & 1;
#pragma unroll
                    for (int pr = 0; pr < kS1Planes + 2; ++pr) {
                        mbar_wait(fullA(aslot), aphase);
                        tc_fence_after();
                        const uint64_t adesc0 = umma_desc_sw128(a_base + aslot * (uint32_t)kS1ABytes);
                        const int kd_hi = pr < 2 ? pr : 2, kd_lo = pr > 3 ? pr - 3 : 0;
                        const int nplanes = kd_hi - kd_lo + 1;
                        const uint32_t idesc = nplanes == 1 ? idesc1 : (nplanes == 2 ? idesc2 : idesc3);
                        const uint32_t tmem_d = tmem_acc0 + (uint32_t)((pr - kd_hi) * p.nt);
#pragma unroll
                        for (int kh = 0; kh < 3; ++kh) {
                            const uint32_t slot_hi = bbase + kh * 3 + (2 - kd_hi);
                            if (pr <= 2) {                      // first use of the tile kd = pr (= kd_hi)
                                mbar_wait(fullB(slot_hi), bphase);
                                tc_fence_after();
                            }
                            const uint64_t adesc = adesc0 + (uint64_t)(kh * (1024 >> 4));
                            const uint64_t bdesc = umma_desc_sw128(b_base + slot_hi * (uint32_t)p.b_bytes);
#pragma unroll
                            for (int k = 0; k < kKChunk / 8; ++k)
                                umma_tf32(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
                            if (pr >= 3) umma_commit(emptyB(bbase + kh * 3 + (2 - (pr - 3))));   // last use of kd = pr-3
                        }
                        umma_commit(emptyA(aslot));
                        if (++aslot == kS1NA) { aslot = 0; aphase ^= 1; }
                    }
                }
                umma_commit(tfull(accbuf));
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue =====================
        const int lane_grp = warp & 3;
        const int m = lane_grp * 32 + lane;
        const int hl = m / kTileW, wl = m % kTileW;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(lane_grp * 32) << 16);
        // zero both accumulator buffers, then hand them to the MMA warp
        for (int c0 = 0; c0 < 2 * kS1Planes * p.nt; c0 += 16) tmem_st16_zero(lane_addr + c0);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(tempty(0));
        mbar_arrive(tempty(1));
        const bool stats = stat_partial != nullptr;       // fused GroupNorm statistics of the output
        float4* const stage = reinterpret_cast<float4*>(stage_all + lane_grp * (kEpiStageBytes / 4));
        StatTotals tot;
        tot.clear();
        long long it = 0;
        for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
            const int accbuf = (int)(it & 1);
            S1Tile tc = s1_decode(p, t);
            const int r = tc.h0 + hl, w = tc.w0 + wl;
            const bool ok_hw = r < p.R && w < p.W;
            mbar_wait(tfull(accbuf), (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            const long long tile_off = (long long)tc.n * p.D * p.H * p.W * p.Cout + tc.nti * p.nt;
            // element offset of warp row rw (0..31) of plane j, or -1 outside the volume
            auto off_of = [&](int rw, int j) -> long long {
                const int m2 = lane_grp * 32 + rw;
                const int r2 = tc.h0 + (m2 >> 3), w2 = tc.w0 + (m2 & 7), pl = tc.d0 + j;
                if (r2 >= p.R || w2 >= p.W || pl >= p.P) return -1;
                const int d = p.swap ? r2 : pl, h = p.swap ? pl : r2;
                return tile_off + (((long long)d * p.H + h) * p.W + w2) * p.Cout;
            };
            int c0 = 0;
            for (; c0 + 32 <= p.nt; c0 += 32) {           // 32-channel chunk outer, the 4 planes inner
                float ssum[32], ssq[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) { ssum[i] = 0.f; ssq[i] = 0.f; }
#pragma unroll 1
                for (int j = 0; j < kS1Planes; ++j) {
                    const bool ok = ok_hw && tc.d0 + j < p.P;
                    const uint32_t taddr = lane_addr + (uint32_t)((accbuf * kS1Planes + j) * p.nt);
                    uint32_t rr[32];
                    tmem_ld32(taddr + c0, rr);
                    tmem_ld_wait();
                    tmem_st32_zero(taddr + c0);
                    epi_chunk32(ef, out, rr, stage, lane, ok,
                                [&](int rw) { const long long o = off_of(rw, j); return o < 0 ? o : o + c0; },
                                ssum, ssq, tc.nti * p.nt + c0, coef_sh);
                }
                if (stats) {
                    warp_transpose_sum32(ssum, lane);
                    warp_transpose_sum32(ssq, lane);
                    tot.add(tc.nti * 2 + (c0 >> 5), ssum[0], ssq[0]);
                }
            }
            if (c0 < p.nt) {                               // 16-channel tail (Nt = 16 or 48; no statistics)
                for (int j = 0; j < kS1Planes; ++j) {
                    const bool ok = ok_hw && tc.d0 + j < p.P;
                    const uint32_t taddr = lane_addr + (uint32_t)((accbuf * kS1Planes + j) * p.nt);
                    uint32_t rr[16];
                    tmem_ld16(taddr + c0, rr);
                    tmem_ld_wait();
                    tmem_st16_zero(taddr + c0);
                    epi_tail16(ef, out, rr, ok, off_of(lane, j) + c0);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(tempty(accbuf));
        }
        if (stats) {                                       // [warp][sum slots 0..3, second-sum slots 4..7][lane]
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                stage_all[(lane_grp * 8 + i) * 32 + lane] = tot.s[i];
                stage_all[(lane_grp * 8 + 4 + i) * 32 + lane] = tot.q[i];
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (stat_partial != nullptr && warp == 2) stat_store_row(stage_all, stat_partial, p.Cout, p.nt, p.n_tiles, lane);
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

// stat_partial != null: also write the per-CTA GroupNorm partial sums of the output, table
// [grid][2][Cout]; *stat_rows (if given) receives the number of rows (= grid), 0 when this launch
// configuration cannot produce them.  query = only compute *stat_rows, launch nothing.
static int conv3d_s1_launch(EncodeTiledFn encode, const float* in, const float* wp, float* out, int N, int Cin,
                            int Cout, int D, int H, int W, cudaStream_t st, const EpiFusion& ef, int* stat_rows,
                            bool query, int* addend_ok) {
    const float* addend = ef.addend;
    float* stat_partial = ef.stat_partial;
    S1Params p{};
    p.N = N; p.Cin = Cin; p.Cout = Cout; p.D = D; p.H = H; p.W = W;
    p.nt = Cout <= 64 ? Cout : Cout / 2;
    p.n_tiles = Cout / p.nt;
    {
        auto padded = [](int rows, int planes) {
            return (long long)((rows + kTileH - 1) / kTileH) * kTileH * ((planes + kS1Planes - 1) / kS1Planes) * kS1Planes;
        };
        p.swap = padded(D, H) < padded(H, D) ? 1 : 0;
        p.R = p.swap ? D : H;
        p.P = p.swap ? H : D;
    }
    p.tiles_w = (W + kTileW - 1) / kTileW;
    p.tiles_h = (p.R + kTileH - 1) / kTileH;
    p.dblocks = (p.P + kS1Planes - 1) / kS1Planes;
    p.kchunks = Cin / kKChunk;
    p.b_bytes = p.nt * 128;
    p.nb = (216 * 1024 - kS1NA * kS1ABytes) / p.b_bytes;
    if (p.nb > kS1MaxNB) p.nb = kS1MaxNB;
    static int nstack = -1;
    if (nstack < 0) { const char* e = getenv("B2_CONV_S1_NSTACK"); nstack = (e && e[0] == '0') ? 0 : 1; }
    const bool use_nstack = nstack && p.nb >= 9 && (3 * p.nt) % 16 == 0 && 3 * p.nt <= 256;
    if (use_nstack) p.nb = p.nb >= 18 ? 18 : 9;     // whole generations (9 tiles) in the ring
    p.tmem_cols = 32;
    while (p.tmem_cols < 2 * kS1Planes * p.nt) p.tmem_cols *= 2;
    p.total_tiles = (long long)p.n_tiles * N * p.dblocks * p.tiles_h * p.tiles_w;
    const int grid = (int)(p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs);
    // statistics need one sample per launch (a CTA's row mixes all its tiles), whole 32-channel chunks,
    // at most 2 x 2 (n tile, chunk) slots and the N-stacked kernel
    const bool stats_ok = use_nstack && N == 1 && p.nt % 32 == 0 && p.nt <= 64 && p.n_tiles <= 2;
    if (stat_rows) *stat_rows = stats_ok ? grid : 0;
    if (addend_ok) *addend_ok = use_nstack ? 1 : 0;
    if (query) return 0;
    if (stat_partial && !stats_ok) { set_error("conv3d(tcgen05,s1): statistics not available for this shape"); return B2_ERR_UNSUPPORTED; }
    if (addend && !use_nstack) { set_error("conv3d(tcgen05,s1): fused addend not available for this shape"); return B2_ERR_UNSUPPORTED; }
    if (ef.stat_mode == 3 && Cout > kEpiMaxC) { set_error("conv3d(tcgen05,s1): stat_mode 3 needs Cout <= %d", kEpiMaxC); return B2_ERR_UNSUPPORTED; }

    CUtensorMap map_a, map_b;
    {
        cuuint64_t gdim[5] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
        cuuint64_t gstr[4] = {(cuuint64_t)Cin * 4, (cuuint64_t)W * Cin * 4, (cuuint64_t)H * W * Cin * 4,
                              (cuuint64_t)D * H * W * Cin * 4};
        if (p.swap) {   // tensor-map dim 2 = rows = D, dim 3 = planes = H
            gdim[2] = (cuuint64_t)D; gdim[3] = (cuuint64_t)H;
            gstr[1] = (cuuint64_t)H * W * Cin * 4; gstr[2] = (cuuint64_t)W * Cin * 4;
        }
        cuuint32_t box[5] = {(cuuint32_t)kKChunk, (cuuint32_t)kTileW, (cuuint32_t)kTileH + 2, 1, 1};
        cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        CUresult r = encode(&map_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)in, gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("conv3d(tcgen05,s1): cuTensorMapEncodeTiled(A) failed: %d", (int)r); return B2_ERR_DRIVER; }
    }
    {
        cuuint64_t gdim[2] = {(cuuint64_t)Cin, (cuuint64_t)27 * Cout};
        cuuint64_t gstr[1] = {(cuuint64_t)Cin * 4};
        cuuint32_t box[2] = {(cuuint32_t)kKChunk, (cuuint32_t)p.nt};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&map_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)wp, gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("conv3d(tcgen05,s1): cuTensorMapEncodeTiled(B) failed: %d", (int)r); return B2_ERR_DRIVER; }
    }
    const int smem = kS1NA * kS1ABytes + p.nb * p.b_bytes + 4 * kEpiStageBytes + 1024;
    {
        static SmemOptIn optin_s1, optin_s1n;
        cudaError_t e = ensure_dynamic_smem(optin_s1, conv3d_s1_tcgen05_kernel, smem);
        if (e == cudaSuccess) e = ensure_dynamic_smem(optin_s1n, conv3d_s1n_tcgen05_kernel, smem);
        if (e != cudaSuccess) { set_error("conv3d(tcgen05,s1): cudaFuncSetAttribute(%d): %s", smem, cudaGetErrorString(e)); return (int)e; }
    }
    if (use_nstack)
        conv3d_s1n_tcgen05_kernel<<<grid, kTcThreads, smem, st>>>(map_a, map_b, out, p, ef);
    else
        conv3d_s1_tcgen05_kernel<<<grid, kTcThreads, smem, st>>>(map_a, map_b, out, p);
    return check_launch("conv3d(tcgen05,s1)");
}

// =============================================================================================
// Transposed-conv (DECONV gather) fast path: "class-stacked" kernel.
// The generic kernel treats the 8 output-parity classes as 8 independent tiles and reloads the A
// box and a weight tile per tap: 96 KB of L2->SM traffic per 2 MFLOP (21 FLOP/B) -> L2-bound at
// ~230 TFLOP/s on the 128->64 layers that write the full-resolution volumes.  But all classes of
// one input block read the SAME input neighbourhood: per axis, input offset 0 serves parity 0
// (tap k=1) and parity 1 (k=2), offset +1 serves parity 1 (k=0).
// Work unit = (input block 16 rows x 8 w, input plane d, output-plane parity pd): its 4
// (row, w)-parity classes c = 2*ph + pw own 4 adjacent TMEM accumulators of Nt columns (double
// buffered = all 512 columns for Nt = 64).  Per (K chunk, plane tap (kd, sd), w shift sw) ONE
// {32 ch, 8 w, 17 rows} halo box is loaded; the row shift sh = 1 is the same smem tile at
// +8 rows (+1024 B, a whole swizzle atom).  Weight tiles of classes that share an input shift
// and sit in adjacent accumulators are stored back to back and fed by ONE MMA:
//   sw = 0:  sh = 0 -> classes 0..3 (kh = ph?2:1, kw = pw?2:1)  N = 4 Nt at column 0
//            sh = 1 -> classes 2,3  (kh = 0,      kw = pw?2:1)  N = 2 Nt at column 2 Nt
//   sw = 1:  sh = 0 -> class 1 (kh = 1, kw = 0), class 3 (kh = 2, kw = 0)      2 x N = Nt
//            sh = 1 -> class 3 (kh = 0, kw = 0)                                N = Nt
// = the 9 (kh, kw) taps of one plane tap.  L2->SM: 2 (or 4) A boxes of 17 KB per K chunk instead
// of 9 (or 18) of 16 KB; smem operand reads per K step drop from 54 KB to 38 KB.
// =============================================================================================
constexpr int kDcARows = (kTileH + 1) * kTileW;        // 136 rows
constexpr int kDcABytes = kDcARows * 128;              // 17408 B = 17 swizzle atoms
constexpr int kDcMaxStages = 6;

struct DcParams {
    int N, Cin, Cout;
    int Pt, Rt, Wt;            // INPUT extents in role order (planes, rows, w)
    int Do, Ho, Wo;            // output dims
    int swap;                  // rows along D, planes along H
    int nt, n_tiles;
    int tiles_w, tiles_h;
    int kchunks;
    int stages, stage_bytes, b_bytes;
    int tmem_cols;
    long long total_units;
};

struct DcUnit { int nti, n, d, pd, h0, w0; };

__device__ __forceinline__ DcUnit dc_decode(const DcParams& p, long long t) {
    DcUnit u;
    u.w0 = (int)(t % p.tiles_w) * kTileW; t /= p.tiles_w;
    u.h0 = (int)(t % p.tiles_h) * kTileH; t /= p.tiles_h;
    u.pd = (int)(t & 1); t >>= 1;
    u.d = (int)(t % p.Pt); t /= p.Pt;
    u.n = (int)(t % p.N); t /= p.N;
    u.nti = (int)t;
    return u;
}

__global__ void __launch_bounds__(kTcThreads, 1)
conv3d_dc_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                         float* __restrict__ out, const DcParams p, const EpiFusion ef) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * kDcMaxStages + 4];
    __shared__ uint32_t tmem_base_slot;
    __shared__ float coef_sh[2 * kEpiMaxC];      // stat_mode 3: forward scale / shift of the preceding norm
    float* __restrict__ stat_partial = ef.stat_partial;
    if (ef.stat_mode == 3)
        for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) {
            coef_sh[i] = ef.gn_coef[i];
            coef_sh[kEpiMaxC + i] = ef.gn_coef[p.Cout + i];
        }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar0 = smem_u32(bars);
    // epilogue staging (4 warps x 2 KB) behind the pipeline stages; reused for the statistics hand-over
    float* const stage_all = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) + p.stages * p.stage_bytes);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (kDcMaxStages + s); };
    auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * kDcMaxStages + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * kDcMaxStages + 2 + a); };

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
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
            for (long long t = blockIdx.x; t < p.total_units; t += gridDim.x) {
                const DcUnit u = dc_decode(p, t);
                const int nptap = 1 + u.pd;
                for (int kc = 0; kc < p.kchunks; ++kc) {
                    for (int jt = 0; jt < nptap; ++jt) {
                        int kd, sd;
                        deconv_axis(u.pd, jt, kd, sd);
                        for (int sw = 0; sw < 2; ++sw) {
                            const int nb = sw == 0 ? 6 : 3;
                            mbar_wait(empty_bar(stage), phase ^ 1);
                            mbar_expect_tx(full_bar(stage), (uint32_t)(kDcABytes + nb * p.b_bytes));
                            const uint32_t sa = smem_base + (uint32_t)stage * p.stage_bytes;
                            tma_load_5d(sa, &map_a, full_bar(stage), kc * kKChunk, u.w0 + sw, u.h0, u.d + sd, u.n);
                            for (int b = 0; b < nb; ++b) {
                                int kh, kw;
                                if (sw == 0) {
                                    if (b < 4) { kh = (b >> 1) ? 2 : 1; kw = (b & 1) ? 2 : 1; }
                                    else       { kh = 0;                kw = (b & 1) ? 2 : 1; }   // b = 4, 5 -> classes 2, 3
                                } else {
                                    kw = 0; kh = b == 0 ? 1 : (b == 1 ? 2 : 0);
                                }
                                const int tap = p.swap ? (kh * 3 + kd) * 3 + kw : (kd * 3 + kh) * 3 + kw;
                                tma_load_2d(sa + kDcABytes + (uint32_t)(b * p.b_bytes), &map_b, full_bar(stage),
                                            kc * kKChunk, tap * p.Cout + u.nti * p.nt);
                            }
                            if (++stage == p.stages) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc1 = umma_idesc_tf32(128, p.nt), idesc2 = umma_idesc_tf32(128, 2 * p.nt),
                           idesc4 = umma_idesc_tf32(128, 4 * p.nt);
            int stage = 0; uint32_t phase = 0;
            long long it = 0;
            for (long long t = blockIdx.x; t < p.total_units; t += gridDim.x, ++it) {
                const int acc = (int)(it & 1);
                mbar_wait(tempty_bar(acc), (uint32_t)(((it >> 1) & 1) ^ 1));
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * 4 * p.nt);
                const DcUnit u = dc_decode(p, t);
                const int nst = p.kchunks * (1 + u.pd);
                for (int s = 0; s < nst; ++s) {
                    // ---- sw = 0 stage: classes 0..3 (sh = 0) and 2,3 (sh = 1)
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    {
                        const uint32_t sa = smem_base + (uint32_t)stage * p.stage_bytes;
                        const uint64_t a0 = umma_desc_sw128(sa), a1 = umma_desc_sw128(sa + 1024);
                        const uint64_t b0 = umma_desc_sw128(sa + kDcABytes),
                                       b4 = umma_desc_sw128(sa + kDcABytes + 4u * p.b_bytes);
#pragma unroll
                        for (int k = 0; k < kKChunk / 8; ++k)
                            umma_tf32(tmem_d, a0 + 2 * k, b0 + 2 * k, idesc4, (s | k) != 0);
#pragma unroll
                        for (int k = 0; k < kKChunk / 8; ++k)
                            umma_tf32(tmem_d + 2 * p.nt, a1 + 2 * k, b4 + 2 * k, idesc2, 1u);
                    }
                    umma_commit(empty_bar(stage));
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    // ---- sw = 1 stage: class 1 and class 3 (sh = 0), class 3 (sh = 1)
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    {
                        const uint32_t sa = smem_base + (uint32_t)stage * p.stage_bytes;
                        const uint64_t a0 = umma_desc_sw128(sa), a1 = umma_desc_sw128(sa + 1024);
                        const uint64_t b0 = umma_desc_sw128(sa + kDcABytes),
                                       b1 = umma_desc_sw128(sa + kDcABytes + (uint32_t)p.b_bytes),
                                       b2 = umma_desc_sw128(sa + kDcABytes + 2u * p.b_bytes);
#pragma unroll
                        for (int k = 0; k < kKChunk / 8; ++k)
                            umma_tf32(tmem_d + p.nt, a0 + 2 * k, b0 + 2 * k, idesc1, 1u);
#pragma unroll
                        for (int k = 0; k < kKChunk / 8; ++k)
                            umma_tf32(tmem_d + 3 * p.nt, a0 + 2 * k, b1 + 2 * k, idesc1, 1u);
#pragma unroll
                        for (int k = 0; k < kKChunk / 8; ++k)
                            umma_tf32(tmem_d + 3 * p.nt, a1 + 2 * k, b2 + 2 * k, idesc1, 1u);
                    }
                    umma_commit(empty_bar(stage));
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                umma_commit(tfull_bar(acc));
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue =====================
        const int lane_grp = warp & 3;
        const int m = lane_grp * 32 + lane;
        const int hl = m / kTileW, wl = m % kTileW;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(lane_grp * 32) << 16);
        const bool stats = stat_partial != nullptr;       // fused GroupNorm statistics of the output
        float4* const stage = reinterpret_cast<float4*>(stage_all + lane_grp * (kEpiStageBytes / 4));
        StatTotals tot;
        tot.clear();
        long long it = 0;
        for (long long t = blockIdx.x; t < p.total_units; t += gridDim.x, ++it) {
            const int acc = (int)(it & 1);
            const DcUnit u = dc_decode(p, t);
            const int h = u.h0 + hl, w = u.w0 + wl;
            const bool ok = h < p.Rt && w < p.Wt;
            const int op = 2 * u.d + u.pd;
            mbar_wait(tfull_bar(acc), (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            const long long unit_off = (long long)u.n * p.Do * p.Ho * p.Wo * p.Cout + u.nti * p.nt;
            // element offset of warp row rw (0..31) of parity class c, or -1 outside the volume
            auto off_of = [&](int rw, int c) -> long long {
                const int m2 = lane_grp * 32 + rw;
                const int h2 = u.h0 + (m2 >> 3), w2 = u.w0 + (m2 & 7);
                if (h2 >= p.Rt || w2 >= p.Wt) return -1;
                const int orow = 2 * h2 + (c >> 1), ow = 2 * w2 + (c & 1);
                const int od = p.swap ? orow : op, oh = p.swap ? op : orow;
                return unit_off + (((long long)od * p.Ho + oh) * p.Wo + ow) * p.Cout;
            };
            int c0 = 0;
            for (; c0 + 32 <= p.nt; c0 += 32) {           // 32-channel chunk outer, the 4 parity classes inner
                float ssum[32], ssq[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) { ssum[i] = 0.f; ssq[i] = 0.f; }
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    const uint32_t taddr = lane_addr + (uint32_t)((acc * 4 + c) * p.nt);
                    uint32_t rr[32];
                    tmem_ld32(taddr + c0, rr);
                    tmem_ld_wait();
                    epi_chunk32(ef, out, rr, stage, lane, ok,
                                [&](int rw) { const long long o = off_of(rw, c); return o < 0 ? o : o + c0; },
                                ssum, ssq, u.nti * p.nt + c0, coef_sh);
                }
                if (stats) {
                    warp_transpose_sum32(ssum, lane);
                    warp_transpose_sum32(ssq, lane);
                    tot.add(u.nti * 2 + (c0 >> 5), ssum[0], ssq[0]);
                }
            }
            if (c0 < p.nt) {                               // 16-channel tail (no statistics)
                for (int c = 0; c < 4; ++c) {
                    const uint32_t taddr = lane_addr + (uint32_t)((acc * 4 + c) * p.nt);
                    uint32_t rr[16];
                    tmem_ld16(taddr + c0, rr);
                    tmem_ld_wait();
                    epi_tail16(ef, out, rr, ok, off_of(lane, c) + c0);
                }
            }
            tc_fence_before();
            mbar_arrive(tempty_bar(acc));
        }
        if (stats) {                                       // [warp][sum slots 0..3, second-sum slots 4..7][lane]
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                stage_all[(lane_grp * 8 + i) * 32 + lane] = tot.s[i];
                stage_all[(lane_grp * 8 + 4 + i) * 32 + lane] = tot.q[i];
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (stat_partial != nullptr && warp == 2) stat_store_row(stage_all, stat_partial, p.Cout, p.nt, p.n_tiles, lane);
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

// =============================================================================================
// Transposed conv, CTA-PAIR variant (cta_group::2).  ncu on the single-CTA kernel above (128 -> 64 at
// 24x48x156): tensor pipe 39.8 %, L2->SM 63 %, and per work unit two thirds of the bytes arriving in smem are
// weight tiles that every unit re-reads -- the kernel is bound by per-SM smem ingress.  Here
// two CTAs of one TPC take two w-adjacent input blocks of the same (plane, parity) and issue ONE
// tcgen05.mma.cta_group::2 of M = 256 per step: each CTA brings its own A box (its 128 pixels) and HALF of
// every weight operand's rows, the tensor cores read the other half from the peer's smem.  Weight bytes
// arriving per SM halve (per K chunk and plane tap: 17 + 24 and 17 + 12 KB instead of 17 + 48 and 17 + 24),
// the accumulator layout in each CTA's TMEM is unchanged (its 128 rows x the 4 class column ranges), so the
// epilogue (incl. the fused statistics / addend) is the single-CTA one.
//   weight operand split (rows of B = N of the MMA): CTA r holds rows [r N/2, (r+1) N/2):
//     sw = 0: N = 4 nt  [cls 0..3] -> CTA r loads the tiles of classes 2r, 2r+1;  N = 2 nt [cls 2,3; sh = 1] -> tile 4+r
//     sw = 1: three MMAs of N = nt -> CTA r loads rows [r nt/2, (r+1) nt/2) of each of the 3 tiles
//   barriers: "full" lives in the leader (rank 0) and counts the leader's expect_tx arrival plus one remote
//   arrival of the peer's producer; both CTAs' TMA loads signal it (cta_group::2 loads).  "empty" / "tmem full"
//   exist in both CTAs and are signalled by multicast tcgen05.commit; "tmem empty" lives in the leader and
//   collects the 2 x 128 epilogue threads of the pair.
// =============================================================================================
constexpr int kDc2MaxStages = 6;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
conv3d_dc2_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                          float* __restrict__ out, const DcParams p, const EpiFusion ef) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * kDc2MaxStages + 4];
    __shared__ uint32_t tmem_base_slot;
    __shared__ float coef_sh[2 * kEpiMaxC];
    float* __restrict__ stat_partial = ef.stat_partial;
    if (ef.stat_mode == 3)
        for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) {
            coef_sh[i] = ef.gn_coef[i];
            coef_sh[kEpiMaxC + i] = ef.gn_coef[p.Cout + i];
        }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar0 = smem_u32(bars);
    float* const stage_all = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) + p.stages * p.stage_bytes);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (kDc2MaxStages + s); };
    auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * kDc2MaxStages + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * kDc2MaxStages + 2 + a); };

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 256); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&tmem_base_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                       // both CTAs' barriers initialised and TMEM allocated
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    const int wpairs = (p.tiles_w + 1) / 2;
    const long long total_pairs = (long long)p.n_tiles * p.N * p.Pt * 2 * p.tiles_h * wpairs;
    const long long pair0 = blockIdx.x >> 1, pstride = gridDim.x >> 1;
    // pair-unit index -> this CTA's unit (w tile 2 * wpair + rank; may lie beyond the volume: loads are zero-filled)
    auto decode = [&](long long t) {
        DcUnit u;
        u.w0 = ((int)(t % wpairs) * 2 + (int)rank) * kTileW; t /= wpairs;
        u.h0 = (int)(t % p.tiles_h) * kTileH; t /= p.tiles_h;
        u.pd = (int)(t & 1); t >>= 1;
        u.d = (int)(t % p.Pt); t /= p.Pt;
        u.n = (int)(t % p.N); t /= p.N;
        u.nti = (int)t;
        return u;
    };
    const uint32_t half_b = (uint32_t)p.b_bytes / 2;          // bytes of half a weight tile (nt / 2 rows)

    if (warp == 0) {
        // ===================== TMA producer (one thread per CTA) =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (long long t = pair0; t < total_pairs; t += pstride) {
                const DcUnit u = decode(t);
                const int nptap = 1 + u.pd;
                for (int kc = 0; kc < p.kchunks; ++kc) {
                    for (int jt = 0; jt < nptap; ++jt) {
                        int kd, sd;
                        deconv_axis(u.pd, jt, kd, sd);
                        for (int sw = 0; sw < 2; ++sw) {
                            mbar_wait(empty_bar(stage), phase ^ 1);
                            const uint32_t fb = mapa_cluster(full_bar(stage), 0);          // the leader's barrier
                            // bytes per CTA: A box + 3 tiles (sw = 0) or 3 half tiles (sw = 1)
                            const uint32_t mine = (uint32_t)kDcABytes + (sw == 0 ? 3u * (uint32_t)p.b_bytes : 3u * half_b);
                            if (leader) mbar_expect_tx(full_bar(stage), 2u * mine);
                            const uint32_t sa = smem_base + (uint32_t)stage * p.stage_bytes;
                            tma_load_5d_pair(sa, &map_a, fb, kc * kKChunk, u.w0 + sw, u.h0, u.d + sd, u.n);
                            const uint32_t sb = sa + kDcABytes;
                            auto tap_of = [&](int kh, int kw) { return p.swap ? (kh * 3 + kd) * 3 + kw : (kd * 3 + kh) * 3 + kw; };
                            auto half_tile = [&](uint32_t dst, int tap, int half) {      // rows [half nt/2, +nt/2) of a tile
                                tma_load_2d_pair(dst, &map_b, fb, kc * kKChunk, tap * p.Cout + u.nti * p.nt + half * (p.nt / 2));
                            };
                            if (sw == 0) {
                                // N = 4 nt operand: classes 2r, 2r+1 (class c: kh = c>>1 ? 2 : 1, kw = c&1 ? 2 : 1)
                                for (int j = 0; j < 2; ++j) {
                                    const int c = 2 * (int)rank + j;
                                    const int tap = tap_of((c >> 1) ? 2 : 1, (c & 1) ? 2 : 1);
                                    half_tile(sb + (uint32_t)(2 * j) * half_b, tap, 0);
                                    half_tile(sb + (uint32_t)(2 * j + 1) * half_b, tap, 1);
                                }
                                // N = 2 nt operand (sh = 1, kh = 0): classes 2, 3 -> tile of class 2 + r
                                const int tap = tap_of(0, rank ? 2 : 1);
                                half_tile(sb + 4u * half_b, tap, 0);
                                half_tile(sb + 5u * half_b, tap, 1);
                            } else {
                                // three N = nt operands (kw = 0; kh = 1, 2, 0): this CTA's half of the rows of each
                                for (int b = 0; b < 3; ++b) {
                                    const int kh = b == 0 ? 1 : (b == 1 ? 2 : 0);
                                    half_tile(sb + (uint32_t)b * half_b, tap_of(kh, 0), (int)rank);
                                }
                            }
                            if (!leader) mbar_arrive_cluster_relaxed(fb);
                            if (++stage == p.stages) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (lane == 0 && leader) {
            const uint32_t idesc1 = umma_idesc_tf32(256, p.nt), idesc2 = umma_idesc_tf32(256, 2 * p.nt),
                           idesc4 = umma_idesc_tf32(256, 4 * p.nt);
            int stage = 0; uint32_t phase = 0;
            long long it = 0;
            for (long long t = pair0; t < total_pairs; t += pstride, ++it) {
                const int acc = (int)(it & 1);
                mbar_wait(tempty_bar(acc), (uint32_t)(((it >> 1) & 1) ^ 1));
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * 4 * p.nt);
                const DcUnit u = decode(t);
                const int nst = p.kchunks * (1 + u.pd);
                for (int s = 0; s < nst; ++s) {
                    // ---- sw = 0 stage
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    {
                        const uint32_t sa = smem_base + (uint32_t)stage * p.stage_bytes;
                        const uint64_t a0 = umma_desc_sw128(sa), a1 = umma_desc_sw128(sa + 1024);
                        const uint64_t b0 = umma_desc_sw128(sa + kDcABytes), b4 = umma_desc_sw128(sa + kDcABytes + 4u * half_b);
#pragma unroll
                        for (int k = 0; k < kKChunk / 8; ++k)
                            umma_tf32_pair(tmem_d, a0 + 2 * k, b0 + 2 * k, idesc4, (s | k) != 0);
#pragma unroll
                        for (int k = 0; k < kKChunk / 8; ++k)
                            umma_tf32_pair(tmem_d + 2 * p.nt, a1 + 2 * k, b4 + 2 * k, idesc2, 1u);
                    }
                    umma_commit_pair(empty_bar(stage));
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    // ---- sw = 1 stage
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    {
                        const uint32_t sa = smem_base + (uint32_t)stage * p.stage_bytes;
                        const uint64_t a0 = umma_desc_sw128(sa), a1 = umma_desc_sw128(sa + 1024);
                        const uint64_t b0 = umma_desc_sw128(sa + kDcABytes), b1 = umma_desc_sw128(sa + kDcABytes + half_b),
                                       b2 = umma_desc_sw128(sa + kDcABytes + 2u * half_b);
#pragma unroll
                        for (int k = 0; k < kKChunk / 8; ++k)
                            umma_tf32_pair(tmem_d + p.nt, a0 + 2 * k, b0 + 2 * k, idesc1, 1u);
#pragma unroll
                        for (int k = 0; k < kKChunk / 8; ++k)
                            umma_tf32_pair(tmem_d + 3 * p.nt, a0 + 2 * k, b1 + 2 * k, idesc1, 1u);
#pragma unroll
                        for (int k = 0; k < kKChunk / 8; ++k)
                            umma_tf32_pair(tmem_d + 3 * p.nt, a1 + 2 * k, b2 + 2 * k, idesc1, 1u);
                    }
                    umma_commit_pair(empty_bar(stage));
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                umma_commit_pair(tfull_bar(acc));
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue (both CTAs: their own 128 rows) =====================
        const int lane_grp = warp & 3;
        const int m = lane_grp * 32 + lane;
        const int hl = m / kTileW, wl = m % kTileW;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(lane_grp * 32) << 16);
        const bool stats = stat_partial != nullptr;
        float4* const stage = reinterpret_cast<float4*>(stage_all + lane_grp * (kEpiStageBytes / 4));
        StatTotals tot;
        tot.clear();
        long long it = 0;
        for (long long t = pair0; t < total_pairs; t += pstride, ++it) {
            const int acc = (int)(it & 1);
            const DcUnit u = decode(t);
            const int h = u.h0 + hl, w = u.w0 + wl;
            const bool ok = h < p.Rt && w < p.Wt;
            const int op = 2 * u.d + u.pd;
            mbar_wait(tfull_bar(acc), (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            const long long unit_off = (long long)u.n * p.Do * p.Ho * p.Wo * p.Cout + u.nti * p.nt;
            auto off_of = [&](int rw, int c) -> long long {
                const int m2 = lane_grp * 32 + rw;
                const int h2 = u.h0 + (m2 >> 3), w2 = u.w0 + (m2 & 7);
                if (h2 >= p.Rt || w2 >= p.Wt) return -1;
                const int orow = 2 * h2 + (c >> 1), ow = 2 * w2 + (c & 1);
                const int od = p.swap ? orow : op, oh = p.swap ? op : orow;
                return unit_off + (((long long)od * p.Ho + oh) * p.Wo + ow) * p.Cout;
            };
            int c0 = 0;
            for (; c0 + 32 <= p.nt; c0 += 32) {
                float ssum[32], ssq[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) { ssum[i] = 0.f; ssq[i] = 0.f; }
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    const uint32_t taddr = lane_addr + (uint32_t)((acc * 4 + c) * p.nt);
                    uint32_t rr[32];
                    tmem_ld32(taddr + c0, rr);
                    tmem_ld_wait();
                    epi_chunk32(ef, out, rr, stage, lane, ok,
                                [&](int rw) { const long long o = off_of(rw, c); return o < 0 ? o : o + c0; },
                                ssum, ssq, u.nti * p.nt + c0, coef_sh);
                }
                if (stats) {
                    warp_transpose_sum32(ssum, lane);
                    warp_transpose_sum32(ssq, lane);
                    tot.add(u.nti * 2 + (c0 >> 5), ssum[0], ssq[0]);
                }
            }
            if (c0 < p.nt) {
                for (int c = 0; c < 4; ++c) {
                    const uint32_t taddr = lane_addr + (uint32_t)((acc * 4 + c) * p.nt);
                    uint32_t rr[16];
                    tmem_ld16(taddr + c0, rr);
                    tmem_ld_wait();
                    epi_tail16(ef, out, rr, ok, off_of(lane, c) + c0);
                }
            }
            tc_fence_before();
            mbar_arrive_cluster_relaxed(mapa_cluster(tempty_bar(acc), 0));     // the leader's MMA thread waits for all 256
        }
        if (stats) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                stage_all[(lane_grp * 8 + i) * 32 + lane] = tot.s[i];
                stage_all[(lane_grp * 8 + 4 + i) * 32 + lane] = tot.q[i];
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (stat_partial != nullptr && warp == 2) stat_store_row(stage_all, stat_partial, p.Cout, p.nt, p.n_tiles, lane);
    cluster_sync_all();                       // no CTA may leave (or free TMEM) while its peer can still signal / read it
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

static int conv3d_dc_launch(EncodeTiledFn encode, const float* in, const float* wp, float* out, int N, int Cin,
                            int Cout, int Di, int Hi, int Wi, cudaStream_t st, const EpiFusion& ef, int* stat_rows,
                            bool query, int* addend_ok) {
    float* stat_partial = ef.stat_partial;
    DcParams p{};
    p.N = N; p.Cin = Cin; p.Cout = Cout;
    p.Do = 2 * Di; p.Ho = 2 * Hi; p.Wo = 2 * Wi;
    p.nt = Cout <= 64 ? Cout : Cout / 2;
    p.n_tiles = Cout / p.nt;
    {
        const long long cost_h = (long long)((Hi + kTileH - 1) / kTileH) * kTileH * Di;
        const long long cost_d = (long long)((Di + kTileH - 1) / kTileH) * kTileH * Hi;
        p.swap = cost_d < cost_h ? 1 : 0;
        p.Rt = p.swap ? Di : Hi;
        p.Pt = p.swap ? Hi : Di;
        p.Wt = Wi;
    }
    p.tiles_w = (p.Wt + kTileW - 1) / kTileW;
    p.tiles_h = (p.Rt + kTileH - 1) / kTileH;
    p.kchunks = Cin / kKChunk;
    p.b_bytes = p.nt * 128;
    // CTA-pair variant (cta_group::2, default): each CTA stages its A box and HALF of the weight rows (3 tiles' worth).
    // Measured (tools/bench_conv3d_shapes.py, graph replay): 0.1806 vs 0.1975 ms on 128 -> 64 at 24x48x156, 0.0634 vs
    // 0.0687 ms on 128 -> 128 at 12x24x78.  (The first version, whose cross-CTA mbarrier arrivals had cluster-scope RELEASE
    // semantics, was 1.5x SLOWER than the single-CTA kernel: every epilogue thread then waits for its global stores to
    // reach L2 before it may hand the accumulator back -- profiles/r2_conv3d_dc_pair_ncu.txt.)
    // B2_CONV_DC_PAIR=0 / b2_set_flag("conv_dc_pair", 0) selects the single-CTA kernel.
    const bool pair = flag_value(kFlagConvDcPair, "B2_CONV_DC_PAIR", 1) && p.nt % 16 == 0;
    p.stage_bytes = kDcABytes + (pair ? 3 : 6) * p.b_bytes;
    p.stages = (212 * 1024) / p.stage_bytes;
    if (p.stages > kDcMaxStages) p.stages = kDcMaxStages;
    p.tmem_cols = 32;
    while (p.tmem_cols < 8 * p.nt) p.tmem_cols *= 2;
    p.total_units = (long long)p.n_tiles * N * p.Pt * 2 * p.tiles_h * p.tiles_w;
    int grid = (int)(p.total_units < kNumSMs ? p.total_units : kNumSMs);
    if (pair) {
        const long long pairs = (long long)p.n_tiles * N * p.Pt * 2 * p.tiles_h * ((p.tiles_w + 1) / 2);
        grid = (int)(2 * (pairs < kNumSMs / 2 ? pairs : kNumSMs / 2));
    }
    const bool stats_ok = N == 1 && p.nt % 32 == 0 && p.nt <= 64 && p.n_tiles <= 2;
    if (stat_rows) *stat_rows = stats_ok ? grid : 0;
    if (addend_ok) *addend_ok = 1;
    if (query) return 0;
    if (stat_partial && !stats_ok) { set_error("conv3d(tcgen05,deconv): statistics not available for this shape"); return B2_ERR_UNSUPPORTED; }
    if (ef.stat_mode == 3 && Cout > kEpiMaxC) { set_error("conv3d(tcgen05,deconv): stat_mode 3 needs Cout <= %d", kEpiMaxC); return B2_ERR_UNSUPPORTED; }

    CUtensorMap map_a, map_b;
    {
        cuuint64_t gdim[5] = {(cuuint64_t)Cin, (cuuint64_t)Wi, (cuuint64_t)Hi, (cuuint64_t)Di, (cuuint64_t)N};
        cuuint64_t gstr[4] = {(cuuint64_t)Cin * 4, (cuuint64_t)Wi * Cin * 4, (cuuint64_t)Hi * Wi * Cin * 4,
                              (cuuint64_t)Di * Hi * Wi * Cin * 4};
        if (p.swap) {
            gdim[2] = (cuuint64_t)Di; gdim[3] = (cuuint64_t)Hi;
            gstr[1] = (cuuint64_t)Hi * Wi * Cin * 4; gstr[2] = (cuuint64_t)Wi * Cin * 4;
        }
        cuuint32_t box[5] = {(cuuint32_t)kKChunk, (cuuint32_t)kTileW, (cuuint32_t)kTileH + 1, 1, 1};
        cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        CUresult r = encode(&map_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)in, gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("conv3d(tcgen05,deconv): cuTensorMapEncodeTiled(A) failed: %d", (int)r); return B2_ERR_DRIVER; }
    }
    {
        cuuint64_t gdim[2] = {(cuuint64_t)Cin, (cuuint64_t)27 * Cout};
        cuuint64_t gstr[1] = {(cuuint64_t)Cin * 4};
        cuuint32_t box[2] = {(cuuint32_t)kKChunk, (cuuint32_t)(pair ? p.nt / 2 : p.nt)};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&map_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)wp, gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("conv3d(tcgen05,deconv): cuTensorMapEncodeTiled(B) failed: %d", (int)r); return B2_ERR_DRIVER; }
    }
    const int smem = p.stages * p.stage_bytes + 4 * kEpiStageBytes + 1024;
    {
        static SmemOptIn optin;
        cudaError_t e = ensure_dynamic_smem(optin, conv3d_dc_tcgen05_kernel, smem);
        if (e != cudaSuccess) { set_error("conv3d(tcgen05,deconv): cudaFuncSetAttribute(%d): %s", smem, cudaGetErrorString(e)); return (int)e; }
    }
    if (pair) {
        static SmemOptIn optin2;
        cudaError_t e2 = ensure_dynamic_smem(optin2, conv3d_dc2_tcgen05_kernel, smem);
        if (e2 != cudaSuccess) { set_error("conv3d(tcgen05,deconv,pair): cudaFuncSetAttribute(%d): %s", smem, cudaGetErrorString(e2)); return (int)e2; }
        conv3d_dc2_tcgen05_kernel<<<grid, kTcThreads, smem, st>>>(map_a, map_b, out, p, ef);
        return check_launch("conv3d(tcgen05,deconv,pair)");
    }
    conv3d_dc_tcgen05_kernel<<<grid, kTcThreads, smem, st>>>(map_a, map_b, out, p, ef);
    return check_launch("conv3d(tcgen05,deconv)");
}

// =============================================================================================
// Generic (one tap per stage) kernel on a CTA PAIR, for the stride-2 convs (and the stride-2-conv-shaped data
// gradients of the transposed convs): two CTAs of one TPC take two w-adjacent output tiles of the same
// (sample, plane, row block); per (tap, K chunk) stage each stages its own A box (16 KB) and HALF of the weight
// tile's rows, the leader issues tcgen05.mma.cta_group::2 of M = 256, N = Cout.  Bytes arriving per SM per stage:
// 16 + Cout*64 B instead of 16 + Cout*128 B (24 vs 32 KB at Cout = 128).  Barrier scheme as in
// conv3d_dc2_tcgen05_kernel (leader-side "full" and "accumulator free", multicast commits, relaxed cross-CTA arrivals).
// =============================================================================================
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
conv3d_g2_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                         float* __restrict__ out, const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 4];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (kMaxStages + s); };
    auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * kMaxStages + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * kMaxStages + 2 + a); };

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 256); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&tmem_base_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    const int wpairs = (p.tiles_w + 1) / 2;
    const long long total_pairs = (long long)p.N * p.Dt * p.tiles_h * wpairs;
    const long long pair0 = blockIdx.x >> 1, pstride = gridDim.x >> 1;
    auto decode = [&](long long t) {
        TileCoord c;
        c.w0 = ((int)(t % wpairs) * 2 + (int)rank) * kTileW; t /= wpairs;
        c.h0 = (int)(t % p.tiles_h) * kTileH; t /= p.tiles_h;
        c.d = (int)(t % p.Dt); t /= p.Dt;
        c.n = (int)t;
        c.cls = 0;
        return c;
    };
    const uint32_t half_b = (uint32_t)(p.Cout / 2) * 128u;
    const int s = p.mode == 1 ? 2 : 1;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            const uint32_t mine = (uint32_t)kABytes + half_b;
            for (long long t = pair0; t < total_pairs; t += pstride) {
                const TileCoord tc = decode(t);
                for (int tap_i = 0; tap_i < 27; ++tap_i) {
                    const int kd = tap_i / 9, kh = (tap_i / 3) % 3, kw = tap_i % 3;
                    const int aw = s * tc.w0 + kw - 1, ah = s * tc.h0 + kh - 1, ad = s * tc.d + kd - 1;
                    const int tap = p.swap ? (kh * 3 + kd) * 3 + kw : (kd * 3 + kh) * 3 + kw;
                    for (int kc = 0; kc < p.kchunks; ++kc) {
                        mbar_wait(empty_bar(stage), phase ^ 1);
                        const uint32_t fb = mapa_cluster(full_bar(stage), 0);
                        if (leader) mbar_expect_tx(full_bar(stage), 2u * mine);
                        const uint32_t sa = smem_base + (uint32_t)stage * p.stage_bytes;
                        tma_load_5d_pair(sa, &map_a, fb, kc * kKChunk, aw, ah, ad, tc.n);
                        tma_load_2d_pair(sa + kABytes, &map_b, fb, kc * kKChunk, tap * p.Cout + (int)rank * (p.Cout / 2));
                        if (!leader) mbar_arrive_cluster_relaxed(fb);
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            const uint32_t idesc = umma_idesc_tf32(256, p.Cout);
            int stage = 0; uint32_t phase = 0;
            long long it = 0;
            for (long long t = pair0; t < total_pairs; t += pstride, ++it) {
                const int acc = (int)(it & 1);
                mbar_wait(tempty_bar(acc), (uint32_t)(((it >> 1) & 1) ^ 1));
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * p.Cout);
                const int nsteps = 27 * p.kchunks;
                for (int st = 0; st < nsteps; ++st) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + (uint32_t)stage * p.stage_bytes;
                    const uint64_t adesc = umma_desc_sw128(sa), bdesc = umma_desc_sw128(sa + kABytes);
#pragma unroll
                    for (int k = 0; k < kKChunk / 8; ++k)
                        umma_tf32_pair(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (st | k) != 0);
                    umma_commit_pair(empty_bar(stage));
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                umma_commit_pair(tfull_bar(acc));
            }
        }
        __syncwarp();
    } else {
        const int lane_grp = warp & 3;
        const int m = lane_grp * 32 + lane;
        const int hl = m / kTileW, wl = m % kTileW;
        long long it = 0;
        for (long long t = pair0; t < total_pairs; t += pstride, ++it) {
            const int acc = (int)(it & 1);
            const TileCoord tc = decode(t);
            const int h = tc.h0 + hl, w = tc.w0 + wl;
            const bool ok = h < p.Ht && w < p.Wt;
            const int od = p.swap ? h : tc.d, oh = p.swap ? tc.d : h;
            float* optr = out + ((((long long)tc.n * p.Do + od) * p.Ho + oh) * p.Wo + w) * p.Cout;
            mbar_wait(tfull_bar(acc), (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(acc * p.Cout);
            for (int c0 = 0; c0 < p.Cout; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(taddr + c0, r);
                tmem_ld_wait();
                if (ok) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4*>(optr + c0 + 4 * j) =
                            make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                        __uint_as_float(r[4 * j + 3]));
                }
            }
            tc_fence_before();
            mbar_arrive_cluster_relaxed(mapa_cluster(tempty_bar(acc), 0));
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

int conv3d_tcgen05_launch(const float* in, const float* wp, float* out, int N, int Cin, int Cout, int Di,
                          int Hi, int Wi, int Do, int Ho, int Wo, int stride, int mode, cudaStream_t st,
                          const EpiFusion& ef, int* stat_rows, bool query, int* addend_ok) {
    const float* addend = ef.addend;
    float* stat_partial = ef.stat_partial;
    if (stat_rows) *stat_rows = 0;
    if (addend_ok) *addend_ok = 0;
    if (Cin % 32 != 0 || Cout % 32 != 0 || Cout > 256 || Cout < 32) {
        if (query) return 0;
        set_error("conv3d(tcgen05): needs Cin %% 32 == 0 and Cout in {32,64,...,256} (got %d -> %d); "
                  "use impl=1 for other widths", Cin, Cout);
        return B2_ERR_UNSUPPORTED;
    }
    EncodeTiledFn encode = get_encode();
    if (!encode && !query) { set_error("conv3d(tcgen05): cuTensorMapEncodeTiled not available from the driver"); return B2_ERR_DRIVER; }
    {
        // stride-1 convs (74 % of the flops) take the halo-reuse super-tile kernel; B2_CONV_S1_SIMPLE=1
        // forces the generic one-tap-per-stage kernel (A/B testing)
        static int simple = -1;
        if (simple < 0) { const char* e = getenv("B2_CONV_S1_SIMPLE"); simple = (e && e[0] == '1') ? 1 : 0; }
        const int nt = Cout <= 64 ? Cout : Cout / 2;
        if (mode == 0 && stride == 1 && !simple && nt % 16 == 0)
            return conv3d_s1_launch(encode, in, wp, out, N, Cin, Cout, Di, Hi, Wi, st, ef, stat_rows, query, addend_ok);
        // transposed convs take the class-stacked kernel; B2_CONV_DC_SIMPLE=1 forces the generic one
        static int dc_simple = -1;
        if (dc_simple < 0) { const char* e = getenv("B2_CONV_DC_SIMPLE"); dc_simple = (e && e[0] == '1') ? 1 : 0; }
        if (mode == 1 && !dc_simple && nt % 16 == 0 && 4 * nt <= 256)
            return conv3d_dc_launch(encode, in, wp, out, N, Cin, Cout, Di, Hi, Wi, st, ef, stat_rows, query, addend_ok);
    }

    if (query) return 0;                                   // the generic kernel has no statistics epilogue
    if (stat_partial || addend) { set_error("conv3d(tcgen05): fused statistics / addend not available for this shape"); return B2_ERR_UNSUPPORTED; }

    TcParams p{};
    p.N = N; p.Cin = Cin; p.Cout = Cout; p.Do = Do; p.Ho = Ho; p.Wo = Wo;
    p.mode = (mode == 1) ? 2 : (stride == 2 ? 1 : 0);
    {
        // tile rows (16) run along H or D, whichever wastes fewer padded voxels (e.g. the 192x20x304
        // voxel grid tiles perfectly with rows along Z=192 but wastes 37 % with rows along Y=20)
        const int td = (p.mode == 2) ? Di : Do, th = (p.mode == 2) ? Hi : Ho;
        const long long cost_h = (long long)((th + kTileH - 1) / kTileH) * kTileH * td;
        const long long cost_d = (long long)((td + kTileH - 1) / kTileH) * kTileH * th;
        p.swap = cost_d < cost_h ? 1 : 0;
        p.Wt = (p.mode == 2) ? Wi : Wo;
        p.Dt = p.swap ? th : td;
        p.Ht = p.swap ? td : th;
    }
    p.tiles_w = (p.Wt + kTileW - 1) / kTileW;
    p.tiles_h = (p.Ht + kTileH - 1) / kTileH;
    p.kchunks = Cin / kKChunk;
    // stride-1 / stride-2 CONV launches that land here run on a CTA pair (see conv3d_g2_tcgen05_kernel)
    const bool gpair = p.mode != 2 && Cout % 64 == 0 && flag_value(kFlagConvG2Pair, "B2_CONV_G2_PAIR", 1);
    p.stage_bytes = kABytes + (gpair ? Cout * 64 : Cout * 128);
    p.stages = (200 * 1024) / p.stage_bytes;
    if (p.stages > kMaxStages) p.stages = kMaxStages;
    p.tmem_cols = 32;
    while (p.tmem_cols < 2 * Cout) p.tmem_cols *= 2;
    p.total_tiles = (long long)(p.mode == 2 ? 8 : 1) * N * p.Dt * p.tiles_h * p.tiles_w;

    CUtensorMap map_a, map_b;
    {
        cuuint64_t gdim[5] = {(cuuint64_t)Cin, (cuuint64_t)Wi, (cuuint64_t)Hi, (cuuint64_t)Di, (cuuint64_t)N};
        cuuint64_t gstr[4] = {(cuuint64_t)Cin * 4, (cuuint64_t)Wi * Cin * 4, (cuuint64_t)Hi * Wi * Cin * 4,
                              (cuuint64_t)Di * Hi * Wi * Cin * 4};
        if (p.swap) {   // tensor-map dim 2 = rows = D, dim 3 = planes = H
            gdim[2] = (cuuint64_t)Di; gdim[3] = (cuuint64_t)Hi;
            gstr[1] = (cuuint64_t)Hi * Wi * Cin * 4; gstr[2] = (cuuint64_t)Wi * Cin * 4;
        }
        const cuuint32_t s = (p.mode == 1) ? 2 : 1;
        // traversal box; with element stride s the box lands ceil(box/s) elements per dim in smem
        cuuint32_t box[5] = {(cuuint32_t)kKChunk, (cuuint32_t)kTileW * s, (cuuint32_t)kTileH * s, s, 1};
        cuuint32_t estr[5] = {1, s, s, s, 1};
        CUresult r = encode(&map_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)in, gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("conv3d(tcgen05): cuTensorMapEncodeTiled(A) failed: %d", (int)r); return B2_ERR_DRIVER; }
    }
    {
        cuuint64_t gdim[2] = {(cuuint64_t)Cin, (cuuint64_t)27 * Cout};
        cuuint64_t gstr[1] = {(cuuint64_t)Cin * 4};
        cuuint32_t box[2] = {(cuuint32_t)kKChunk, (cuuint32_t)(gpair ? Cout / 2 : Cout)};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&map_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)wp, gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("conv3d(tcgen05): cuTensorMapEncodeTiled(B) failed: %d", (int)r); return B2_ERR_DRIVER; }
    }

    const int smem = p.stages * p.stage_bytes + 1024;
    {
        static SmemOptIn optin;
        cudaError_t e = ensure_dynamic_smem(optin, conv3d_tcgen05_kernel, smem);
        if (e != cudaSuccess) { set_error("conv3d(tcgen05): cudaFuncSetAttribute(%d): %s", smem, cudaGetErrorString(e)); return (int)e; }
    }
    int grid = (int)(p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs);
    if (gpair) {
        const long long pairs = (long long)N * p.Dt * p.tiles_h * ((p.tiles_w + 1) / 2);
        grid = (int)(2 * (pairs < kNumSMs / 2 ? pairs : kNumSMs / 2));
        static SmemOptIn optin2;
        cudaError_t e2 = ensure_dynamic_smem(optin2, conv3d_g2_tcgen05_kernel, smem);
        if (e2 != cudaSuccess) { set_error("conv3d(tcgen05,pair): cudaFuncSetAttribute(%d): %s", smem, cudaGetErrorString(e2)); return (int)e2; }
        conv3d_g2_tcgen05_kernel<<<grid, kTcThreads, smem, st>>>(map_a, map_b, out, p);
        return check_launch("conv3d(tcgen05,pair)");
    }
    conv3d_tcgen05_kernel<<<grid, kTcThreads, smem, st>>>(map_a, map_b, out, p);
    return check_launch("conv3d(tcgen05)");
}

}  // namespace b2
