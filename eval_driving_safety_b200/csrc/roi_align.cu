// RoIAlign forward and DETERMINISTIC (gather-form) backward for the Stereo R-CNN
// attack (BASELINE config 5).  Replaces upstream model.roi_layers.ROIAlign
// (maskrcnn-benchmark ROIAlign_cuda.cu, atomicAdd backward) constructed at
// attack/Stereo-RCNN/stereo_rcnn.py:44-45 and called per FPN level at :110-141.
// Legacy (unaligned) coordinates, adaptive sampling ratio (sampling_ratio = 0).
#include "common.cuh"

namespace b2 {

struct RoiGeom {
    float start_w, start_h, bin_w, bin_h;
    int grid_w, grid_h;
};

__device__ __forceinline__ RoiGeom roi_geom(const float* roi, float scale, int P) {
    RoiGeom r;
    r.start_w = roi[1] * scale;
    r.start_h = roi[2] * scale;
    float end_w = roi[3] * scale, end_h = roi[4] * scale;
    float rw = fmaxf(end_w - r.start_w, 1.f), rh = fmaxf(end_h - r.start_h, 1.f);
    r.bin_w = rw / (float)P;
    r.bin_h = rh / (float)P;
    r.grid_w = (int)ceilf(rw / (float)P);
    r.grid_h = (int)ceilf(rh / (float)P);
    return r;
}

// 1-D half of bilinear_interpolate(): returns false when the sample is dropped.
__device__ __forceinline__ bool axis_interp(float p, int size, int& lo, int& hi, float& wlo, float& whi) {
    if (p < -1.0f || p > (float)size) return false;
    if (p <= 0.f) p = 0.f;
    lo = (int)p;
    if (lo >= size - 1) { hi = lo = size - 1; p = (float)lo; } else { hi = lo + 1; }
    whi = p - (float)lo;
    wlo = 1.f - whi;
    return true;
}

// FPN level of a RoI, attack/Stereo-RCNN/stereo_rcnn.py:113-119: round(log(sqrt(h w) / 224) + 4) clamped to
// [2, 5] with h = y2 - y1 + 1, w = x2 - x1 + 1 -- NATURAL log and round-half-to-even, as torch.log / torch.round.
__device__ __forceinline__ int roi_fpn_level(const float* roi) {
    const float h = __fadd_rn(__fsub_rn(roi[4], roi[2]), 1.f), w = __fadd_rn(__fsub_rn(roi[3], roi[1]), 1.f);
    float l = rintf(__fadd_rn(logf(__fdiv_rn(sqrtf(__fmul_rn(h, w)), 224.0f)), 4.f));
    l = fminf(fmaxf(l, 2.f), 5.f);
    return (int)l;
}

// The four pyramid levels of one view (stereo_rcnn.py:110-141): level l = 2..5 lives at index l - 2.
struct PyrLevels {
    const float* feat[4];
    float* gfeat[4];
    int H[4], W[4];
    float scale[4];
    long long pix_off[5];      // backward: prefix sums of H*W over the levels
};

template <bool PYR>
__global__ void __launch_bounds__(256)
roi_align_fwd_kernel(const float* __restrict__ feat, const float* __restrict__ rois,
                     float* __restrict__ out, int R, int C, int H, int W, int P, float scale, const PyrLevels lv) {
    const int64_t total = (int64_t)R * C * P * P;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int pw = (int)(i % P), ph = (int)((i / P) % P);
        int c = (int)((i / ((int64_t)P * P)) % C), r = (int)(i / ((int64_t)P * P * C));
        const float* roi = rois + r * 5;
        if (PYR) {                                   // the RoI's own level: one launch instead of one per level
            const int l = roi_fpn_level(roi) - 2;
            feat = lv.feat[l]; H = lv.H[l]; W = lv.W[l]; scale = lv.scale[l];
        }
        RoiGeom g = roi_geom(roi, scale, P);
        int b = (int)roi[0];
        const float* f = feat + ((int64_t)b * C + c) * H * W;
        float acc = 0.f;
        for (int iy = 0; iy < g.grid_h; ++iy) {
            float y = g.start_h + ph * g.bin_h + ((float)iy + .5f) * g.bin_h / (float)g.grid_h;
            int yl, yh; float wyl, wyh;
            bool oky = axis_interp(y, H, yl, yh, wyl, wyh);
            for (int ix = 0; ix < g.grid_w; ++ix) {
                float x = g.start_w + pw * g.bin_w + ((float)ix + .5f) * g.bin_w / (float)g.grid_w;
                int xl, xh; float wxl, wxh;
                if (!oky || !axis_interp(x, W, xl, xh, wxl, wxh)) continue;
                float v1 = __ldg(f + yl * W + xl), v2 = __ldg(f + yl * W + xh);
                float v3 = __ldg(f + yh * W + xl), v4 = __ldg(f + yh * W + xh);
                acc += wyl * wxl * v1 + wyl * wxh * v2 + wyh * wxl * v3 + wyh * wxh * v4;
            }
        }
        out[i] = acc / (float)(g.grid_h * g.grid_w);
    }
}

// Gather backward.  One thread per feature PIXEL (y, x) and CHANNEL CHUNK (blockIdx.y): the (RoI, bin, weight)
// contributions that reach the pixel do not depend on the channel, so they are enumerated once per thread (RoIs in
// index order, samples in (iy, ix) order) into a small per-thread list and then replayed for the chunk's channels --
// the summation order is fixed (deterministic, no float atomics).  A list that fills up is flushed (accumulating
// into gfeat) and refilled.
// The channel chunks are what keeps the kernel balanced: a pixel under ~100 overlapping proposals has ~700 entries, and
// replaying them for all 256 channels in ONE thread made the launch take 24 .. 92 ms depending on how the RoI set
// clusters (measured at 600x1987, R = 256, eight different sets); with 32-channel chunks the longest thread does an
// eighth of that and there are eight times as many blocks to schedule.  The RoI table (level and box) is staged in
// shared memory once per block instead of being re-derived (logf / sqrtf / div) by every thread for every RoI.
constexpr int kRoiEntries = 96;
constexpr int kRoiTile = 512;            // RoIs staged in smem per pass
constexpr int kRoiChunk = 32;            // channels per thread

struct RoiEntry { int off; float w; };   // off = r*C*P*P + (ph*P + pw)*bs ; the channel adds c*cs

__device__ __forceinline__ void roi_flush(const RoiEntry* e, int n, const float* __restrict__ gout,
                                          float* __restrict__ gp, int c0, int c1, int cs, int64_t cstride, bool first) {
    for (int c = c0; c < c1; ++c) {
        float acc = first ? 0.f : gp[(int64_t)c * cstride];
        const float* gc = gout + (int64_t)c * cs;
        for (int k = 0; k < n; ++k) acc += e[k].w * __ldg(gc + e[k].off);
        gp[(int64_t)c * cstride] = acc;
    }
}

template <bool PYR>
__global__ void __launch_bounds__(128)
roi_align_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ rois,
                     float* __restrict__ gfeat, int R, int B, int C, int H, int W, int P, float scale, const PyrLevels lv,
                     int cs, int bs) {
    __shared__ float4 box_sh[kRoiTile];          // x1, y1, x2, y2 (image coordinates)
    __shared__ int key_sh[kRoiTile];             // batch index << 8 | pyramid level - 2
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int level = 0;
    bool active = true;
    if (PYR) {                                       // pixel of which level?  (batch of 1)
        if (i >= lv.pix_off[4]) { active = false; i = 0; }
        level = i >= lv.pix_off[3] ? 3 : (i >= lv.pix_off[2] ? 2 : (i >= lv.pix_off[1] ? 1 : 0));
        i -= lv.pix_off[level];
        gfeat = lv.gfeat[level]; H = lv.H[level]; W = lv.W[level]; scale = lv.scale[level];
    } else if (i >= (int64_t)B * H * W) {
        active = false; i = 0;
    }
    const int x = (int)(i % W), y = (int)((i / W) % H), b = (int)(i / ((int64_t)W * H));
    float* gp = gfeat + (int64_t)b * C * H * W + (int64_t)y * W + x;
    const int64_t cstride = (int64_t)H * W;
    const int PP = P * P;
    const int c0 = blockIdx.y * kRoiChunk, c1 = min(C, c0 + kRoiChunk);
    const int want = (b << 8) | level;
    RoiEntry ent[kRoiEntries];
    int n = 0;
    bool first = true;
    for (int r0 = 0; r0 < R; r0 += kRoiTile) {
        const int rn = min(kRoiTile, R - r0);
        __syncthreads();
        for (int t = threadIdx.x; t < rn; t += blockDim.x) {
            const float* roi = rois + (int64_t)(r0 + t) * 5;
            box_sh[t] = make_float4(__ldg(roi + 1), __ldg(roi + 2), __ldg(roi + 3), __ldg(roi + 4));
            key_sh[t] = ((int)__ldg(roi) << 8) | (PYR ? roi_fpn_level(roi) - 2 : 0);
        }
        __syncthreads();
        if (!active) continue;
        for (int t = 0; t < rn; ++t) {
            if (key_sh[t] != want) continue;
            const float4 bx = box_sh[t];
            const float sw = bx.x * scale, sh = bx.y * scale;
            const float ew = bx.z * scale, eh = bx.w * scale;
            const float rw = fmaxf(ew - sw, 1.f), rh = fmaxf(eh - sh, 1.f);
            if ((float)y < sh - 2.f || (float)y > sh + rh + 1.f || (float)x < sw - 2.f || (float)x > sw + rw + 1.f) continue;
            const int r = r0 + t;
            float rr[5] = {(float)b, bx.x, bx.y, bx.z, bx.w};
            const RoiGeom g = roi_geom(rr, scale, P);
            const float sy = g.bin_h / (float)g.grid_h, sx = g.bin_w / (float)g.grid_w;
            const int ny = P * g.grid_h, nx = P * g.grid_w;
            const int jy0 = max(0, (int)floorf(((float)y - 1.f - g.start_h) / sy - 0.5f) - 1);
            const int jy1 = min(ny - 1, (int)ceilf(((float)y + 1.f - g.start_h) / sy - 0.5f) + 1);
            const int jx0 = max(0, (int)floorf(((float)x - 1.f - g.start_w) / sx - 0.5f) - 1);
            const int jx1 = min(nx - 1, (int)ceilf(((float)x + 1.f - g.start_w) / sx - 0.5f) + 1);
            if (jy0 > jy1 || jx0 > jx1) continue;
            const float inv = 1.f / (float)(g.grid_h * g.grid_w);
            for (int jy = jy0; jy <= jy1; ++jy) {
                const int ph = jy / g.grid_h, iy = jy % g.grid_h;
                const float py = g.start_h + ph * g.bin_h + ((float)iy + .5f) * g.bin_h / (float)g.grid_h;
                int yl, yh; float wyl, wyh;
                if (!axis_interp(py, H, yl, yh, wyl, wyh)) continue;
                if (yl != y && yh != y) continue;
                const float wy = (yl == y ? wyl : 0.f) + (yh == y ? wyh : 0.f);
                for (int jx = jx0; jx <= jx1; ++jx) {
                    const int pw = jx / g.grid_w, ix = jx % g.grid_w;
                    const float px = g.start_w + pw * g.bin_w + ((float)ix + .5f) * g.bin_w / (float)g.grid_w;
                    int xl, xh; float wxl, wxh;
                    if (!axis_interp(px, W, xl, xh, wxl, wxh)) continue;
                    if (xl != x && xh != x) continue;
                    const float wx = (xl == x ? wxl : 0.f) + (xh == x ? wxh : 0.f);
                    if (n == kRoiEntries) { roi_flush(ent, n, gout, gp, c0, c1, cs, cstride, first); first = false; n = 0; }
                    ent[n].off = r * C * PP + (ph * P + pw) * bs;
                    ent[n].w = wy * wx * inv;
                    ++n;
                }
            }
        }
    }
    if (active && (n > 0 || first)) roi_flush(ent, n, gout, gp, c0, c1, cs, cstride, first);
}

// Warp-per-pixel gather backward (C a multiple of 32, <= 256: the FPN widths).  A block owns 32 consecutive pixels of
// the (concatenated) pixel range, each warp walks 8 of them; for one pixel the 32 lanes test 32 RoIs at a time against
// its position (ballot), the hits are visited in ascending RoI order, and for every (RoI, sample) that touches the
// pixel lane l adds weight * gout[channel l + 32 j] for j < C / 32 -- the same entries in the same order as the
// kernel above, so the result is bit-identical to it, but the index arithmetic of an entry is done once for all
// channels, the loads of one entry are C contiguous floats when gout is channels-last (layout 1: [R,P,P,C]), and the
// work of a pixel under many overlapping proposals is spread over a warp.  The block's 32 x C results go through a
// shared-memory tile so that the stores run along the pixels (128 B rows of the NCHW gradient map).
// Measured (600x1987 pair, R = 256, C = 256, P = 7, four levels, one launch): 3.0 ms with the chunked kernel above,
// 0.93 ms with every lane scanning the whole sample window of a hit (issue-bound, ncu), ~0.2 ms with one window
// candidate per lane (below).
constexpr int kRoiWarpMaxC = 256;

template <bool PYR>
__global__ void __launch_bounds__(128)
roi_align_bwd_warp_kernel(const float* __restrict__ gout, const float* __restrict__ rois, float* __restrict__ gfeat0,
                          int R, int C, int H0, int W0, int P, float scale0, const PyrLevels lv, long long npix,
                          long long cs, long long bs) {
    __shared__ float4 box_sh[kRoiTile];
    __shared__ int key_sh[kRoiTile];
    __shared__ float tile[kRoiWarpMaxC][33];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long base = (long long)blockIdx.x * 32;
    const int PP = P * P, nacc = C / 32;
    for (int c = threadIdx.x; c < C * 32; c += blockDim.x) tile[c >> 5][c & 31] = 0.f;

    auto locate = [&](long long i, int& level, int& H, int& W, float& scale, float*& gf, int& x, int& y) {
        level = 0; H = H0; W = W0; scale = scale0; gf = gfeat0;
        if (PYR) {
            level = i >= lv.pix_off[3] ? 3 : (i >= lv.pix_off[2] ? 2 : (i >= lv.pix_off[1] ? 1 : 0));
            i -= lv.pix_off[level];
            gf = lv.gfeat[level]; H = lv.H[level]; W = lv.W[level]; scale = lv.scale[level];
        }
        x = (int)(i % W); y = (int)(i / W);
    };

    for (int r0 = 0; r0 < R; r0 += kRoiTile) {
        const int rn = min(kRoiTile, R - r0);
        __syncthreads();
        for (int t = threadIdx.x; t < rn; t += blockDim.x) {
            const float* roi = rois + (long long)(r0 + t) * 5;
            box_sh[t] = make_float4(__ldg(roi + 1), __ldg(roi + 2), __ldg(roi + 3), __ldg(roi + 4));
            key_sh[t] = ((int)__ldg(roi) << 8) | (PYR ? roi_fpn_level(roi) - 2 : 0);
        }
        __syncthreads();
        for (int q = 0; q < 8; ++q) {
            const int slot = warp * 8 + q;
            const long long i = base + slot;
            if (i >= npix) break;
            int level, H, W, x, y; float scale; float* gf;
            locate(i, level, H, W, scale, gf, x, y);
            float acc[kRoiWarpMaxC / 32];                    // continues the sum of the previous RoI tiles
#pragma unroll
            for (int j = 0; j < kRoiWarpMaxC / 32; ++j) acc[j] = j < nacc ? tile[j * 32 + lane][slot] : 0.f;
            for (int g0 = 0; g0 < rn; g0 += 32) {
                const int t = g0 + lane;
                bool hit = false;
                if (t < rn && key_sh[t] == level) {          // batch of 1: key = level
                    const float4 bx = box_sh[t];
                    const float sw = bx.x * scale, sh = bx.y * scale;
                    const float rw = fmaxf(bx.z * scale - sw, 1.f), rh = fmaxf(bx.w * scale - sh, 1.f);
                    hit = !((float)y < sh - 2.f || (float)y > sh + rh + 1.f || (float)x < sw - 2.f || (float)x > sw + rw + 1.f);
                }
                unsigned mask = __ballot_sync(0xffffffffu, hit);
                while (mask) {
                    const int tt = g0 + __ffs(mask) - 1;
                    mask &= mask - 1;
                    const float4 bx = box_sh[tt];
                    float rr[5] = {0.f, bx.x, bx.y, bx.z, bx.w};
                    const RoiGeom g = roi_geom(rr, scale, P);
                    const float sy = g.bin_h / (float)g.grid_h, sx = g.bin_w / (float)g.grid_w;
                    const int ny = P * g.grid_h, nx = P * g.grid_w;
                    const int jy0 = max(0, (int)floorf(((float)y - 1.f - g.start_h) / sy - 0.5f) - 1);
                    const int jy1 = min(ny - 1, (int)ceilf(((float)y + 1.f - g.start_h) / sy - 0.5f) + 1);
                    const int jx0 = max(0, (int)floorf(((float)x - 1.f - g.start_w) / sx - 0.5f) - 1);
                    const int jx1 = min(nx - 1, (int)ceilf(((float)x + 1.f - g.start_w) / sx - 0.5f) + 1);
                    if (jy0 > jy1 || jx0 > jx1) continue;
                    const float inv = 1.f / (float)(g.grid_h * g.grid_w);
                    const float* groi = gout + (long long)(r0 + tt) * C * PP + lane * cs;
                    // the (jy, jx) candidates of the window are evaluated one per lane, the contributing ones are then
                    // visited in ascending (jy, jx) order (ballot) with weight and bin broadcast from their lane
                    const int nwx = jx1 - jx0 + 1, ncand = (jy1 - jy0 + 1) * nwx;
                    for (int cb = 0; cb < ncand; cb += 32) {
                        const int ci = cb + lane;
                        float wgt = 0.f;
                        int bin = 0;
                        bool valid = false;
                        if (ci < ncand) {
                            const int jy = jy0 + ci / nwx, jx = jx0 + ci % nwx;
                            const int ph = jy / g.grid_h, iy = jy % g.grid_h;
                            const float py = g.start_h + ph * g.bin_h + ((float)iy + .5f) * g.bin_h / (float)g.grid_h;
                            int yl, yh; float wyl, wyh;
                            if (axis_interp(py, H, yl, yh, wyl, wyh) && (yl == y || yh == y)) {
                                const float wy = (yl == y ? wyl : 0.f) + (yh == y ? wyh : 0.f);
                                const int pw = jx / g.grid_w, ix = jx % g.grid_w;
                                const float px = g.start_w + pw * g.bin_w + ((float)ix + .5f) * g.bin_w / (float)g.grid_w;
                                int xl, xh; float wxl, wxh;
                                if (axis_interp(px, W, xl, xh, wxl, wxh) && (xl == x || xh == x)) {
                                    const float wx = (xl == x ? wxl : 0.f) + (xh == x ? wxh : 0.f);
                                    wgt = wy * wx * inv;
                                    bin = ph * P + pw;
                                    valid = true;
                                }
                            }
                        }
                        unsigned vm = __ballot_sync(0xffffffffu, valid);
                        while (vm) {
                            const int src = __ffs(vm) - 1;
                            vm &= vm - 1;
                            const float wv = __shfl_sync(0xffffffffu, wgt, src);
                            const int bv = __shfl_sync(0xffffffffu, bin, src);
                            const float* ge = groi + (long long)bv * bs;
#pragma unroll
                            for (int j = 0; j < kRoiWarpMaxC / 32; ++j)
                                if (j < nacc) acc[j] += wv * __ldg(ge + (long long)j * 32 * cs);
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < kRoiWarpMaxC / 32; ++j)
                if (j < nacc) tile[j * 32 + lane][slot] = acc[j];
        }
    }
    __syncthreads();
    // stores along the pixels: lane = pixel of the block, warps stride over the channels
    const long long i = base + lane;
    if (i < npix) {
        int level, H, W, x, y; float scale; float* gf;
        locate(i, level, H, W, scale, gf, x, y);
        float* gp = gf + (long long)y * W + x;
        const long long cstride = (long long)H * W;
        for (int c = warp; c < C; c += 4) gp[c * cstride] = tile[c][lane];
    }
}

}  // namespace b2

using namespace b2;

extern "C" int b2_roi_align_fwd(const float* feat, const float* rois, float* out, int R, int C, int H,
                                int W, int P, float scale, void* stream) {
    B2_REQUIRE(feat && out && (rois || R == 0), "roi_align_fwd: null pointer");
    B2_REQUIRE(P >= 1 && C >= 1 && H >= 1 && W >= 1, "roi_align_fwd: bad dims");
    int64_t total = (int64_t)R * C * P * P;
    if (total == 0) return 0;
    roi_align_fwd_kernel<false><<<stream_grid(total, 256, kNumSMs * 32), 256, 0, (cudaStream_t)stream>>>(
        feat, rois, out, R, C, H, W, P, scale, PyrLevels{});
    return check_launch("roi_align_fwd");
}

// gout_layout 0: [R,C,P,P] (what the upstream op hands back); 1: [R,P,P,C] (channels-last: contiguous channel rows)
static bool roi_warp_kernel_ok(int C, int gout_layout) {
    (void)gout_layout;                     // both kernels read either layout
    return flag_value(kFlagRoiBwdWarp, "B2_ROI_BWD_WARP", 1) && C % 32 == 0 && C <= kRoiWarpMaxC;
}

extern "C" int b2_roi_align_bwd(const float* gout, const float* rois, float* gfeat, int R, int C, int H,
                                int W, int P, float scale, int gout_layout, void* stream) {
    B2_REQUIRE(gfeat && (R == 0 || (gout && rois)), "roi_align_bwd: null pointer");
    B2_REQUIRE(P >= 1 && C >= 1 && H >= 1 && W >= 1, "roi_align_bwd: bad dims");
    B2_REQUIRE((int64_t)R * C * P * P < ((int64_t)1 << 31), "roi_align_bwd: gout has more than 2^31 elements");
    B2_REQUIRE(gout_layout == 0 || gout_layout == 1, "roi_align_bwd: gout_layout must be 0 ([R,C,P,P]) or 1 ([R,P,P,C])");
    const int64_t npix = (int64_t)H * W;     // batch of 1 (the attack scripts run batch size 1)
    if (roi_warp_kernel_ok(C, gout_layout)) {
        const long long cs = gout_layout ? 1 : (long long)P * P, bs = gout_layout ? C : 1;
        roi_align_bwd_warp_kernel<false><<<(unsigned)((npix + 31) / 32), 128, 0, (cudaStream_t)stream>>>(
            gout, rois, gfeat, R, C, H, W, P, scale, PyrLevels{}, npix, cs, bs);
        return check_launch("roi_align_bwd");
    }
    roi_align_bwd_kernel<false><<<dim3((unsigned)((npix + 127) / 128), (C + kRoiChunk - 1) / kRoiChunk), 128, 0,
                                  (cudaStream_t)stream>>>(gout, rois, gfeat, R, 1, C, H, W, P, scale, PyrLevels{},
                                                          gout_layout ? 1 : P * P, gout_layout ? C : 1);
    return check_launch("roi_align_bwd");
}

static int fill_levels(PyrLevels& lv, const float* const* feats, float* const* gfeats, const int* Hs, const int* Ws,
                       float im_h, const char* who) {
    lv.pix_off[0] = 0;
    for (int l = 0; l < 4; ++l) {
        if (Hs[l] < 1 || Ws[l] < 1 || (feats && !feats[l]) || (gfeats && !gfeats[l])) {
            set_error("%s: bad pyramid level %d", who, l + 2);
            return B2_ERR_BAD_ARG;
        }
        lv.feat[l] = feats ? feats[l] : nullptr;
        lv.gfeat[l] = gfeats ? gfeats[l] : nullptr;
        lv.H[l] = Hs[l]; lv.W[l] = Ws[l];
        lv.scale[l] = (float)((double)Hs[l] / (double)im_h);   // stereo_rcnn.py:131: feat_maps[i].size(2) / im_info[0][0]
        lv.pix_off[l + 1] = lv.pix_off[l] + (long long)Hs[l] * Ws[l];
    }
    return 0;
}

extern "C" int b2_roi_align_pyramid_fwd(const float* const* feats, const int* Hs, const int* Ws, const float* rois,
                                        float* out, int R, int C, int P, float im_h, void* stream) {
    B2_REQUIRE(feats && Hs && Ws && out && (rois || R == 0), "roi_align_pyramid_fwd: null pointer");
    B2_REQUIRE(P >= 1 && C >= 1 && im_h > 0.f, "roi_align_pyramid_fwd: bad dims");
    PyrLevels lv{};
    if (int e = fill_levels(lv, feats, nullptr, Hs, Ws, im_h, "roi_align_pyramid_fwd")) return e;
    int64_t total = (int64_t)R * C * P * P;
    if (total == 0) return 0;
    roi_align_fwd_kernel<true><<<stream_grid(total, 256, kNumSMs * 32), 256, 0, (cudaStream_t)stream>>>(
        nullptr, rois, out, R, C, 0, 0, P, 0.f, lv);
    return check_launch("roi_align_pyramid_fwd");
}

extern "C" int b2_roi_align_pyramid_bwd(const float* gout, const float* rois, float* const* gfeats, const int* Hs,
                                        const int* Ws, int R, int C, int P, float im_h, int gout_layout, void* stream) {
    B2_REQUIRE(gfeats && Hs && Ws && (R == 0 || (gout && rois)), "roi_align_pyramid_bwd: null pointer");
    B2_REQUIRE(P >= 1 && C >= 1 && im_h > 0.f, "roi_align_pyramid_bwd: bad dims");
    B2_REQUIRE((int64_t)R * C * P * P < ((int64_t)1 << 31), "roi_align_pyramid_bwd: gout has more than 2^31 elements");
    B2_REQUIRE(gout_layout == 0 || gout_layout == 1, "roi_align_pyramid_bwd: gout_layout must be 0 ([R,C,P,P]) or 1 ([R,P,P,C])");
    PyrLevels lv{};
    if (int e = fill_levels(lv, nullptr, gfeats, Hs, Ws, im_h, "roi_align_pyramid_bwd")) return e;
    const int64_t npix = lv.pix_off[4];
    if (roi_warp_kernel_ok(C, gout_layout)) {
        const long long cs = gout_layout ? 1 : (long long)P * P, bs = gout_layout ? C : 1;
        roi_align_bwd_warp_kernel<true><<<(unsigned)((npix + 31) / 32), 128, 0, (cudaStream_t)stream>>>(
            gout, rois, nullptr, R, C, 0, 0, P, 0.f, lv, npix, cs, bs);
        return check_launch("roi_align_pyramid_bwd");
    }
    roi_align_bwd_kernel<true><<<dim3((unsigned)((npix + 127) / 128), (C + kRoiChunk - 1) / kRoiChunk), 128, 0,
                                 (cudaStream_t)stream>>>(gout, rois, nullptr, R, 1, C, 0, 0, P, 0.f, lv,
                                                         gout_layout ? 1 : P * P, gout_layout ? C : 1);
    return check_launch("roi_align_pyramid_bwd");
}
