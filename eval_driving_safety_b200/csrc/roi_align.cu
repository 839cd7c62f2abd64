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

__global__ void __launch_bounds__(256)
roi_align_fwd_kernel(const float* __restrict__ feat, const float* __restrict__ rois,
                     float* __restrict__ out, int R, int C, int H, int W, int P, float scale) {
    const int64_t total = (int64_t)R * C * P * P;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int pw = (int)(i % P), ph = (int)((i / P) % P);
        int c = (int)((i / ((int64_t)P * P)) % C), r = (int)(i / ((int64_t)P * P * C));
        const float* roi = rois + r * 5;
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

// Gather backward.  One thread per feature element (c, y, x) of batch 0..B-1; RoIs
// are visited in index order and samples in (iy, ix) order -> fixed summation order.
// For a pixel y the contributing sample rows are those whose (clamped) position lies
// in (y-1, y+1]; candidates are enumerated generously and filtered with the exact
// forward arithmetic.
constexpr int kRoiChunk = 128;

__global__ void __launch_bounds__(256)
roi_align_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ rois,
                     float* __restrict__ gfeat, int R, int B, int C, int H, int W, int P, float scale) {
    __shared__ float s_roi[kRoiChunk][5];
    const int64_t total = (int64_t)B * C * H * W;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < total;
    int x = 0, y = 0, c = 0, b = 0;
    if (live) {
        x = (int)(i % W); y = (int)((i / W) % H);
        c = (int)((i / ((int64_t)W * H)) % C); b = (int)(i / ((int64_t)W * H * C));
    }
    float acc = 0.f;
    for (int r0 = 0; r0 < R; r0 += kRoiChunk) {
        int nr = min(kRoiChunk, R - r0);
        __syncthreads();
        for (int k = threadIdx.x; k < nr * 5; k += blockDim.x) s_roi[k / 5][k % 5] = rois[(r0 + k / 5) * 5 + k % 5];
        __syncthreads();
        if (!live) continue;
        for (int rr = 0; rr < nr; ++rr) {
            const float* roi = s_roi[rr];
            if ((int)roi[0] != b) continue;
            RoiGeom g = roi_geom(roi, scale, P);
            float sy = g.bin_h / (float)g.grid_h, sx = g.bin_w / (float)g.grid_w;
            int ny = P * g.grid_h, nx = P * g.grid_w;
            // sample j sits at start + (j + .5) * s ; want positions in [y-1, y+1]
            int jy0 = max(0, (int)floorf(((float)y - 1.f - g.start_h) / sy - 0.5f) - 1);
            int jy1 = min(ny - 1, (int)ceilf(((float)y + 1.f - g.start_h) / sy - 0.5f) + 1);
            if (jy0 > jy1) continue;
            int jx0 = max(0, (int)floorf(((float)x - 1.f - g.start_w) / sx - 0.5f) - 1);
            int jx1 = min(nx - 1, (int)ceilf(((float)x + 1.f - g.start_w) / sx - 0.5f) + 1);
            if (jx0 > jx1) continue;
            float inv = 1.f / (float)(g.grid_h * g.grid_w);
            const float* go = gout + ((int64_t)(r0 + rr) * C + c) * P * P;
            for (int jy = jy0; jy <= jy1; ++jy) {
                int ph = jy / g.grid_h, iy = jy % g.grid_h;
                float py = g.start_h + ph * g.bin_h + ((float)iy + .5f) * g.bin_h / (float)g.grid_h;
                int yl, yh; float wyl, wyh;
                if (!axis_interp(py, H, yl, yh, wyl, wyh)) continue;
                float wy = (yl == y ? wyl : 0.f) + (yh == y ? wyh : 0.f);
                if (yl != y && yh != y) continue;
                for (int jx = jx0; jx <= jx1; ++jx) {
                    int pw = jx / g.grid_w, ix = jx % g.grid_w;
                    float px = g.start_w + pw * g.bin_w + ((float)ix + .5f) * g.bin_w / (float)g.grid_w;
                    int xl, xh; float wxl, wxh;
                    if (!axis_interp(px, W, xl, xh, wxl, wxh)) continue;
                    if (xl != x && xh != x) continue;
                    float wx = (xl == x ? wxl : 0.f) + (xh == x ? wxh : 0.f);
                    acc += wy * wx * __ldg(go + ph * P + pw) * inv;
                }
            }
        }
    }
    if (live) gfeat[i] = acc;
}

}  // namespace b2

using namespace b2;

extern "C" int b2_roi_align_fwd(const float* feat, const float* rois, float* out, int R, int C, int H,
                                int W, int P, float scale, void* stream) {
    B2_REQUIRE(feat && out && (rois || R == 0), "roi_align_fwd: null pointer");
    B2_REQUIRE(P >= 1 && C >= 1 && H >= 1 && W >= 1, "roi_align_fwd: bad dims");
    int64_t total = (int64_t)R * C * P * P;
    if (total == 0) return 0;
    roi_align_fwd_kernel<<<stream_grid(total, 256, kNumSMs * 32), 256, 0, (cudaStream_t)stream>>>(
        feat, rois, out, R, C, H, W, P, scale);
    return check_launch("roi_align_fwd");
}

extern "C" int b2_roi_align_bwd(const float* gout, const float* rois, float* gfeat, int R, int C, int H,
                                int W, int P, float scale, void* stream) {
    B2_REQUIRE(gfeat && (R == 0 || (gout && rois)), "roi_align_bwd: null pointer");
    B2_REQUIRE(P >= 1 && C >= 1 && H >= 1 && W >= 1, "roi_align_bwd: bad dims");
    int64_t total = (int64_t)C * H * W;   // batch of 1 (the attack scripts run batch size 1)
    roi_align_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        gout, rois, gfeat, R, 1, C, H, W, P, scale);
    return check_launch("roi_align_bwd");
}
