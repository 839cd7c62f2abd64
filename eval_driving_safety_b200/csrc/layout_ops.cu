// Channel concatenation / split and batch-prefix gradient merge on channels-last maps: the glue between the
// extractor's SPP branches, its two heads and the fused detection heads (upstream feature_extraction.forward
// torch.cat / slicing, reached from attack/DSGN/pgd_attack.py:308 forward and :336 backward).  Stock autograd runs
// these as lazy strided views that every consumer then copies, zero-filled gradient buffers and separate adds (24
// copies + 11 fills + 8 adds per pair-iteration measured); here each direction is ONE streaming launch.
// Roofline: HBM streaming, bytes = everything read once + everything written once.
#include "common.cuh"

namespace b2 {

constexpr int kMaxPieces = 8;

struct Pieces {
    const float* src[kMaxPieces];   // concat: sources (nullptr = zeros);  split: unused
    float* dst[kMaxPieces];         // split: destinations (nullptr = skip); concat: unused
    int off4[kMaxPieces + 1];       // float4 offset of piece k inside a row of the wide tensor
    int n;
};

// wide[row][off_k + c] = piece_k[row][c]   (rows = N*H*W pixels; all widths multiples of 4 floats)
__global__ void __launch_bounds__(256)
channel_concat_kernel(Pieces p, float4* __restrict__ wide, int64_t rows, int c4_total) {
    const int64_t total = rows * c4_total;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / c4_total;
        const int c = (int)(i - row * c4_total);
        int k = 0;
#pragma unroll
        for (int j = 1; j < kMaxPieces; ++j) k += (j < p.n && c >= p.off4[j]) ? 1 : 0;
        const int w4 = p.off4[k + 1] - p.off4[k];
        const float4* s = reinterpret_cast<const float4*>(p.src[k]);
        wide[i] = s ? ldg_stream(s + row * w4 + (c - p.off4[k])) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

__global__ void __launch_bounds__(256)
channel_split_kernel(Pieces p, const float4* __restrict__ wide, int64_t rows, int c4_total) {
    const int64_t total = rows * c4_total;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / c4_total;
        const int c = (int)(i - row * c4_total);
        int k = 0;
#pragma unroll
        for (int j = 1; j < kMaxPieces; ++j) k += (j < p.n && c >= p.off4[j]) ? 1 : 0;
        float4* d = reinterpret_cast<float4*>(p.dst[k]);
        if (!d) continue;
        const int w4 = p.off4[k + 1] - p.off4[k];
        d[row * w4 + (c - p.off4[k])] = ldg_stream(wide + i);
    }
}

// out[i] = a[i] + (i < n_prefix ? b[i] : 0)
__global__ void __launch_bounds__(256)
add_prefix_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ out, int64_t n4,
                  int64_t n4_prefix) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 v = ldg_stream(a + i);
        if (i < n4_prefix) {
            const float4 w = ldg_stream(b + i);
            v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
        }
        out[i] = v;
    }
}

static int pieces_setup(const char* who, Pieces& p, const int* widths, int n, int* c_total) {
    if (n < 1 || n > kMaxPieces) { set_error("%s: 1..%d pieces (got %d)", who, kMaxPieces, n); return B2_ERR_BAD_ARG; }
    int off = 0;
    for (int k = 0; k < n; ++k) {
        if (widths[k] <= 0 || widths[k] % 4) { set_error("%s: piece widths must be positive multiples of 4 (got %d)", who, widths[k]); return B2_ERR_UNSUPPORTED; }
        p.off4[k] = off / 4;
        off += widths[k];
    }
    for (int k = n; k <= kMaxPieces; ++k) p.off4[k] = off / 4;
    p.n = n;
    *c_total = off;
    return 0;
}

}  // namespace b2

using namespace b2;

extern "C" int b2_channel_concat(const float* const* srcs, const int* widths, int n_pieces, float* wide, int64_t rows,
                                 void* stream) {
    B2_REQUIRE(srcs && widths && wide, "channel_concat: null pointer");
    B2_REQUIRE(rows >= 0, "channel_concat: negative size");
    Pieces p{};
    int c_total = 0;
    if (int e = pieces_setup("channel_concat", p, widths, n_pieces, &c_total)) return e;
    for (int k = 0; k < n_pieces; ++k) {
        B2_REQUIRE(!srcs[k] || aligned16(srcs[k]), "channel_concat: pointers must be 16-byte aligned");
        p.src[k] = srcs[k];
    }
    B2_REQUIRE(aligned16(wide), "channel_concat: pointers must be 16-byte aligned");
    if (rows == 0) return 0;
    const int64_t total = rows * (c_total / 4);
    channel_concat_kernel<<<stream_grid(total, 256 * 4, kNumSMs * 8), 256, 0, (cudaStream_t)stream>>>(p, (float4*)wide, rows, c_total / 4);
    return check_launch("channel_concat");
}

extern "C" int b2_channel_split(const float* wide, float* const* dsts, const int* widths, int n_pieces, int64_t rows,
                                void* stream) {
    B2_REQUIRE(wide && dsts && widths, "channel_split: null pointer");
    B2_REQUIRE(rows >= 0, "channel_split: negative size");
    Pieces p{};
    int c_total = 0;
    if (int e = pieces_setup("channel_split", p, widths, n_pieces, &c_total)) return e;
    for (int k = 0; k < n_pieces; ++k) {
        B2_REQUIRE(!dsts[k] || aligned16(dsts[k]), "channel_split: pointers must be 16-byte aligned");
        p.dst[k] = dsts[k];
    }
    B2_REQUIRE(aligned16(wide), "channel_split: pointers must be 16-byte aligned");
    if (rows == 0) return 0;
    const int64_t total = rows * (c_total / 4);
    channel_split_kernel<<<stream_grid(total, 256 * 4, kNumSMs * 8), 256, 0, (cudaStream_t)stream>>>(p, (const float4*)wide, rows, c_total / 4);
    return check_launch("channel_split");
}

extern "C" int b2_add_prefix(const float* a, const float* b, float* out, int64_t count, int64_t count_prefix,
                             void* stream) {
    B2_REQUIRE(a && out && (b || count_prefix == 0), "add_prefix: null pointer");
    B2_REQUIRE(count >= 0 && count_prefix >= 0 && count_prefix <= count, "add_prefix: need 0 <= count_prefix <= count");
    B2_REQUIRE(count % 4 == 0 && count_prefix % 4 == 0, "add_prefix: counts must be multiples of 4");
    B2_REQUIRE(aligned16(a) && aligned16(out) && (!b || aligned16(b)), "add_prefix: pointers must be 16-byte aligned");
    if (count == 0) return 0;
    add_prefix_kernel<<<stream_grid(count / 4, 256 * 4, kNumSMs * 8), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)a, (const float4*)b, (float4*)out, count / 4, count_prefix / 4);
    return check_launch("add_prefix");
}
