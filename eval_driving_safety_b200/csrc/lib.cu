// libb2attack.so: version, error reporting and the conv3d dispatcher.
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include "common.cuh"

namespace b2 {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int conv3d_simt_launch(const float* in, const float* wp, float* out, int N, int Cin, int Cout, int Di,
                       int Hi, int Wi, int Do, int Ho, int Wo, int stride, int mode, cudaStream_t st);
int conv3d_tcgen05_launch(const float* in, const float* wp, float* out, int N, int Cin, int Cout, int Di,
                          int Hi, int Wi, int Do, int Ho, int Wo, int stride, int mode, cudaStream_t st,
                          const EpiFusion& ef, int* stat_rows, bool query, int* addend_ok);

static int g_flags[kNumFlags] = {-1, -1, -1, -1, -1};

int flag_value(Flag f, const char* env_name, int dflt) {
    if (g_flags[f] >= 0) return g_flags[f];
    const char* e = getenv(env_name);
    if (e && e[0] >= '0' && e[0] <= '9') return atoi(e);
    return dflt;
}

}  // namespace b2

extern "C" int b2_version(void) { return 200; }  // 0.2.0

extern "C" int b2_set_flag(const char* name, int value) {
    B2_REQUIRE(name, "set_flag: null name");
    int idx = -1;
    if (!strcmp(name, "conv_dc_pair")) idx = b2::kFlagConvDcPair;
    else if (!strcmp(name, "conv2d_halo")) idx = b2::kFlagConv2dHalo;
    else if (!strcmp(name, "conv_s2_pair")) idx = b2::kFlagConvG2Pair;
    else if (!strcmp(name, "depth_head_x4")) idx = b2::kFlagDepthHeadX4;
    else if (!strcmp(name, "roi_bwd_warp")) idx = b2::kFlagRoiBwdWarp;
    B2_REQUIRE(idx >= 0, "set_flag: unknown flag '%s'", name);
    b2::g_flags[idx] = value < 0 ? -1 : value;
    return 0;
}

extern "C" const char* b2_last_error(void) { return b2::g_err; }

static int conv3d_dispatch(const float* in, const float* wp, float* out, int N, int Cin, int Cout, int Di,
                           int Hi, int Wi, int stride, int mode, int impl, void* stream, const b2::EpiFusion& ef,
                           int* stat_rows, bool query, int* addend_ok) {
    const float* addend = ef.addend;
    float* stat_partial = ef.stat_partial;
    B2_REQUIRE(query || (in && wp && out), "conv3d: null pointer");
    B2_REQUIRE(N >= 0 && Cin > 0 && Cout > 0 && Di > 0 && Hi > 0 && Wi > 0, "conv3d: bad dims");
    B2_REQUIRE(mode == 0 || mode == 1, "conv3d: mode must be 0 (CONV) or 1 (DECONV)");
    B2_REQUIRE((mode == 0 && (stride == 1 || stride == 2)) || (mode == 1 && stride == 2),
               "conv3d: unsupported stride %d for mode %d", stride, mode);
    int Do, Ho, Wo;
    if (mode == 0) {
        Do = (Di - 1) / stride + 1; Ho = (Hi - 1) / stride + 1; Wo = (Wi - 1) / stride + 1;
    } else {
        Do = 2 * Di; Ho = 2 * Hi; Wo = 2 * Wi;
    }
    if (stat_rows) *stat_rows = 0;
    if (addend_ok) *addend_ok = 0;
    if (N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (impl == 1) {
        B2_REQUIRE(!stat_partial && !addend, "conv3d: the SIMT implementation has no fused epilogue");
        if (query) return 0;
        return b2::conv3d_simt_launch(in, wp, out, N, Cin, Cout, Di, Hi, Wi, Do, Ho, Wo, stride, mode, st);
    }
    if (impl == 0)
        return b2::conv3d_tcgen05_launch(in, wp, out, N, Cin, Cout, Di, Hi, Wi, Do, Ho, Wo, stride, mode, st,
                                         ef, stat_rows, query, addend_ok);
    b2::set_error("conv3d: unknown impl %d", impl);
    return B2_ERR_BAD_ARG;
}

extern "C" int b2_conv3d(const float* in, const float* wp, float* out, int N, int Cin, int Cout, int Di,
                         int Hi, int Wi, int stride, int mode, int impl, void* stream) {
    const b2::EpiFusion none{nullptr, nullptr, 0, nullptr, nullptr};
    return conv3d_dispatch(in, wp, out, N, Cin, Cout, Di, Hi, Wi, stride, mode, impl, stream, none, nullptr, false, nullptr);
}

extern "C" int b2_conv3d_fusion_caps(int N, int Cin, int Cout, int Di, int Hi, int Wi, int stride, int mode,
                                     int* stat_rows, int* addend_ok) {
    int rows = 0, aok = 0;
    const b2::EpiFusion none{nullptr, nullptr, 0, nullptr, nullptr};
    int rc = conv3d_dispatch(nullptr, nullptr, nullptr, N, Cin, Cout, Di, Hi, Wi, stride, mode, 0, nullptr, none, &rows,
                             true, &aok);
    if (rc != 0) { rows = 0; aok = 0; }
    if (stat_rows) *stat_rows = rows;
    if (addend_ok) *addend_ok = aok;
    return 0;
}

extern "C" int b2_conv3d_fused(const float* in, const float* wp, float* out, const float* addend, float* stat_partial,
                               int stat_mode, const float* gn_x, const float* gn_coef, int N, int Cin, int Cout,
                               int Di, int Hi, int Wi, int stride, int mode, void* stream) {
    B2_REQUIRE(stat_partial ? (stat_mode >= 1 && stat_mode <= 3) : stat_mode == 0,
               "conv3d_fused: stat_mode must be 1..3 with a statistics table and 0 without");
    B2_REQUIRE(stat_mode < 2 || gn_x, "conv3d_fused: stat_mode 2/3 needs gn_x");
    B2_REQUIRE(stat_mode != 3 || (gn_coef && Cout <= 256), "conv3d_fused: stat_mode 3 needs gn_coef and Cout <= 256");
    const b2::EpiFusion ef{addend, stat_partial, stat_mode, gn_x, gn_coef};
    return conv3d_dispatch(in, wp, out, N, Cin, Cout, Di, Hi, Wi, stride, mode, 0, stream, ef, nullptr, false, nullptr);
}
