/*
 * b2attack.h -- C ABI of libb2attack.so: hand-written sm_100a kernels for the
 * attack hot path of DexterJZ/eval_driving_safety (PGD / FGSM / patch attacks
 * against DSGN and Stereo R-CNN).
 *
 * The reference has no FFI of its own for this path: it is inline PyTorch code
 * in attack/DSGN/{pgd,patch}_attack.py and attack/Stereo-RCNN/ (*.py) that reaches
 * ATen kernels and the upstream extensions dsgn._C / model.roi_layers.  Each
 * entry point below names the reference lines (or upstream op) it replaces.
 * INTEGRATION.md shows the ctypes / autograd.Function binding a maintainer adds.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes; all tensors fp32, contiguous in the stated layout;
 *     every pointer is DEVICE memory unless marked "host".
 *   - the caller owns every buffer including workspaces; the library never
 *     allocates or frees device memory, never synchronises, and launches on the
 *     cudaStream_t passed as `void* stream` (graph-capturable).
 *   - return 0 on success, non-zero (cudaError_t or B2_ERR_*) otherwise; the text
 *     of the last error on the calling thread is b2_last_error().  Never throws.
 *   - no global state except a per-device table of granted cudaFuncSetAttribute dynamic-smem
 *     opt-ins; the intended model is one process per GPU (torchrun), but a process may drive
 *     several devices; calls on one stream are ordered by that stream.
 */
#ifndef B2ATTACK_H
#define B2ATTACK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2_ERR_BAD_ARG   10001
#define B2_ERR_UNSUPPORTED 10002
#define B2_ERR_DRIVER    10003

int b2_version(void);
const char* b2_last_error(void);
/* Kernel-variant switches for A/B measurements and for testing both variants in one process:
 *   "conv_dc_pair"  1 = CTA-pair (tcgen05 cta_group::2) transposed-conv kernel (default), 0 = single-CTA;
 *   "conv_s2_pair"  1 = CTA-pair variant of the one-tap-per-stage kernel that serves the stride-2 convs (default), 0 = single-CTA;
 *   "conv2d_halo"   1 = halo-reuse stride-1 3x3 2-D conv kernel (default), 0 = one tap per stage;
 *   "depth_head_x4" 1 = unrolled depth-head kernels when every upsampling ratio is 4 (default), 0 = the generic ones;
 *   "roi_bwd_warp"  1 = warp-per-pixel RoIAlign backward for C % 32 == 0, C <= 256 (default), 0 = thread per pixel and
 *                   channel chunk.
 * value < 0 returns the flag to its default (environment variable B2_<NAME>, else the built-in default). */
int b2_set_flag(const char* name, int value);

/* ------------------------------------------------------------------------- *
 * (4) Perturbation update -- replaces the ~26 ATen launches of
 *     attack/DSGN/pgd_attack.py:339-354 (and :196-207 de/normalise) and of
 *     attack/Stereo-RCNN/pgd_attack.py:177-217 with ONE launch over all sets.
 *
 *   v   = denorm ? x*std_c + mean_c : x
 *   adv = v + alpha*sign(g);  eta = clamp(adv - clean, -eps, eps)
 *   o   = clamp(clean + eta, lo_c, hi_c);  out = denorm ? (o - mean_c)/std_c : o
 *
 * n_sets (<=4) tensors (e.g. left and right image batches) of shape
 * [n_img, C, hw] each (C <= 4); x/g/clean/out are HOST arrays of n_sets device
 * pointers.  mean/std/lo/hi: HOST arrays of C floats.  out may alias x.
 * Bit-exact with the reference's fp32 operation order (no FMA contraction).
 * ------------------------------------------------------------------------- */
int b2_pgd_update(const float* const* x, const float* const* g, const float* const* clean,
                  float* const* out, int n_sets, int n_img, int C, int64_t hw,
                  float alpha, float eps, int denorm,
                  const float* mean, const float* std_, const float* lo, const float* hi,
                  void* stream);

/* L2 variant (north_star; not in the reference, which only has the L-inf clamp at
 * attack/DSGN/pgd_attack.py:346-347).  Per image: adv = v + alpha*g/||g||2,
 * eta = (adv-clean)*min(1, eps/||adv-clean||2).  One tensor set [n_img,C,hw].
 * workspace: b2_pgd_update_l2_workspace_bytes(n_img) bytes.  Deterministic
 * (fixed-order two-stage reductions). */
int64_t b2_pgd_update_l2_workspace_bytes(int n_img);
int b2_pgd_update_l2(const float* x, const float* g, const float* clean, float* out,
                     int n_img, int C, int64_t hw, float alpha, float eps, int denorm,
                     const float* mean, const float* std_, const float* lo, const float* hi,
                     void* workspace, void* stream);

/* Patch blend, attack/DSGN/patch_attack.py:326-333,369-376 (and Stereo-RCNN
 * patch_attack.py:225-230): img = (1-m)*img + m*pad(patch) with the circular mask
 * dist((y,x),(cy,cx)) <= radius computed in-kernel (integer dy^2+dx^2 <= r^2 is
 * exact).  In place on img [n_img,C,H,W]; only the (2r+1)^2 box
 * is touched.  patch [C,2r+1,2r+1]; centers: HOST int array [n_img][2] = (cy,cx). */
int b2_patch_apply(float* img, const float* patch, int n_img, int C, int H, int W,
                   const int* centers, int radius, void* stream);

/* Patch update, attack/DSGN/patch_attack.py:416-430:
 *   d = clamp(0.5*alpha*(gL[box_L] + gR[box_R]), -eps, eps)
 *   patch_out = patch - d, then optional per-channel clamp [lo_c,hi_c]
 *   (attack/Stereo-RCNN/patch_attack.py:272-281; lo/hi HOST arrays or NULL).
 * If delta_out != NULL the clipped step d is written there and patch is left
 * untouched (multi-GPU: all-reduce the deltas, then b2_patch_axpy). */
int b2_patch_update(float* patch, const float* gL, const float* gR, int C, int H, int W,
                    int cyL, int cxL, int cyR, int cxR, int radius, float alpha, float eps,
                    const float* lo, const float* hi, float* delta_out, void* stream);

/* Graph-replayable variants: the patch centres are read from DEVICE memory at run time (patch_apply_dev:
 * int32 [n_img][2] = (cy,cx); patch_update_dev: int32 [4] = (cyL,cxL,cyR,cxR)), so one captured CUDA graph of the
 * patch iteration serves images with different patch positions.  A box that leaves the frame cannot be rejected on
 * the host here: patch_update_dev then leaves the patch unchanged (writes a zero step to delta_out). */
int b2_patch_apply_dev(float* img, const float* patch, int n_img, int C, int H, int W,
                       const int* centers_dev, int radius, void* stream);
int b2_patch_update_dev(float* patch, const float* gL, const float* gR, int C, int H, int W,
                        const int* centers_dev, int radius, float alpha, float eps,
                        const float* lo, const float* hi, float* delta_out, void* stream);

/* Second half of the split patch update (multi-GPU universal patch, BASELINE config 4): after the ranks have
 * all-reduced (sum) the clipped steps written through delta_out, patch = clamp(patch - delta, lo_c, hi_c)
 * (lo/hi HOST arrays or NULL = no clamp).  patch, delta [C,dim,dim].  With one rank
 * b2_patch_update(delta_out) + b2_patch_axpy == b2_patch_update(NULL) bit for bit. */
int b2_patch_axpy(float* patch, const float* delta, int C, int dim, const float* lo, const float* hi, void* stream);

/* ------------------------------------------------------------------------- *
 * (1) Plane-sweep cost volume -- replaces upstream dsgn._C
 *     build_cost_volume_{forward,backward} reached from
 *     attack/DSGN/pgd_attack.py:308 / :336.
 * layout 0 (NCDHW, the upstream layout): left/right [N,C,H,W] -> cost [N,2C,D,H,W]
 * layout 1 (channels-last, the fast internal layout): left/right [N,H,W,C],
 *          cost [N,D,H,W,2C]; C % 4 == 0.
 * shifts: device [N,D] plane disparities in feature px (>= 0, fractional ok; a negative value is
 *         treated as 0 by every kernel, so no gather can leave its row).
 * Backward is gather-form (no atomics) and therefore bitwise deterministic.
 * ------------------------------------------------------------------------- */
int b2_cost_volume_fwd(const float* left, const float* right, const float* shifts, float* cost,
                       int N, int C, int D, int H, int W, int layout, void* stream);
int64_t b2_cost_volume_bwd_workspace_bytes(int N, int C, int H, int W);
int b2_cost_volume_bwd(const float* gcost, const float* shifts, float* gleft, float* gright,
                       int N, int C, int D, int H, int W, int layout, void* workspace, void* stream);
/* workspace (channels-last layout only; may be NULL = slower plane-strided kernel): per-segment partial
 * sums of the row-staged backward, combined in a fixed order. */

/* ------------------------------------------------------------------------- *
 * (2) grid_sample lifting (frustum -> voxel), bilinear/trilinear, zeros padding
 *     -- replaces ATen grid_sampler_{2d,3d}[_backward] reached through
 *     F.grid_sample inside StereoNet.forward (attack/DSGN/pgd_attack.py:308,336).
 * Channels-last tensors.  The output may be a channel slice of a wider tensor:
 * element (voxel v, channel c) lives at out[v*out_cstride + out_coff + c].
 *   3-D: in [N,D,H,W,C], grid [N,Z,Y,X,3] (x->W, y->H, z->D), out voxels N*Z*Y*X
 *   2-D: in [N,H,W,C],   grid [N,Ho,Wo,2]
 * C % 4 == 0 and C <= 128.
 * ------------------------------------------------------------------------- */
int b2_grid_sample3d_fwd(const float* in, const float* grid, float* out,
                         int N, int C, int D, int H, int W, int64_t nvox_per_n,
                         int out_cstride, int out_coff, int align_corners, void* stream);
int b2_grid_sample2d_fwd(const float* in, const float* grid, float* out,
                         int N, int C, int H, int W, int64_t nvox_per_n,
                         int out_cstride, int out_coff, int align_corners, void* stream);

/* Deterministic backward w.r.t. the input: a CSR "input cell -> (output voxel,
 * weight)" plan is built once per grid (the grid is fixed by the calibration, so
 * it amortises over all PGD iterations) and the backward is a pure gather.
 *   step 1  b2_grid_plan_count : counts[cell] (int32, zero-initialised by caller)
 *   (caller: exclusive prefix sum counts -> row_ptr[ncell+1])
 *   step 2  b2_grid_plan_fill  : entries[row_ptr[cell] ...] = (voxel, weight);
 *           cursor = zeroed int32[ncell] scratch
 *   step 3  b2_grid_plan_sort  : sort every row by voxel id (fixes summation order)
 *   bwd     b2_grid_sample_bwd : gin[cell, :] = sum_e w_e * gout[voxel_e, :]
 * ndim = 2 or 3; for ndim 2 pass D = 1.  cells = N*D*H*W; entries are int2
 * {voxel (global, n-major), float bits of weight}. */
int b2_grid_plan_count(const float* grid, int32_t* counts, int ndim, int N, int D, int H, int W,
                       int64_t nvox_per_n, int align_corners, void* stream);
int b2_grid_plan_fill(const float* grid, const int32_t* row_ptr, int32_t* cursor, void* entries,
                      int ndim, int N, int D, int H, int W, int64_t nvox_per_n,
                      int align_corners, void* stream);
int b2_grid_plan_sort(const int32_t* row_ptr, void* entries, int64_t ncell, void* stream);
int b2_grid_sample_bwd(const float* gout, const int32_t* row_ptr, const void* entries, float* gin,
                       int64_t ncell, int C, int gout_cstride, int gout_coff, int long_rows, void* stream);
/* long_rows != 0: a whole warp walks each CSR row (rows of >~ 32 entries, e.g. the 2-D lifting). */

/* Fused lifting forward of DSGN (frustum -> voxel): out[v, 0:64] = trilinear sample of psv [N,D,H,W,64] at
 * grid[v] = (x,y,z), out[v, 64:96] = bilinear sample of img [N,Hi,Wi,32] at grid[v].xy; out [N*nvox_per_n, 96].
 * Same values, bit for bit, as b2_grid_sample3d_fwd + b2_grid_sample2d_fwd into the two channel slices; one
 * launch, the grid read once, 8 lanes per voxel (the generic kernels repeat the corner arithmetic in every one of
 * their 16 lanes).  C3 = 64, C2 = 32 only. */
int b2_lift_fwd(const float* psv, const float* img, const float* grid, float* out, int N, int C3, int C2,
                int D, int H, int W, int Hi, int Wi, int64_t nvox_per_n, int align_corners, void* stream);

/* ------------------------------------------------------------------------- *
 * (3) 3-D convolutions of the hourglass stacks -- replaces cuDNN Conv3d /
 *     ConvTranspose3d forward and data-gradient reached from
 *     attack/DSGN/pgd_attack.py:308 / :336 (weights are frozen in an attack, so
 *     no weight gradient is ever needed).
 * Channels-last activations [N,D,H,W,C]; 3x3x3 kernel, padding 1; weights
 * pre-packed as wp[27][Cout][Cin] (tap = (kd*3+kh)*3+kw).
 *   mode 0 CONV  : out[o] = sum_k in[o*stride + k - 1] . wp[k]      (stride 1|2)
 *   mode 1 DECONV: out[o] = sum_k [(o+1-k) even] in[(o+1-k)/2] . wp[k]  (stride 2,
 *                  i.e. ConvTranspose3d(k3,s2,p1,output_padding=1) and the data
 *                  gradient of a stride-2 conv)
 * Data gradients are the same two modes with repacked weights (see ops.py).
 * impl 0 = tcgen05/TMEM/TMA implicit GEMM (TF32 inputs, fp32 accumulate);
 * impl 1 = fp32 SIMT implicit GEMM (verification mode; any Cin,Cout % 4 == 0).
 * in dims are the INPUT spatial dims; out dims are derived:
 *   CONV: (d+2-3)/stride+1 ; DECONV: 2*d.
 * ------------------------------------------------------------------------- */
int b2_conv3d(const float* in, const float* wp, float* out, int N, int Cin, int Cout,
              int Di, int Hi, int Wi, int stride, int mode, int impl, void* stream);

/* conv3d (impl 0) with a fused epilogue; every extra is optional (NULL / 0 = off):
 *   addend       [same shape as out]: out = conv(in) + addend -- the other gradient that autograd
 *                would add to a data gradient in a separate full pass (a tensor with two consumers);
 *   stat_partial [rows][2][Cout] + stat_mode: per-channel sums over the voxels each CTA writes, in a
 *                fixed order, one table row per CTA, v = the written value (after the addend):
 *       1  GroupNorm FORWARD statistics of this conv's output (upstream convbn_3d = conv + norm):
 *          (sum v, sum v^2); b2_groupnorm_fwd_ext then skips its statistics pass;
 *       2  GroupNorm BACKWARD sums of the norm that produced this conv's INPUT in the forward, when
 *          this launch is that conv's data gradient (v = gradient w.r.t. the norm's output, gn_x =
 *          the norm's input, same shape as out): (sum v*x, sum v); b2_groupnorm_bwd_ext then skips
 *          its statistics pass (a read of gy and of x);
 *       3  as 2 for a norm followed by ReLU: v counts only where fmaf(x, scale_c, shift_c) > 0,
 *          gn_coef = that norm's forward scale[Cout], shift[Cout] (the tail of its stats row).
 * b2_conv3d_fusion_caps reports what the kernel serving this shape can do: *stat_rows = rows of
 * the table (0: not available -- N != 1, stride-2 CONV, widths that are not multiples of 32),
 * *addend_ok = 1/0 (stride-1 CONV and DECONV kernels). */
int b2_conv3d_fusion_caps(int N, int Cin, int Cout, int Di, int Hi, int Wi, int stride, int mode,
                          int* stat_rows, int* addend_ok);
int b2_conv3d_fused(const float* in, const float* wp, float* out, const float* addend, float* stat_partial,
                    int stat_mode, const float* gn_x, const float* gn_coef, int N, int Cin, int Cout,
                    int Di, int Hi, int Wi, int stride, int mode, void* stream);

/* ------------------------------------------------------------------------- *
 * 2-D convolutions of the feature extractor and the BEV head (upstream
 * dsgn feature_extraction / bev_conv / bbox_* layers, stock nn.Conv2d -> cuDNN in
 * the reference) reached from attack/DSGN/pgd_attack.py:308 / :336: forward and
 * data gradient on the tcgen05 tensor cores.
 * Channels-last activations [N,H,W,C]; kernel 1x1 or 3x3, padding = dilation*(k/2).
 *   mode 0 CONV  : out[o] = sum_k in[o*stride + (k - k/2)*dilation] . wp[k]   (stride 1|2; dilation 2 only
 *                  with stride 1)
 *   mode 1 DECONV: ConvTranspose2d(k, stride 2, padding k/2, output_padding 1) = the data gradient of
 *                  a stride-2 conv on an even-sized input; out dims = 2 * in dims.
 * wp: packed weights [S][k*k][Cout][Cin] with S = 2 when split != 0: slab 0 = w_hi (low 13 mantissa
 * bits cleared, exactly TF32-representable), slab 1 = w_lo = w - w_hi; S = 1 (plain w) otherwise.
 * split != 0: error-compensated 3xTF32 (x_hi*w_hi + x_hi*w_lo + x_lo*w_hi, operands split inside
 * the kernel) -- fp32-class accuracy, which is what the reference computes in; split = 0: plain TF32;
 * split = 2: as 1, and the activation tile is also rewritten as its truncated value (verification only);
 * split = 3: the three products accumulate into ONE accumulator, small terms first (A/B of the summation order).
 * bias [Cout] (or NULL) and addend [same layout as out] (or NULL; the second gradient of a tensor with
 * two consumers) are added in the epilogue.  Cin % 32 == 0, Cout % 16 == 0, Cout <= 256.
 * ------------------------------------------------------------------------- */
int b2_conv2d(const float* in, const float* wp, const float* bias, const float* addend, float* out,
              int N, int Cin, int Cout, int Hi, int Wi, int ksize, int stride, int dilation, int mode,
              int split, void* stream);

/* b2_conv2d whose epilogue also adds up GroupNorm sums of the rows it writes (upstream convbn = Conv2d +
 * GroupNorm, the pattern of every extractor / BEV layer), stat_mode as b2_conv3d_fused:
 *   1  (sum v, sum v^2) of this conv's output v (after bias / addend) -> b2_groupnorm_fwd_ext skips its statistics pass;
 *   2  this launch is the data gradient that produces gy w.r.t. the output of a GroupNorm whose input was gn_x
 *      [same layout as out]: (sum gy*x, sum gy) -> b2_groupnorm_bwd_ext skips its statistics pass;
 *   3  as 2 for a norm followed by ReLU: gy counts only where fmaf(x, scale_c, shift_c) > 0; gn_coef = the norm's
 *      forward scale[Cout], shift[Cout] of sample 0 (tail of its stats row), sample n at + n * gn_coef_stride floats.
 * stat_partial: [N][rows][2][Cout] floats, rows = b2_conv2d_stat_rows(...) (one row per CTA and sample, fixed
 * summation order); rows == 0: the kernel serving this shape has no statistics epilogue (DECONV, or a channel
 * tile wider than 128) and stat_mode must be 0. */
int b2_conv2d_fused(const float* in, const float* wp, const float* bias, const float* addend, float* out,
                    int N, int Cin, int Cout, int Hi, int Wi, int ksize, int stride, int dilation, int mode,
                    int split, int stat_mode, float* stat_partial, const float* gn_x, const float* gn_coef,
                    int gn_coef_stride, void* stream);
int b2_conv2d_stat_rows(int N, int Cin, int Cout, int Hi, int Wi, int ksize, int stride, int dilation, int mode,
                        int split, int* rows);

/* First extractor layer, Conv2d(3 -> Cout, k3, stride 2, pad 1), exact fp32: img [N,3,H,W] (NCHW, the
 * attack's image tensor) -> out [N,Ho,Wo,Cout] channels-last, and its data gradient back to the NCHW
 * pixels (the gradient tensor b2_pgd_update consumes).  w [Cout][3][3][3] (the nn.Conv2d layout). */
int b2_conv2d_first_fwd(const float* img, const float* w, float* out, int N, int Cout, int H, int W,
                        void* stream);
int b2_conv2d_first_dgrad(const float* gout, const float* w, float* gimg, int N, int Cout, int H, int W,
                          void* stream);

/* Cout == 1 head (classif1's last layer) and its data gradient: bandwidth-bound,
 * SIMT.  w1 [27][Cin].  fwd: in [N,D,H,W,Cin] -> out [N,D,H,W];
 * dgrad: gout [N,D,H,W] -> gin [N,D,H,W,Cin]. */
int64_t b2_conv3d_c1_workspace_bytes(int N, int D, int H, int W);   /* 27 tap planes, fwd only */
int b2_conv3d_c1_fwd(const float* in, const float* w1, float* out, int N, int Cin,
                     int D, int H, int W, void* workspace, void* stream);
int b2_conv3d_c1_dgrad(const float* gout, const float* w1, float* gin, int N, int Cin,
                       int D, int H, int W, void* stream);

/* GroupNorm (+ residual add) (+ ReLU) on channels-last 3-D volumes, forward and
 * data gradient -- the norm/activation that follows every conv3d (upstream
 * convbn_3d).  y = act(GN(x)*gamma + beta (+ res)).
 *   stats  [N, 2G + 2C]: per sample (mean, rstd) x G, then scale[C], shift[C]; written by fwd,
 *          read by bwd
 *   workspace: b2_groupnorm_workspace_bytes(N, C) bytes
 * bwd: gx (and gres = masked gy when has_res; may alias gy).  relu = 0 none; 1 = mask from the
 * saved forward output y; 2 = mask recomputed from x with the forward's own scale/shift (exactly
 * the forward's fmaf, valid when the forward had no residual) -> y may be NULL and need not be
 * kept alive. */
int64_t b2_groupnorm_workspace_bytes(int N, int C);
int b2_groupnorm_fwd(const float* x, const float* res, const float* gamma, const float* beta,
                     float* y, float* stats, int N, int C, int64_t S, int G, float eps,
                     int relu, void* workspace, void* stream);
/* forward with the per-channel partial sums supplied by the producer (b2_conv3d_fused, b2_conv2d_fused):
 * ext_partial [N][ext_rows][2][C] (sum, sum of squares). */
int b2_groupnorm_fwd_ext(const float* x, const float* res, const float* gamma, const float* beta,
                         float* y, float* stats, int N, int C, int64_t S, int G, float eps,
                         int relu, const float* ext_partial, int ext_rows, void* workspace,
                         void* stream);
int b2_groupnorm_bwd(const float* gy, const float* x, const float* y, const float* gamma,
                     const float* stats, float* gx, float* gres, int N, int C, int64_t S, int G,
                     int relu, void* workspace, void* stream);
/* backward with (sum gz*x, sum gz) per channel supplied by the kernel that produced gy
 * (b2_conv3d_fused / b2_conv2d_fused stat_mode 2 / 3): ext_partial [N][ext_rows][2][C], relu 0 or 2. */
int b2_groupnorm_bwd_ext(const float* gy, const float* x, const float* y, const float* gamma,
                         const float* stats, float* gx, float* gres, int N, int C, int64_t S, int G,
                         int relu, const float* ext_partial, int ext_rows, void* workspace,
                         void* stream);

/* Voxel -> BEV hand-off: AvgPool3d((1,p,1)) over the height axis + fold (C, Y/p) into channels
 * (upstream StereoNet, behind attack/DSGN/pgd_attack.py:308/:336).  v [N,Z,Y,X,C] channels-last ->
 * bev [N,Z,X,C*(Y/p)] channels-last with channel = c*(Y/p) + yy; bwd is the exact adjoint. */
int b2_bev_pool_fwd(const float* v, float* bev, int N, int C, int Z, int Y, int X, int p, void* stream);
int b2_bev_pool_bwd(const float* gbev, float* gv, int N, int C, int Z, int Y, int X, int p, void* stream);

/* ------------------------------------------------------------------------- *
 * Channel concatenation / split of channels-last maps and the gradient merge of a tensor that is also
 * consumed through a batch prefix: the torch.cat of the SPP branches and the slicing around the extractor's
 * two heads / the detection heads (upstream feature_extraction.forward, reached from
 * attack/DSGN/pgd_attack.py:308 / :336).  One streaming launch per direction instead of autograd's strided
 * views + per-consumer copies + zero fills + adds.
 *   channel_concat: wide[row][off_k + c] = srcs[k][row][c]; srcs / widths are HOST arrays of n_pieces (<= 8) device
 *                   pointers / channel counts (multiples of 4); a NULL source contributes zeros; rows = N*H*W.
 *   channel_split : the reverse; a NULL destination is skipped.
 *   add_prefix    : out[i] = a[i] + (i < count_prefix ? b[i] : 0)   (a: gradient of the whole batch, b: gradient
 *                   that arrived through the view of its first samples).
 * ------------------------------------------------------------------------- */
int b2_channel_concat(const float* const* srcs, const int* widths, int n_pieces, float* wide, int64_t rows,
                      void* stream);
int b2_channel_split(const float* wide, float* const* dsts, const int* widths, int n_pieces, int64_t rows,
                     void* stream);
int b2_add_prefix(const float* a, const float* b, float* out, int64_t count, int64_t count_prefix, void* stream);

/* Depth head of the PSV branch (SURVEY 8f "next" row 1), fused and deterministic: trilinear
 * upsample of the 1-channel cost volume cost [N,D,Hc,Wc] to (J,H,W) (align_corners=False),
 * softmax over the J planes, expectation over z_j = z0 + (j+0.5)*dz  ->  depth [N,H,W].
 * Replaces F.interpolate + F.softmax + (prob*z).sum of upstream StereoNet, whose output is read
 * at attack/DSGN/pgd_attack.py:310-317.  bwd: gdepth [N,H,W] -> gcost [N,D,Hc,Wc];
 * workspace b2_depth_head_workspace_bytes(N,D,H,W) bytes.
 * sm_stats [N,H,W,2] (optional, NULL = off): the forward saves the softmax (max, sum) per pixel and
 * the backward, given them and the forward's depth, skips its own softmax pass (same values, so
 * the gradient is bit-identical either way). */
int b2_depth_head_fwd(const float* cost, float* depth, float* sm_stats, int N, int D, int Hc, int Wc,
                      int H, int W, int J, float z0, float dz, void* stream);
int64_t b2_depth_head_workspace_bytes(int N, int D, int H, int W);
int b2_depth_head_bwd(const float* cost, const float* gdepth, float* gcost, const float* sm_stats,
                      const float* depth, int N, int D, int Hc, int Wc, int H, int W, int J, float z0,
                      float dz, void* workspace, void* stream);

/* ------------------------------------------------------------------------- *
 * RoIAlign forward / deterministic backward (Stereo R-CNN, config 5) -- replaces
 * upstream model.roi_layers.ROIAlign constructed at
 * attack/Stereo-RCNN/stereo_rcnn.py:44-45 and dispatched per FPN level at
 * :110-141.  feat [1,C,H,W] (NCHW, as the upstream op), rois [R,5]
 * (batch,x1,y1,x2,y2), out [R,C,P,P]; legacy (unaligned) sampling, adaptive
 * sampling ratio (0).  Backward is gather-form over feature pixels (fixed summation order, no atomics);
 * gout_layout 0 = [R,C,P,P] as the upstream op hands it back, 1 = [R,P,P,C] (channels-last memory: the
 * C values of one bin are contiguous, which is what the warp-per-pixel kernel reads fastest).
 * ------------------------------------------------------------------------- */
int b2_roi_align_fwd(const float* feat, const float* rois, float* out, int R, int C, int H, int W,
                     int P, float scale, void* stream);
int b2_roi_align_bwd(const float* gout, const float* rois, float* gfeat, int R, int C, int H, int W,
                     int P, float scale, int gout_layout, void* stream);

/* The whole FPN dispatch of _StereoRCNN.PyramidRoI_Feat (attack/Stereo-RCNN/stereo_rcnn.py:110-141) in ONE launch
 * per direction: every RoI's level = clamp(round(log(sqrt(h w) / 224) + 4), 2, 5) (natural log, :113-119) is
 * evaluated in the kernel, the RoI is pooled from feats[level - 2] with scale = Hs[level - 2] / im_h (:131), and its
 * row of out [R,C,P,P] is written in the ORIGINAL RoI order -- no per-level index lists, concatenation or re-sort
 * (:121-141), and no host synchronisation.  feats / gfeats: HOST arrays of 4 device pointers (levels 2..5, each
 * [1,C,Hs[l],Ws[l]]); Hs, Ws: HOST int[4].  Backward: gather form over the pixels of all four maps, deterministic;
 * every gfeats[l] is fully written. */
int b2_roi_align_pyramid_fwd(const float* const* feats, const int* Hs, const int* Ws, const float* rois, float* out,
                             int R, int C, int P, float im_h, void* stream);
int b2_roi_align_pyramid_bwd(const float* gout, const float* rois, float* const* gfeats, const int* Hs, const int* Ws,
                             int R, int C, int P, float im_h, int gout_layout, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B2ATTACK_H */
