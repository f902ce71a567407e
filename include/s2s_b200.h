/* stain2stain_b200 -- C ABI of the B200 (sm_100a) kernels behind the flow-matching hot path.
 *
 * The reference (nirschl-lab/stain2stain) has no FFI: its hot path is stock PyTorch called from
 *   src/models/conditional_flow_matching.py:53-74   (model_step: FM sample -> net(t, xt) -> MSE)
 *   src/models/conditional_flow_matching.py:133-170 (generate: NeuralODE.trajectory)
 * and the arithmetic lives in torchcfm 1.0.7 (UNetModel) / torchdyn 1.0.6 (odeint).  Each entry point below names the
 * ATen/torchcfm operation(s) of that path it replaces.  The host side (stain2stain_b200/*.py) binds these with ctypes;
 * INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - all pointers are DEVICE pointers (caller-owned, no allocation inside), `stream` is a cudaStream_t passed as void*;
 *   - activations are 16-bit NHWC (fp16 forward / bf16 gradients, see S2S_FMT_*), parameters/gradients fp32 in the reference's own layouts (OIHW conv weights),
 *     image-space tensors (x0, x1, xt, v) fp32 NCHW exactly as the reference passes them;
 *   - every function is stream-ordered and re-entrant, returns 0 on success or a negative code; s2s_last_error()
 *     returns a thread-local message for the last failure.  Nothing falls back to the CPU.
 */
#ifndef S2S_B200_H
#define S2S_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S2S_OK 0
#define S2S_ERR_INVALID (-1) /* unsupported shape / argument               */
#define S2S_ERR_CUDA (-2)    /* CUDA runtime / driver error                */
#define S2S_ERR_TMAP (-3)    /* cuTensorMapEncodeTiled rejected a tensor   */

/* 16-bit storage formats ("fmt" arguments).  Forward activations / forward weights default to fp16, gradients to
 * bf16; both feed the same tcgen05 kind::f16 MMA (formats are set per operand in the instruction descriptor). */
#define S2S_FMT_BF16 0
#define S2S_FMT_F16 1

const char* s2s_last_error(void);
int s2s_abi_version(void);
int s2s_num_sms(void);

/* One input of a fused convolution: bf16 NHWC [B, Hout*stride, Wout*stride, C]. taps = 9 (3x3, pad 1) or 1 (1x1). */
typedef struct {
    const void* x;
    int C;
    int taps;
    int stride;
} s2s_conv_src;

/* torch.nn.Conv2d weights (OIHW fp32) -> bf16 K-major GEMM operand rows [ld_k].
 * transpose_flip = 0: forward operand  dst[co][k_off + tap*ci_count + (ci - ci_begin)] = w[co][ci][tap]
 * transpose_flip = 1: dgrad operand    dst[ci - ci_begin][k_off + tap*Cout + co]       = w[co][ci][taps-1-tap]
 * Replaces: the implicit weight handling of F.conv2d / conv backward (torchcfm unet.py conv_nd). */
int s2s_pack_conv_weight(const float* w_oihw, int Cout, int Cin, int taps, int ci_begin, int ci_count, void* dst_bf16,
                         int ld_k, int k_off, int transpose_flip, int fmt, void* stream);

/* Fused implicit-GEMM convolution (tcgen05 + TMA).
 *   acc[b,y,x,n] = sum_s sum_tap sum_c srcs[s][b, y*stride+dy, x*stride+dx, c] * w[n][k]     (k = segment, tap, c order)
 *   out_bf16 (NHWC, Cout % 64 == 0)  = acc + bias + residual                                  when out_bf16 != NULL
 *   out_f32  (NCHW, Cout <= 16)      = axpy_x + axpy_a * (acc + bias)   (axpy_x NULL: acc+bias) when out_f32  != NULL
 * Replaces: F.conv2d (3x3 s1/s2, 1x1), the ResBlock `skip_connection(x) + h` add, torch.cat before a conv (two srcs),
 * the head conv + Euler update `x + dt * v` of torchdyn's fixed-step solver.
 * stats_out (may be NULL; fp32 [B][s2s_conv_stat_tiles(Hout,Wout,Cout)][Cout][2]): per 128-pixel sub-tile (sum, sumsq)
 * of the STORED 16-bit outputs, taken from the epilogue's staging tile -- the GroupNorm statistics of the consumer
 * without a pass over the tensor (fold with s2s_gn_coef_parts).  Only the CTA-pair path provides them. */
int s2s_conv_fwd(const s2s_conv_src* srcs, int nsrc, int B, int Hout, int Wout, const void* w_packed, int Ktot,
                 int Cout, const float* bias, const void* residual, void* out_bf16, float* out_f32,
                 const float* axpy_x, float axpy_a, float* stats_out, int a_fmt, int w_fmt, int out_fmt, int res_fmt,
                 void* stream);
/* s2s_conv_fwd with the normalisation that PRECEDES the conv fused into its operand path (inference): segment i with
 * norms[i].coef != NULL is the RAW tensor, and the kernel rewrites every landed shared-memory tile in place as
 * act(x * A + Bc) with coef = fp32 [B][ld][2] (A, Bc) as produced by s2s_gn_coef / s2s_gn_coef_parts (GroupNorm32 +
 * FiLM folded), channel c of the segment at coef[b][off + c]; act must be 1 (SiLU) and a_fmt fp16: the prologue
 * evaluates silu(z) = z/2 + z/2*tanh(z/2) with one tanh.approx.f16x2 per channel pair (|error| <= |z|/2 * 2^-11).  Replaces, for the conv that follows:
 * torchcfm ResBlock `in_layers[0:2]` / `out_layers[0:2]` (GroupNorm32 -> (FiLM) -> SiLU) in eval mode, i.e. the whole
 * s2s_gn_apply pass.  Only for geometries where s2s_conv_norm_fusable() != 0 (halo-tiled CTA-pair kernel, 3x3 stride-1
 * normalised segments); everything else about the call is s2s_conv_fwd's. */
typedef struct {
    const float* coef;
    int ld;
    int off;
} s2s_conv_norm;
int s2s_conv_norm_fusable(const s2s_conv_src* srcs, int nsrc, int Cout);
int s2s_conv_fwd_norm(const s2s_conv_src* srcs, const s2s_conv_norm* norms, int act, int nsrc, int B, int Hout, int Wout,
                      const void* w_packed, int Ktot, int Cout, const float* bias, const void* residual, void* out_bf16,
                      float* stats_out, int a_fmt, int w_fmt, int out_fmt, int res_fmt, void* stream);

/* Sub-tiles per sample of the epilogue statistics for this output geometry; 0 = not available (use s2s_gn_stats).
 * s2s_conv_stat_tiles: the plain CTA-pair kernel's geometry; s2s_conv_stat_tiles_for: the geometry of the kernel
 * s2s_conv_fwd will pick for these segments (the halo-tiled pair kernel takes stride-1 convs with a 3x3 segment). */
int s2s_conv_stat_tiles(int Hout, int Wout, int Cout);
int s2s_conv_stat_tiles_for(const s2s_conv_src* srcs, int nsrc, int Hout, int Wout, int Cout);
/* _for = _geom unless the experiment switch S2S_EPI_STATS_MINK is set (> 0): then convs with a GEMM K below it answer 0 and
 * the consumer runs s2s_gn_stats (measured neutral, off by default); _geom answers what the kernel CAN emit regardless. */
int s2s_conv_stat_tiles_geom(const s2s_conv_src* srcs, int nsrc, int Hout, int Wout, int Cout);

/* Weight gradient of one conv segment (tcgen05, split-K over pixels, fp32 reductions):
 *   dw[tap][m][n_off + n] += sum_{b,y,x} dy[b,y,x,m] * x[b, y*stride+dy, x*stride+dx, n]
 * dy: 16-bit NHWC [B,Hout,Wout,Cm]; x: 16-bit NHWC [B,Hout*stride,Wout*stride,Cq]; dw: fp32 [taps][Cm][ldn].
 * dy_fmt must equal x_fmt (the MMA takes one operand format; mixed fp16 x bf16 is an illegal instruction on sm_100a).
 * Replaces: the weight half of conv2d backward (cudnn wgrad). */
int s2s_conv_wgrad(const void* dy, int Cm, const void* x, int Cq, int taps, int stride, int B, int Hout, int Wout,
                   float* dw, int ldn, int n_off, int dy_fmt, int x_fmt, void* stream);

/* s2s_pack_conv_weight for many weights in ONE launch (every cached GEMM operand is stale after an optimizer step).
 * jobs_dev: device table of jobs; work_dev: device list of (job index, tile index) int pairs, one CTA each, tile index in
 * [0, s2s_pack_tiles(Cout, ci_count, transpose_flip)) of that job (a tile = 16 destination rows x 64 destination-contiguous
 * elements x all taps, staged through shared memory so that both the OIHW reads and the operand writes are coalesced).
 * The zero padding of the destinations is not touched. */
typedef struct {
    const float* w;
    void* dst16;
    int Cout, Cin, taps, ci_begin, ci_count, ld_k, k_off, transpose_flip, fmt, mode; /* mode: see s2s_pack_conv_weight_mode */
} s2s_pack_job;
int s2s_pack_tiles(int Cout, int ci_count, int transpose_flip);
int s2s_pack_conv_weight_multi(const s2s_pack_job* jobs_dev, const int* work_dev, int n_work, void* stream);

/* s2s_pack_conv_weight with a tap-combination mode.  mode 0: plain.  mode 1 + phase (phase = py*2 + px; 3x3 weights):
 * operand of the phase-decomposed Upsample conv -- 4 logical taps ti = ai*2 + bi, each the fp32 SUM of the 3x3 taps
 * (rows R(py, ai) x columns R(px, bi), R(0,0)={0}, R(0,1)={1,2}, R(1,0)={0,1}, R(1,1)={2}) that read the same
 * low-resolution pixel:  forward  dst[co][k_off + ti*ci_count + ci] ; transpose_flip = 1  dst[ci][k_off + ti*Cout + co]. */
int s2s_pack_conv_weight_mode(const float* w_oihw, int Cout, int Cin, int taps, int ci_begin, int ci_count, void* dst16,
                              int ld_k, int k_off, int transpose_flip, int fmt, int mode, void* stream);

/* Upsample(nearest x2) -> conv3x3(pad 1) WITHOUT the upsampled tensor (torchcfm unet.py Upsample.forward:
 * F.interpolate(x, scale_factor=2, mode="nearest"); self.conv(x) -- SURVEY.md A.2, row a12).  Four phase launches of the
 * halo-tiled CTA-pair conv over the LOW-resolution input x [B,H,W,C], each a 2x2 conv with tap-summed weights writing the
 * pixels (2y+py, 2x+px) of out [B,2H,2W,Cout] through a strided TMA store map: 4/9 of the reference's MACs.
 *   w_packed (fwd)   16-bit [Cout][16*C]: phase p at columns [p*4*C, (p+1)*4*C)   (s2s_pack_conv_weight_mode, mode 1+p)
 *   w_packed (dgrad) 16-bit [Cin][16*Cm]: same column layout, transpose_flip = 1
 *   dw16             fp32 [16 = phase*4 + ti][Cm][Cq], zero-initialised by the caller; s2s_upconv_unpack_wgrad folds it
 *                    into the OIHW gradient [Cm][Cq][3][3] (every 3x3 tap belongs to exactly one logical tap per phase)
 *   stats_out        optional fp32 [B][s2s_upconv_stat_tiles(H,W,Cout)][Cout][2], as in s2s_conv_fwd
 * s2s_upconv_supported: 1 when both channel counts are multiples of 128 (else use s2s_upsample2x + s2s_conv_fwd). */
int s2s_upconv_supported(int C, int Cout);
int s2s_upconv_stat_tiles(int H, int W, int Cout);
int s2s_upconv_fwd(const void* x, int B, int H, int W, int C, const void* w_packed, int Cout, const float* bias, void* out,
                   float* stats_out, int a_fmt, int w_fmt, int out_fmt, void* stream);
int s2s_upconv_dgrad(const void* dy, int B, int H, int W, int Cm, const void* w_packed, int Cin, void* dx, int a_fmt,
                     int w_fmt, int out_fmt, void* stream);
int s2s_upconv_wgrad(const void* dy, int Cm, const void* x, int Cq, int B, int H, int W, float* dw16, int dy_fmt, int x_fmt,
                     void* stream);
int s2s_upconv_unpack_wgrad(const float* dw16, int M, int N, float* grad_oihw, void* stream);

/* Data gradient of Downsample = conv3x3(stride 2, pad 1) (torchcfm unet.py Downsample.op) without zero insertion: the four
 * input-pixel phases (2y+py, 2x+px) are 1-, 2-, 2- and 4-tap convs over the low-resolution gradient dy [B,H,W,Cm], stored
 * into dx [B,2H,2W,Cin] through strided TMA maps.  w_packed = the ordinary dgrad operand of s2s_pack_conv_weight
 * (transpose_flip = 1): 16-bit [Cin][9*Cm].  Needs Cm % 128 == 0 and Cin % 128 == 0 (s2s_upconv_supported). */
int s2s_downconv_dgrad(const void* dy, int B, int H, int W, int Cm, const void* w_packed, int Cin, void* dx, int a_fmt,
                       int w_fmt, int out_fmt, void* stream);

/* dw fp32 [taps][M][ldn] -> grad_oihw[m][n_begin + n][tap] = beta * grad + dw[tap][m][n_off + n] */
int s2s_unpack_wgrad(const float* dw, int taps, int M, int ldn, int n_off, int n_count, float* grad_oihw,
                     int Cin_total, int n_begin, float beta, void* stream);

/* fp32 NCHW [B,3,H,W] -> bf16 NHWC [B,H,W,64] 3x3 patches (channel tap*3+c, zero padded), sgn = +1 (stem conv
 * operand) or -1 (head-conv backward operand).  With x1 != NULL the source is the flow-matching interpolant
 * xt = (1 - t_b) x0 + t_b x1 (torchcfm ConditionalFlowMatcher.sample_xt, sigma = 0), optionally also written to
 * xt_out (fp32 NCHW). */
int s2s_patch27_pack(const float* x0, const float* x1, const float* t, int B, int H, int W, int sgn, void* dst_bf16,
                     float* xt_out, int fmt, void* stream);

/* Number of pixel chunks the normalisation kernels cut one sample into (a function of B and HW only). */
int s2s_gn_chunks(int B, int HW);

/* GroupNorm(32, C) statistics of a 16-bit NHWC tensor, deterministic two-stage reduction:
 * stats[b][chunk][c_off + c] = (sum, sumsq) over the chunk's pixels.  stats: fp32 [B][s2s_gn_chunks][Ctot][2] (no
 * zeroing needed; every source of a concat writes its own channel range).  Replaces pass 1 of ATen native_group_norm. */
int s2s_gn_stats(const void* x, int B, int HW, int C, float* stats, int Ctot, int c_off, int x_fmt, void* stream);

/* Folds the chunk partials, then per-(sample, channel) coefficients:
 * A = rstd*gamma*(1+scale), Bc = (beta - mean*rstd*gamma)*(1+scale) + shift.
 * film: fp32 [B][2C] = ResBlock emb_layers output (scale | shift) or NULL.  coef: [B][C][2], mean_rstd: [B][G][2]. */
int s2s_gn_coef(const float* stats, const float* gamma, const float* beta, const float* film, int B, int C, int G,
                int HW, float eps, float* coef, float* mean_rstd, void* stream);

/* The same fold for partial statistics that arrive per source of a channel concat (conv epilogue statistics):
 * stats_i: fp32 [B][nchunks][Ci][2]; stats1 may be NULL (C1 = 0). */
int s2s_gn_coef_parts(const float* stats0, int C0, const float* stats1, int C1, int nchunks, const float* gamma,
                      const float* beta, const float* film, int B, int G, int HW, float eps, float* coef,
                      float* mean_rstd, void* stream);

/* y[b,p,c_off+c] = dropout(act(x[b,p,c]*A + Bc)); y row stride ld_out channels (concat written in place).
 * The `silu` argument of the four streaming kernels is the activation: 0 = none, 1 = SiLU, 2 = ReLU.
 * Replaces: GroupNorm32 apply + `* (1 + scale) + shift` + SiLU + Dropout (+ torch.cat) of torchcfm ResBlock.
 * y2_bf16 (may be NULL): the same values additionally stored as bf16 with the same geometry -- the weight-gradient
 * GEMM operand of the consuming conv (one MMA cannot mix fp16 x bf16), for +2 B/element instead of a conversion pass.
 * mask_out (may be NULL; uint8 [B*HW*ld_out/8]): the dropout keep bits, one byte per 8 channels, so that backward
 * reads 1 bit / element instead of re-evaluating the Philox hash (mask_in of the backward kernels; NULL = re-hash). */
int s2s_gn_apply(const void* x, int B, int HW, int C, const float* coef, int Ctot, int c_off, void* y, void* y2_bf16,
                 int ld_out, int silu, float drop_p, uint64_t seed, void* mask_out, int x_fmt, int y_fmt, void* stream);
/* s2s_gn_apply for launches replayed from a CUDA graph: seed_step_dev (may be NULL) points at a device-resident uint64 step
 * counter that is mixed into `seed` (seed + counter * 0x9E3779B97F4A7C15), so every replay draws a fresh dropout mask
 * although the host-side argument is frozen at capture. */
int s2s_gn_apply_step(const void* x, int B, int HW, int C, const float* coef, int Ctot, int c_off, void* y, void* y2_bf16,
                      int ld_out, int silu, float drop_p, uint64_t seed, const uint64_t* seed_step_dev, void* mask_out,
                      int x_fmt, int y_fmt, void* stream);

/* Backward of the fused normalisation.  g = dL/dy (bf16 NHWC, row stride ld_g; g_fmt must be S2S_FMT_BF16 for the two
 * streaming kernels -- gradients are bf16 everywhere in the engine, anything else is refused with S2S_ERR_INVALID).
 *   reduce: red_part[b][chunk][c_off+c] = (sum dz, sum dz*xhat) over the chunk   (fp32 [B][s2s_gn_chunks][Ctot][2])
 *   coef  : folds red_part into red[B][C][2]; pqr[b][c] = (P,Q,R,0); dgamma/dbeta += ...; dfilm[b][2C] = (dscale | dshift)
 *   apply : dx[b,p,c] = dz*P + x*Q + R (+ add[b,p,c])                      (bf16 NHWC [B,HW,C]) */
int s2s_gn_bwd_reduce(const void* x, const void* g, int ld_g, int B, int HW, int C, const float* coef,
                      const float* mean_rstd, int G, int Ctot, int c_off, float* red, int silu, float drop_p,
                      uint64_t seed, const void* mask_in, int x_fmt, int g_fmt, void* stream);
/* s2s_gn_bwd_reduce that additionally stores x as bf16 NHWC [B,HW,C] (x_bf16_out, may be NULL): the weight-gradient
 * operand of a 1x1 skip conv over the raw block input, for +2 B/element instead of an s2s_convert16 pass. */
int s2s_gn_bwd_reduce_x2(const void* x, const void* g, int ld_g, int B, int HW, int C, const float* coef,
                         const float* mean_rstd, int G, int Ctot, int c_off, float* red, int silu, float drop_p,
                         uint64_t seed, const void* mask_in, void* x_bf16_out, int x_fmt, int g_fmt, void* stream);
int s2s_gn_bwd_coef(const float* red_part, float* red, const float* mean_rstd, const float* gamma, const float* beta,
                    const float* film, int B, int C, int G, int HW, float* pqr, float* dgamma, float* dbeta,
                    float* dfilm, void* stream);
int s2s_gn_bwd_apply(const void* x, const void* g, int ld_g, int B, int HW, int C, const float* coef, const float* pqr,
                     int Ctot, int c_off, const void* add, void* dx, int silu, float drop_p, uint64_t seed,
                     const void* mask_in, int x_fmt, int g_fmt, void* stream);

/* nearest x2 upsample (F.interpolate(scale_factor=2, mode="nearest")), its adjoint, and zero insertion (the adjoint of
 * a stride-2 subsampling), all bf16 NHWC with C % 8 == 0.  H, W are the SMALL spatial dims. */
int s2s_upsample2x(const void* in, void* out, int B, int H, int W, int C, void* stream);
int s2s_sumpool2x(const void* in, void* out, int B, int H, int W, int C, int fmt, void* stream);
int s2s_zero_insert2x(const void* in, void* out, int B, int H, int W, int C, void* stream);

/* out[c] += sum over pixels of bf16 NHWC x (conv bias gradient). */
int s2s_channel_sum(const void* x, long long npix, int C, float* out, int fmt, void* stream);

/* loss += mean((v - (x1 - x0))^2); dv = 2 (v - (x1 - x0)) / n (may be NULL).  fp32 NCHW, n elements.
 * Replaces: `ut = x1 - x0`, `torch.mean((vt - ut) ** 2)` and its backward (conditional_flow_matching.py:66,72). */
int s2s_fm_loss(const float* v, const float* x0, const float* x1, long long n, float* loss, float* dv, void* stream);

/* ---- image-space kernels either side of the UNet (f1 input pipeline, f3 mask variants, f4 output) ------- */

/* Stem-conv operand for 3- or 4-channel inputs: x0/x1 fp32 NCHW [B,Cx,H,W], extra fp32 [B,1,H,W] or NULL (the
 * condition mask that conditional_flow_matching_conditional_mask.py:54-66 concatenates as a 4th channel -- here it is
 * never concatenated, and never interpolated).  dst: 16-bit NHWC [B,H,W,64], column tap*CT + c, CT = Cx + (extra!=NULL). */
int s2s_patch_pack(const float* x0, const float* x1, const float* t, const float* extra, int B, int Cx, int H, int W,
                   void* dst16, int fmt, void* stream);

/* Mask-weighted flow-matching loss (src/models/conditional_flow_matching_masked.py:76-92): w = 1 + lam*mask[b,0,p];
 * sums[0] += sum w (v-(x1-x0))^2, sums[1] += sum w (over B*C*HW); dv = 2 w (v-(x1-x0)) (may be NULL; the caller scales
 * by 1/(sums[1]+1e-8)).  sums: fp32 [2], zeroed by the caller. */
int s2s_fm_loss_weighted(const float* v, const float* x0, const float* x1, const float* mask, float lam, int B, int C,
                         int HW, float* sums, float* dv, void* stream);

/* ROI Charbonnier term (src/models/conditional_flow_matching_ROI_loss.py:73-97) on xt = t x1 + (1-t) x0 against x1:
 * sums[0] += sum sqrt((xt-x1)^2 + eps^2) * m, sums[1] += sum m (per pixel).  sums: fp32 [2], zeroed by the caller. */
int s2s_roi_charbonnier(const float* x0, const float* x1, const float* t, const float* mask, int B, int C, int HW,
                        float eps, float* sums, void* stream);

/* Input side (src/data/paired_data_module.py:171-199): uint8 HWC tile pair [B,Hs,Ws,3] (+ optional uint8 mask
 * [B,Hs,Ws]) -> crop(top,left,S,S) -> hflip -> vflip -> to_tensor -> Normalize(0.5,0.5) -> fp32 NCHW [B,3,S,S],
 * bit-identical to the torchvision chain.  params: int32 [B][4] = (top, left, hflip, vflip); flags bit 0: the colour
 * bytes are in cv2.imread (BGR) order; flags bit 1: the mask is written as its raw byte value (class ids,
 * src/data/paired_data_multiclassmask.py:113-128) instead of mask / 255.  tgt/out1 and mask/outm may be NULL. */
int s2s_tile_prep(const uint8_t* src, const uint8_t* tgt, const uint8_t* mask, const int* params, int B, int Hs, int Ws,
                  int S, int flags, float* out0, float* out1, float* outm, void* stream);

/* One pass of Pillow's antialiased 8-bit resampling (TF.resize on a PIL image, paired_data_module.py:201-203):
 * out = clip8((2^21 + sum_k in[first+k]*kk[k]) >> 22).  bounds int32 [n_out][2] = (first, count), kk int32
 * [n_out][ksize] (host-built like Pillow's precompute_coeffs / normalize_coeffs_8bpc).  in uint8 [B,Hin,Win,C];
 * vertical == 0: out [B,Hin,n_out,C]; vertical != 0: out [B,n_out,Win,C]. */
int s2s_resample_u8(const uint8_t* in, int B, int Hin, int Win, int C, const int* bounds, const int* kk, int ksize,
                    int n_out, int vertical, uint8_t* out, void* stream);

/* Output side (src/infer_simple_flowmatching.py:37-38, 86-88): fp32 NCHW [B,C,H,W] -> uint8 NHWC [B,H,W,C] =
 * floor(clamp(x*0.5+0.5, 0, 1)*255 + 0.5). */
int s2s_denorm_u8(const float* x, int B, int C, int HW, uint8_t* out, void* stream);

/* 16-bit storage format conversion of n elements (n % 8 == 0), e.g. fp16 saved activations -> bf16 wgrad operands
 * (the two operands of one tcgen05 kind::f16 MMA must share a format). */
int s2s_convert16(const void* in, void* out, long long n, int in_fmt, int out_fmt, void* stream);

int s2s_nchw_f32_to_nhwc16(const float* in, void* out, int B, int C, int HW, int fmt, void* stream);
int s2s_nhwc16_to_nchw_f32(const void* in, float* out, int B, int C, int HW, int fmt, void* stream);

/* ---- multitask model (config M): src/models/components/shared_encoder.py, task_decoders.py ------------------------ */

/* Train-mode torch.nn.BatchNorm2d folds.  The streaming passes are s2s_gn_stats / s2s_gn_apply / s2s_gn_bwd_reduce /
 * s2s_gn_bwd_apply with act = 2 (ReLU) and G = C; these two kernels fold their per-chunk partials over the BATCH:
 *   bn_coef    : stats [B][nchunks][C][2] -> coef [B][C][2] (A = gamma*rstd, Bc = beta - mean*A, same for every b),
 *                mean_rstd [B][C][2], running_mean/var <- (1-momentum)*running + momentum*(mean, unbiased var) (may be NULL)
 *   bn_bwd_coef: red [B][nchunks][C][2] -> pqr [B][C][4], dgamma += sum dz*xhat, dbeta += sum dz
 * Replaces: nn.BatchNorm2d (+ nn.ReLU) forward / backward in DoubleConv (shared_encoder.py:14-21, task_decoders.py:14-21). */
int s2s_bn_coef(const float* stats, int B, int nchunks, int C, int HW, const float* gamma, const float* beta, float eps,
                float momentum, float* running_mean, float* running_var, float* coef, float* mean_rstd, void* stream);
int s2s_bn_bwd_coef(const float* red, int B, int nchunks, int C, int HW, const float* mean_rstd, const float* gamma,
                    float* pqr, float* dgamma, float* dbeta, void* stream);

/* Two-level folds and torch.nn.SyncBatchNorm (what `sync_batchnorm: True` of configs/trainer/ddp.yaml:9 makes Lightning
 * install in place of every BatchNorm2d).
 *   bn_fold         : parts [nparts][C][2] -> sums [nslices][C][2], slice s = the parts [s*per, (s+1)*per) folded in a fixed
 *                     order (per = ceil(nparts / nslices); s2s_bn_fold_slices(nparts) is the slice count the host wrappers use:
 *                     the statistics of a 512^2 layer are 32768 partials per channel, far too many for the C/32 CTAs of the
 *                     coefficient kernels).  nslices = 1: this rank's per-channel totals, which SyncBatchNorm all-reduces.
 *   bn_coef_sums    : bn_coef over `nparts` rows of sums with an explicit element count per channel (the GLOBAL count for
 *                     SyncBatchNorm: coef / mean_rstd for the B local samples, running statistics from the global mean /
 *                     unbiased variance, identical on every rank)
 *   bn_bwd_coef_sums: likewise for (sum dz, sum dz*xhat) -> pqr; dgamma / dbeta += the folded sums (SyncBatchNorm callers pass
 *                     scratch vectors here and take the LOCAL parameter gradients from their own bn_fold result, as torch's
 *                     SyncBatchNorm does -- DDP averages them). */
int s2s_bn_fold_slices(int nparts);
int s2s_bn_fold(const float* parts, int nparts, int C, float* sums, int nslices, void* stream);
int s2s_bn_coef_sums(const float* sums, int nparts, int C, long long count, int B, const float* gamma, const float* beta,
                     float eps, float momentum, float* running_mean, float* running_var, float* coef, float* mean_rstd,
                     void* stream);
int s2s_bn_bwd_coef_sums(const float* sums, int nparts, int C, long long count, int B, const float* mean_rstd,
                         const float* gamma, float* pqr, float* dgamma, float* dbeta, void* stream);

/* nn.MaxPool2d(2) (shared_encoder.py:32-34) and its backward (gradient to the first maximum in row-major order, as ATen);
 * 16-bit NHWC, H, W = OUTPUT spatial dims. */
int s2s_maxpool2x(const void* in, void* out, int B, int H, int W, int C, int fmt, void* stream);
int s2s_maxpool2x_bwd(const void* x, const void* g, void* dx, int B, int H, int W, int C, int x_fmt, int g_fmt,
                      void* stream);

/* nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True) (task_decoders.py:34) and its adjoint; 16-bit NHWC,
 * H, W = INPUT (small) spatial dims. */
int s2s_bilinear2x(const void* in, void* out, int B, int H, int W, int C, int fmt, void* stream);
int s2s_bilinear2x_bwd(const void* g, void* din, int B, int H, int W, int C, int fmt, void* stream);

/* fp32 NCHW [B, C, HW] -> 16-bit NHWC [B, HW, Cpad], zero channels beyond C (narrow image-space gradients as GEMM operands). */
int s2s_nchw_f32_to_nhwc16_pad(const float* in, void* out, int B, int C, int Cpad, int HW, int fmt, void* stream);

/* Segmentation loss pieces: softmax Dice (sums over the WHOLE batch) + cross-entropy on fp32 NCHW logits [B, C, HW]
 * (C <= 8), int64 targets [B, HW].  sums: double[3C + 2] = I_c | P_c | T_c | sum(-log p_t) | #non-ignored, accumulated
 * (zero it first).  bwd writes dlogits = *gscale * (w_dice * dDice + w_ce * dCE).
 * Replaces: MulticlassDiceLoss.forward + nn.CrossEntropyLoss (conditional_flow_matching_multitask_multiclassloss.py:31-83, 231-236). */
int s2s_seg_loss_sums(const float* logits, const long long* target, int B, int C, int HW, long long ignore_index,
                      double* sums, void* stream);
int s2s_seg_loss_bwd(const float* logits, const long long* target, int B, int C, int HW, long long ignore_index,
                     const double* sums, float smooth, float w_dice, float w_ce, const float* gscale, float* dlogits,
                     void* stream);

/* The UNet's output projection (torchcfm UNetModel.out[2] = conv3x3(C -> out_channels)): a 16-bit NHWC
 * [B,H,W,C] -> out fp32 NCHW [B,Cout,H,W] = bias + conv3x3(a, w_oihw) with the fp32 nn.Conv2d weight itself (converted to the
 * activation format inside the kernel), Cout <= 8, C % 16 == 0.  Optional fused Euler update: out = axpy_x + axpy_a * (...)
 * (out may alias axpy_x) -- torchdyn's fixed-step `x + dt * f(t, x)`.  Reads the input once (halo tile in shared memory)
 * instead of once per filter tap. */
int s2s_head_conv(const void* a, int B, int H, int W, int C, const float* w_oihw, int Cout, const float* bias, float* out,
                  const float* axpy_x, float axpy_a, int a_fmt, void* stream);

/* ---- attention core (rows a14, f3) ----------------------------------------------------------------------------------- */

/* torchcfm AttentionBlock's QKVAttentionLegacy / QKVAttention: per (sample, head)
 *   a = softmax_fp32(q k^T / sqrt(ch)) v        over T = H*W tokens
 * reading q / k / v in place from the qkv 1x1 conv's NHWC output qkv [B][T][3*heads*ch] (new_order = 0: channel =
 * head*3*ch + {q,k,v}*ch + c -- the legacy interleave every reference config uses; new_order = 1: {q,k,v}*C + head*ch + c)
 * and writing out [B][T][heads*ch] (channel head*ch + c).  lse (may be NULL for inference): fp32 [B*heads][T] log2-domain
 * logsumexp kept for the backward.  Flash-style: the T x T matrix never leaves registers; fp32 softmax statistics.
 * Backward: d_qkv [B][T][3*heads*ch] (g_fmt) from d_out [B][T][heads*ch] (g_fmt), the saved qkv / out (a_fmt) and lse;
 * dvec: fp32 [B*heads][T] scratch.  Three launches (row sums, dK/dV per key block, dQ per query block); no atomics.
 * ch must be 32 or 64 (s2s_attn_supported).  Replaces: the reshape / split / einsum / softmax / einsum chain of
 * QKVAttentionLegacy.forward and its autograd. */
int s2s_attn_supported(int ch);
int s2s_attn_fwd(const void* qkv, int B, int T, int heads, int ch, int new_order, void* out, float* lse, int a_fmt,
                 void* stream);
int s2s_attn_bwd(const void* qkv, const void* out, const void* d_out, const float* lse, float* dvec, void* d_qkv, int B, int T,
                 int heads, int ch, int new_order, int a_fmt, int g_fmt, void* stream);

/* ---- embedding path (row a9): timestep embedding, time_embed MLP, label embedding, the 22 FiLM linears -------------- */

/* One fp32 GEMM  C[m][n] = bias[n] + add[m][n] + sum_k A[m*sam + k*sak] * B[k*sbk + n*sbn]  (element strides; bias / add
 * may be NULL), optionally with a second output C2 = silu(C).  Many of them run in ONE launch (s2s_linear_multi: the job
 * table is a kernel parameter, so the launch is capturable in a CUDA graph; more than s2s_linear_max_jobs() jobs are
 * split into several launches).  Plain fp32 FMA -- the reference's own arithmetic for these layers.
 * Replaces: nn.Linear / nn.Embedding-add / nn.SiLU of torchcfm UNetModel.time_embed, label_emb and every
 * ResBlock.emb_layers, forward and backward (dX = dY W, dW = dY^T X, db = 1^T dY are the same GEMM
 * with other strides). */
typedef struct {
    const float* A;
    const float* B;
    float* C;
    const float* bias;
    const float* add;
    float* C2;
    int M, N, K;
    int ldc, ld_add;
    int sam, sak, sbk, sbn;
} s2s_gemm_job;
int s2s_linear_max_jobs(void);
int s2s_linear_multi(const s2s_gemm_job* jobs_host, int njobs, void* stream);
/* out[i] = (sum_j parts[j*n + i]) * silu'(z[i])  (z NULL: plain sum): folds the per-ResBlock partial gradients of the shared
 * embedding and takes them through the SiLU in one pass. */
int s2s_sum_parts_silu_bwd(const float* parts, int nparts, long long n, const float* z, float* out, void* stream);
/* guided-diffusion timestep_embedding: emb[b] = [cos(t_b f_i) | sin(t_b f_i)], f_i = exp(-ln(max_period) i / (dim/2)). */
int s2s_timestep_embedding(const float* t, int B, int dim, float max_period, float* emb, void* stream);

/* One parameter tensor of the fused optimizer: fp32 param / grad / exp_avg / exp_avg_sq of n elements. */
typedef struct {
    float* p;
    const float* g;
    float* m;
    float* v;
    long long n;
} s2s_adam_tensor;

/* Elements of one (tensor, chunk) work item of s2s_adam_multi. */
int s2s_adam_chunk(void);

/* torch.optim.Adam step (L2 weight decay: g += wd * p; bias-corrected; eps outside the sqrt) over ALL tensors in one
 * launch.  tensors_dev: device array of s2s_adam_tensor; work_dev: device array of n_work int pairs (tensor index,
 * chunk index), one CTA each.  `step` is the 1-based step count, grad_scale multiplies every gradient first.
 * Replaces: torch.optim.Adam / _multi_tensor_adam (configs/model/conditional_flow_matching.yaml:3-7,
 * src/models/conditional_flow_matching.py:112-131). */
int s2s_adam_multi(const s2s_adam_tensor* tensors_dev, const int* work_dev, int n_work, double lr, double beta1,
                   double beta2, double eps, double weight_decay, int step, double grad_scale, void* stream);

/* s2s_adam_multi for launches replayed from a CUDA graph: step_dev (may be NULL -> `step` is used) points at the 1-based
 * step count in device memory (int64); the bias corrections are evaluated on the device from the live counter. */
int s2s_adam_multi_step(const s2s_adam_tensor* tensors_dev, const int* work_dev, int n_work, double lr, double beta1,
                        double beta2, double eps, double weight_decay, int step, const long long* step_dev,
                        double grad_scale, void* stream);

/* dst_i[0..n) = src_i[0..n) for many fp32 tensors in ONE launch (table / work list as s2s_adam_multi, chunks of
 * s2s_adam_chunk() elements): gathers the per-parameter gradients into one flat buffer, so that the data-parallel training
 * step (one gradient all-reduce per step) issues a single NCCL all-reduce.
 * Replaces: DistributedDataParallel's bucket copies (configs/trainer/ddp.yaml strategy: ddp). */
typedef struct {
    const float* src;
    float* dst;
    long long n;
} s2s_copy_tensor;
int s2s_copy_multi(const s2s_copy_tensor* tensors_dev, const int* work_dev, int n_work, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* S2S_B200_H */
