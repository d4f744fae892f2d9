/*
 * mspi_b200 — C ABI of the B200 (sm_100a) kernels behind MSPI's batched clip forward.
 *
 * The reference (oraclefina/MSPI) is pure Python/PyTorch and has no FFI of its own;
 * every entry point below replaces the PyTorch library call(s) that the reference
 * makes at the cited file:line (paths relative to the reference root).  All pointers
 * are DEVICE pointers unless stated otherwise, sizes are element counts, `stream` is a
 * cudaStream_t passed as void*.  Every function returns 0 on success and a negative
 * code on failure; mspi_last_error() returns the message for the calling thread.
 *
 * Activation layout is channels-last ("NDHWC": [N][T][H][W][C], C contiguous), bf16
 * unless a dtype field says otherwise.  There is no CPU fallback: with no usable GPU
 * every compute entry point fails with MSPI_ERR_CUDA.
 */
#ifndef MSPI_B200_H
#define MSPI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSPI_OK 0
#define MSPI_ERR_ARG (-1)
#define MSPI_ERR_CUDA (-2)
#define MSPI_ERR_UNSUPPORTED (-3)

/* dtype codes */
#define MSPI_BF16 0
#define MSPI_F32 1
/* activation codes (epilogues) */
#define MSPI_ACT_NONE 0
#define MSPI_ACT_RELU 1
#define MSPI_ACT_GELU 2    /* exact erf GELU, nn.GELU() default (model_utils.py:325) */
#define MSPI_ACT_SIGMOID 3 /* nn.Sigmoid (model_utils.py:164) */
#define MSPI_ACT_SWISH 4   /* x * sigmoid(x), SlowFast/resnet_helper.py:76-101 (X3D); not an epilogue of mspi_conv_gemm */

#define MSPI_MAX_TAPS 64

const char* mspi_last_error(void);
/* Library/ABI version and the SM architecture the kernels were compiled for ("sm_100a"). */
int mspi_version(void);
const char* mspi_arch(void);
/* Number of kernels launched by this library since load (process-wide counter). */
int64_t mspi_launch_count(void);
/* Programmatic dependent launch for the encoder kernels (conv_gemm, fused MLP, depthwise 7x7): the next kernel's prologue
 * overlaps the previous kernel's drain; every such kernel executes griddepcontrol.wait before its first global access.
 * Default on (environment MSPI_PDL=0 turns it off); returns the previous setting.  Set it BEFORE a plan is captured into a
 * CUDA graph: the captured edges keep the kind they were captured with. */
int mspi_set_pdl(int on);
/* Profiling aid for the depthwise 7x7 + LayerNorm kernel (MSPI_DW_DEBUG=1 selects its instrumented instance): cycles the
 * warps spent [0] waiting for their tile, [1] in the stencil, [2] in the LayerNorm phase, [3] the number of warp samples,
 * [4] in the barrier after the stencil, [5] storing the result tile, summed since the last reset.  out8 has 8 entries.
 * Synchronises the device.  No counterpart in the reference. */
int mspi_debug_dw_phase_cycles(uint64_t* out8, int reset);
/* Same for the epilogue warps of mspi_conv_gemm, filled only by a study build of the library (-DMSPI_GEMM_STUDY, selected at
 * run time through the MSPI_LIB environment variable): cycles [0] waiting for the accumulator, [1] in tcgen05.ld, [2] in the
 * scale/shift/activation math, [3] waiting for the staging buffer, [4] staging + bulk-store issue, [5] chunks, [6] tiles,
 * [7] the whole epilogue loop, summed over warps.  The release build returns zeros. */
int mspi_debug_gemm_epilogue_cycles(uint64_t* out8, int reset);

/* ------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution / linear layer on tcgen05 tensor cores (TMA-fed, TMEM accumulators).
 * Replaces: nn.Conv3d / nn.Conv2d / nn.Linear + BatchNorm(eval) + ReLU/GELU/Sigmoid + residual
 *   backbones/s3d.py:41-52,95-116 (BasicConv3d, SepConv3d), backbones/resnet.py:17-54,
 *   model/model_utils.py:43-46,92-94,155-170,324-327,367-377,404-435,437-504.
 *
 * The activation tensor is described as a 5-D view (inner -> outer) C, d1, d2, d3, d4 with
 * arbitrary element strides (stride of C must be 1).  One M-tile is a box of up to 128 output
 * positions (box[1]*box[2]*box[3]*box[4] <= 128); for every filter tap the same box shifted by
 * tap_off[tap][*] is fetched by one TMA bulk-tensor load, out-of-bounds elements read as zero
 * (this is the convolution's zero padding).  K is ordered (tap, cin); the weight matrix is
 * [cout_padded][ntaps*cin_pad], K contiguous, zero padded.
 *   y[pos][n] = act( scale[n] * sum_k A[pos][k] W[n][k] + shift[n] (+ residual[pos][n]) )
 * A plain GEMM is the special case ntaps=1, a_dims={K,M,1,1,1}, box={*,128,1,1,1}.
 */
typedef struct {
  int32_t a_dtype;      /* MSPI_BF16 (kind::f16) or MSPI_F32 (kind::tf32); weights use the same */
  int32_t a_dims[5];    /* extents of the activation view; a_dims[0] = input channels visible */
  int64_t a_strides[5]; /* element strides; a_strides[0] == 1; byte strides must be 16B multiples */
  int32_t box[5];       /* box[0] ignored (K chunk is 128 bytes); box[1..4] output-position box */
  int32_t ntaps;
  int32_t tap_off[MSPI_MAX_TAPS][4]; /* coordinate offset of each tap along d1..d4 */
  int32_t cin_pad;      /* K extent of one tap inside the weight matrix (multiple of the K chunk) */
  int32_t cout;         /* valid output channels */
  int32_t w_rows;       /* rows of the weight matrix (>= cout) */
  int32_t bn;           /* N tile: multiple of 16, 16..256 */
  int32_t o_dims[4];    /* output extents along d1..d4 */
  int64_t o_strides[4]; /* element strides of the output rows (channel stride 1) */
  int64_t r_strides[4]; /* same for the residual tensor */
  int32_t o_dtype;      /* MSPI_BF16 or MSPI_F32 */
  int32_t r_dtype;      /* dtype of the residual */
  int32_t act;          /* MSPI_ACT_* */
  int32_t has_residual;
  int32_t res_after_act; /* 0: act(v + res)   1: act(v) + res */
  int32_t k_row_bytes;  /* 0 = 128.  Bytes of K fetched per row per step (one TMA box, one swizzled smem tile): 128, 64 or
                           32.  The small sizes serve the Cin=3 stems, where one filter row over the padded 4-channel
                           frames of mspi_clip_to_padded_nhwc4 is a 64-byte (S3D, 8 px) or 32-byte (ConvNeXt, 4 px)
                           contiguous run; cin_pad is then a multiple of k_row_bytes / elsize. */
  int32_t w_batch_dims[2]; /* {0,0}: one weight matrix (normal).  {D3, D4} (== o_dims[2], o_dims[3]): batched GEMM — the B
                              operand of an M tile at (d3, d4) is matrix (d3, d4) of a [D4][D3][w_rows][K] family addressed
                              by w_strides (elements): [0] between rows, [1] between d3 matrices, [2] between d4 matrices;
                              K = a_dims[0], ntaps = 1, box[3] = box[4] = 1.  Used for attention (model_utils.py:97-109):
                              scores = Q K^T per (head, sample) and out = P V. */
  int64_t w_strides[3];
} MspiConvDesc;

int mspi_conv_gemm(const MspiConvDesc* d, const void* x, const void* w, const float* scale,
                   const float* shift, const void* residual, void* y, void* stream);
/* Same convolution with a LayerNorm over channels as its epilogue (ConvNeXt stem: Conv2d 4x4/s4 + LayerNorm2d,
 * timm convnext_tiny stem, model_utils.py:361):  y = LN(acc * scale + shift) * ln_weight + ln_bias, bf16.
 * One N tile (cout == bn, 64 / 128 / 192 columns); a GEMM row holds ln_groups (1, 2, 4) groups of cout / ln_groups
 * channels, each normalised on its own (the stem computes two output pixels per row).  Statistics in fp32, two passes
 * over the accumulators; the fp32 conv output is never written. */
int mspi_conv_gemm_ln(const MspiConvDesc* d, const void* x, const void* w, const float* scale, const float* shift,
                      const float* ln_weight, const float* ln_bias, float ln_eps, int ln_groups, void* y, void* stream);

/* (1,3,3) convolution, stride 1, zero padding (0,1,1), Cin = Cout = c = 8 or 16, + per-channel shift (+ ReLU), bf16 NDHWC
 * channel-slice views in and out: SlowFast fast-pathway BottleneckTransform.branch2.b with dim_inner 8 / 16
 * (backbones/SlowFast/resnet_helper.py:323-341) on CUDA cores — nine K = 8 taps are no work for a tensor-core tile.
 * w_packed: fp32 [9 taps (kh, kw)][ci][co] with the BatchNorm scale already multiplied in; shift: fp32 [c] or null. */
int mspi_conv133_small(const void* x, int64_t x_cstride, const float* w_packed, const float* shift, void* y,
                       int64_t y_cstride, int64_t planes, int h, int w, int c, int act, void* stream);

/* ------------------------------------------------------------------------------------------
 * Weight gradient of mspi_conv_gemm's convolution on tcgen05 tensor cores (training step, engine_train.py:74:
 * loss.backward() through nn.Conv3d / nn.Linear):
 *   dw[co*s_co + tap*s_tap + ci*s_ci] += sum over output positions p of dy[p][co] * x[p + tap_off[tap]][ci]
 * `d` is the FORWARD descriptor of the layer (a_dims/a_strides/box/tap_off describe x, o_dims/o_strides describe dy,
 * d->o_dtype == d->a_dtype: both operands bf16 or both fp32/tf32; box positions a multiple of 16 (bf16) / 8 (tf32)).
 * The position axis is the GEMM's reduction axis (both operands "MN-major"), split across CTAs; partial sums are
 * added into the fp32 gradient tensor with red.global.add.f32, so dw must be initialised (zero or the running sum).
 * s_co / s_ci / s_tap: element strides of dw, e.g. (cin*taps, taps, 1) for PyTorch's [Cout][Cin][kt][kh][kw]. */
int mspi_conv_wgrad(const MspiConvDesc* d, const void* x, const void* dy, float* dw, int64_t s_co, int64_t s_ci,
                    int64_t s_tap, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused ConvNeXt MLP (timm convnext Mlp + layer scale + residual, model_utils.py:361), C = 96 or 192:
 *   y[m][:] = residual[m][:] + scale * ( W2 . gelu(W1 . x[m][:] + b1) ) + shift        (scale = gamma, shift = gamma*b2)
 * x bf16 [m][c] (row stride c), w1 bf16 [4c][c_pad] (K zero padded to c_pad = multiple of 64), w2 bf16 [c][4c],
 * b1 fp32 [4c], scale/shift fp32 [c], residual / y bf16 with row strides res_stride / y_stride (elements).
 * Two chained tcgen05 GEMMs per 128-row tile; the 4c-wide hidden tile lives in TMEM / shared memory only. */
int mspi_mlp_fused(const void* x, const void* w1, const float* b1, const void* w2, const float* scale,
                   const float* shift, const void* residual, void* y, int64_t m, int c, int c_pad,
                   int64_t res_stride, int64_t y_stride, void* stream);
/* Same block with the NEXT layer's LayerNorm fused into the store: y = LN_c(r + gamma * mlp(x)) * ln_weight + ln_bias, the
 * statistics taken over the c channels of the bf16-rounded block output.  Serves the last block of ConvNeXt stages 0 and 1,
 * whose output is read only by the following stage's downsample.0 LayerNorm2d (timm ConvNeXtStage.downsample,
 * model_utils.py:361): the un-normalised block output is never written. */
int mspi_mlp_fused_ln(const void* x, const void* w1, const float* b1, const void* w2, const float* scale,
                      const float* shift, const void* residual, void* y, int64_t m, int c, int c_pad,
                      int64_t res_stride, int64_t y_stride, const float* ln_weight, const float* ln_bias, float ln_eps,
                      void* stream);

/* ------------------------------------------------------------------------------------------
 * Patch gather (explicit im2col) for the layers whose input has too few channels for a TMA
 * K-chunk (Cin = 3 or 1) or an odd-size strided grid: S3D stem conv_s (backbones/s3d.py:383),
 * ConvNeXt stem 4x4/s4 (timm convnext_tiny, model_utils.py:361), ResNet conv1 and its stride-2
 * convs (backbones/resnet.py:79, 8-14).  Writes bf16 rows [M][k_pad], K ordered (kt,kh,kw,c),
 * zero padded to k_pad.  src layout: 0 = fp32 NCDHW (the model's input contract), 1 = bf16 NDHWC.
 */
typedef struct {
  int32_t src_layout;
  int32_t n, c, t, h, w;      /* input extents */
  int64_t src_cstride;        /* NDHWC only: elements between consecutive pixels (>= c) */
  int32_t kt, kh, kw;
  int32_t st, sh, sw;
  int32_t pt, ph, pw;
  int32_t ot, oh, ow;         /* output extents */
  int32_t k_pad;              /* row length of dst (>= kt*kh*kw*c, multiple of 8) */
} MspiPatchDesc;
int mspi_patch_gather(const MspiPatchDesc* d, const void* src, void* dst, void* stream);

/* fp32 NCDHW -> bf16 NDHWC (clips, model_utils.py:557 rearrange; audio [B,1,F,T]) */
int mspi_ncdhw_to_ndhwc(const float* src, void* dst, int n, int c, int thw, int64_t dst_cstride,
                        void* stream);
/* fp32 NCDHW clip [N][3][T][H][W] (the forward's input contract, model_utils.py:556) -> bf16 frames
 * [N*T][hp][wp][4] with the image at rows pad_t.., columns pad_l.. and channel 3 = 0.  Only the interior is
 * written; the caller zeroes the buffer once.  With 8-byte pixels every row of a stem filter (7 taps at stride 2
 * for S3D, s3d.py:383; 4 taps at stride 4 for ConvNeXt, model_utils.py:361) is one contiguous, 16-byte aligned
 * run of the padded frame, which is what lets the stems run as implicit GEMMs straight off TMA loads
 * (MspiConvDesc.k_row_bytes).  w % 4 == 0, pad_l and wp even. */
int mspi_clip_to_padded_nhwc4(const float* src, void* dst, int n, int t, int h, int w, int pad_t, int pad_l,
                              int hp, int wp, void* stream);
/* Same, for a subset / re-ordering of the frames and a time-padded destination: destination frame
 * b*dst_frames_per_clip + dst_frame0 + j (j < t_out) receives source frame frame_map[j] (HOST array).  Serves the
 * SlowFast pathways: slow = frames [0,4,12,T-1] (model_utils.py:523); fast = all frames with two zero frames at
 * either end of every clip, the temporal padding of its (5,7,7) stem (stem_helper.py:128-204). */
int mspi_clip_frames_to_padded_nhwc4(const float* src, void* dst, int n, int t, int h, int w, int pad_t, int pad_l,
                                     int hp, int wp, const int32_t* frame_map, int t_out, int dst_frames_per_clip,
                                     int dst_frame0, void* stream);

/* uint8 frames [N][T][H][W][3] (decoder output; the reference converts them on the host: inference.py:154-165 Resize ->
 * ToTensor (/255) -> Normalize(mean, std)) -> the same bf16 zero-padded 4-channel frames as mspi_clip_to_padded_nhwc4, the
 * normalisation done in fp32 with IEEE division, so the result is bit-identical to converting the reference's normalised
 * fp32 clip.  frame_map == NULL: all t frames in order; otherwise as mspi_clip_frames_to_padded_nhwc4.  mean3 / std3 are
 * HOST pointers to three floats. */
int mspi_clip_u8_to_padded_nhwc4(const void* src, void* dst, int n, int t, int h, int w, int pad_t, int pad_l, int hp,
                                 int wp, const int32_t* frame_map, int t_out, int dst_frames_per_clip, int dst_frame0,
                                 const float* mean3, const float* std3, void* stream);

/* dst row i = src row idx[i], rows of row_bytes (multiple of 16) bytes; idx is a DEVICE array of n_rows indices, clamped to
 * [0, src_rows).  Builds the per-window [B][T] views of the per-frame image-encoder feature cache: consecutive sliding
 * windows of inference.py:120-150 share 15 of their 16 frames, and the image encoder (model_utils.py:357-385) is per frame. */
int mspi_gather_rows(const void* src, const int32_t* idx, void* dst, int64_t n_rows, int64_t row_bytes, int64_t src_rows,
                     void* stream);
/* bf16/fp32 NDHWC (channel slice) -> fp32 NCDHW, used to hand taps back to PyTorch callers */
int mspi_ndhwc_to_ncdhw(const void* src, int src_dtype, int64_t src_cstride, float* dst, int n,
                        int c, int thw, void* stream);

/* ------------------------------------------------------------------------------------------
 * MaxPool3d, channels-last bf16, -inf padding.  Replaces nn.MaxPool3d / nn.MaxPool2d at
 * backbones/s3d.py:134,384-401, backbones/resnet.py:83, model_utils.py:189,206.
 */
typedef struct {
  int32_t n, t, h, w, c;
  int64_t in_cstride, out_cstride; /* elements between pixels (channel-slice views) */
  int32_t kt, kh, kw, st, sh, sw, pt, ph, pw;
  int32_t ot, oh, ow;
} MspiPoolDesc;
int mspi_maxpool3d(const MspiPoolDesc* d, const void* x, void* y, void* stream);

/* ------------------------------------------------------------------------------------------
 * Bilinear (1,k,k) upsample, align_corners=False, channels-last.  Replaces
 * nn.Upsample(mode='trilinear', scale_factor=(1,k,k)) at model_utils.py:158,208,486-488,498.
 *   y = act( up(x) (+ y if accumulate) )        in/out dtype bf16 or fp32; act NONE or RELU
 * (the ReLU lets a temporal conv be applied BEFORE the upsample: both are linear, they commute,
 *  and the non-linearity that followed the conv moves here — readout.7-9, model_utils.py:498-500)
 */
typedef struct {
  int32_t nt;        /* N*T planes */
  int32_t h, w, c;   /* input plane */
  int32_t k;         /* integer scale */
  int64_t in_cstride, out_cstride;
  int32_t in_dtype, out_dtype;
  int32_t accumulate;
  int32_t act;
} MspiUpDesc;
int mspi_upsample_bilinear(const MspiUpDesc* d, const void* x, void* y, void* stream);

/* ------------------------------------------------------------------------------------------
 * Depthwise convolution (+ optional LayerNorm over C), channels-last.
 * Replaces ConvNextBlock.dwconv_t/dwconv_s + LayerNorm3d (model_utils.py:321-323,293-303)
 * and timm ConvNeXt conv_dw + norm (model_utils.py:361).
 *   y = LN_C( dwconv(x) + bias ) * ln_w + ln_b        (LN skipped when ln_w == NULL)
 * weights: fp32 [kt*kh*kw][C] (tap-major), bias fp32 [C].
 */
typedef struct {
  int32_t n, t, h, w, c;
  int32_t kt, kh, kw; /* odd, "same" padding */
  float ln_eps;
  int32_t out_dtype;
  int32_t in_dtype;
} MspiDwDesc;
int mspi_dwconv_ln(const MspiDwDesc* d, const void* x, const float* wgt, const float* bias,
                   const float* ln_w, const float* ln_b, void* y, void* stream);

/* ------------------------------------------------------------------------------------------
 * X3D-L blocks (backbones/X3D.py:111-250, SlowFast/resnet_helper.py:47-73,213-351, stem_helper.py:207-290).
 * Depthwise (kt,kh,kw) convolution with "same" padding, spatial stride (sh,sw), BatchNorm folded into the weights
 * (wgt fp32 [kt*kh*kw][c] = w * bn_scale) and `shift` (fp32 [c]), then act (NONE / RELU / SWISH); bf16 NDHWC in/out.
 * Serves X3DTransform.b + b_bn (3x3x3, stride (1,s,s)) and the stem's depthwise (5,1,1) conv + bn + relu. */
typedef struct {
  int32_t n, t, h, w, c;
  int64_t in_cstride, out_cstride;
  int32_t kt, kh, kw, sh, sw;
  int32_t oh, ow;
  int32_t act;
} MspiDw3dDesc;
int mspi_dwconv3d_bn(const MspiDw3dDesc* d, const void* x, const float* wgt, const float* shift, void* y, void* stream);
/* The same layer AND the per-(sample, channel) mean of its output, fp32 [n][c] — the squeeze of the SE block that follows
 * X3DTransform.b (resnet_helper.py:47-73, 327-333) — accumulated by the depthwise kernel itself where its shared-memory tile
 * form applies (16 partial means per (sample, channel) in `work`, then summed: the blocks of one sample would otherwise queue
 * on one atomic unit per channel), else by mspi_channel_mean. */
int mspi_dwconv3d_bn_mean(const MspiDw3dDesc* d, const void* x, const float* wgt, const float* shift, void* y,
                          float* mean_out, float* work /* scratch, fp32 [n][16][c] */, void* stream);
/* Squeeze-Excitation (resnet_helper.py:47-73): out[n][c] = mean over rows of x[n][rows][c] (bf16, pixel stride cstride); */
int mspi_channel_mean(const void* x, float* out, int n, int64_t rows, int c, int64_t cstride, void* stream);
/* gate[n][c] = sigmoid(w2 relu(w1 mean[n] + b1) + b2), w1 fp32 [cfc][c], w2 fp32 [c][cfc]; */
int mspi_se_gate(const float* mean, const float* w1, const float* b1, const float* w2, const float* b2, float* gate,
                 int n, int c, int cfc, void* stream);
/* y = act(x * gate[n][c]) over bf16 [n][rows][c] (x == y allowed): the SE scaling fused with the Swish that follows it. */
int mspi_scale_act(const void* x, const float* gate, void* y, int n, int64_t rows, int c, int act, void* stream);

/* LayerNorm over the last dim of [rows][c]; in/out dtype selectable; optional ReLU and an
 * optional additive table pos[(row % pos_rows)][c] (sinusoid position table, model_utils.py:18-29,
 * 270-274).  Replaces nn.LayerNorm at model_utils.py:231-233,138-144,406-434 and LayerNorm2d. */
typedef struct {
  int64_t rows;
  int32_t c;
  int64_t in_rstride, out_rstride;
  int32_t in_dtype, out_dtype;
  float eps;
  int32_t relu;
  int32_t pos_rows; /* 0 = no table */
  int64_t rows_per_group, out_gstride; /* output row r -> (r / rows_per_group)*out_gstride + (r % rows_per_group)*out_rstride */
} MspiLnDesc;
int mspi_layernorm(const MspiLnDesc* d, const void* x, const float* w, const float* b,
                   const float* pos, void* y, void* stream);

/* Multi-head self attention over short sequences (model_utils.py:97-109): qkv (bf16 or fp32, `dtype`)
 * [B][N][3][heads][hd] -> out (same dtype) [B][N][heads*hd]; softmax(q k^T * scale) v in fp32. */
int mspi_attention(const void* qkv, void* out, int dtype, int b, int n, int heads, int hd, float scale,
                   void* stream);

/* Attention as two batched tensor-core GEMMs (MspiConvDesc.w_batch_dims) plus this glue, fp32:
 *   scores[b][h][q][k] = scale * Q K^T  (mspi_conv_gemm)  ->  mspi_softmax_rows in place  ->  out = P V (mspi_conv_gemm)
 * mspi_transpose_v writes the K-major copy of V the second GEMM needs: qkv [B][N][3][heads][hd] -> vt [B][heads][hd][n_pad]
 * (columns >= n are written as zero). */
int mspi_softmax_rows(float* s, int64_t rows, int n, int64_t stride, void* stream);
int mspi_transpose_v(const float* qkv, float* vt, int b, int n, int heads, int hd, int n_pad, void* stream);

/* SA gating + top-down sums (model_utils.py:167-170,566-568):
 *   y = x * sigmoid_mask + x ; the mask logits are fp32 [N*T*H*W] (one channel); x, y bf16 or fp32 (`dtype`). */
int mspi_sa_gate(const void* x, int64_t x_cstride, const float* mask_logits, void* y, int64_t y_cstride,
                 int64_t pixels, int c, int dtype, void* stream);

/* The same gate fused with the top-down sums of model_utils.py:566-568 (fp32):
 *   y = x * sigmoid_mask + x + sum_i up_{k_i}(src_i),   up = bilinear (1,k,k) upsample, align_corners=False
 * x, y: [nt][h][w][c]; src_i: [nt][h/k_i][w/k_i][c] (pixel stride src_cstrides[i]); srcs / src_cstrides / src_scales are
 * HOST arrays of nsrc <= 3 entries.  y is written once instead of once per term.  mask_logits == NULL drops the gate
 * (y = x + sum up(src); x == y allowed): the readout.0 partial products are accumulated that way (model_utils.py:570). */
int mspi_sa_gate_fused(const float* x, int64_t x_cstride, const float* mask_logits, float* y, int64_t y_cstride, int nt,
                       int h, int w, int c, int nsrc, const float* const* srcs, const int64_t* src_cstrides,
                       const int32_t* src_scales, void* stream);

/* Elementwise y = a + b over bf16 (used for ViT residuals when not fused) */
int mspi_add_bf16(const void* a, const void* b, void* y, int64_t n, void* stream);

/* Mean over tokens: x (bf16 or fp32) [B][rows][C] (row range [r0, r1)) -> fp32 [B][C]
 * (AdaptiveAvgPool3d/2d to 1, model_utils.py:401-402,545-546). */
int mspi_token_mean(const void* x, int x_dtype, float* y, int b, int rows, int r0, int r1, int c, void* stream);

/* Strided row copy with dtype conversion: dst[g][r][0..c) = src[g][r][0..c) for g < groups,
 * r < rows; strides in elements.  Moves the fused visual tokens into the channel slice that
 * torch.cat([v4, vis_sync]) occupies (model_utils.py:541-543,559). */
int mspi_cast_rows(const void* src, int src_dtype, int64_t src_rstride, int64_t src_gstride, void* dst,
                   int dst_dtype, int64_t dst_rstride, int64_t dst_gstride, int groups, int rows, int c,
                   void* stream);

/* SimSiam negative cosine loss (model_utils.py:285-290,551):
 *   out[0] = 0.5*( -mean_b cos(p_v, z_a) - mean_b cos(p_a, z_v) ), all fp32 [B][C]. */
int mspi_simsiam_loss(const float* p_v, const float* z_a, const float* p_a, const float* z_v,
                      float* out, int b, int c, void* stream);

/* out[b] = x[b] - logsumexp(x[b]) over H*W pixels, fp32 (model_utils.py:571-572). */
int mspi_logsoftmax2d(const float* x, float* y, int b, int64_t pixels, void* stream);

/* Saliency metrics (utils/compute_saliency_metrics.py:9-108, utils/loss.py:26-49).
 * pred: fp32 [B][P]; if pred_is_log != 0 the kernel uses exp(pred) (SalLoss passes the log map).
 * gt: fp32 [B][P]; fix: fp32 [B][P] or NULL.  out fp32[5] = {kld, cc, sim, nss, loss}, batch
 * means; loss = kld - cc (- 0.1*nss when fix != NULL); nss = 0 when fix == NULL.
 * work: fp32 scratch of at least 16*B floats. */
int mspi_saliency_metrics(const float* pred, int pred_is_log, const float* gt, const float* fix,
                          float* out, float* work, int b, int64_t pixels, void* stream);

/* Post-processing of the inference driver (inference.py:65-91): 11x11 Gaussian blur (sigma 2.0, reflect-101) of the LOG map
 * -> exp -> bilinear resize to (oh, ow) (cv2 INTER_LINEAR) -> per-map min-max -> round(255 x) -> uint8 [b][oh][ow].
 * work: fp32 scratch of b*h*w + b*oh*ow + 2*b elements. */
int mspi_postprocess_maps(const float* log_maps, uint8_t* out, float* work, int b, int h, int w, int oh, int ow,
                          void* stream);

/* Log power spectrogram front end (inference.py:24-63, avsp_dataloader.py:51-80):
 * wave fp32 [B][n] -> out fp32 [B][257][frames_out]; STFT n_fft=512, hop=160, periodic Hann,
 * center=True reflect pad, |.|^2, log(x+1e-6), per-frame standardise over the 257 bins
 * (unbiased std, +1e-6), columns beyond the signal's frames filled with 0.02, cropped to
 * frames_out. */
int mspi_logspec(const float* wave, float* out, int b, int n, int frames_out, void* stream);

/* ==========================================================================================
 * Training step (engine_train.py:27-76, train.py:151-166): model.train() + frozen_encoder(), SalLoss + gamma*loss_va,
 * loss.backward(), AdamW.  Data and weight gradients of every Conv3d / Linear are mspi_conv_gemm (on the transposed /
 * flipped weights) and mspi_conv_wgrad; the entry points below are everything else.  All tensors fp32, channels-last
 * [pixels][C] views with an element stride between pixels ("cstride", so channel slices of concatenated buffers work in
 * place).  Gradient outputs: `accumulate` 0 overwrites, 1 adds; scatter-type adjoints (max-pool, upsample) and parameter
 * gradients always add, so their targets must hold zero or the running sum.
 * ========================================================================================== */

/* BatchNorm3d in training mode (torch.nn.BatchNorm semantics; backbones/s3d.py:45,99,103, model_utils.py:493,496):
 * y = act((x - mean_batch) * invstd_batch * weight + bias), biased batch variance; running buffers move by `momentum`
 * towards (mean, unbiased variance), *num_batches_tracked += 1 (both optional).  save_mean / save_invstd [c] are kept for
 * the backward.  work: double [2c] that MUST BE ZERO on entry (the batch sums are accumulated into it and left there; the
 * training plan clears all of its BatchNorm scratch with one memset per step).  Two launches: sums, normalise.  relu: 0 / 1. */
int mspi_bn_train_fwd(const float* x, int64_t x_cstride, float* y, int64_t y_cstride, int64_t pixels, int c,
                      const float* weight, const float* bias, float eps, float momentum, float* running_mean,
                      float* running_var, int64_t* num_batches_tracked, float* save_mean, float* save_invstd, double* work,
                      int relu, void* stream);
/* Its backward: g = dy masked by the ReLU (y > 0) when relu; dweight += sum g*xhat; dbias += sum g;
 * dx (+)= weight*invstd*(g - mean(g) - xhat*mean(g*xhat)).  work: double [2c], zero on entry (not the forward's). */
int mspi_bn_train_bwd(const float* x, int64_t x_cstride, const float* y, int64_t y_cstride, const float* dy,
                      int64_t dy_cstride, float* dx, int64_t dx_cstride, int64_t pixels, int c, const float* weight,
                      const float* save_mean, const float* save_invstd, float* dweight, float* dbias, double* work, int relu,
                      int accumulate, void* stream);

/* y = act(x) elementwise (exact-erf GELU of nn.GELU, model_utils.py:45,325; kept separate from the GEMM epilogue in training
 * because the backward needs the pre-activation). */
int mspi_act_fwd(const float* x, float* y, int64_t n, int act, void* stream);
/* g = dy * act'(ref): ref is the activation's OUTPUT for RELU, its INPUT for GELU, unused for NONE.  dz = g when dz != NULL
 * (may alias dy); dbias[ch] += sum over pixels of g when dbias != NULL (the bias gradient of the conv / linear before it). */
int mspi_act_bwd(const float* dy, int64_t dy_cstride, const float* ref, int64_t ref_cstride, float* dz, int64_t dz_cstride,
                 int64_t pixels, int c, int act, float* dbias, void* stream);

/* fp32 MaxPool3d (same descriptor as mspi_maxpool3d) and its adjoint: each output element adds its gradient to the first
 * maximal input of its window in (t,h,w) scan order, the index torch's forward records. */
int mspi_maxpool3d_f32(const MspiPoolDesc* d, const float* x, float* y, void* stream);
int mspi_maxpool3d_bwd(const MspiPoolDesc* d, const float* x, const float* dy, int64_t dy_cstride, float* dx,
                       int64_t dx_cstride, void* stream);

/* Adjoint of mspi_upsample_bilinear (fp32; `d` is the forward descriptor): dx += up^T(g), g = dy masked by y > 0 when
 * d->act == MSPI_ACT_RELU (y = the forward output, else unused).  d->accumulate is ignored: y = y_old + up(x) passes dy to
 * y_old unchanged. */
int mspi_upsample_bilinear_bwd(const MspiUpDesc* d, const float* dy, const float* y, float* dx, void* stream);

/* LayerNorm backward over rows (nn.LayerNorm / LayerNorm3d, model_utils.py:138-144,231-233,293-303,406-434).  Row r of dy
 * (and of y_relu, the forward output, when the LayerNorm was followed by a ReLU) lives at
 * (r / rows_per_group)*dy_gstride + (r % rows_per_group)*dy_rstride.  dx (+)= ...; dw += sum g*xhat; db += sum g. */
int mspi_layernorm_bwd(const float* x, int64_t x_rstride, const float* dy, int64_t dy_rstride, int64_t rows_per_group,
                       int64_t dy_gstride, const float* y_relu, const float* w, float eps, float* dx, int64_t dx_rstride,
                       int64_t rows, int c, int accumulate, float* dw, float* db, void* stream);

/* Softmax backward in place: dp <- scale * p * (dp - sum_k dp*p) per row (attention, model_utils.py:104-106). */
int mspi_softmax_bwd_rows(const float* p, float* dp, int64_t rows, int n, int64_t stride, float scale, void* stream);

/* Small strided batched fp32 GEMM, C[b1][b0][i][j] (+)= alpha * sum_k A[..][i][k] B[..][k][j]; strides (elements) are
 * {i, k, batch0, batch1} for A, {k, j, batch0, batch1} for B, {i, j, batch0, batch1} for C.  The attention backward
 * (dV = P^T dO, dP = dO V^T, dQ = dS K, dK = dS^T Q over 372 tokens x 4 heads) reads four differently transposed views. */
typedef struct {
  int32_t m, n, k, batch0, batch1, accumulate;
  float alpha;
  int64_t a_strides[4], b_strides[4], c_strides[4];
} MspiSgemmDesc;
int mspi_sgemm_strided(const MspiSgemmDesc* d, const float* a, const float* b, float* c, void* stream);

/* SA gating backward (model_utils.py:167-170), y = x*sigmoid(l) + x: dx (+)= dy*(1 + s); dlogits = s(1-s) sum_c dy*x. */
int mspi_sa_gate_bwd(const float* x, int64_t x_cstride, const float* mask_logits, const float* dy, int64_t dy_cstride,
                     float* dx, int64_t dx_cstride, float* dlogits, int64_t pixels, int c, int accumulate, void* stream);

/* Backward of the 32 -> 1 channel (1,3,3) convolutions (SA.conv_mask.2, model_utils.py:163; readout.12, :503), whose
 * one-channel gradient cannot feed a tensor-core tile: x [planes][h][w][32] (pixel stride x_cstride), dy [planes][h][w],
 * w / dw in PyTorch layout [1][32][1][3][3].  dx (+)= conv^T(dy) (skipped when dx == NULL); dw += ...; db += sum dy. */
/* Forward of those layers (also used by the inference plan): y [planes][h][w] fp32 = bias + conv(x, w); x bf16 or fp32
 * [planes][h][w][32] with pixel stride x_cstride, w fp32 [1][32][1][3][3].  One output channel would waste a tensor-core
 * tile and make the implicit GEMM re-fetch the input per tap; this reads every input line once and reduces with shuffles. */
int mspi_conv_c1_fwd(const void* x, int x_dtype, int64_t x_cstride, const float* w, const float* bias, float* y,
                     int64_t planes, int h, int wd, int cin, void* stream);
int mspi_conv_c1_bwd(const float* x, int64_t x_cstride, const float* dy, const float* w, float* dx, int64_t dx_cstride,
                     float* dw, float* db, int64_t planes, int h, int wd, int cin, int accumulate, void* stream);

/* Depthwise-conv weight / bias gradient (ConvNextBlock.dwconv_t (7,1,1) / dwconv_s (1,7,7), model_utils.py:321-322):
 * dw [c][kt*kh*kw] (PyTorch layout) += sum_p dy[p][ch] x[p+off][ch]; db += sum_p dy.  x, dy whole [n,t,h,w,c] buffers.
 * (The data gradient is mspi_dwconv_ln on the flipped filter.) */
int mspi_dwconv_wgrad(const MspiDwDesc* d, const float* x, const float* dy, float* dw, float* db, void* stream);

/* Saliency loss + its gradient (utils/loss.py:26-49, engine_train.py:38): loss = mean_b[KLD - CC](exp(log_map), gt) +
 * gamma*loss_va.  dlogits [b][pixels] = scale * dloss/d(readout logits) (through exp and the log-softmax);
 * out = {loss, kld, cc, loss_va}; loss_va: device scalar or NULL; work: 2*b floats. */
int mspi_salloss_bwd(const float* log_map, const float* gt, const float* loss_va, float gamma, float* dlogits, float* out,
                     float* work, int b, int64_t pixels, float scale, void* stream);
/* SimSiam loss backward (model_utils.py:285-290,551): dp_v, dp_a [b][c] = scale * dL/dp (z is detached). */
int mspi_simsiam_bwd(const float* p_v, const float* z_a, const float* p_a, const float* z_v, float* dp_v, float* dp_a,
                     int b, int c, float scale, void* stream);
/* Adjoint of mspi_token_mean: dx[b][r][:] += dy[b][:] / (r1 - r0) for r in [r0, r1). */
int mspi_token_mean_bwd(const float* dy, float* dx, int b, int rows, int r0, int r1, int c, void* stream);
/* dst[g][r][0..c) (+)= src[g][r][0..c): residual / concat gradient routing (strides in elements, multiples of 4). */
int mspi_add_rows(const float* src, int64_t src_rstride, int64_t src_gstride, float* dst, int64_t dst_rstride,
                  int64_t dst_gstride, int groups, int rows, int c, int accumulate, void* stream);

/* 4-D strided permute copy, fp32 -> fp32 / bf16: re-packs the updated fp32 master weights into the K-major GEMM matrices
 * (forward, and transposed + tap-flipped for the data gradient) after every optimiser step. */
typedef struct {
  int32_t n[4];
  int64_t src_strides[4], dst_strides[4];
  int32_t dst_dtype, accumulate;
} MspiPermDesc;
int mspi_permute_copy(const MspiPermDesc* d, const float* src, void* dst, void* stream);
/* The same for a whole list of jobs in ONE launch (the ~280 weight re-packs of a training step).  table: DEVICE array of
 * 16 x int64 per job {src pointer, dst pointer, n[4], src_strides[4], dst_strides[4], dst_is_bf16, total elements};
 * block_job / block_chunk: DEVICE arrays mapping each of the nblocks thread blocks to its job and to its 2048-element
 * chunk of that job. */
#define MSPI_PACK_CHUNK 2048
int mspi_permute_copy_batched(const int64_t* table, const int32_t* block_job, const int32_t* block_chunk, int nblocks,
                              void* stream);

/* torch.optim.AdamW single step over a flat fp32 buffer (train.py:157-158); g is multiplied by grad_scale first
 * (1/world_size after a sum all-reduce).  step_dev (device int32, optional) overrides `step` so a CUDA graph can replay. */
int mspi_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                    float weight_decay, int step, const int32_t* step_dev, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MSPI_B200_H */
