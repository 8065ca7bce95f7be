/*
 * mstr_b200.h - C ABI of the B200 (sm_100a) word-crop recognition forward pass.
 *
 * The reference (JoshuaPlacidi/multimodal_scene_text_recognition) has no FFI layer: its hot path is
 * a set of Python nn.Module classes whose arithmetic is dispatched to PyTorch ATen.  Each entry
 * point below replaces the ATen work done by one reference call site (cited per function, paths
 * relative to the reference root).  The Python modules in
 * multimodal_scene_text_recognition_b200/modules/ keep the reference's class names and forward
 * signatures and bind these symbols with ctypes.
 *
 * Conventions
 *   - every function returns 0 on success or a negative mstr_status code; no exceptions, no
 *     host synchronisation; all work is enqueued on `stream` (a cudaStream_t passed as void*);
 *   - all pointers are DEVICE pointers borrowed for the duration of the call, caller-allocated;
 *   - activations are NHWC ([B][H][W][C], C contiguous) unless stated otherwise;
 *   - `mstr_last_error_string()` describes the most recent failure on the calling thread.
 */
#ifndef MSTR_B200_H_
#define MSTR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* mstr_stream_t; /* cudaStream_t */

enum mstr_status {
  MSTR_OK = 0,
  MSTR_ERR_BAD_ARG = -1,      /* null pointer, non-positive size, misaligned pointer */
  MSTR_ERR_UNSUPPORTED = -2,  /* shape / mode the kernels are not written for */
  MSTR_ERR_CUDA = -3,         /* a CUDA runtime / driver call failed (see error string) */
  MSTR_ERR_NO_DEVICE = -4     /* no sm_100 device */
};

enum mstr_dtype { MSTR_F32 = 0, MSTR_BF16 = 1 };

/* Library / ABI version (major*1000 + minor). */
int mstr_version(void);
const char* mstr_last_error_string(void);
/* Number of kernels this library has launched since load (bench.py's "gpu_launches"). */
int64_t mstr_launch_count(void);

/* Asynchronous host<->device copies on `stream` (one cudaMemcpyAsync; the host buffer should be pinned).
 * Plumbing for the input pipeline that stands in for the reference's per-batch `.to(device)` calls
 * (training_functions.py:166-176). */
int mstr_copy_h2d_async(void* dst_device, const void* src_host, int64_t bytes, mstr_stream_t stream);
int mstr_copy_d2h_async(void* dst_host, const void* src_device, int64_t bytes, mstr_stream_t stream);

/* ---- TPS rectification ------------------------------------------------------------------
 * Replaces GridGenerator.build_P_prime (modules/transformation.py:154-163: two repeat() copies
 * and two torch.bmm) fused with F.grid_sample(bilinear, padding_mode='border',
 * align_corners=True) (modules/transformation.py:33-34).
 *   img      [B][C][H][W] f32  (NCHW, as the reference's DataLoader yields it)
 *   c_prime  [B][F][2]    f32  fiducial points from the localisation network
 *   basis    mode 0: P_hat^T        [F+3][H*W] f32  (reference order: T = inv_delta_C [C';0], P' = P_hat T)
 *            mode 1: (P_hat inv_delta_C[:, :F])^T  [F][H*W] f32 (folded constant, fp64-precomputed)
 *   inv_delta_C  [F+3][F+3] f32, used by mode 0 only (may be null in mode 1)
 *   out      [B][C][H][W] f32 rectified image
 *   grid_out optional [B][H*W][2] f32 sampling grid (null to skip)
 * img and out may both be null when grid_out is given: only the grid is produced (that is
 * GridGenerator.build_P_prime called on its own).
 */
int mstr_tps_rectify(const float* img, const float* c_prime, const float* basis,
                     const float* inv_delta_C, float* out, float* grid_out,
                     int B, int C, int H, int W, int F, int mode, mstr_stream_t stream);

/* ---- first-layer convolution (C_in = 1..4) -------------------------------------------------
 * Replaces Conv2d(C_in, C_out, 3, 1, 1, bias=False) + BatchNorm2d(eval) + ReLU
 * [+ MaxPool2d(2,2)] for the two layers that read the image directly
 * (modules/feature_extraction.py:58-60,111-113 conv0_1/bn0_1; modules/transformation.py:49-51
 * conv.0/conv.1/ReLU/MaxPool).  SIMT: K = 9*C_in is too small for tensor cores.
 *   x [B][C_in][H][W] f32 NCHW;  w [C_out][C_in][3][3] f32;  scale/shift [C_out] f32 (folded BN)
 *   pool = 0: out NHWC [B][H][W][C_out] of out_dtype;  pool = 1: fused 2x2 max-pool, out [B][H/2][W/2][C_out];
 *   pool = 2: bf16 only, out is the padded-flat activation [B][H+2][W+2][C_out] with a ZERO one-pixel ring that
 *             mstr_conv3x3_flat_bf16 consumes.
 */
int mstr_conv3x3_first(const float* x, const float* w, const float* scale, const float* shift,
                       void* out, int out_dtype, int B, int C_in, int H, int W, int C_out,
                       int pool, mstr_stream_t stream);

/* ---- implicit-GEMM convolution + folded BN + residual + ReLU --------------------------------
 * Replaces nn.Conv2d(bias=False) -> nn.BatchNorm2d(eval) [-> += residual] [-> ReLU]
 * (modules/feature_extraction.py:32-47,109-150; modules/transformation.py:52-57) and, with
 * R=S=1,H=W=1, nn.Linear (modules/encoders.py:15,28; modules/decoders.py:71,81).
 *   x        NHWC [B][H][W][C_in]
 *   w        [C_out][R][S][C_in]  (K-major, K = (r*S+s)*C_in + c)
 *   scale/shift [C_out] f32: y = acc*scale + shift   (scale may be null => 1)
 *   residual optional NHWC [B][P][Q][C_out], added before the ReLU
 *   out      NHWC [B][P][Q][C_out];  P = (H+2*pad_h-R)/stride_h+1, Q likewise
 * mstr_conv_gemm_f32 : fp32 operands, fp32 FMA (exact parity mode), x/w/residual/out f32.
 * mstr_conv_gemm_bf16: bf16 operands on tcgen05 tensor cores (TMA im2col loads, TMEM fp32
 *                      accumulators); residual bf16; out_bf16 and/or out_f32 may be given.
 *                      Requires C_in % 32 == 0, C_out % 16 == 0.
 */
int mstr_conv_gemm_f32(const float* x, const float* w, const float* scale, const float* shift,
                       const float* residual, float* out, int B, int H, int W, int C_in,
                       int C_out, int R, int S, int stride_h, int stride_w, int pad_h, int pad_w,
                       int relu, mstr_stream_t stream);
int mstr_conv_gemm_bf16(const void* x, const void* w, const float* scale, const float* shift,
                        const void* residual, void* out_bf16, float* out_f32, int B, int H, int W,
                        int C_in, int C_out, int R, int S, int stride_h, int stride_w, int pad_h,
                        int pad_w, int relu, mstr_stream_t stream);

/* Convolution with a fused 1x1 branch (a BasicBlock's conv2/bn2 + downsample conv/bn + add + ReLU,
 * modules/feature_extraction.py:39-47 with :94-100):
 *   out = act( conv_RxS(x, w_cat[:, :R*S*C_in]) + x2 . w_cat[:, R*S*C_in:]^T ) * scale + shift [+ residual] )
 * x2 is NHWC [B][P][Q][C_in2] on the OUTPUT grid; both branches accumulate in one TMEM accumulator, so the two
 * BatchNorm scales must be folded into w_cat by the caller (scale is then null) and shift = shift_a + shift_b.
 * Requires C_in % 64 == 0, C_in2 % 64 == 0.
 */
int mstr_conv_gemm_bf16_dual(const void* x, const void* x2, const void* w_cat, const float* scale, const float* shift,
                             const void* residual, void* out_bf16, float* out_f32, int B, int H, int W, int C_in,
                             int C_in2, int C_out, int R, int S, int stride_h, int stride_w, int pad_h, int pad_w,
                             int relu, mstr_stream_t stream);
/* Test hook: which kernel mstr_conv_gemm_bf16 uses where both apply (C_in % 64 == 0, C_out % 256 == 0).  0 (default): the
 * CTA-pair kernel, except for problems of at most one wave of 128 x 128 tiles, which take the single-CTA kernel;
 * 1: the single-CTA kernel always; 2: CTA pairs always.  The two give bit-identical results. */
int mstr_debug_gemm_route(int route);

/* TF32 parity mode on the tensor cores (error-compensated 3xTF32, see csrc/conv_tf32.cu): fp32 NHWC
 * activations and fp32 results like mstr_conv_gemm_f32, ~fp32 accuracy, tcgen05.mma.kind::tf32.
 *   w_hi = w with the low 13 mantissa bits cleared, w_lo = w - w_hi (both f32, [C_out][R][S][C_in])
 * Requires C_in % 32 == 0, C_out % 128 == 0.
 */
int mstr_conv_gemm_tf32(const float* x, const float* w_hi, const float* w_lo, const float* scale,
                        const float* shift, const float* residual, float* out, int B, int H, int W, int C_in,
                        int C_out, int R, int S, int stride_h, int stride_w, int pad_h, int pad_w, int relu,
                        mstr_stream_t stream);

/* ---- pooling / layout helpers -------------------------------------------------------------
 * mstr_maxpool_nhwc : nn.MaxPool2d with -inf padding (feature_extraction.py:66,72,78).
 * mstr_mean_hw      : AdaptiveAvgPool2d(1) of the localisation net (transformation.py:57):
 *                     NHWC [B][H][W][C] -> [B][C] f32.
 * mstr_locnet_head  : the head of the localisation net in one launch (transformation.py:57,60-61,76-81):
 *                     AdaptiveAvgPool2d(1) -> Linear(C = 512, H1 = 256) + ReLU -> Linear(256, O = 2F), every product an
 *                     fp32 FMA.  feat NHWC [B][HW][C] (bf16 or f32); w1t = localization_fc1 weight TRANSPOSED to
 *                     [C][H1], w2t = localization_fc2 weight transposed to [H1][O]; out [B][O] f32.
 * mstr_mean_h       : model.py:92-93 permute+AdaptiveAvgPool2d((None,1))+squeeze:
 *                     NHWC [B][H][W][C] -> [B][W][C] (mean over H) in out_dtype (+ optional f32).
 * mstr_cast         : f32 <-> bf16 element cast.
 */
int mstr_maxpool_nhwc(const void* x, void* out, int dtype, int B, int H, int W, int C, int kh,
                      int kw, int sh, int sw, int ph, int pw, mstr_stream_t stream);
/* 2x2 / stride 2 max-pool whose INPUT is a padded-flat bf16 activation [B][H+2][W+2][C] (one-pixel ring, skipped), the
 * layout mstr_conv3x3_flat_bf16 writes; out [B][H/2][W/2][C] bf16 (modules/feature_extraction.py:66,116). */
int mstr_maxpool2x2_ring_bf16(const void* x_ring, void* out, int B, int H, int W, int C, mstr_stream_t stream);

/* 3x3 / stride 1 / pad 1 convolution + folded BatchNorm + ReLU over a PADDED-FLAT activation (csrc/conv_flat.cu):
 * replaces conv0_2 + bn0_2 + ReLU (modules/feature_extraction.py:61-63,113-115) in the bf16 mode.
 *   x   [B][H+2][W+2][C_in] bf16 whose one-pixel ring is zero (mstr_conv3x3_first with pool = 2 writes it)
 *   w   [C_out][3][3][C_in] bf16 (device)
 *   scale_host, shift_host [C_out] f32 in HOST memory (NULL = 1 / 0): the folded BatchNorm is passed to the kernel by
 *        value (kernel parameters are constant-bank operands of the epilogue's FMAs), read during this call
 *   out [B][H+2][W+2][C_out] bf16, ring written as zeros
 * Every activation row is loaded once per tile of 256 positions; the nine taps are shared-memory descriptor offsets.
 * C_in = 32 and C_out = 64 only. */
int mstr_conv3x3_flat_bf16(const void* x, const void* w, const float* scale_host, const float* shift_host, void* out,
                           int B, int H, int W, int C_in, int C_out, int relu, mstr_stream_t stream);
/* The same with the 2x2 / stride 2 max-pool that follows it (maxpool1, modules/feature_extraction.py:66,116) in the
 * epilogue: out is the POOLED plain NHWC activation [B][H/2][W/2][C_out] bf16; the pre-pool tensor never reaches HBM.
 * Bit-identical to mstr_conv3x3_flat_bf16 followed by mstr_maxpool2x2_ring_bf16.  H, W even; W <= 128 or W % 128 == 0. */
int mstr_conv3x3_flat_pool_bf16(const void* x, const void* w, const float* scale_host, const float* shift_host, void* out,
                                int B, int H, int W, int C_in, int C_out, int relu, mstr_stream_t stream);
/* Test hook of mstr_conv3x3_flat_bf16: 0 (default) leaves the "matrix base offset" field of the row-shifted A
 * descriptors zero - correct on B200, the swizzle is applied to absolute address bits; 1 sets it to the phase of the
 * start address (wrong result; kept as the experiment that pins the rule, tests/test_gpu_kernels.py). */
int mstr_debug_flat_desc_mode(int mode);
int mstr_mean_hw(const void* x, float* out, int dtype, int B, int H, int W, int C,
                 mstr_stream_t stream);
int mstr_locnet_head(const void* feat, int dtype, const float* w1t, const float* b1, const float* w2t, const float* b2,
                     float* out, int B, int HW, int C, int H1, int O, mstr_stream_t stream);
int mstr_mean_h(const void* x, void* out, float* out_f32, int dtype, int B, int H, int W, int C,
                mstr_stream_t stream);
int mstr_cast(const void* x, void* out, int src_dtype, int dst_dtype, int64_t n,
              mstr_stream_t stream);

/* ---- semantic vectors -----------------------------------------------------------------------
 * Replaces Linear_Embedding.forward (modules/semantic_vectors.py:33-49).
 *   ids   [B][n] f32 object ids (cast like .long());  table [rows][D] f32
 *   ious  [B][n] f32, mode 1 ('scene') only: out = softmax(ious, dim=1)[..., None] * table[ids]
 *   mode 0 = 'overlap' (plain gather).   out [B][n][D] f32.
 */
int mstr_embed_semantics(const float* ids, const float* table, const float* ious, float* out,
                         int B, int n, int D, int rows, int mode, mstr_stream_t stream);

/* ---- BiLSTM recurrence ----------------------------------------------------------------------
 * Replaces the recurrent part of nn.LSTM(bidirectional=True, batch_first=True)
 * (modules/encoders.py:14,24); the input projection x W_ih^T + b_ih + b_hh is a GEMM done with
 * mstr_conv_gemm_* beforehand.
 *   gates_x [B][T][2][4*Hd] f32   (dir 0 = forward, 1 = reverse; gate order i,f,g,o)
 *   w_hh    [2][4*Hd][Hd]   f32
 *   out_f32 / out_bf16  [B][T][2*Hd]  ([fwd | bwd]); either may be null.  Hd must be 256.
 */
int mstr_bilstm_recurrent(const float* gates_x, const float* w_hh, float* out_f32, void* out_bf16,
                          int B, int T, int Hd, mstr_stream_t stream);
/* Tensor-core variant for the bf16 throughput mode: h_t is kept in shared memory as bf16 (the
 * operand of tcgen05.mma), c_t in TMEM as f32, W_hh AND the gates_x tiles streamed by TMA (gates_x is
 * written into the TMEM accumulators before each chunk's MMAs), the output leaves by TMA store; gate
 * non-linearities use tanh.approx.  One launch for all T steps of both directions.
 *   w_hh_packed [2][4*Hd][Hd] bf16 with the rows of each direction re-ordered so that packed row
 *               chunk*128 + g*32 + j  holds original row  g*Hd + chunk*32 + j  (chunk = 0..7,
 *               g = gate i,f,g,o, j = 0..31): one 128-column accumulator chunk then carries all four
 *               gates of the same 32 hidden units.
 *   gates_x     [B][T][2][4*Hd] of gates_dtype: MSTR_F32, or MSTR_BF16 - the input-projection GEMM then writes (and
 *               this kernel reads) half the bytes; the values are widened to f32 when they enter the accumulators
 *   out_bf16    [B][T][2*Hd] bf16.   Hd must be 256.
 */
int mstr_bilstm_recurrent_bf16(const void* gates_x, int gates_dtype, const void* w_hh_packed, void* out_bf16, int B, int T,
                               int Hd, mstr_stream_t stream);
/* Two kernels implement mstr_bilstm_recurrent_bf16 with bit-identical results: one CTA per crop tile streaming W_hh
 * (csrc/bilstm_tc.cu) and a cluster of 8 CTAs per crop tile with stationary W_hh slices (csrc/bilstm_cluster.cu).
 * which = 0 (default) picks by occupancy, 1 / 2 force the first / second (used by the parity tests). */
int mstr_set_bilstm_kernel(int which);

/* ---- attention decoder ---------------------------------------------------------------------
 * Replaces Attention.forward (modules/decoders.py:26-64) with AttentionCell.forward (:79-101),
 * the generator (:17,:59) and the argmax feedback (:61-62), all steps in one persistent launch.
 *   enc   [B][T][Hd] f32 encoder output (batch_H);  proj [B][T][Hd] f32 = i2h(enc) (loop invariant, :81)
 *   w_h2h [Hd][Hd], b_h2h [Hd], w_score [Hd]
 *   w_ctx [4Hd][Hd]   = rnn.weight_ih[:, :Hd];  w_tok [C][4Hd] = rnn.weight_ih[:, Hd:]^T (one-hot == row gather)
 *   w_hh  [4Hd][Hd],  bias [4Hd] = rnn.bias_ih + rnn.bias_hh
 *   w_gen [C][Hd], b_gen [C]
 *   text  optional [B][steps] int64: teacher-forced tokens (is_train=True, :41-50); null => greedy (:52-62)
 *   h0/c0 optional [B][Hd] f32 initial state (null = zeros, :36-38); h_out/c_out optional final state
 *   logits [B][steps][C] f32 (must be zero-initialised by the caller when early_exit != 0)
 *   ids    [B][steps] int32 argmax per step; lengths [B] int32 = index of the first [s] (id 1) or steps
 *   alpha_out optional [B][steps][T] f32 attention weights (the `alpha` AttentionCell returns, :87,:101)
 *   early_exit != 0: a crop stops after it emits [s]; the reference always runs all steps
 *                    (decoders.py:34,56) but its callers cut at the first [s]
 *                    (training_functions.py:188-189).
 */
int mstr_attn_decode(const float* enc, const float* proj, const float* w_h2h, const float* b_h2h,
                     const float* w_score, const float* w_ctx, const float* w_tok,
                     const float* w_hh, const float* bias, const float* w_gen, const float* b_gen,
                     const int64_t* text, const float* h0, const float* c0, float* logits,
                     int32_t* ids, int32_t* lengths, float* alpha_out, float* h_out, float* c_out,
                     int B, int T, int Hd, int C, int steps, int early_exit, mstr_stream_t stream);

/* Tensor-core variant for the bf16 throughput mode (modules/decoders.py:26-64, :79-101).  ONE call = ONE
 * persistent cooperative launch that runs every step of the greedy (or teacher-forced) loop: per step a tcgen05 GEMM
 * for [h2h | generator] with the argmax in its epilogue, the Bahdanau attention over all SMs, and a tcgen05 GEMM for
 * the LSTMCell gates with the cell update in its epilogue, separated by grid-wide barriers (csrc/attn_tc.cu).
 * bf16 operands (enc, proj, h, ctx, weights), f32 accumulation, cell state, logits.
 *   enc_bf16, proj_bf16 [B][T][Hd] bf16
 *   w_head [384][Hd] bf16: rows 0..255 = h2h.weight, rows 256..256+C-1 = generator.weight, rest zero
 *   w_cat  [4*Hd][2*Hd] bf16: [rnn.weight_ih[:, :Hd] | rnn.weight_hh] with rows re-ordered so that every 64 rows hold
 *          the four gates of 16 hidden units (chunk*64 + g*16 + j <- g*Hd + chunk*16 + j): the cell GEMM then runs on
 *          64- or 128-column tiles
 *   w_tok  [C][4*Hd] f32 = rnn.weight_ih[:, Hd:]^T (original gate-major columns); bias [4*Hd] f32, or NULL when
 *          the caller has already added it to every row of w_tok (one load less per gate and unit)
 *   workspace: mstr_attn_decode_bf16_workspace_bytes(B, Hd) bytes of device memory (16-byte aligned)
 *   logits must be zero-initialised by the caller; ids / lengths / alpha_out as in mstr_attn_decode.
 */
int64_t mstr_attn_decode_bf16_workspace_bytes(int B, int Hd);
int mstr_attn_decode_bf16(const void* enc_bf16, const void* proj_bf16, const void* w_head, const float* b_h2h,
                          const float* b_gen, const float* w_score, const void* w_cat, const float* w_tok,
                          const float* bias, const int64_t* text, void* workspace, float* logits, int32_t* ids,
                          int32_t* lengths, float* alpha_out, int B, int T, int Hd, int C, int steps,
                          int early_exit, mstr_stream_t stream);

/* Driver of mstr_attn_decode_bf16: on = -1 (default) or 1: ONE persistent cooperative launch for all steps (faster at
 * every batch size measured, 256 .. 4096 crops); 0: the chain of three launches per step (programmatic dependent
 * launch), kept as the cross-check.  Both give bit-identical results (tests/test_gpu_kernels.py). */
int mstr_set_decode_persistent(int on);
/* Diagnostic: %globaltimer (ns) of CTA 0 at the start of the last persistent decode on the current device and at
 * every grid-barrier entry and exit (two values per barrier, three barriers per step) -> out_host[n]; optionally the
 * barrier-entry time of every CTA at the first 48 barriers -> per_cta_host[48 * 160] (barrier-major), or, with
 * n_cta = 192, the tile-level trace [hp, cell, logits][tile 0..3][16 slots] of step 6.  Synchronises
 * the device. */
int mstr_debug_decode_trace(unsigned long long* out_host, int n, unsigned long long* per_cta_host, int n_cta);

/* Input side of the path: the loader's transforms.Resize((32,100)) + ToTensor()
 * (coco_dataset.py:27,40-41) on a ragged batch of grayscale uint8 crops, bit-exact with Pillow's antialiased
 * fixed-point bilinear resample (what torchvision's Resize calls for PIL images).
 *   pixels    packed uint8 crops (row-major, no padding), device memory
 *   offsets   [B] int64 byte offset of each crop in `pixels`; sizes_hw [B][2] int32 (height, width), device
 *   out       [B][1][out_h][out_w] float32 in [0,1], device
 *   status    optional device int: set to (crop index + 1) if a crop is empty or larger than
 *             max_in_h x max_in_w (its output is then filled with NaN); may be NULL
 *   max_in_h, max_in_w  upper bounds of the crop sizes (they size the coefficient tables) */
int mstr_resize_to_tensor_u8(const void* pixels, const int64_t* offsets, const int32_t* sizes_hw, float* out,
                             int* status, int B, int out_h, int out_w, int max_in_h, int max_in_w,
                             mstr_stream_t stream);

/* ---- transformer encoder / decoder family ------------------------------------------
 * The reference's TF_Encoder (modules/encoders.py:68-143: PositionalEncoding, 6 custom post-norm layers, LayerNorm),
 * TF_Decoder (modules/decoders.py:103-275: nn.Embedding + PositionalEncoding, 6 decoder layers, greedy loop) and
 * Linear_Decoder (:278-296) are GEMMs (mstr_conv_gemm_* with 1x1 kernels: bias, ReLU and residual in the epilogue)
 * plus the kernels below.  dtype = MSTR_F32 or MSTR_BF16 is the storage type of the activations; all arithmetic is
 * fp32. */
/* nn.LayerNorm over the last dimension: out[r] = (x[r] - mean) / sqrt(var + eps) * gamma + beta (biased variance).
 * x, out [rows][D] of `dtype`; D % 32 == 0, D <= 1024. */
int mstr_layer_norm(const void* x, const float* gamma, const float* beta, void* out, int64_t rows, int D, float eps,
                    int dtype, mstr_stream_t stream);
/* PositionalEncoding.forward in eval mode (encoders.py:162-165): out[b][t][:] = x[b][t][:] + pe[t][:];
 * x [B][T][D] (f32 or bf16), pe [>=T][D] f32, out f32 or bf16 (f32 -> f32, f32 -> bf16, bf16 -> bf16). */
int mstr_add_pos_encoding(const void* x, int x_dtype, const float* pe, void* out, int out_dtype, int B, int T, int D,
                          mstr_stream_t stream);
/* nn.Embedding + PositionalEncoding of the decoder targets (decoders.py:117,256-262):
 * out[b][t][:] = emb[tokens[b][t]][:] + pe[t0 + t][:];  tokens [B][T] int32, emb [C][D] f32, out [B][T][D]. */
int mstr_embed_tokens(const int32_t* tokens, const float* emb, const float* pe, void* out, int out_dtype, int B, int T,
                      int D, int C, int t0, mstr_stream_t stream);
/* Core of nn.MultiheadAttention for whole sequences (encoders.py:186-187 self attention; decoders.py:441-442 with
 * the square subsequent mask when causal != 0; :450-451 attention over the memory): query i of crop b at
 * q + (b*Lq + i)*ldq, key j at k + (b*Lk + j)*ldkv, likewise v (elements; head h at column h*dh) - for a packed in_proj
 * output [B][L][3E] that is q = qkv, k = qkv + E, v = qkv + 2E, ldq = ldkv = 3E.
 * out [B][Lq][heads*dh] = softmax(q k^T / sqrt(dh)) v per head, before out_proj.  Lq, Lk <= 32, dh = 32 or 64. */
int mstr_mha(const void* q, int ldq, const void* k, const void* v, int ldkv, void* out, int B, int Lq, int Lk,
             int heads, int dh, int causal, int dtype, mstr_stream_t stream);
/* The same for ONE query per crop against Lk keys / values (greedy decode with a K/V cache, and the memory attention
 * decoders.py:450-451): q [B][ldq]; key j of crop b, head h at k_cache + b*kv_batch_stride + j*ldkv + h*dh (elements),
 * likewise v_cache; when k_new / v_new ([B][ld_new], head h at h*dh) are given they are first stored as row Lk-1 of
 * the caches.  out [B][heads*dh].  Lk <= 32, dh = 32 or 64. */
int mstr_mha_step(const void* q, int ldq, void* k_cache, void* v_cache, int ldkv, int64_t kv_batch_stride,
                  const void* k_new, const void* v_new, int ld_new, void* out, int B, int Lk, int heads, int dh,
                  int dtype, mstr_stream_t stream);
/* get_relevant_semantic of the reference's PRE_ENCODER_MLP / PRE_DECODER_MLP fusion (modules/encoders.py:89-111,
 * modules/decoders.py:147-170): for every (column t, object o) pair of a crop an MLP on cat(column, object) gives a score,
 * the scores are soft-maxed over the objects and weight the sum of the object vectors.
 * mstr_pair_add_relu:    h[b][t][o][:] = relu(a[b][t][:] + s[b][o][:]) - the MLP's first layer from the two halves of its
 *                        weight applied once per column (a, with the bias) and once per object (s);  a [B][T][C],
 *                        s [B][O][C], h [B][T][O][C] of `dtype`.
 * mstr_relevance_reduce: score_o = h2[b][t][o][:] . w3 + b3 (the MLP's last layer), p = softmax over o,
 *                        out[b][t][:] = sum_o p_o sem[b][o][:];  h2 [B][T][O][C] of `dtype`, w3 [C] f32, sem [B][O][E] f32,
 *                        out [B][T][E] of `dtype`;  O <= 64, E % 32 == 0, E <= 512. */
int mstr_pair_add_relu(const void* a, const void* s, void* h, int B, int T, int O, int C, int dtype, mstr_stream_t stream);
int mstr_relevance_reduce(const void* h2, const float* w3, float b3, const float* sem, void* out, int B, int T, int O, int C,
                          int E, int dtype, mstr_stream_t stream);
/* Exact GELU, 0.5 x (1 + erf(x / sqrt(2))): the activation of BertModel's intermediate layer (the reference's Oscar_Bert,
 * modules/encoders.py:34-66).  x, out [n] of `dtype` (may alias); n a multiple of 8 (bf16) / 4 (f32). */
int mstr_gelu(const void* x, void* out, int64_t n, int dtype, mstr_stream_t stream);
/* t_output[-1].max(1) of the greedy loop (decoders.py:276-279): logits [B][ld] f32 (ld >= C; columns >= C ignored);
 * copies the C logits of crop b to logits_out + b*out_row_stride (optional), stores the argmax (lowest index on ties)
 * to ids[b*ids_row_stride] and next_tok[b*next_stride] (both optional). */
int mstr_argmax_tokens(const float* logits, int ld, float* logits_out, int64_t out_row_stride, int32_t* ids,
                       int ids_row_stride, int32_t* next_tok, int next_stride, int B, int C, mstr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MSTR_B200_H_ */
