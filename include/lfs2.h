/*
 * lfs2.h -- C ABI of the B200-native (sm_100a) mel-generation path of LightningFastSpeech2.
 *
 * The reference (MiniXC/LightningFastSpeech2) has NO native/FFI layer: its boundary is the
 * Python module API of litfass.fastspeech2.  This header is therefore the
 * NEW seam underneath that API: each entry point replaces the ATen call sequence of one
 * reference module/forward, cited per function as "replaces <file>:<lines>" (paths relative
 * to the reference root).  The host-side mirror that calls these (same class names, ctor
 * arguments and state_dict keys as the reference) lives in
 * lightningfastspeech2_b200/fastspeech2/{model,fastspeech2}.py; the ctypes binding a
 * reference maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter name ends in _host;
 *   - activations are row-major "channels-last": (B, T, d) float32, row m = b*T + t;
 *   - masks are uint8 (torch.bool storage), 1 = PAD, exactly like the reference's
 *     src_mask / tgt_mask / key_padding_mask;
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream);
 *     all calls are asynchronous on it and never synchronise the device;
 *   - no entry point allocates or frees device memory: outputs and workspaces are caller-owned;
 *   - return value: 0 = OK, negative = LFS2_ERR_*; lfs2_last_error() gives a thread-local
 *     message.  Nothing throws, aborts or falls back to the CPU.
 */
#ifndef LFS2_H_
#define LFS2_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define LFS2_API __attribute__((visibility("default")))
#else
#define LFS2_API
#endif

#define LFS2_OK 0
#define LFS2_ERR_INVALID_ARG (-1) /* bad shape / null pointer / misalignment            */
#define LFS2_ERR_UNSUPPORTED (-2) /* configuration outside what the kernels implement */
#define LFS2_ERR_CUDA (-3)        /* a CUDA runtime call or launch failed             */

LFS2_API int lfs2_version(void);
LFS2_API const char* lfs2_last_error(void);
/* compute capability major*10+minor of the current device, or negative error */
LFS2_API int lfs2_device_arch(void);

/* ---- A1: front end ------------------------------------------------------------------
 * replaces litfass/fastspeech2/model.py:137-143 (SpeakerEmbedding.forward):
 *   spk[b,:] = relu(W (d,in_dim) . dvec[b,:] + bias) */
LFS2_API int lfs2_speaker_proj(const float* dvec, const float* w, const float* bias, float* spk,
                      int batch, int in_dim, int d, void* stream);

/* replaces litfass/fastspeech2/fastspeech2.py:651-660 (src_mask, phone_embedding,
 * positional_encoding [model.py:53-55], + speaker term):
 *   src_mask[b,t] = phones[b,t]==0 ; x[b,t,:] = emb[phones[b,t],:] + pe[t,:] + spk[b,:]
 * phones is int64 as emitted by TTSDataset._collate_fn (dataset/datasets.py:852-882). */
LFS2_API int lfs2_embed_pe_spk(const int64_t* phones, const float* emb, const float* pe, const float* spk,
                      float* x, uint8_t* src_mask, int batch, int t, int d, int vocab, void* stream);

/* replaces fastspeech2.py:703-718: x[b,t,:] += pe[t,:] + spk[b,:]   (in place) */
LFS2_API int lfs2_add_pe_spk(float* x, const float* pe, const float* spk, int batch, int t, int d, void* stream);

/* ---- A2: FFTBlock (ConformerEncoderLayer.forward, model.py:108-122) ------------------
 * Linear / 1x1 conv: c (m,n) = a (m,k) . w (n,k)^T + bias (n) [, relu]
 * replaces the in_proj / out_proj GEMMs of nn.MultiheadAttention (model.py:111-114), the
 * pointwise Conv1d(.,.,1) layers (model.py:82,92), Linear(d,80) (fastspeech2.py:723). */
LFS2_API int lfs2_linear(const float* a, const float* w, const float* bias, float* c,
                int m, int n, int k, int relu, void* stream);

/* Dense Conv1d(d -> n, kernel ksize, zero "same" padding at the ends of the padded
 * sequence) on channels-last data, as an implicit GEMM over taps:
 *   c[b,t,:] = sum_j x[b,t+j-(ksize-1)/2,:] . wp[:, j*d:(j+1)*d]^T + bias [, relu]
 * wp is the (n, ksize*d) tap-major repack of the reference's (n, d, ksize) weight.
 * replaces model.py:95-106 (dense FFN conv) and model.py:529-536 (dense predictor conv). */
LFS2_API int lfs2_conv1d_dense(const float* x, const float* wp, const float* bias, float* c,
                      int batch, int t, int d, int n, int ksize, int relu, void* stream);

/* Depthwise Conv1d(d, d, ksize, groups=d), channels-last; wt is the (ksize, d) transpose
 * of the reference's (d,1,ksize) weight.  replaces model.py:75-81 and model.py:545-551. */
LFS2_API int lfs2_dwconv1d(const float* x, const float* wt, const float* bias, float* out,
                  int batch, int t, int d, int ksize, void* stream);
/* out = LayerNorm(x + y) with x (the residual stream) and the result as bf16 hi/lo planes, y fp32 or NULL: the
 * FFTBlock of widths without a LayerNorm GEMM epilogue (d != 256) stays in plane form from block to block
 * (model.py:114-115 at d = 768) */
LFS2_API int lfs2_add_layernorm_planes(const void* x_hi, const void* x_lo, const float* y, const float* gamma,
                                       const float* beta, void* out_hi, void* out_lo, int m, int d, float eps,
                                       void* stream);
/* The same over (batch, t, d) tensors with PAD-row skipping: rows of the 128-row groups of utterance b that start at or
 * after row_limit[b] + limit_extra are neither read nor written (row_limit NULL: every row).  The d != 256 FFTBlock of
 * FastSpeech2.skip_pad_rows (reference model.py:113-122 on the rows a valid frame can depend on). */
LFS2_API int lfs2_add_layernorm_planes_limited(const void* x_hi, const void* x_lo, const float* y, const float* gamma,
                                               const float* beta, void* out_hi, void* out_lo, int batch, int t, int d,
                                               float eps, const int* row_limit, int limit_extra, void* stream);
/* same, reading the input as fp32 (x) OR as bf16 hi/lo planes (x_hi, x_lo; x = NULL), and writing
 * the result as fp32 (out, may be NULL) and/or as hi/lo planes (the A operand of the following
 * pointwise lfs2_gemm_tc) */
LFS2_API int lfs2_dwconv1d_planes(const float* x, const void* x_hi, const void* x_lo, const float* wt,
                                  const float* bias, float* out, void* out_hi, void* out_lo, int batch, int t,
                                  int d, int ksize, void* stream);

/* Multi-head self attention core on a packed qkv (B,T,3d) tensor [q | k | v], heads =
 * contiguous d/nhead column blocks, q scaled by (d/nhead)^-1/2, PAD keys get -inf,
 * softmax over keys in fp32, ctx (B,T,d) = P.V.  Fully masked rows give NaN like the
 * reference.  replaces torch _sa_block / nn.MultiheadAttention as used at model.py:111-114. */
LFS2_API int lfs2_attention(const float* qkv, const uint8_t* key_padding_mask, float* ctx,
                   int batch, int t, int d, int nhead, void* stream);

/* out[m,:] = LayerNorm(x[m,:] (+ y[m,:] if y) ; gamma, beta, eps)   (y may be NULL)
 * replaces norm1/norm2 + residual (model.py:114-115) and nn.LayerNorm(filter) (model.py:538,556). */
LFS2_API int lfs2_add_layernorm(const float* x, const float* y, const float* gamma, const float* beta,
                       float* out, int m, int d, float eps, void* stream);

/* ---- A3/A4: variance predictors ------------------------------------------------------
 * out[m] = mask[m] ? 0 : dot(z[m,:], w) + bias[0]
 * replaces Linear(filter,1) + squeeze + masked_fill (model.py:512-518). */
LFS2_API int lfs2_rowdot_mask(const float* z, const float* w, const float* bias, const uint8_t* mask,
                     float* out, int m, int f, void* stream);

/* idx = bucketize(val*std+mean, bins[nbins-1], right=False) (or idx_forced if non-NULL);
 * x[m,:] += emb[idx,:]; optionally acc[m,:] (+)= emb[idx,:] and idx_out[m] = idx.
 * acc_mode: 0 = no acc, 1 = acc = emb, 2 = acc += emb.
 * replaces model.py:421-422 / 434-438 (VarianceEncoder) and model.py:329-333. */
LFS2_API int lfs2_bucket_embed_add(float* x, const float* val, float std, float mean, const float* bins,
                          int nbins, const float* emb, const int64_t* idx_forced, int64_t* idx_out,
                          float* acc, int acc_mode, int m, int d, void* stream);

/* PriorEmbedding.forward (model.py:146-164, used at fastspeech2.py:687-692): per-utterance scalar prior ->
 * out[b,:] = relu(emb[bucketize(prior[b], bins[nbins-1], right=False), :]) (B,d); idx_out (B) optional.
 * The broadcast add over time is lfs2_add_pe_spk with a zero positional table. */
LFS2_API int lfs2_prior_embed(const float* prior, const float* bins, int nbins, const float* emb, float* out,
                              int64_t* idx_out, int batch, int d, void* stream);

/* ---- A5: inference durations ----------------------------------------------------------
 * dur = int32(clamp(round_half_even(exp(log_dur) - 1), 0)); per utterance, if
 * sum(dur[valid]) <= n_valid/2 then dur[valid] = 1.   replaces model.py:300-309. */
LFS2_API int lfs2_duration_round_guard(const float* log_dur, const uint8_t* src_mask, int32_t* dur,
                              int batch, int tp, void* stream);

/* ---- A6: LengthRegulator.forward (model.py:349-370) ------------------------------------
 * scan: cum[b,p] = inclusive prefix sum of dur[b,:] (int64), lengths[b] = cum[b,Tp-1],
 *       *max_len = max_b lengths[b].  dur is int32 or int64 (dur_is_i64).
 * The caller reads *max_len back (the one host sync of the path), sets
 * L = min(max_len, (int)max_length) and allocates out/mask. */
LFS2_API int lfs2_length_regulate_scan(const void* dur, int dur_is_i64, int64_t* cum, int64_t* lengths,
                              int64_t* max_len, int batch, int tp, void* stream);
/* scatter: out[b,t,:] = t < lengths[b] ? x[b, #{p: cum[b,p] <= t}, :] : +0 ; mask[b,t] = t >= lengths[b]
 * rows are copied as raw bytes (row_bytes = d*sizeof(elem), multiple of 16): bit-exact for any dtype. */
LFS2_API int lfs2_length_regulate_scatter(const void* x, const int64_t* cum, const int64_t* lengths,
                                 void* out, uint8_t* mask, int batch, int tp, int l, int row_bytes,
                                 void* stream);

/* same with l output frames of which only those below cap (<= l) can be valid: frames in [cap, l) are PAD
 * (+0, mask = 1) for every utterance.  Used by the length-bucketed synthesis, which appends the decoder's
 * conv halo as explicit PAD frames to each bucket (cap = the reference's L). */
LFS2_API int lfs2_length_regulate_scatter_ex(const void* x, const int64_t* cum, const int64_t* lengths,
                                    void* out, uint8_t* mask, int batch, int tp, int l, int cap,
                                    int row_bytes, void* stream);

/* Ragged read-back of a synthesis batch: the valid rows of x (batch, l, width) fp32 (width % 4 == 0) packed back to back
 * in utterance order into out (sum_b min(lengths[b], l), width): what the reference's caller keeps of a batch
 * (synthesis/generator.py:164-170 cuts every mel at ~tgt_mask).  lengths = the LengthRegulator's frame counts. */
LFS2_API int lfs2_pack_valid_rows(const float* x, const int64_t* lengths, float* out, int batch, int l, int width,
                                  void* stream);

/* ---- tensor-core (tcgen05) GEMM / Conv1d with fused epilogues ---------------------------
 * Operands are bf16 "hi/lo" planes of fp32 values (x = hi + lo, see lfs2_split_bf16):
 *   a_hi/a_lo : (batch, t, d)      row-major bf16   (activations)
 *   w_hi/w_lo : (n, taps*d)        row-major bf16   (weights, tap-major for taps > 1)
 * acc[b,tt,:] = sum_j a[b, tt + j - (taps-1)/2, :] . w[:, j*d:(j+1)*d]^T   (rows outside [0,t) are zero)
 * npass = 3: hi.hi + lo.hi + hi.lo (fp32-parity, ~2^-16 relative); npass = 1: hi.hi only.
 * Residual: if res_hi/res_lo (batch, t, n) are given (with ident_hi = bf16 identity (n, n)), the
 * tensor core also accumulates res_hi.I + res_lo.I, i.e. acc += residual, before the epilogue.
 * Epilogue, in fp32:  v = acc + bias ; relu ? max(v,0) ; if gamma: v = LayerNorm(v) (needs n == 256);
 * written EITHER to out_f32 (batch*t, n) OR as hi/lo planes (out_hi, out_lo), by coalesced TMA stores.
 * taps = 1 replaces the Linear / pointwise Conv1d GEMMs (model.py:82,92,111-114,552;
 * fastspeech2.py:723); taps = k replaces the dense Conv1d(d, n, k) (model.py:95-106,529-536);
 * residual + LayerNorm replaces norm1/norm2 (model.py:114-115); ReLU + LayerNorm the predictor
 * ReLU -> LayerNorm (model.py:537-538,555-556). */
LFS2_API int lfs2_gemm_tc(const void* a_hi, const void* a_lo, int batch, int t, int d, int taps,
                          const void* w_hi, const void* w_lo, int n, const float* bias, int relu,
                          const void* res_hi, const void* res_lo, const void* ident_hi,
                          const float* gamma, const float* beta, float eps,
                          float* out_f32, void* out_hi, void* out_lo, int npass, void* stream);

/* lfs2_gemm_tc / lfs2_dwconv1d_planes restricted to the rows a caller needs: for utterance b every 128-row tile
 * that starts at or after row_limit[b] + limit_extra is skipped and its output rows are left untouched (A is
 * tiled per utterance: batch x t).  Used by the variance predictors, whose outputs on PAD rows are masked to 0 by
 * construction (model.py:518), so skipping rows farther than the conv halo beyond an utterance's end changes no
 * result bit.  row_limit = lfs2_mask_lengths of the padding mask.  The depthwise conv reads the input rows past the
 * last kept tile (t_row >= roundup128(row_limit[b] + limit_extra)) as zeros instead of from memory, so a chain of
 * row-limited kernels never consumes a row that none of them wrote. */
LFS2_API int lfs2_gemm_tc_limited(const void* a_hi, const void* a_lo, int batch, int t, int d, int taps,
                                  const void* w_hi, const void* w_lo, int n, const float* bias, int relu,
                                  const void* res_hi, const void* res_lo, const void* ident_hi,
                                  const float* gamma, const float* beta, float eps, float* out_f32, void* out_hi,
                                  void* out_lo, int npass, const int* row_limit, int limit_extra,
                                  void* workspace, void* stream);
/* General form of lfs2_gemm_tc(_limited):
 *   dilation   tap spacing of the Conv1d (rows t + (j - (taps-1)/2) * dilation; HiFi-GAN's dilated ResBlock convs,
 *              third_party/hifigan/models.py:24-57)
 *   activation 0 none | 1 ReLU | 2 leaky ReLU with `slope` (models.py:86-93), applied after the bias
 *   residual   res_hi/res_lo + ident_hi as in lfs2_gemm_tc; allowed without LayerNorm when npass = 3 (x = conv(.) + x)
 *   out_kind   LFS2_OUT_PLANES: out0/out1 = bf16 hi/lo planes; LFS2_OUT_F32: out0 = fp32; LFS2_OUT_F16: out0 = ONE
 *              fp16 plane (saturating conversion), the operand format of lfs2_attention_tc_ex(..., fp16)
 *   row_mask   NULL or (batch, t) bytes: rows with a non-zero byte are written as zeros (PAD frames of a ragged batch)
 *   npass      3: hi.hi + lo.hi + hi.lo; 1: hi.hi; 2: a_hi is ONE fp16 plane of a (a_lo unused) against fp16 hi/lo
 *              weight planes (lfs2_split_f16), a.w_hi + a.w_lo -- 11 significant bits on the activation side, full weights, two thirds
 *              of the tensor work (no LayerNorm / residual epilogue in this recipe)
 *   column tiles of 256 / 128 / 64 outputs (n % 16 == 0). */
#define LFS2_OUT_PLANES 0
#define LFS2_OUT_F32 1
#define LFS2_OUT_F16 2
#define LFS2_OUT_BF16 3 /* out0 = ONE bf16 plane: a result that only feeds single-pass (npass = 1) products */
LFS2_API int lfs2_gemm_tc_ex(const void* a_hi, const void* a_lo, int batch, int t, int d, int taps, int dilation,
                             const void* w_hi, const void* w_lo, int n, const float* bias, int activation, float slope,
                             const void* res_hi, const void* res_lo, const void* ident_hi, const float* gamma,
                             const float* beta, float eps, void* out0, void* out1, int out_kind, int npass,
                             const int* row_limit, int limit_extra, void* workspace, const uint8_t* row_mask,
                             void* stream);
/* One VarianceConvolutionLayer of a depthwise predictor stack (model.py:524-561, filter = channels = 256) with what
 * FOLLOWS it fused into the LayerNorm epilogue:  z = LayerNorm(relu(a . w^T + bias)), then
 *   next_dw_w != NULL: out = depthwise3(z) -- the NEXT layer's k = 3 depthwise conv (next_dw_w (3, 256) tap-major, next_dw_b
 *                      (256)), rows outside the utterance being zeros like Conv1d's padding -- as bf16 hi/lo planes
 *                      (z[r-1], z[r+1] come from the neighbouring epilogue threads; tiles advance by 126 rows);
 *   head_w != NULL:    head_out[b, t] = z . head_w + head_b, 0 where head_mask -- the predictor's Linear(256, 1) + masked_fill
 *                      (model.py:512-518); z itself is never written.
 * a = depthwise3(previous layer) as bf16 hi/lo planes (batch, t, 256).  row_limit / limit_extra / workspace as in
 * lfs2_gemm_tc_limited (workspace: lfs2_predictor_layer_tc_workspace_bytes; the stencil form has its own 126-row tile list). */
LFS2_API long long lfs2_predictor_layer_tc_workspace_bytes(int batch, int t);
LFS2_API int lfs2_predictor_layer_tc(const void* a_hi, const void* a_lo, int batch, int t, const void* w_hi,
                                     const void* w_lo, const float* bias, const float* gamma, const float* beta, float eps,
                                     int npass, const float* next_dw_w, const float* next_dw_b, void* out_hi, void* out_lo,
                                     const float* head_w, const float* head_b, const uint8_t* head_mask, float* head_out,
                                     const int* row_limit, int limit_extra, void* workspace, void* stream);
/* workspace of lfs2_gemm_tc_limited when row_limit != NULL: the compact list of active row tiles.  A later call with
 * row_limit == NULL and the same workspace reuses that list (same batch, t and limit: the layers of one predictor). */
LFS2_API long long lfs2_gemm_tc_limited_workspace_bytes(int batch, int t);
LFS2_API int lfs2_dwconv1d_planes_limited(const float* x, const void* x_hi, const void* x_lo, const float* wt,
                                          const float* bias, float* out, void* out_hi, void* out_lo, int batch,
                                          int t, int d, int ksize, const int* row_limit, int limit_extra,
                                          void* stream);
/* same with one more optional output: out_f16 = the result as ONE fp16 plane (saturating), the activation operand of a
 * 2-pass GEMM (npass = 2 of lfs2_gemm_tc_ex / lfs2_ffn_fused_tc_ex); odd kernel sizes <= 25 */
LFS2_API int lfs2_dwconv1d_planes_ex(const float* x, const void* x_hi, const void* x_lo, const float* wt,
                                     const float* bias, float* out, void* out_hi, void* out_lo, void* out_f16, int batch,
                                     int t, int d, int ksize, const int* row_limit, int limit_extra, void* stream);
/* x[r, :] = +0.0 for every row r with mask[r] != 0 (x: (rows, width) fp32, width % 4 == 0): the mel frames the
 * reference's consumers drop with tgt_mask (generator.py:164) after a PAD-row skipping synthesis call. */
LFS2_API int lfs2_zero_masked_rows(float* x, const uint8_t* mask, long long rows, int width, void* stream);
/* lengths[b] = 1 + index of the last row with pad_mask[b, .] == 0 (0 if all PAD); pad_mask NULL -> t */
LFS2_API int lfs2_mask_lengths(const uint8_t* pad_mask, int* lengths, int batch, int t, void* stream);

/* Fused position-wise FFN tail of the depthwise FFTBlock (model.py:118-122 after the depthwise conv; model width 256):
 *   out = LayerNorm( res + relu(u . w1^T + b1) . w2^T + b2 ; gamma, beta, eps )
 * u, res, out: (m, 256) bf16 hi/lo planes; w1 (f, 256), w2 (256, f) hi/lo planes (w2 = the folded
 * conv2.1 . blockdiag(conv2.0) matrix, b2 its folded bias); ident_hi = bf16 identity (256, 256).  The f-wide
 * intermediate stays in tensor memory (f % 256 == 0, f <= 2048).  npass as in lfs2_gemm_tc. */
LFS2_API int lfs2_ffn_fused_tc(const void* u_hi, const void* u_lo, int m, const void* w1_hi, const void* w1_lo, int f,
                               const float* b1, const void* w2_hi, const void* w2_lo, const float* b2,
                               const void* res_hi, const void* res_lo, const void* ident_hi, const float* gamma,
                               const float* beta, float eps, void* out_hi, void* out_lo, int npass, void* stream);
/* Same over the (batch, t) rows of a padded batch, skipping the 128-row tiles no utterance needs: utterance b needs its
 * rows t_row < roundup128(row_limit[b] + limit_extra) (the rule of lfs2_gemm_tc_limited); rows of skipped tiles are not
 * written.  workspace: lfs2_ffn_fused_tc_limited_workspace_bytes(batch, t) bytes; row_limit NULL with a workspace
 * filled by an earlier call = reuse that tile list; both NULL = every row.  Used by the PAD-row skipping synthesis
 * path (PAD rows past an utterance's end + conv halo cannot reach its valid frames). */
LFS2_API long long lfs2_ffn_fused_tc_limited_workspace_bytes(int batch, int t);
LFS2_API int lfs2_ffn_fused_tc_limited(const void* u_hi, const void* u_lo, int batch, int t, const void* w1_hi,
                                       const void* w1_lo, int f, const float* b1, const void* w2_hi, const void* w2_lo,
                                       const float* b2, const void* res_hi, const void* res_lo, const void* ident_hi,
                                       const float* gamma, const float* beta, float eps, void* out_hi, void* out_lo,
                                       int npass, const int* row_limit, int limit_extra, void* workspace, void* stream);
/* Same with the 2-pass recipe and an extra output.  npass = 2: u_hi is ONE fp16 plane of u (u_lo unused), w1 / w2 are
 * fp16 hi/lo planes (lfs2_split_f16) and the intermediate is packed as fp16: every product is a.w_hi + a.w_lo (11 significant bits on the activation side, full
 * weights; two thirds of the tensor work of npass = 3).  out_f16 (or NULL): the output rows also as one fp16 plane, the
 * activation operand of the next block's 2-pass QKV GEMM (lfs2_gemm_tc_ex, npass = 2). */
LFS2_API int lfs2_ffn_fused_tc_ex(const void* u_hi, const void* u_lo, int batch, int t, const void* w1_hi,
                                  const void* w1_lo, int f, const float* b1, const void* w2_hi, const void* w2_lo,
                                  const float* b2, const void* res_hi, const void* res_lo, const void* ident_hi,
                                  const float* gamma, const float* beta, float eps, void* out_hi, void* out_lo,
                                  void* out_f16, int npass, const int* row_limit, int limit_extra, void* workspace,
                                  void* stream);

/* tensor-core multi-head self attention (head_dim 128) on the bf16 hi/lo planes of the packed
 * qkv (B,T,3d) tensor [q | k | v] written by lfs2_gemm_tc: flash-style streaming softmax,
 * S = Q.K^T and O = P.V on tcgen05 with Q/P read from tensor memory, K/V tiles by TMA.
 * Semantics identical to lfs2_attention (q scaled by head_dim^-1/2, PAD keys -inf, fp32
 * softmax, fully masked rows NaN).  npass = 3: hi.hi + lo.hi + hi.lo (fp32-parity mode);
 * npass = 1: bf16 operands.  Outputs: ctx as hi/lo planes (B,T,d) and/or fp32.
 * workspace: lfs2_attention_tc_workspace_bytes(batch) bytes of device memory.
 * replaces torch _sa_block / nn.MultiheadAttention as used at model.py:111-114. */
LFS2_API int lfs2_attention_tc_workspace_bytes(int batch);
LFS2_API int lfs2_attention_tc(const void* qkv_hi, const void* qkv_lo, const uint8_t* key_padding_mask,
                               void* ctx_hi, void* ctx_lo, float* ctx_f32, void* workspace, int batch, int t,
                               int d, int nhead, int npass, void* stream);
/* Same, skipping the 128-row query tiles that start at or after row_limit[b] + limit_extra (their ctx rows are not
 * written); row_limit NULL = every row.  Keys are unaffected (PAD keys are masked anyway). */
LFS2_API int lfs2_attention_tc_limited(const void* qkv_hi, const void* qkv_lo, const uint8_t* key_padding_mask,
                                       void* ctx_hi, void* ctx_lo, float* ctx_f32, void* workspace, int batch, int t,
                                       int d, int nhead, int npass, const int* row_limit, int limit_extra,
                                       void* stream);
/* Same with the operand format of qkv made explicit: LFS2_OPERAND_BF16 = bf16 hi (+ lo for npass = 3) planes;
 * LFS2_OPERAND_F16 (npass = 1) = qkv_hi is ONE fp16 plane (written by lfs2_gemm_tc_ex(..., LFS2_OUT_F16)), Q.K^T and
 * P.V run as single fp16 passes with P held as fp16: 11 significant bits on every operand.  Measured on the fp64
 * goldens this keeps the mel within 1e-4 of the reference (budget 1e-3) at a third of the tensor work of npass = 3;
 * it is what compute mode "fp32" uses by default (tools/precision_emulation.py).  ctx is written as bf16
 * hi/lo planes and/or fp32 in every case. */
#define LFS2_OPERAND_BF16 0
#define LFS2_OPERAND_F16 1
LFS2_API int lfs2_attention_tc_ex(const void* qkv_hi, const void* qkv_lo, int operand_format,
                                  const uint8_t* key_padding_mask, void* ctx_hi, void* ctx_lo, float* ctx_f32,
                                  void* workspace, int batch, int t, int d, int nhead, int npass, const int* row_limit,
                                  int limit_extra, void* stream);
/* Wide heads (head_dim 256 or 384: the 76 M configuration has d = 768, 2 heads): the same flash attention with the O
 * accumulator in 128 * (head_dim / 128) tensor-memory columns, Q resident in shared memory (SS-form Q.K^T accumulated
 * over the head's 128-column chunks) and K / V streamed in [64 keys x 128 columns] slots.  qkv is ONE 16-bit plane
 * (B, T, 3d): bf16 (compute mode "bf16") or fp16 (compute mode "fp32"); ctx leaves as bf16 hi/lo planes and/or fp32.
 * No (T x T) tensor reaches HBM.  workspace: lfs2_attention_tc_workspace_bytes(batch). */
LFS2_API int lfs2_attention_tc_wide(const void* qkv, int operand_format, const uint8_t* key_padding_mask, void* ctx_hi,
                                    void* ctx_lo, float* ctx_f32, void* workspace, int batch, int t, int d, int nhead,
                                    const int* row_limit, int limit_extra, void* stream);

/* out = hi + lo (fp32) for n values (n % 4 == 0): the inverse of lfs2_split_bf16 up to 2^-17 relative */
LFS2_API int lfs2_merge_planes(const void* hi, const void* lo, float* out, long long n, void* stream);

/* hi = bf16(x), lo = bf16(x - hi) for n fp32 values (n % 4 == 0) */
LFS2_API int lfs2_split_bf16(const float* x, void* hi, void* lo, long long n, void* stream);
/* hi = fp16(x) (saturating), lo = fp16(x - hi): the WEIGHT planes of the 2-pass recipe (npass = 2), whose activation
 * operand is one fp16 plane -- a tensor-core instruction cannot mix an fp16 A with a bf16 B */
LFS2_API int lfs2_split_f16(const float* x, void* hi, void* lo, long long n, void* stream);
/* same, and (f16 != NULL) the values also as ONE fp16 plane (saturating): the activation operand of npass = 2 GEMMs */
LFS2_API int lfs2_split_bf16_ex(const float* x, void* hi, void* lo, void* f16, long long n, void* stream);

/* ---- general batched tcgen05 GEMM with per-operand majorness (weight gradients, attention products) ----
 * An operand is a window of a dense row-major bf16 tensor (d2, d1, d0) (d0 innermost, d0 % 8 == 0), given
 * as hi/lo planes.  mn_major = 0 ("K-major"): tensor rows = the operand's m (n) index, columns = the
 * contraction index k.  mn_major = 1 ("MN-major"): tensor ROWS = the contraction index, columns = the
 * m (n) index, i.e. the operand is used transposed without a transposed copy.  For batch element
 * z = b * nhead + h the window starts at column col0 + h * hstride of tensor slice (per_z ? z : b). */
typedef struct lfs2_operand {
  int mn_major;
  int d0, d1, d2;
  int col0, hstride, per_z;
} lfs2_operand;
/*   C[z] (m x n, fp32, row pitch ldc, at c + b*c_bstride + h*c_hstride)  (+)=  A[z] (m x k) . B[z] (n x k)^T
 * npass = 3: hi.hi + lo.hi + hi.lo; npass = 1: hi.hi.  accumulate != 0 adds into C with fp32 atomics and
 * lets the kernel split the contraction across CTAs (C must be initialised); 0 overwrites C.
 * replaces (as gradients) the Linear / Conv1d(.,.,1) weight gradients torch.autograd computes for
 * model.py:82,92,111-114,552 and fastspeech2.py:723, and the bmm's of torch MHA (model.py:111-114). */
LFS2_API int lfs2_gemm_tc2(const void* a_hi, const void* a_lo, const lfs2_operand* a, const void* b_hi,
                           const void* b_lo, const lfs2_operand* b, float* c, int ldc, long long c_bstride,
                           long long c_hstride, int m, int n, int k, int nbatch, int nhead, int npass,
                           int accumulate, void* stream);

/* GEMM-decomposed attention, row-wise pieces (see csrc/attention_mat.cu).  Z = batch*nhead, tp = t rounded
 * up to a multiple of 8.
 * softmax: s (Z,t,tp) raw Q.K^T logits -> P = softmax(s*scale with PAD keys masked) as bf16 hi/lo planes
 *          (Z,t,tp) (p_lo may be NULL), lse (Z,t) of the scaled logits; pad columns / PAD keys get 0. */
LFS2_API int lfs2_attn_softmax_planes(const float* s, const uint8_t* key_padding_mask, void* p_hi, void* p_lo,
                                      float* lse, int batch, int nhead, int t, int tp, float scale, void* stream);
/* the same with attention-probability dropout (torch MHA dropout, model.py:111) fused: additionally writes
 * PM = P o mask / (1 - p) (pm_hi / pm_lo, mask as lfs2_dropout_planes over the (Z,t,tp) index space): PM feeds P.V and
 * P^T.dO, P stays for the softmax backward */
LFS2_API int lfs2_attn_softmax_planes_drop(const float* s, const uint8_t* key_padding_mask, void* p_hi, void* p_lo,
                                           void* pm_hi, void* pm_lo, float* lse, int batch, int nhead, int t, int tp,
                                           float scale, float drop_p, unsigned long long drop_seed,
                                           unsigned int drop_site, void* stream);
/* delta (Z,t) = rowsum(dctx o ctx) per head */
LFS2_API int lfs2_attn_delta(const float* dctx, const float* ctx, float* delta, int batch, int t, int d, int nhead,
                             void* stream);
/* dS = scale * P o (dP - delta) as hi/lo planes (Z,t,tp) (lo pointers may be NULL) */
LFS2_API int lfs2_attn_ds_planes(const void* p_hi, const void* p_lo, const float* dp, const float* delta,
                                 void* ds_hi, void* ds_lo, int batch, int nhead, int t, int tp, float scale,
                                 void* stream);

/* the same with the dropout mask of the forward applied to dP first (dP_eff = mask/(1-p) o dP) */
LFS2_API int lfs2_attn_ds_planes_drop(const void* p_hi, const void* p_lo, const float* dp, const float* delta,
                                      void* ds_hi, void* ds_lo, int batch, int nhead, int t, int tp, float scale,
                                      float drop_p, unsigned long long drop_seed, unsigned int drop_site,
                                      void* stream);

/* ================= train-step config: backward kernels, loss, optimizer ===================
 * The reference obtains all of these from torch.autograd / torch.optim; the citations name the
 * forward construct each gradient belongs to.  Parameter-gradient outputs ACCUMULATE (+=) into
 * caller-owned buffers (zeroed once per step by lfs2_adamw_step) using fp32 atomics. */

/* lfs2_add_layernorm that also saves the pre-norm sum z = x (+ y) (m,d) and the per-row
 * (mean, rstd) pairs stats (m,2) needed by lfs2_layernorm_bwd; z_out / stats may be NULL. */
LFS2_API int lfs2_add_layernorm_train(const float* x, const float* y, const float* gamma, const float* beta,
                                      float* out, float* z_out, float* stats, int m, int d, float eps,
                                      float drop_p, unsigned long long drop_seed, unsigned int drop_site,
                                      void* stream);
/* (drop_p > 0 fuses dropout1 / dropout2 of the FFTBlock, model.py:114-115: z = x + dropout(y), mask as lfs2_dropout) */
/* backward of out = LayerNorm(z; gamma, beta) (model.py:114-115, 538, 556):
 *   dz = rstd * (dy*gamma - mean(dy*gamma) - xhat * mean(dy*gamma*xhat)) (+ add if non-NULL)
 *   dgamma += sum_m dy * xhat ; dbeta += sum_m dy */
LFS2_API int lfs2_layernorm_bwd(const float* dy, const float* z, const float* stats, const float* gamma,
                                const float* add, float* dz, float* dgamma, float* dbeta, int m, int d,
                                void* stream);

/* same, plus dz_drop = dropout(dz) with the mask of (drop_p, drop_seed, drop_site): the gradient entering the branch
 * whose output was dropped in the forward pass (dz itself continues along the residual) */
LFS2_API int lfs2_layernorm_bwd_drop(const float* dy, const float* z, const float* stats, const float* gamma,
                                     const float* add, float* dz, float* dz_drop, float* dgamma, float* dbeta,
                                     int m, int d, float drop_p, unsigned long long drop_seed,
                                     unsigned int drop_site, void* stream);
/* same with the incoming gradient given as two summands, dy + dy2 (dy2 may be NULL): the residual joins of the FFTBlock
 * backward (x + sublayer(x), model.py:113-115) are added on load instead of by a kernel of their own */
LFS2_API int lfs2_layernorm_bwd_ex(const float* dy, const float* dy2, const float* z, const float* stats,
                                   const float* gamma, const float* add, float* dz, float* dz_drop, float* dgamma,
                                   float* dbeta, int m, int d, float drop_p, unsigned long long drop_seed,
                                   unsigned int drop_site, void* stream);

/* weight gradient of Linear / pointwise Conv1d / one tap of a dense Conv1d:
 *   c[n,k] += sum_r a[r, n] * b[r + shift, k]      (r over m rows; a is dY, b is the layer input)
 * t > 0: rows are frames of utterances of length t and b-rows shifted outside [0,t) count as zero
 * (Conv1d "same" padding); t = 0 requires shift = 0.  lda/ldb/ldc = leading dimensions. */
LFS2_API int lfs2_gemm_tn(const float* a, const float* b, float* c, int m, int n, int k, int lda, int ldb,
                          int ldc, int t, int shift, void* stream);
/* bias gradient: out[n] += sum_m a[m,n] */
LFS2_API int lfs2_colsum(const float* a, float* out, int m, int n, void* stream);
/* dx = y > 0 ? dy : 0 (y = the ReLU output; dx may alias dy) */
LFS2_API int lfs2_relu_bwd(const float* dy, const float* y, float* dx, long long n, void* stream);
/* dx = y > 0 ? dy * scale : 0: ReLU followed by dropout (model.py:120) in one pass -- y is the saved DROPPED
 * activation, so y > 0 is relu-mask and keep-mask at once and scale = 1/(1-p) */
LFS2_API int lfs2_relu_bwd_scaled(const float* dy, const float* y, float* dx, long long n, float scale, void* stream);
/* The same gradient written as bf16 hi/lo planes (the operand format of the input- and weight-gradient GEMMs that
 * consume it) with the bias gradient in the same pass: dx = y > 0 ? dy * scale : 0, db[c] += sum_r dx[r, c] (db may be
 * NULL).  y = the ReLU output either as fp32 (y_f32) or as the hi plane of its bf16 split (y_hi) -- exactly one of the
 * two; rows x cols row-major, cols % 4 == 0.  Replaces relu_bwd + split_bf16 + colsum of the train step
 * (autograd of F.relu / nn.Dropout / Conv1d bias in reference model.py:117-121, 539-557). */
LFS2_API int lfs2_relu_bwd_planes(const float* dy, const float* y_f32, const void* y_hi, void* dx_hi, void* dx_lo,
                                  float* db, int rows, int cols, float scale, void* stream);
/* Per-step weight re-formatting of the train step in one launch.  Each entry describes one fp32 weight matrix
 * (rows, cols; row-major) and where its operand forms go: bf16 hi/lo planes of W (same layout; hi == NULL: skip) and of
 * W^T ((cols, rows) row-major; hi_t == NULL: skip) -- what lfs2_split_bf16(W) and lfs2_split_bf16(lfs2_transpose(W))
 * produce, i.e. the forward and the input-gradient operand of Linear / 1x1 Conv1d weights (autograd of reference
 * model.py:82,92,111-114,552).  tile_begin = index of the entry's first 32 x 32 tile in the launch (entries sorted by it,
 * the first one 0); total_tiles = sum of ceil(rows/32) * ceil(cols/32).  `entries` is DEVICE memory. */
typedef struct lfs2_prep_entry {
  const float* src;
  void* hi;
  void* lo;
  void* hi_t;
  void* lo_t;
  int rows, cols;
  int tile_begin;
  int reserved;
} lfs2_prep_entry;
LFS2_API int lfs2_weight_planes_batched(const lfs2_prep_entry* entries, int n_entries, int total_tiles, void* stream);
/* dst += src */
LFS2_API int lfs2_add_inplace(float* dst, const float* src, long long n, void* stream);
/* out (cols, rows) = in (rows, cols)^T -- transposed weight copies for the input-gradient GEMMs */
LFS2_API int lfs2_transpose(const float* in, float* out, int rows, int cols, void* stream);

/* depthwise Conv1d weight/bias gradient (model.py:75-81, 545-551):
 *   dwt[j,c] += sum_{b,t} dy[b,t,c] * x[b,t+j-(ksize-1)/2,c] ; dbias[c] += sum dy   (dbias may be NULL)
 * the input gradient is lfs2_dwconv1d itself with the taps reversed and a zero bias. */
LFS2_API int lfs2_dwconv1d_bwd_w(const float* dy, const float* x, float* dwt, float* dbias, int batch, int t,
                                 int d, int ksize, void* stream);

/* lfs2_attention that also writes lse (B, nhead, T) = log-sum-exp of the scaled, masked logits */
LFS2_API int lfs2_attention_lse(const float* qkv, const uint8_t* key_padding_mask, float* ctx, float* lse,
                                int batch, int t, int d, int nhead, void* stream);
/* backward of the attention core: dqkv (B,T,3d) = [dq | dk | dv] from qkv, ctx = forward output,
 * dctx and the saved lse; probabilities are recomputed tile by tile.  workspace:
 * lfs2_attention_bwd_workspace_bytes(batch, t, nhead) bytes. */
LFS2_API long long lfs2_attention_bwd_workspace_bytes(int batch, int t, int nhead);
LFS2_API int lfs2_attention_bwd(const float* qkv, const float* ctx, const float* dctx, const float* lse,
                                const uint8_t* key_padding_mask, float* dqkv, void* workspace, int batch,
                                int t, int d, int nhead, void* stream);

/* LengthRegulator backward (model.py:349-370): dx[b,p,:] = sum of dout[b,f,:] over the frames
 * f in [cum[b,p-1], min(cum[b,p], l)) phone p was repeated into (cum from lfs2_length_regulate_scan). */
LFS2_API int lfs2_length_regulate_bwd(const float* dout, const int64_t* cum, float* dx, int batch, int tp,
                                      int l, int d, void* stream);

/* nn.Embedding weight gradient (fastspeech2.py:653; model.py:401,422): demb[idx[m],:] += dx[m,:];
 * rows with idx == skip_idx get none (padding_idx; pass -1 for no padding row). */
LFS2_API int lfs2_embedding_bwd(const float* dx, const int64_t* idx, float* demb, int m, int d, int nrows_emb,
                                long long skip_idx, void* stream);
/* Decoder input in one pass (model.py:263-266 for the LAST frame-level variance encoder + fastspeech2.py:716-721 +
 * lfs2_split_bf16): y[b,t,:] = ((x[b,t,:] + emb[bucket(val[b,t]),:]) + pe[t,:]) + spk[b,:], written only as bf16 hi/lo
 * planes (and, out_f16 != NULL, one fp16 plane): the operands of the first decoder block.  Arguments val .. acc_mode as
 * in lfs2_bucket_embed_add (acc receives / accumulates the embedding term); emb == NULL skips the bucket term.  Same
 * operation order as lfs2_bucket_embed_add -> lfs2_add_pe_spk -> lfs2_split_bf16, so the planes are bit-identical. */
LFS2_API int lfs2_decoder_input_planes(const float* x, const float* val, float std, float mean, const float* bins,
                                       int nbins, const float* emb, const int64_t* idx_forced, int64_t* idx_out,
                                       float* acc, int acc_mode, const float* pe, const float* spk, int batch, int t,
                                       int d, void* out_hi, void* out_lo, void* out_f16, void* stream);
/* ---- stochastic duration predictor, inference direction (third_party/stochastic_duration_predictor/
 * sdp.py:11-164, 254-269, 330-349, transforms.py:50-212; model.py:299-309, 463-480) ----------------------------------
 * Phoneme-level, channels-last (B, T, C) fp32; the 1x1 convolutions between these kernels are lfs2_linear.
 * lfs2_sdp_dwconv: depthwise Conv1d with a tap dilation, rows whose pad_mask byte is non-zero read as zeros (x * x_mask,
 *   sdp.py:62); wt (ksize, c) tap-major.
 * lfs2_sdp_ln_gelu: out = [res +] gelu_erf(LayerNorm_c(y; gamma, beta, eps))   (sdp.py:63-69).
 * lfs2_sdp_flow_pre: h[m,:] = z[m, channel] * w[:] + bias[:] + g[m,:]          (sdp.py:141-143 with the conditioning).
 * lfs2_sdp_spline_inverse: one spline coupling flow in reverse, in place on the flow state z (m, 2): channel x1_channel
 *   goes through the inverse of the 10-bin monotone rational-quadratic spline on [-tail_bound, tail_bound] (identity
 *   outside) whose 29 parameters are h[m, 0:29] (row stride h_stride; widths and heights divided by sqrt(hidden_channels)),
 *   PAD rows are zeroed (sdp.py:147-164, transforms.py:50-212 with inverse=True).
 * lfs2_sdp_affine_reverse: z[m, c ^ flip] = (z[m, c ^ flip] - translation[c]) * exp(-log_scale[c]), PAD rows zero (sdp.py:93-95).
 * lfs2_sdp_durations: dur = int32(clamp(ceil(exp(logw + 1e-9)), 0)), 0 where logw == 0, and the all-ones guard (model.py:302-309). */
LFS2_API int lfs2_sdp_dwconv(const float* x, const uint8_t* pad_mask, const float* wt, const float* bias, float* out,
                             int batch, int t, int c, int ksize, int dilation, void* stream);
LFS2_API int lfs2_sdp_ln_gelu(const float* y, const float* gamma, const float* beta, float eps, const float* res, float* out,
                              int m, int c, void* stream);
LFS2_API int lfs2_sdp_flow_pre(const float* z, int channel, const float* w, const float* bias, const float* g, float* out,
                               int m, int c, void* stream);
LFS2_API int lfs2_sdp_spline_inverse(float* z, int x1_channel, const float* h, int h_stride, const uint8_t* pad_mask,
                                     int hidden_channels, float tail_bound, int m, void* stream);
LFS2_API int lfs2_sdp_affine_reverse(float* z, const float* translation, const float* log_scale, const uint8_t* pad_mask,
                                     int flip, int m, void* stream);
LFS2_API int lfs2_sdp_durations(const float* logw, const uint8_t* src_mask, int32_t* dur, int batch, int tp, void* stream);

/* out-of-place lfs2_bucket_embed_add: x[m,:] = x_in[m,:] + emb[idx,:] (the input stays intact for
 * the predictor's backward pass) */
LFS2_API int lfs2_bucket_embed_add_oop(const float* x_in, float* x, const float* val, float std, float mean,
                                       const float* bins, int nbins, const float* emb,
                                       const int64_t* idx_forced, int64_t* idx_out, float* acc, int acc_mode,
                                       int m, int d, void* stream);

/* backward of lfs2_rowdot_mask (model.py:512-518): g = mask[m] ? 0 : dout[m];
 *   dz[m,:] = g * w ; dw += sum_m g * z[m,:] ; db += sum_m g */
LFS2_API int lfs2_rowdot_mask_bwd(const float* dout, const float* z, const float* w, const uint8_t* mask,
                                  float* dz, float* dw, float* db, int m, int f, void* stream);

/* gradient of the broadcast speaker term (model.py:137-143 used at fastspeech2.py:658,707):
 *   out[b,:] += sum_t dx[b,t,:] */
LFS2_API int lfs2_sum_over_time(const float* dx, float* out, int batch, int t, int d, void* stream);

/* conv2 of the depthwise FFTBlock (model.py:85-93): grouped 1x1 conv2.0 (weight w20 (F, g, 1),
 * groups = d, g = F/d) followed by pointwise conv2.1 (weight w21 (d_out, F, 1)) folded into
 *   w_eff (d_out, F) = w21 . blockdiag(w20) ; b_eff = b21 + w21 . b20
 * and the chain rule back to the four reference parameters. */
LFS2_API int lfs2_fold_pw_fwd(const float* w21, const float* w20, const float* b20, const float* b21,
                              float* w_eff, float* b_eff, int d_out, int groups, int g, void* stream);
LFS2_API int lfs2_fold_pw_bwd(const float* dw_eff, const float* db_eff, const float* w21, const float* w20,
                              const float* b20, float* dw21, float* dw20, float* db20, float* db21, int d_out,
                              int groups, int g, void* stream);

/* dropout (nn.Dropout at model.py:42,55,111-122,539,557): y = x * keep / (1 - p), keep(i) a pure function
 * of (seed, site, i) via Philox4x32-10 -- calling it again on a gradient with the same (seed, site)
 * IS the backward pass (no stored mask).  y may alias x.  The stream differs from PyTorch's generator. */
LFS2_API int lfs2_dropout(const float* x, float* y, long long n, float p, unsigned long long seed,
                          unsigned int site, void* stream);
/* same on bf16 hi/lo planes (attention probabilities); lo pointers may be NULL */
LFS2_API int lfs2_dropout_planes(const void* in_hi, const void* in_lo, void* out_hi, void* out_lo, long long n,
                                 float p, unsigned long long seed, unsigned int site, void* stream);

/* ---- A9: FastSpeech2Loss default branches (loss.py:156-187) ---------------------------
 * loss = mean over rows with pad_mask == 0 (and all `inner` columns) of |pred - tgt| (kind 0)
 * or (pred - tgt)^2 (kind 1); tgt = target, or log(target_i64 + 1) for the duration loss.
 * *loss_out = loss ; if total_out: *total_out += weight * loss (loss.py:204-211) ;
 * if dpred: dpred = d(weight * loss)/d pred (zero on PAD rows).  workspace: 2 floats. */
LFS2_API int lfs2_masked_loss(const float* pred, const float* target, const int64_t* target_i64,
                              const uint8_t* pad_mask, int rows, int inner, int kind, float weight,
                              float* loss_out, float* total_out, float* dpred, float* workspace, void* stream);

/* x[i] *= *scalar (device scalar): applies an upstream d(total)/d(total) factor to the stored loss gradients */
LFS2_API int lfs2_scale_by(float* x, const float* scalar, long long n, void* stream);

/* ---- A10: AdamW + Noam (fastspeech2.py:1166-1182, noam.py:20-25) -----------------------
 * out[0] += sum x^2 (global gradient norm for clipping) */
LFS2_API int lfs2_sumsq(const float* x, float* out, long long n, void* stream);
/* one AdamW step over flat buffers, torch.optim.AdamW operation order; the caller passes the
 * Noam-scheduled lr.  g is scaled by grad_scale (1/world_size) and, if max_norm > 0, by
 * min(1, max_norm / (sqrt(*gnorm_sq) * grad_scale + 1e-6)); zero_grad != 0 clears g afterwards. */
LFS2_API int lfs2_adamw_step(float* p, float* g, float* m, float* v, long long n, float lr, float beta1,
                             float beta2, float eps, float weight_decay, int step, float grad_scale,
                             float max_norm, const float* gnorm_sq, int zero_grad, void* stream);

/* ---- HiFi-GAN generator (reference litfass/third_party/hifigan/models.py:112-174) ----
 * The convolutions run on lfs2_gemm_tc_ex (channels-last bf16 hi/lo planes; dilated taps; leaky-ReLU / residual
 * epilogues; ConvTranspose1d(stride u, kernel 2u, padding u/2) as a 3-tap polyphase convolution onto u*C_out columns);
 * these are the element-wise stages between them.  n = element count (multiple of 8). */
/* y = leaky_relu(x, slope) on planes (models.py:87,89) */
LFS2_API int lfs2_lrelu_planes(const void* in_hi, const void* in_lo, void* out_hi, void* out_lo, long long n,
                               float slope, void* stream);
/* y = leaky_relu(((a + b) + c) * scale, slope): the multi-receptive-field average of the three ResBlocks followed by
 * the next stage's activation (models.py:157-166) */
LFS2_API int lfs2_mean3_lrelu_planes(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo,
                                     const void* c_hi, const void* c_lo, void* out_hi, void* out_lo, long long n,
                                     float scale, float slope, void* stream);
/* mel (batch, c, t) fp32 channels-first (what Generator.forward takes, hifigan/__init__.py:37) -> (batch, t, c_padded)
 * channels-last planes; channels >= c and frames t >= lengths[b] (lengths may be NULL) are zero */
LFS2_API int lfs2_mel_to_planes(const float* mel, const int* lengths, void* out_hi, void* out_lo, int batch, int c,
                                int t, int c_padded, void* stream);
/* out[b, t] = tanh(Conv1d(c -> 1, ksize, "same")(leaky_relu(x, slope)))[b, t] (models.py:167-169); w is (ksize * c)
 * tap-major; samples t >= lengths[b] are read as zero padding and written as 0 (lengths may be NULL) */
LFS2_API int lfs2_conv_post_tanh(const void* x_hi, const void* x_lo, const float* w, const float* bias,
                                 const int* lengths, float slope, float* out, int batch, int t, int c, int ksize,
                                 void* stream);

/* ---- FastDiff variance adaptor glue (reference litfass/fastspeech2/fastdiff_variances.py) ----
 * The noise-predicting networks reuse the predictor-stack kernels (lfs2_dwconv1d_planes + lfs2_gemm_tc with the
 * ReLU + LayerNorm epilogue + lfs2_rowdot_mask); these are the element-wise pieces around them. */
/* out (batch, dim) = [sin(steps[b] * e_i) | cos(steps[b] * e_i)], e_i = 10000^(-i / (dim/2 - 1))
 * (third_party/fastdiff/module/util.py:318-343) */
LFS2_API int lfs2_diffusion_step_embed(const float* steps, float* out, int batch, int dim, void* stream);
/* x <- x * sigmoid(x) in place (FastDiff.py swish; fastdiff_variances.py:196-197) */
LFS2_API int lfs2_swish(float* x, long long n, void* stream);
/* out[b,t,:] = (xt[b,t] * w_in + b_in + c[b,t,:]) + noise_embed[b,:]: the scalar track lifted to d channels by
 * Linear(1, d), plus the condition, plus the projected step embedding (fastdiff_variances.py:199-208) */
LFS2_API int lfs2_diffusion_input(const float* xt, const float* w_in, const float* b_in, const float* c,
                                  const float* noise_embed, float* out, int batch, int t, int d, void* stream);
/* out[b,t] = (a[b] * x[b,t] + e[b] * y[b,t]) * s[b] + g[b] * z[b,t] + add with per-utterance coefficient vectors (a, y/e,
 * s, z/g optional); positions with zero_mask[b,t] != 0 (optional) are written as 0: q(x_t | x_0) = alpha_t x_0 + delta_t z
 * (:185-190), the DDPM reverse update x <- (x - k eps) / sqrt(1 - beta) + sigma noise (util.py:224-228) and the affine
 * de-normalisation of the duration track (fastdiff_variances.py:108) */
LFS2_API int lfs2_diffusion_mix(const float* x, const float* y, const float* z, const float* a, const float* e,
                                const float* s, const float* g, float add, const uint8_t* zero_mask, float* out,
                                int batch, int t, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LFS2_H_ */
