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

/* out = hi + lo (fp32) for n values (n % 4 == 0): the inverse of lfs2_split_bf16 up to 2^-17 relative */
LFS2_API int lfs2_merge_planes(const void* hi, const void* lo, float* out, long long n, void* stream);

/* hi = bf16(x), lo = bf16(x - hi) for n fp32 values (n % 4 == 0) */
LFS2_API int lfs2_split_bf16(const float* x, void* hi, void* lo, long long n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LFS2_H_ */
