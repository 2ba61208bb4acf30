// Train-step tail of the path: FastSpeech2Loss default branches (reference
// litfass/fastspeech2/loss.py:156-187, 204-211) and the AdamW + Noam update (reference
// fastspeech2.py:1166-1182, noam.py:20-25) as HBM-bound streaming kernels.
//
//   masked_loss   mean over VALID rows of |pred - tgt| (L1) or (pred - tgt)^2 (MSE), plus, in the same
//                 pass, d(weight * loss)/d pred -- the loss is a leaf of the graph, so its gradient
//                 is produced together with its value; loss value and the weighted total stay on
//                 the device (no host sync).
//   adamw_step    one fused pass over the flat parameter / gradient / moment buffers (28 B/param):
//                 optional global-norm clipping, 1/world_size gradient averaging, decoupled weight
//                 decay, bias-corrected update in torch.optim.AdamW's operation order; zeroes the
//                 gradient buffer for the next step.
#include <math.h>

#include "common.cuh"

namespace lfs2 {

// ws[0] = number of valid rows
__global__ void mask_count_kernel(const uint8_t* __restrict__ pad_mask, float* __restrict__ ws, int rows) {
  int cnt = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += gridDim.x * blockDim.x)
    cnt += pad_mask ? !pad_mask[i] : 1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(ws, (float)cnt);
}

// element e of row r: diff = pred - tgt ; ws[1] += sum f(diff) ; dpred = weight / (count * inner) * f'(diff)
__global__ void __launch_bounds__(256)
masked_loss_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, const int64_t* __restrict__ tgt_i64,
                   const uint8_t* __restrict__ pad_mask, float* __restrict__ dpred, float* __restrict__ ws,
                   size_t total, int inner, int kind, float weight) {
  const float inv = 1.f / (ws[0] * (float)inner);  // 1/0 = inf -> NaN loss like torch's mean over nothing
  float s = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t row = i / inner;
    float g = 0.f;
    if (!(pad_mask && pad_mask[row])) {
      // duration target: log(duration + 1), computed like torch.log(int64 + 1) -> fp32
      const float tv = tgt_i64 ? logf((float)(tgt_i64[i] + 1)) : tgt[i];
      const float diff = pred[i] - tv;
      if (kind == 0) {
        s += fabsf(diff);
        g = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
      } else {
        s = fmaf(diff, diff, s);
        g = 2.f * diff;
      }
    }
    if (dpred) dpred[i] = g * weight * inv;
  }
  s = warp_sum(s);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < 8; ++i) tot += part[i];
    atomicAdd(ws + 1, tot);
  }
}

__global__ void loss_finalize_kernel(const float* __restrict__ ws, float* __restrict__ loss_out,
                                     float* __restrict__ total_out, int inner, float weight) {
  const float l = ws[1] / (ws[0] * (float)inner);
  *loss_out = l;
  if (total_out) *total_out += weight * l;
}

// out[0] += sum x^2
__global__ void __launch_bounds__(256) sumsq_kernel(const float4* __restrict__ x, float* __restrict__ out, size_t n4) {
  float s = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = x[i];
    s += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  s = warp_sum(s);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < 8; ++i) tot += part[i];
    atomicAdd(out, tot);
  }
}

// x *= *s  (chain-rule factor of an upstream scalar gradient that lives on the device)
__global__ void scale_by_kernel(float* __restrict__ x, const float* __restrict__ s, size_t n) {
  const float f = *s;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] *= f;
}

struct AdamWParams {
  float lr, beta1, beta2, eps, weight_decay;
  float bc1, bc2_sqrt;   // 1 - beta1^step, sqrt(1 - beta2^step)
  float grad_scale;      // 1 / world_size (gradient averaging after the sum all-reduce)
  float max_norm;        // <= 0: no clipping
  int zero_grad;
};

__global__ void __launch_bounds__(256)
adamw_kernel(float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
             const float* __restrict__ gnorm_sq, size_t n4, AdamWParams a) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float gs = a.grad_scale;
  if (a.max_norm > 0.f && gnorm_sq) {  // torch.nn.utils.clip_grad_norm_: coef = min(1, max_norm / (norm + 1e-6))
    const float norm = sqrtf(*gnorm_sq) * a.grad_scale;
    gs *= fminf(1.f, a.max_norm / (norm + 1e-6f));
  }
  float4 pv = p[i], gv = g[i], mv = m[i], vv = v[i];
  float* pp = reinterpret_cast<float*>(&pv);
  float* gp = reinterpret_cast<float*>(&gv);
  float* mp = reinterpret_cast<float*>(&mv);
  float* vp = reinterpret_cast<float*>(&vv);
  const float decay = 1.f - a.lr * a.weight_decay;
  const float step_size = a.lr / a.bc1;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float gj = gp[j] * gs;
    pp[j] *= decay;
    mp[j] = mp[j] + (gj - mp[j]) * (1.f - a.beta1);               // exp_avg.lerp_(grad, 1 - beta1)
    vp[j] = vp[j] * a.beta2 + (1.f - a.beta2) * gj * gj;          // exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
    const float denom = sqrtf(vp[j]) / a.bc2_sqrt + a.eps;
    pp[j] -= step_size * (mp[j] / denom);
  }
  p[i] = pv;
  m[i] = mv;
  v[i] = vv;
  if (a.zero_grad) g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

}  // namespace lfs2

using namespace lfs2;

extern "C" {

int lfs2_masked_loss(const float* pred, const float* target, const int64_t* target_i64, const uint8_t* pad_mask,
                     int rows, int inner, int kind, float weight, float* loss_out, float* total_out, float* dpred,
                     float* workspace, void* stream) {
  LFS2_REQUIRE(pred && (target || target_i64) && loss_out && workspace, LFS2_ERR_INVALID_ARG, "masked_loss: null pointer");
  LFS2_REQUIRE(rows > 0 && inner > 0, LFS2_ERR_INVALID_ARG, "masked_loss: bad shape");
  LFS2_REQUIRE(kind == 0 || kind == 1, LFS2_ERR_UNSUPPORTED, "masked_loss: kind must be 0 (l1) or 1 (mse)");
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(workspace, 0, 2 * sizeof(float), s) != cudaSuccess) {
    set_error("masked_loss: memset failed");
    return LFS2_ERR_CUDA;
  }
  int cblocks = ceil_div(rows, 256);
  if (cblocks > num_sms()) cblocks = num_sms();
  mask_count_kernel<<<cblocks, 256, 0, s>>>(pad_mask, workspace, rows);
  const size_t total = (size_t)rows * inner;
  int blocks = ceil_div((long long)total, 256 * 4);
  if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
  if (blocks < 1) blocks = 1;
  masked_loss_kernel<<<blocks, 256, 0, s>>>(pred, target, target_i64, pad_mask, dpred, workspace, total, inner, kind,
                                            weight);
  loss_finalize_kernel<<<1, 1, 0, s>>>(workspace, loss_out, total_out, inner, weight);
  LFS2_CHECK_LAUNCH("masked_loss");
  return LFS2_OK;
}

int lfs2_scale_by(float* x, const float* scalar, long long n, void* stream) {
  LFS2_REQUIRE(x && scalar, LFS2_ERR_INVALID_ARG, "scale_by: null pointer");
  if (n <= 0) return LFS2_OK;
  int blocks = ceil_div(n, 256 * 4);
  if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
  scale_by_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, scalar, (size_t)n);
  LFS2_CHECK_LAUNCH("scale_by");
  return LFS2_OK;
}

int lfs2_sumsq(const float* x, float* out, long long n, void* stream) {
  LFS2_REQUIRE(x && out, LFS2_ERR_INVALID_ARG, "sumsq: null pointer");
  if (n == 0) return LFS2_OK;
  LFS2_REQUIRE(n > 0 && n % 4 == 0 && aligned16(x), LFS2_ERR_UNSUPPORTED, "sumsq: n %% 4 == 0 and 16-byte alignment");
  size_t n4 = (size_t)n / 4;
  int blocks = ceil_div((long long)n4, 256 * 4);
  if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
  sumsq_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)x, out, n4);
  LFS2_CHECK_LAUNCH("sumsq");
  return LFS2_OK;
}

int lfs2_adamw_step(float* p, float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                    float weight_decay, int step, float grad_scale, float max_norm, const float* gnorm_sq,
                    int zero_grad, void* stream) {
  LFS2_REQUIRE(p && g && m && v, LFS2_ERR_INVALID_ARG, "adamw_step: null pointer");
  if (n == 0) return LFS2_OK;
  LFS2_REQUIRE(n > 0 && n % 4 == 0, LFS2_ERR_UNSUPPORTED, "adamw_step: n must be a positive multiple of 4 (pad the flat buffer)");
  LFS2_REQUIRE(step >= 1, LFS2_ERR_INVALID_ARG, "adamw_step: step counts from 1");
  LFS2_REQUIRE(aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v), LFS2_ERR_INVALID_ARG,
               "adamw_step: buffers must be 16-byte aligned");
  LFS2_REQUIRE(max_norm <= 0.f || gnorm_sq, LFS2_ERR_INVALID_ARG, "adamw_step: clipping needs the squared gradient norm");
  AdamWParams a;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
  a.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  a.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  a.grad_scale = grad_scale; a.max_norm = max_norm; a.zero_grad = zero_grad;
  size_t n4 = (size_t)n / 4;
  adamw_kernel<<<ceil_div((long long)n4, 256), 256, 0, (cudaStream_t)stream>>>((float4*)p, (float4*)g, (float4*)m,
                                                                             (float4*)v, gnorm_sq, n4, a);
  LFS2_CHECK_LAUNCH("adamw_step");
  return LFS2_OK;
}

}  // extern "C"
