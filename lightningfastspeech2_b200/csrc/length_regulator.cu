// LengthRegulator (reference litfass/fastspeech2/model.py:349-370) as two HBM-bound kernels:
//   scan    : per-utterance inclusive prefix sum of the durations (warp-shuffle scan, int64)
//   scatter : every output frame finds its source phone by binary search in the prefix sums
//             (idx = #{p : cum[p] <= t}) and copies the row as raw 16-byte words
// Index math is integer-only and rows are copied bit-for-bit, so the result is bit-exact
// against the reference's repeat_interleave / pad_sequence loop for any element type.
//
// Algorithmic traffic: B*Tp*(row_bytes + 8) read + B*L*(row_bytes + 1) written.
#include "common.cuh"

namespace lfs2 {

constexpr int kScanThreads = 256;

template <typename DurT>
__global__ void lr_scan_kernel(const DurT* __restrict__ dur, int64_t* __restrict__ cum,
                               int64_t* __restrict__ lengths, int64_t* __restrict__ max_len, int tp) {
  int b = blockIdx.x;
  const DurT* d = dur + (size_t)b * tp;
  int64_t* c = cum + (size_t)b * tp;
  int per = (tp + kScanThreads - 1) / kScanThreads;
  int lo = threadIdx.x * per;
  int hi = min(lo + per, tp);
  long long local = 0;
  for (int i = lo; i < hi; ++i) {
    long long v = (long long)d[i];
    local += v > 0 ? v : 0;  // the reference rejects negative repeats; they count as 0 here
  }
  // inclusive warp scan of the per-thread sums
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  long long incl = local;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    long long n = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += n;
  }
  __shared__ long long warp_tot[kScanThreads / 32];
  if (lane == 31) warp_tot[w] = incl;
  __syncthreads();
  long long base = 0;
  for (int i = 0; i < w; ++i) base += warp_tot[i];
  long long run = base + incl - local;  // exclusive prefix of this thread's chunk
  for (int i = lo; i < hi; ++i) {
    long long v = (long long)d[i];
    run += v > 0 ? v : 0;
    c[i] = run;
  }
  if (threadIdx.x == kScanThreads - 1) {
    long long total = base + incl;
    lengths[b] = total;
    atomicMax((long long*)max_len, total);
  }
}

constexpr int kLrFrames = 32;    // output frames per CTA
constexpr int kLrThreads = 256;
constexpr int kLrUnroll = 4;

__global__ void __launch_bounds__(kLrThreads)
lr_scatter_kernel(const int4* __restrict__ x, const int64_t* __restrict__ cum, const int64_t* __restrict__ lengths,
                  int4* __restrict__ out, uint8_t* __restrict__ mask, int tp, int l, int cap,
                  int cpr /*16B chunks per row*/) {
  int b = blockIdx.y;
  int t0 = blockIdx.x * kLrFrames;
  __shared__ int s_idx[kLrFrames];
  long long len = lengths[b];
  if (threadIdx.x < kLrFrames) {
    int t = t0 + threadIdx.x;
    int idx = -1;
    if (t < l && t < cap && (long long)t < len) {  // frames at or beyond the reference's cut (cap) are PAD
      const int64_t* c = cum + (size_t)b * tp;
      int lo = 0, hi = tp;  // first p with cum[p] > t  == number of p with cum[p] <= t
      while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (c[mid] <= (long long)t) lo = mid + 1;
        else hi = mid;
      }
      idx = lo;
    }
    s_idx[threadIdx.x] = idx;
    if (t < l) mask[(size_t)b * l + t] = (idx < 0);
  }
  __syncthreads();
  int nframes = min(kLrFrames, l - t0);
  int total = nframes * cpr;
  const int4* xb = x + (size_t)b * tp * cpr;
  int4* ob = out + ((size_t)b * l + t0) * cpr;
  for (int q0 = threadIdx.x; q0 < total; q0 += kLrThreads * kLrUnroll) {
    int4 v[kLrUnroll];
#pragma unroll
    for (int u = 0; u < kLrUnroll; ++u) {
      int q = q0 + u * kLrThreads;
      v[u] = make_int4(0, 0, 0, 0);
      if (q < total) {
        int f = q / cpr, col = q - f * cpr;
        int idx = s_idx[f];
        if (idx >= 0) v[u] = ld_stream16(xb + (size_t)idx * cpr + col);
      }
    }
#pragma unroll
    for (int u = 0; u < kLrUnroll; ++u) {
      int q = q0 + u * kLrThreads;
      if (q < total) st_stream16(ob + q, v[u]);
    }
  }
}

}  // namespace lfs2

using namespace lfs2;

// Ragged read-back: the valid rows of a padded (batch, l, width) fp32 tensor packed back to back in utterance order
// (what the reference's caller keeps of a synthesis batch: generator.py:164-170 cuts every mel at ~tgt_mask).  Utterance
// b owns rows [off_b, off_b + n_b) of `out`, n_b = min(lengths[b], l), off_b = n_0 + ... + n_{b-1}.
__global__ void pack_valid_rows_kernel(const float4* __restrict__ x, const int64_t* __restrict__ lengths,
                                       float4* __restrict__ out, int l, int w4, int rows_per_block) {
  const int b = blockIdx.y;
  long long off = 0;
  for (int i = 0; i < b; ++i) off += min((long long)lengths[i], (long long)l);   // batch <= a few hundred: trivial
  const int n = (int)min((long long)lengths[b], (long long)l);
  const int r0 = blockIdx.x * rows_per_block, r1 = min(r0 + rows_per_block, n);
  const size_t src = ((size_t)b * l + r0) * w4, dst = ((size_t)off + r0) * w4;
  const int total = (r1 - r0) * w4;
  for (int i = threadIdx.x; i < total; i += blockDim.x) out[dst + i] = x[src + i];
}

extern "C" {

int lfs2_pack_valid_rows(const float* x, const int64_t* lengths, float* out, int batch, int l, int width, void* stream) {
  LFS2_REQUIRE(x && lengths && out, LFS2_ERR_INVALID_ARG, "pack_valid_rows: null pointer");
  if (batch == 0 || l == 0) return LFS2_OK;
  LFS2_REQUIRE(batch > 0 && batch <= 65535 && l > 0 && width > 0 && width % 4 == 0, LFS2_ERR_UNSUPPORTED,
               "pack_valid_rows: need 0 < batch <= 65535 and width %% 4 == 0");
  LFS2_REQUIRE(aligned16(x) && aligned16(out), LFS2_ERR_INVALID_ARG, "pack_valid_rows: pointers must be 16-byte aligned");
  const int rows_per_block = 64;
  dim3 grid(ceil_div(l, rows_per_block), batch);
  pack_valid_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)x, lengths, (float4*)out, l, width / 4,
                                                                 rows_per_block);
  LFS2_CHECK_LAUNCH("pack_valid_rows");
  return LFS2_OK;
}

int lfs2_length_regulate_scan(const void* dur, int dur_is_i64, int64_t* cum, int64_t* lengths, int64_t* max_len,
                              int batch, int tp, void* stream) {
  LFS2_REQUIRE(dur && cum && lengths && max_len, LFS2_ERR_INVALID_ARG, "length_regulate_scan: null pointer");
  LFS2_REQUIRE(batch > 0 && tp > 0, LFS2_ERR_INVALID_ARG, "length_regulate_scan: bad shape");
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(max_len, 0, sizeof(int64_t), s) != cudaSuccess) {
    set_error("length_regulate_scan: memset failed");
    return LFS2_ERR_CUDA;
  }
  if (dur_is_i64)
    lr_scan_kernel<long long><<<batch, kScanThreads, 0, s>>>((const long long*)dur, cum, lengths, max_len, tp);
  else
    lr_scan_kernel<int><<<batch, kScanThreads, 0, s>>>((const int*)dur, cum, lengths, max_len, tp);
  LFS2_CHECK_LAUNCH("length_regulate_scan");
  return LFS2_OK;
}

int lfs2_length_regulate_scatter(const void* x, const int64_t* cum, const int64_t* lengths, void* out,
                                 uint8_t* mask, int batch, int tp, int l, int row_bytes, void* stream) {
  return lfs2_length_regulate_scatter_ex(x, cum, lengths, out, mask, batch, tp, l, l, row_bytes, stream);
}

int lfs2_length_regulate_scatter_ex(const void* x, const int64_t* cum, const int64_t* lengths, void* out,
                                    uint8_t* mask, int batch, int tp, int l, int cap, int row_bytes, void* stream) {
  LFS2_REQUIRE(cum && lengths, LFS2_ERR_INVALID_ARG, "length_regulate_scatter: null pointer");
  LFS2_REQUIRE(batch > 0 && tp > 0 && l >= 0, LFS2_ERR_INVALID_ARG, "length_regulate_scatter: bad shape");
  if (l == 0) return LFS2_OK;
  LFS2_REQUIRE(x && out && mask, LFS2_ERR_INVALID_ARG, "length_regulate_scatter: null pointer");
  LFS2_REQUIRE(row_bytes > 0 && row_bytes % 16 == 0, LFS2_ERR_UNSUPPORTED,
               "length_regulate_scatter: row_bytes=%d must be a multiple of 16", row_bytes);
  LFS2_REQUIRE(aligned16(x) && aligned16(out), LFS2_ERR_INVALID_ARG,
               "length_regulate_scatter: x/out must be 16-byte aligned");
  LFS2_REQUIRE(batch <= 65535, LFS2_ERR_UNSUPPORTED, "length_regulate_scatter: batch > 65535");
  dim3 grid(ceil_div(l, kLrFrames), batch);
  lr_scatter_kernel<<<grid, kLrThreads, 0, (cudaStream_t)stream>>>((const int4*)x, cum, lengths, (int4*)out, mask,
                                                                  tp, l, cap, row_bytes / 16);
  LFS2_CHECK_LAUNCH("length_regulate_scatter");
  return LFS2_OK;
}

}  // extern "C"
