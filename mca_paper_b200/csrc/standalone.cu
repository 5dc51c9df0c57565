// Kernels behind the FREE-STANDING calls of the reference surface: the functional
// contrastive_loss_with_temperature(...) -> ContrastiveLossOutput with its logits matrices and cross_entropy_kwargs
// (utils/contrastive_loss_with_temperature.py:40-108), and Attention.forward(..., return_attn=True) (model.py:96-103).
// Inside MCA.forward none of these run: the fused all-pairs loss (loss.cu) and the block-sparse attention never
// materialise logits or probabilities.  Small, latency-bound kernels; one CTA per output row.
#include <math_constants.h>

#include "mca_b200.h"
#include "ptx.cuh"
#include "runtime.h"

namespace mca {

constexpr int SA_THREADS = 128;

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < nw; ++w) t += red[w];  // fixed order: bit-reproducible
  return t;
}

__device__ __forceinline__ float block_max(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = -CUDART_INF_F;
  for (int w = 0; w < nw; ++w) t = fmaxf(t, red[w]);
  return t;
}

// logits[i, j] = exp(*logit_scale) * <a_i, b_j>   (contrastive_loss_with_temperature.py:71,82-88)
__global__ void __launch_bounds__(SA_THREADS)
scaled_logits_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ logit_scale,
                     int n_b, int d, float* __restrict__ logits) {
  extern __shared__ float s_a[];
  const int i = blockIdx.x;
  for (int c = threadIdx.x; c < d; c += blockDim.x) s_a[c] = a[static_cast<long long>(i) * d + c];
  __syncthreads();
  const float T = expf(*logit_scale);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int j = warp; j < n_b; j += nw) {
    const float* bj = b + static_cast<long long>(j) * d;
    float acc = 0.f;
    for (int c = lane; c < d; c += 32) acc = fmaf(s_a[c], bj[c], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) logits[static_cast<long long>(i) * n_b + j] = acc * T;
  }
}

// F.cross_entropy(logits, labels, label_smoothing=eps, reduction='none') per row, plus the row log-sum-exp:
//   loss_i = (1 - eps) * (lse_i - z_i[y_i]) + eps * (lse_i - mean_j z_ij)
__global__ void __launch_bounds__(SA_THREADS)
cross_entropy_fwd_kernel(const float* __restrict__ logits, long long ld, const long long* __restrict__ labels, int cols,
                         float eps, float* __restrict__ row_loss, float* __restrict__ row_lse) {
  __shared__ float red[SA_THREADS / 32];
  const int r = blockIdx.x;
  const float* z = logits + r * ld;
  float mx = -CUDART_INF_F;
  for (int j = threadIdx.x; j < cols; j += blockDim.x) mx = fmaxf(mx, z[j]);
  mx = block_max(mx, red);
  float se = 0.f, sz = 0.f;
  for (int j = threadIdx.x; j < cols; j += blockDim.x) {
    se += expf(z[j] - mx);
    sz += z[j];
  }
  se = block_sum(se, red);
  sz = block_sum(sz, red);
  if (threadIdx.x == 0) {
    const float lse = mx + logf(se);
    const long long y = labels[r];
    const float nll = lse - z[y];
    row_lse[r] = lse;
    row_loss[r] = (1.f - eps) * nll + eps * (lse - sz / static_cast<float>(cols));
  }
}

// dz_ij = g_i * (softmax_ij - (1 - eps) * [j == y_i] - eps / cols); out = dz * out_scale (out_scale = exp(*logit_scale)
// when given, so that the operand gradients are plain products with dlogits), dscale += sum_ij dz_ij * z_ij
// (d logits / d logit_scale = logits).
__global__ void __launch_bounds__(SA_THREADS)
cross_entropy_bwd_kernel(const float* __restrict__ logits, long long ld, const long long* __restrict__ labels, int cols,
                         float eps, const float* __restrict__ row_lse, const float* __restrict__ g_row,
                         const float* __restrict__ logit_scale, float* __restrict__ dlogits, float* __restrict__ dscale) {
  __shared__ float red[SA_THREADS / 32];
  const int r = blockIdx.x;
  const float* z = logits + r * ld;
  const float lse = row_lse[r], g = g_row[r];
  const long long y = labels[r];
  const float T = logit_scale != nullptr ? expf(*logit_scale) : 1.f;
  const float sm = eps / static_cast<float>(cols);
  float ds = 0.f;
  for (int j = threadIdx.x; j < cols; j += blockDim.x) {
    const float p = expf(z[j] - lse);
    const float dz = g * (p - (j == y ? 1.f - eps : 0.f) - sm);
    ds = fmaf(dz, z[j], ds);
    dlogits[static_cast<long long>(r) * cols + j] = dz * T;
  }
  if (dscale != nullptr) {
    ds = block_sum(ds, red);
    if (threadIdx.x == 0) atomicAdd(dscale, ds);
  }
}

// Attention probabilities for return_attn=True (model.py:96,102-103), recomputed from the saved row log-sum-exp:
// attn[b,h,i,j] = exp(q_i . k_j - lse_i) for statically allowed, live keys, 0 elsewhere; a row with no live allowed key
// (lse = +inf) is uniform over all N keys (the -finfo.max fill, quirk Q4).  qkv = (Q*scale | K | V) bf16 [B*N, 3*H*64].
__global__ void __launch_bounds__(SA_THREADS)
attn_probs_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ lse, const uint32_t* __restrict__ rowbits,
                  const uint8_t* __restrict__ keygrp, const uint8_t* __restrict__ padding, int N, int H,
                  float* __restrict__ probs) {
  __shared__ float s_q[64];
  const int i = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int ld = 3 * H * 64;
  const long long row = static_cast<long long>(b) * N + i;
  if (threadIdx.x < 64) s_q[threadIdx.x] = __bfloat162float(qkv[row * ld + h * 64 + threadIdx.x]);
  __syncthreads();
  const float l = lse[(static_cast<long long>(b) * H + h) * N + i];
  const uint32_t rb = rowbits[i];
  float* out = probs + ((static_cast<long long>(b) * H + h) * N + i) * N;
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    float p;
    if (isinf(l)) {
      p = 1.0f / static_cast<float>(N);
    } else if (((rb >> keygrp[j]) & 1u) == 0 || padding[static_cast<long long>(b) * N + j]) {
      p = 0.f;
    } else {
      const uint4* kp = reinterpret_cast<const uint4*>(qkv + (static_cast<long long>(b) * N + j) * ld + H * 64 + h * 64);
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 w = kp[c];
        acc = fmaf(s_q[8 * c + 0], bf16_lo(w.x), acc), acc = fmaf(s_q[8 * c + 1], bf16_hi(w.x), acc);
        acc = fmaf(s_q[8 * c + 2], bf16_lo(w.y), acc), acc = fmaf(s_q[8 * c + 3], bf16_hi(w.y), acc);
        acc = fmaf(s_q[8 * c + 4], bf16_lo(w.z), acc), acc = fmaf(s_q[8 * c + 5], bf16_hi(w.z), acc);
        acc = fmaf(s_q[8 * c + 6], bf16_lo(w.w), acc), acc = fmaf(s_q[8 * c + 7], bf16_hi(w.w), acc);
      }
      p = expf(acc - l);
    }
    out[j] = p;
  }
}

}  // namespace mca

using namespace mca;

extern "C" int mca_scaled_logits_f32(const float* a, const float* b_all, const float* logit_scale, int n_a, int n_b, int d,
                                     float* logits, void* stream) {
  if (n_a <= 0 || n_b <= 0 || d <= 0 || d > 8192) return MCA_ERR_SHAPE;
  scaled_logits_kernel<<<n_a, SA_THREADS, d * sizeof(float), reinterpret_cast<cudaStream_t>(stream)>>>(a, b_all, logit_scale,
                                                                                                      n_b, d, logits);
  return check_launch();
}

extern "C" int mca_cross_entropy_fwd(const float* logits, long long ld, const long long* labels, int rows, int cols,
                                     float label_smoothing, float* row_loss, float* row_lse, void* stream) {
  if (rows < 0 || cols <= 0 || ld < cols) return MCA_ERR_SHAPE;
  if (rows == 0) return MCA_OK;
  cross_entropy_fwd_kernel<<<rows, SA_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(logits, ld, labels, cols,
                                                                                           label_smoothing, row_loss, row_lse);
  return check_launch();
}

extern "C" int mca_cross_entropy_bwd(const float* logits, long long ld, const long long* labels, int rows, int cols,
                                     float label_smoothing, const float* row_lse, const float* g_row,
                                     const float* logit_scale, float* dlogits, float* dscale, void* stream) {
  if (rows < 0 || cols <= 0 || ld < cols) return MCA_ERR_SHAPE;
  if (rows == 0) return MCA_OK;
  cross_entropy_bwd_kernel<<<rows, SA_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      logits, ld, labels, cols, label_smoothing, row_lse, g_row, logit_scale, dlogits, dscale);
  return check_launch();
}

extern "C" int mca_attn_probs(const void* qkv, const float* lse, const uint32_t* rowbits, const uint8_t* keygrp,
                              const uint8_t* padding, int B, int N, int H, float* probs, void* stream) {
  if (B <= 0 || N <= 0 || H <= 0 || B > 65535 || H > 65535) return MCA_ERR_SHAPE;
  attn_probs_kernel<<<dim3(N, H, B), SA_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv), lse, rowbits, keygrp, padding, N, H, probs);
  return check_launch();
}
