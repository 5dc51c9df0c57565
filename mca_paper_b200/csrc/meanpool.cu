// Mean pooling of the EAO baseline: MeanTokenProjectionPool(token_types=None, projection=False), model.py:235-280 as
// EAO.single_pass calls it (model.py:553-556,562-563) — per sample and pass, the mean of the final-normed tokens whose
// key-padding bit is clear, zeros when the pass has no live token (model.py:270-271); the projection is an Identity.
// All passes of a sample lie back to back in one packed sequence (plan.EAOPlan), so one launch pools every pass.
// Bandwidth-bound: the forward reads the bf16 tokens once (1 KB per token), the backward writes the fp32 gradient once.
#include "mca_b200.h"
#include "ptx.cuh"
#include "runtime.h"

namespace mca {

constexpr int MP_SPLIT = 8;  // token slices per (pass, sample): enough blocks to fill the GPU at B = 8

// grid (MP_SPLIT, R, B), 256 threads = 512 columns as bf16 pairs: slice sums and live counts into scratch
__global__ void __launch_bounds__(256)
mean_pool_partial_kernel(const __nv_bfloat16* __restrict__ x, const uint8_t* __restrict__ padding,
                         const int* __restrict__ pass_start, int N, int R, float* __restrict__ part,
                         float* __restrict__ pcnt) {
  const int r = blockIdx.y, b = blockIdx.z, sl = blockIdx.x;
  const int s0 = pass_start[r], s1 = pass_start[r + 1];
  const int per = (s1 - s0 + MP_SPLIT - 1) / MP_SPLIT;
  const int t0 = min(s1, s0 + sl * per), t1 = min(s1, t0 + per);
  const uint8_t* pd = padding + static_cast<long long>(b) * N;
  const uint32_t* xb = reinterpret_cast<const uint32_t*>(x + static_cast<long long>(b) * N * 512) + threadIdx.x;
  float a0 = 0.f, a1 = 0.f;
  int live = 0;
  for (int t = t0; t < t1; ++t) {
    if (pd[t]) continue;
    const uint32_t v = xb[static_cast<long long>(t) * 256];
    a0 += __uint_as_float(v << 16);
    a1 += __uint_as_float(v & 0xffff0000u);
    ++live;
  }
  const long long slot = (static_cast<long long>(b) * R + r) * MP_SPLIT + sl;
  *reinterpret_cast<float2*>(part + slot * 512 + threadIdx.x * 2) = make_float2(a0, a1);
  if (threadIdx.x == 0) pcnt[slot] = static_cast<float>(live);
}

// grid (R, B): pooled[b, r, :] = sum of the slices / live count (zeros when the pass has no live token); cnt[b, r] kept
// for the backward
__global__ void __launch_bounds__(256)
mean_pool_final_kernel(const float* __restrict__ part, const float* __restrict__ pcnt, int R, float* __restrict__ pooled,
                       float* __restrict__ cnt) {
  const long long br = static_cast<long long>(blockIdx.y) * R + blockIdx.x;
  float c = 0.f, a0 = 0.f, a1 = 0.f;
#pragma unroll
  for (int sl = 0; sl < MP_SPLIT; ++sl) {
    c += pcnt[br * MP_SPLIT + sl];
    const float2 v = *reinterpret_cast<const float2*>(part + (br * MP_SPLIT + sl) * 512 + threadIdx.x * 2);
    a0 += v.x, a1 += v.y;
  }
  const float inv = c > 0.f ? 1.f / c : 0.f;
  *reinterpret_cast<float2*>(pooled + br * 512 + threadIdx.x * 2) = make_float2(a0 * inv, a1 * inv);
  if (threadIdx.x == 0) cnt[br] = c;
}

// dx[b, t, :] = padded ? 0 : dpooled[b, pass(t), :] / cnt[b, pass(t)]; one warp per token
__global__ void __launch_bounds__(256)
mean_pool_bwd_kernel(const float* __restrict__ dpooled, const uint8_t* __restrict__ padding,
                     const int* __restrict__ tok_pass, const float* __restrict__ cnt, int B, int N, int R,
                     float* __restrict__ dx) {
  const int lane = threadIdx.x & 31;
  const long long rows = static_cast<long long>(B) * N;
  for (long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5); row < rows;
       row += static_cast<long long>(gridDim.x) * 8) {
    const int b = static_cast<int>(row / N), t = static_cast<int>(row % N);
    float4* o = reinterpret_cast<float4*>(dx + row * 512);
    if (padding[row]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) o[lane + 32 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
      continue;
    }
    const int r = tok_pass[t];
    const float inv = 1.f / cnt[b * R + r];   // >= 1 live token: this one
    const float4* g = reinterpret_cast<const float4*>(dpooled + (static_cast<long long>(b) * R + r) * 512);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 v = g[lane + 32 * i];
      v.x *= inv, v.y *= inv, v.z *= inv, v.w *= inv;
      o[lane + 32 * i] = v;
    }
  }
}

}  // namespace mca

using namespace mca;

extern "C" int mca_mean_pool_scratch_floats(int B, int R) { return B * R * MP_SPLIT * (512 + 1); }

extern "C" int mca_mean_pool_fwd(const void* x_bf16, const uint8_t* padding, const int* pass_start, int B, int N, int R,
                                 int d, float* pooled, float* cnt, float* scratch, void* stream_) {
  if (B <= 0 || N <= 0 || R <= 0 || d != 512) return MCA_ERR_SHAPE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  float* part = scratch;
  float* pcnt = scratch + static_cast<long long>(B) * R * MP_SPLIT * 512;
  mean_pool_partial_kernel<<<dim3(MP_SPLIT, R, B), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x_bf16), padding,
                                                                pass_start, N, R, part, pcnt);
  mean_pool_final_kernel<<<dim3(R, B), 256, 0, st>>>(part, pcnt, R, pooled, cnt);
  return check_launch();
}

extern "C" int mca_mean_pool_bwd(const float* dpooled, const uint8_t* padding, const int* tok_pass, const float* cnt, int B,
                                 int N, int R, int d, float* dx, void* stream_) {
  if (B <= 0 || N <= 0 || R <= 0 || d != 512) return MCA_ERR_SHAPE;
  const long long rows = static_cast<long long>(B) * N;
  const long long want = (rows + 7) / 8;
  const int blocks = static_cast<int>(want < 148LL * 16 ? want : 148LL * 16);
  mean_pool_bwd_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(dpooled, padding, tok_pass, cnt, B, N, R, dx);
  return check_launch();
}
