// Fused "grad-norm -> clip -> AdamW" over the flat fp32 parameter / gradient / moment buffers
// (reference: train_accel_gpu.py:80 AdamW(lr) with torch defaults, :116-118 clip_grad_norm_(clip) + step,
// :81-86,119 cosine schedule with warm-up).  Bandwidth-bound: pass 1 reads 4 B/param, pass 2 reads p,g,m,v and
// writes p,m,v = 28 B/param.  The step counter and learning-rate schedule live on the device so the whole
// training step replays from a CUDA graph without host-side scalars.
#include "mca_b200.h"
#include "ptx.cuh"
#include "runtime.h"

namespace mca {

__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ g, long long n, double* __restrict__ out) {
  __shared__ float red[8];
  float acc = 0.f;
  const long long n4 = n / 4;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 q = g4[i];
    acc += q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float v = g[n4 * 4 + threadIdx.x];
    acc += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(out, static_cast<double>(t));
  }
}

// state: [0] = step (as double, incremented here by block 0 AFTER use via the `commit` kernel), see below
__device__ __forceinline__ float scheduled_lr(const mca_adamw_cfg& c, long long step /*1-based*/) {
  if (c.lr_mode == 0) return c.lr;
  // transformers.get_cosine_schedule_with_warmup evaluated at (step-1): scheduler.step() follows optimizer.step()
  const double cur = static_cast<double>(step - 1);
  if (cur < c.warmup_steps) return c.lr * static_cast<float>(cur / fmax(1.0, static_cast<double>(c.warmup_steps)));
  const double prog = (cur - c.warmup_steps) / fmax(1.0, static_cast<double>(c.total_steps - c.warmup_steps));
  return c.lr * static_cast<float>(fmax(0.0, 0.5 * (1.0 + cos(3.14159265358979323846 * prog))));
}

__global__ void __launch_bounds__(256)
clip_adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                  long long n, const double* __restrict__ sumsq, const long long* __restrict__ step_ptr,
                  float grad_scale, mca_adamw_cfg c) {
  __shared__ float s_coef, s_lr, s_bc1, s_bc2s;
  if (threadIdx.x == 0) {
    const long long step = *step_ptr + 1;
    const float total = static_cast<float>(sqrt(*sumsq)) * grad_scale;
    float coef = 1.0f;
    if (c.max_norm > 0.f) coef = fminf(1.0f, c.max_norm / (total + 1e-6f));
    s_coef = coef * grad_scale;
    s_lr = scheduled_lr(c, step);
    s_bc1 = static_cast<float>(1.0 - pow(static_cast<double>(c.beta1), static_cast<double>(step)));
    s_bc2s = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(c.beta2), static_cast<double>(step))));
  }
  __syncthreads();
  const float coef = s_coef, lr = s_lr, bc1 = s_bc1, bc2s = s_bc2s;
  const float decay = 1.0f - lr * c.weight_decay;
  const float step_size = lr / bc1;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float gi = g[i] * coef;
    const float mi = c.beta1 * m[i] + (1.0f - c.beta1) * gi;
    const float vi = c.beta2 * v[i] + (1.0f - c.beta2) * gi * gi;
    const float denom = sqrtf(vi) / bc2s + c.eps;
    p[i] = p[i] * decay - step_size * (mi / denom);
    m[i] = mi;
    v[i] = vi;
  }
}

__global__ void bump_step_kernel(long long* step_ptr, double* sumsq, float* total_norm_out, float grad_scale) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    *step_ptr += 1;
    if (total_norm_out != nullptr) *total_norm_out = static_cast<float>(sqrt(*sumsq)) * grad_scale;
  }
}

}  // namespace mca

using namespace mca;

extern "C" int mca_clip_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                                   long long n, double* sumsq_scratch, long long* step_dev, float* total_norm_out,
                                   float grad_scale, const mca_adamw_cfg* cfg_host, void* stream_) {
  if (n <= 0) return MCA_ERR_SHAPE;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (cudaMemsetAsync(sumsq_scratch, 0, sizeof(double), stream) != cudaSuccess) return MCA_ERR_CUDA;
  const int blocks = num_sms() * 4;
  sumsq_kernel<<<blocks, 256, 0, stream>>>(grads, n, sumsq_scratch);
  clip_adamw_kernel<<<blocks, 256, 0, stream>>>(params, grads, exp_avg, exp_avg_sq, n, sumsq_scratch, step_dev,
                                               grad_scale, *cfg_host);
  bump_step_kernel<<<1, 32, 0, stream>>>(step_dev, sumsq_scratch, total_norm_out, grad_scale);
  return check_launch();
}
