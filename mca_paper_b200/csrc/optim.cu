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
  // transformers.get_scheduler(name) evaluated at (step-1): scheduler.step() follows optimizer.step().
  // 1 = "cosine", 2 = "constant_with_warmup", 3 = "linear" (all with linear warm-up over warmup_steps)
  const double cur = static_cast<double>(step - 1) * static_cast<double>(c.sched_stride > 1 ? c.sched_stride : 1);
  if (cur < c.warmup_steps) return c.lr * static_cast<float>(cur / fmax(1.0, static_cast<double>(c.warmup_steps)));
  if (c.lr_mode == 2) return c.lr;
  if (c.lr_mode == 3)
    return c.lr * static_cast<float>(fmax(0.0, (static_cast<double>(c.total_steps) - cur) /
                                                   fmax(1.0, static_cast<double>(c.total_steps - c.warmup_steps))));
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
  // float4 lanes: four independent 16-byte loads in flight per thread and iteration (the flat buffers are 256-byte
  // aligned and their length is a multiple of 64 elements; a scalar tail keeps the kernel general)
  const long long n4 = n / 4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 gq = reinterpret_cast<const float4*>(g)[i];
    float4 mq = reinterpret_cast<float4*>(m)[i];
    float4 vq = reinterpret_cast<float4*>(v)[i];
    float4 pq = reinterpret_cast<float4*>(p)[i];
    const float gs[4] = {gq.x * coef, gq.y * coef, gq.z * coef, gq.w * coef};
    float* mm = reinterpret_cast<float*>(&mq);
    float* vv = reinterpret_cast<float*>(&vq);
    float* pp = reinterpret_cast<float*>(&pq);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      mm[k] = c.beta1 * mm[k] + (1.0f - c.beta1) * gs[k];
      vv[k] = c.beta2 * vv[k] + (1.0f - c.beta2) * gs[k] * gs[k];
      const float denom = sqrtf(vv[k]) / bc2s + c.eps;
      pp[k] = pp[k] * decay - step_size * (mm[k] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pq;
    reinterpret_cast<float4*>(m)[i] = mq;
    reinterpret_cast<float4*>(v)[i] = vq;
  }
  for (long long i = n4 * 4 + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
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

// ---- data-parallel optimiser over peer memory (one NVLink/NVSwitch domain), ZeRO-1 style: rank r owns the elements
// [off, off + n) of the flat buffers.  Instead of an NCCL all-reduce followed by a full AdamW on every rank:
//   dp_reduce_shard : own shard of the gradient = sum over ranks, PULLED from the peers' gradient buffers (coalesced
//                     float4 NVLink reads), written back into the local gradient buffer; partial sum of squares
//   (barrier carrying the partial sums of squares to every rank)
//   dp_adamw_shard  : global norm from the G partials, clip + AdamW on the own shard only (1/G of the optimiser
//                     traffic), the new parameters PUSHED into every rank's parameter buffer (the all-gather, as posted
//                     NVLink stores from inside the update kernel)
// Every rank ends the step with bit-identical parameters.
__global__ void __launch_bounds__(256)
dp_reduce_shard_kernel(const float* const* __restrict__ grads_peers, const float* grads_mc, float* __restrict__ grads_local,
                       long long off, long long n, int world, double* __restrict__ sumsq_local) {
  __shared__ float red[8];
  float acc = 0.f;
  const long long n4 = n / 4;  // shards are multiples of 4 elements
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (grads_mc != nullptr) {
    // four independent in-switch reductions in flight per thread (a multimem.ld_reduce round trip crosses the NVSwitch)
    for (; i + 3 * stride < n4; i += 4 * stride) {
      float4 t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(t[u].x), "=f"(t[u].y), "=f"(t[u].z), "=f"(t[u].w)
                     : "l"(grads_mc + off + 4 * (i + u * stride))
                     : "memory");
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        reinterpret_cast<float4*>(grads_local + off)[i + u * stride] = t[u];
        acc += t[u].x * t[u].x + t[u].y * t[u].y + t[u].z * t[u].z + t[u].w * t[u].w;
      }
    }
  }
  for (; i < n4; i += stride) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (grads_mc != nullptr) {
      // NVSwitch in-fabric reduction: ONE load through the multicast address returns the sum over every rank's copy
      // (1/G of the NVLink traffic of pulling G - 1 peers)
      asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                   : "l"(grads_mc + off + 4 * i)
                   : "memory");
    } else {
      for (int g = 0; g < world; ++g) {
        const float4 q = reinterpret_cast<const float4*>(grads_peers[g] + off)[i];
        t.x += q.x, t.y += q.y, t.z += q.z, t.w += q.w;
      }
    }
    reinterpret_cast<float4*>(grads_local + off)[i] = t;
    acc += t.x * t.x + t.y * t.y + t.z * t.z + t.w * t.w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(sumsq_local, static_cast<double>(t));
  }
}

__global__ void __launch_bounds__(256)
dp_adamw_shard_kernel(float* const* __restrict__ params_peers, float* params_mc, int world, int rank, const float* __restrict__ g,
                      float* __restrict__ m, float* __restrict__ v, long long off, long long n,
                      const double* __restrict__ sumsq_slots, const long long* __restrict__ step_ptr, float grad_scale,
                      mca_adamw_cfg c, float* __restrict__ total_norm_out) {
  __shared__ float s_coef, s_lr, s_bc1, s_bc2s;
  if (threadIdx.x == 0) {
    const long long step = *step_ptr + 1;
    double ss = 0.0;
    for (int r = 0; r < world; ++r) ss += sumsq_slots[r];  // same order on every rank
    const float total = static_cast<float>(sqrt(ss)) * grad_scale;
    if (blockIdx.x == 0 && total_norm_out != nullptr) *total_norm_out = total;
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
  const float* p_own = params_peers[rank] + off;
  const long long n4 = n / 4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 gq = reinterpret_cast<const float4*>(g + off)[i];
    float4 mq = reinterpret_cast<const float4*>(m + off)[i];
    float4 vq = reinterpret_cast<const float4*>(v + off)[i];
    float4 pq = reinterpret_cast<const float4*>(p_own)[i];
    const float gs[4] = {gq.x * coef, gq.y * coef, gq.z * coef, gq.w * coef};
    float* mm = reinterpret_cast<float*>(&mq);
    float* vv = reinterpret_cast<float*>(&vq);
    float* pp = reinterpret_cast<float*>(&pq);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      mm[k] = c.beta1 * mm[k] + (1.0f - c.beta1) * gs[k];
      vv[k] = c.beta2 * vv[k] + (1.0f - c.beta2) * gs[k] * gs[k];
      const float denom = sqrtf(vv[k]) / bc2s + c.eps;
      pp[k] = pp[k] * decay - step_size * (mm[k] / denom);
    }
    reinterpret_cast<float4*>(m + off)[i] = mq;
    reinterpret_cast<float4*>(v + off)[i] = vq;
    if (params_mc != nullptr) {
      // one multicast store: the switch replicates it into every rank's parameter buffer (this rank's included)
      asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(params_mc + off + 4 * i), "f"(pq.x),
                   "f"(pq.y), "f"(pq.z), "f"(pq.w)
                   : "memory");
    } else {
      for (int r = 0; r < world; ++r) reinterpret_cast<float4*>(params_peers[r] + off)[i] = pq;
    }
  }
}

__global__ void bump_step_only_kernel(long long* step_ptr) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *step_ptr += 1;
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

static int dp_reduce_shard_impl(const float* const* grads_peers_dev, const float* grads_mc, float* grads_local,
                               long long shard_off, long long shard_n, int world, double* sumsq_local, void* stream_);

extern "C" int mca_dp_reduce_shard(const float* const* grads_peers_dev, float* grads_local, long long shard_off,
                                   long long shard_n, int world, double* sumsq_local, void* stream_) {
  return dp_reduce_shard_impl(grads_peers_dev, nullptr, grads_local, shard_off, shard_n, world, sumsq_local, stream_);
}

extern "C" int mca_dp_reduce_shard_mc(const float* grads_multicast, float* grads_local, long long shard_off,
                                      long long shard_n, int world, double* sumsq_local, void* stream_) {
  if (grads_multicast == nullptr) return MCA_ERR_ARG;
  return dp_reduce_shard_impl(nullptr, grads_multicast, grads_local, shard_off, shard_n, world, sumsq_local, stream_);
}

static int dp_reduce_shard_impl(const float* const* grads_peers_dev, const float* grads_mc, float* grads_local,
                               long long shard_off, long long shard_n, int world, double* sumsq_local, void* stream_) {
  if (shard_n < 0 || (shard_n % 4) != 0 || (shard_off % 4) != 0 || world < 1) return MCA_ERR_SHAPE;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (cudaMemsetAsync(sumsq_local, 0, sizeof(double), stream) != cudaSuccess) return MCA_ERR_CUDA;
  if (shard_n == 0) return MCA_OK;
  dp_reduce_shard_kernel<<<num_sms() * 4, 256, 0, stream>>>(grads_peers_dev, grads_mc, grads_local, shard_off, shard_n, world,
                                                            sumsq_local);
  return check_launch();
}

static int dp_adamw_shard_impl(float* const* params_peers_dev, float* params_mc, int world, int rank, const float* grads_local,
                              float* exp_avg, float* exp_avg_sq, long long shard_off, long long shard_n,
                              const double* sumsq_slots, long long* step_dev, float* total_norm_out, float grad_scale,
                              const mca_adamw_cfg* cfg_host, void* stream_);

extern "C" int mca_dp_adamw_shard(float* const* params_peers_dev, int world, int rank, const float* grads_local,
                                  float* exp_avg, float* exp_avg_sq, long long shard_off, long long shard_n,
                                  const double* sumsq_slots, long long* step_dev, float* total_norm_out,
                                  float grad_scale, const mca_adamw_cfg* cfg_host, void* stream_) {
  return dp_adamw_shard_impl(params_peers_dev, nullptr, world, rank, grads_local, exp_avg, exp_avg_sq, shard_off, shard_n,
                             sumsq_slots, step_dev, total_norm_out, grad_scale, cfg_host, stream_);
}

extern "C" int mca_dp_adamw_shard_mc(float* const* params_peers_dev, float* params_multicast, int world, int rank,
                                     const float* grads_local, float* exp_avg, float* exp_avg_sq, long long shard_off,
                                     long long shard_n, const double* sumsq_slots, long long* step_dev,
                                     float* total_norm_out, float grad_scale, const mca_adamw_cfg* cfg_host, void* stream_) {
  if (params_multicast == nullptr) return MCA_ERR_ARG;
  return dp_adamw_shard_impl(params_peers_dev, params_multicast, world, rank, grads_local, exp_avg, exp_avg_sq, shard_off,
                             shard_n, sumsq_slots, step_dev, total_norm_out, grad_scale, cfg_host, stream_);
}

static int dp_adamw_shard_impl(float* const* params_peers_dev, float* params_mc, int world, int rank, const float* grads_local,
                              float* exp_avg, float* exp_avg_sq, long long shard_off, long long shard_n,
                              const double* sumsq_slots, long long* step_dev, float* total_norm_out, float grad_scale,
                              const mca_adamw_cfg* cfg_host, void* stream_) {
  if (shard_n < 0 || (shard_n % 4) != 0 || (shard_off % 4) != 0 || world < 1 || rank < 0 || rank >= world)
    return MCA_ERR_SHAPE;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (shard_n > 0)
    dp_adamw_shard_kernel<<<num_sms() * 4, 256, 0, stream>>>(params_peers_dev, params_mc, world, rank, grads_local, exp_avg,
                                                             exp_avg_sq, shard_off, shard_n, sumsq_slots, step_dev,
                                                             grad_scale, *cfg_host, total_norm_out);
  bump_step_only_kernel<<<1, 32, 0, stream>>>(step_dev);
  return check_launch();
}
