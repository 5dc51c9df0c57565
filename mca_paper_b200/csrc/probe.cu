// Linear probe on frozen embeddings (lp_accel_gpu.py:22-35 dataset, :100-104 nn.Linear head, :160-231 loop): one EPOCH of
// mini-batch training — forward, loss, backward, clip_grad_norm_, AdamW, LR schedule, every step — is ONE launch of one
// thread-block cluster.
//
// The whole model is L x 512 weights + L biases, so every CTA of the cluster keeps its own copy of the parameters and the
// AdamW moments in shared memory for the entire epoch.  Per step the batch rows are dealt round-robin to the warps of the 8
// CTAs; a warp reads a row ONCE (2 KB, coalesced), forms the L predictions with shuffles, the loss derivative, and
// accumulates dW in registers.  Each CTA then writes its partial gradient into slot `rank` of EVERY CTA's exchange buffer
// through distributed shared memory (st.shared::cluster), one cluster barrier, and every CTA sums the 8 slots in the same
// order: all copies apply bit-identical clip + AdamW updates and never diverge.  No global-memory traffic except the
// embedding rows, no host involvement between steps (the reference runs ~10 kernel launches and a Python DataLoader
// iteration per step).
#include <math_constants.h>

#include "mca_b200.h"
#include "ptx.cuh"
#include "runtime.h"

namespace mca {

constexpr int PB_D = 512;        // embedding width
constexpr int PB_CL = 8;         // CTAs per cluster
constexpr int PB_THREADS = 256;  // 8 warps
constexpr int PB_MAXL = 8;       // outputs (the MOSEI labels: 1 sentiment + 6 emotions)

__device__ __forceinline__ void pb_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t pb_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void pb_st_peer(float* local_ptr, uint32_t rank, float v) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local_ptr)), "r"(rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
}

// transformers.get_scheduler(name) after `cur` scheduler steps (same modes as optim.cu)
__device__ __forceinline__ float pb_lr(const mca_adamw_cfg& c, long long step /*1-based*/) {
  if (c.lr_mode == 0) return c.lr;
  const double cur = static_cast<double>(step - 1);
  if (cur < c.warmup_steps) return c.lr * static_cast<float>(cur / fmax(1.0, static_cast<double>(c.warmup_steps)));
  if (c.lr_mode == 2) return c.lr;
  if (c.lr_mode == 3)
    return c.lr * static_cast<float>(fmax(0.0, (static_cast<double>(c.total_steps) - cur) /
                                                   fmax(1.0, static_cast<double>(c.total_steps - c.warmup_steps))));
  const double prog = (cur - c.warmup_steps) / fmax(1.0, static_cast<double>(c.total_steps - c.warmup_steps));
  return c.lr * static_cast<float>(fmax(0.0, 0.5 * (1.0 + cos(3.14159265358979323846 * prog))));
}

struct ProbeArgs {
  const float* X;      // [n_rows_total, 512] embeddings
  const float* Y;      // [n_rows_total, L] labels
  const int* order;    // [n] dataset indices in visiting order (the sampler's permutation; identity for evaluation)
  int n, B, L, loss;   // loss: 0 L1, 1 MSE, 2 BCE-with-logits, 3 cross-entropy with probability targets
  int train;
  float* state;        // [3][L*512 + L]: parameters (W row-major [L,512], then bias), exp_avg, exp_avg_sq
  long long* step;     // optimiser steps taken so far
  mca_adamw_cfg cfg;
  float* pred;         // [n_rows_total, L] predictions, stored at the dataset index (for the epoch metrics)
  double* loss_sum;    // += sum over batches of the batch-mean loss (the reference's epoch_loss)
  float* last_grad_norm;
};

// shared memory (floats): P = L*512 + L (+ 1 loss slot in the exchange)
//   w[P] m[P] v[P] acc[P + 1] xchg[PB_CL][P + 1] red[16]
__global__ void __cluster_dims__(PB_CL, 1, 1) __launch_bounds__(PB_THREADS, 1) probe_epoch_kernel(const ProbeArgs a) {
  extern __shared__ __align__(16) float ps[];
  const int L = a.L, P = L * PB_D + L, PX = P + 1;
  const int PA = (P + 3) & ~3, PXA = (PX + 3) & ~3;   // section strides (16-byte aligned: acc is accessed as float4)
  float* w = ps;
  float* m = w + PA;
  float* v = m + PA;
  float* acc = v + PA;
  float* xchg = acc + PXA;
  float* red = xchg + PB_CL * PXA;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = pb_cluster_rank();
  for (int i = tid; i < P; i += PB_THREADS) {
    w[i] = a.state[i];
    m[i] = a.state[P + i];
    v[i] = a.state[2 * P + i];
  }
  long long step = *a.step;
  double loss_epoch = 0.0;
  float gnorm = 0.f;
  __syncthreads();
  const int n_steps = (a.n + a.B - 1) / a.B;
  for (int s = 0; s < n_steps; ++s) {
    const int r0 = s * a.B, nb = min(a.B, a.n - r0);
    const float inv_elems = 1.0f / static_cast<float>(a.loss == 3 ? nb : nb * L);
    // ---- forward + loss + local gradient: rows dealt round-robin to the 64 warps of the cluster
    float gw[PB_MAXL][16];
    float gb[PB_MAXL];
    float lsum = 0.f;
#pragma unroll
    for (int l = 0; l < PB_MAXL; ++l) {
      gb[l] = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) gw[l][i] = 0.f;
    }
    for (int r = static_cast<int>(rank) * 8 + warp; r < nb; r += PB_CL * 8) {
      const int idx = a.order[r0 + r];
      const float4* xr = reinterpret_cast<const float4*>(a.X + static_cast<long long>(idx) * PB_D);
      float x[16];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 q = xr[lane + 32 * i];
        x[4 * i] = q.x, x[4 * i + 1] = q.y, x[4 * i + 2] = q.z, x[4 * i + 3] = q.w;
      }
      float p[PB_MAXL], y[PB_MAXL], d[PB_MAXL];
#pragma unroll
      for (int l = 0; l < PB_MAXL; ++l) {
        p[l] = 0.f, y[l] = 0.f, d[l] = 0.f;
        if (l < L) {
          float t = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 q = *reinterpret_cast<const float4*>(w + l * PB_D + (lane + 32 * i) * 4);
            t += x[4 * i] * q.x + x[4 * i + 1] * q.y + x[4 * i + 2] * q.z + x[4 * i + 3] * q.w;
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
          p[l] = t + w[L * PB_D + l];
          y[l] = a.Y[static_cast<long long>(idx) * L + l];
        }
      }
      // loss of this row (every lane computes the same scalars) and d loss / d prediction
      float rl = 0.f;
      if (a.loss == 3) {  // F.cross_entropy with probability targets: -sum_c y_c log softmax(p)_c, mean over rows
        float mx = -CUDART_INF_F, se = 0.f, ysum = 0.f;
#pragma unroll
        for (int l = 0; l < PB_MAXL; ++l)
          if (l < L) mx = fmaxf(mx, p[l]);
#pragma unroll
        for (int l = 0; l < PB_MAXL; ++l)
          if (l < L) se += expf(p[l] - mx), ysum += y[l];
        const float lse = mx + logf(se);
#pragma unroll
        for (int l = 0; l < PB_MAXL; ++l)
          if (l < L) {
            rl -= y[l] * (p[l] - lse);
            d[l] = (expf(p[l] - lse) * ysum - y[l]) * inv_elems;
          }
      } else {
#pragma unroll
        for (int l = 0; l < PB_MAXL; ++l)
          if (l < L) {
            const float e = p[l] - y[l];
            if (a.loss == 0) {
              rl += fabsf(e);
              d[l] = (e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f)) * inv_elems;
            } else if (a.loss == 1) {
              rl += e * e;
              d[l] = 2.f * e * inv_elems;
            } else {  // BCEWithLogits: max(p,0) - p*y + log1p(exp(-|p|))
              rl += fmaxf(p[l], 0.f) - p[l] * y[l] + log1pf(expf(-fabsf(p[l])));
              d[l] = (1.f / (1.f + expf(-p[l])) - y[l]) * inv_elems;
            }
          }
      }
      if (lane == 0) {
        lsum += rl;
        for (int l = 0; l < L; ++l) a.pred[static_cast<long long>(idx) * L + l] = p[l];
      }
      if (a.train) {
#pragma unroll
        for (int l = 0; l < PB_MAXL; ++l)
          if (l < L) {
            gb[l] += d[l];
#pragma unroll
            for (int i = 0; i < 16; ++i) gw[l][i] = fmaf(d[l], x[i], gw[l][i]);
          }
      }
    }
    // ---- CTA partial: the 8 warps add their registers into acc one after the other (fixed order)
    for (int i = tid; i < PX; i += PB_THREADS) acc[i] = 0.f;
    __syncthreads();
    for (int wv = 0; wv < 8; ++wv) {
      if (warp == wv) {
        if (a.train) {
#pragma unroll
          for (int l = 0; l < PB_MAXL; ++l)
            if (l < L) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float4* dst = reinterpret_cast<float4*>(acc + l * PB_D + (lane + 32 * i) * 4);
                float4 q = *dst;
                q.x += gw[l][4 * i], q.y += gw[l][4 * i + 1], q.z += gw[l][4 * i + 2], q.w += gw[l][4 * i + 3];
                *dst = q;
              }
              if (lane == 0) acc[L * PB_D + l] += gb[l];
            }
        }
        if (lane == 0) acc[P] += lsum;
      }
      __syncthreads();
    }
    // ---- all-gather of the partials through distributed shared memory: slot `rank` of every CTA's exchange buffer
    for (int i = tid; i < PX; i += PB_THREADS) {
      const float t = acc[i];
#pragma unroll
      for (uint32_t dst = 0; dst < PB_CL; ++dst) pb_st_peer(xchg + rank * PXA + i, dst, t);
    }
    pb_cluster_sync();
    // ---- every CTA: same sum order -> same gradient -> same update
    float ss = 0.f;
    for (int i = tid; i < PX; i += PB_THREADS) {
      float t = 0.f;
#pragma unroll
      for (int c = 0; c < PB_CL; ++c) t += xchg[c * PXA + i];
      acc[i] = t;
      if (i < P) ss += t * t;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    // the exchange buffers may be overwritten by the next step's partials only after every CTA has read them
    pb_cluster_sync();
    loss_epoch += static_cast<double>(acc[P] * inv_elems);
    if (a.train) {
      float tot = 0.f;
      for (int wv = 0; wv < 8; ++wv) tot += red[wv];
      gnorm = sqrtf(tot);
      const float coef = a.cfg.max_norm > 0.f ? fminf(1.0f, a.cfg.max_norm / (gnorm + 1e-6f)) : 1.0f;
      ++step;
      const float lr = pb_lr(a.cfg, step);
      const float bc1 = static_cast<float>(1.0 - pow(static_cast<double>(a.cfg.beta1), static_cast<double>(step)));
      const float bc2s = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(a.cfg.beta2), static_cast<double>(step))));
      const float decay = 1.0f - lr * a.cfg.weight_decay, step_size = lr / bc1;
      for (int i = tid; i < P; i += PB_THREADS) {
        const float g = acc[i] * coef;
        const float mi = a.cfg.beta1 * m[i] + (1.0f - a.cfg.beta1) * g;
        const float vi = a.cfg.beta2 * v[i] + (1.0f - a.cfg.beta2) * g * g;
        m[i] = mi, v[i] = vi;
        w[i] = w[i] * decay - step_size * (mi / (sqrtf(vi) / bc2s + a.cfg.eps));
      }
    }
    __syncthreads();
  }
  if (rank == 0) {
    if (a.train) {
      for (int i = tid; i < P; i += PB_THREADS) {
        a.state[i] = w[i];
        a.state[P + i] = m[i];
        a.state[2 * P + i] = v[i];
      }
    }
    if (tid == 0) {
      if (a.train) *a.step = step;
      *a.loss_sum += loss_epoch;
      if (a.last_grad_norm != nullptr && a.train) *a.last_grad_norm = gnorm;
    }
  }
}

// Pearson correlation of two fp32 vectors (torchmetrics.PearsonCorrCoef over one epoch's (prediction, label) stream,
// lp_accel_gpu.py:148-149,197-198): one block, fp64 moments.
__global__ void __launch_bounds__(256) probe_pcc_kernel(const float* __restrict__ p, const float* __restrict__ y, long long n,
                                                        float* __restrict__ out) {
  __shared__ double red[5][8];
  double s[5] = {0, 0, 0, 0, 0};
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const double a = p[i], b = y[i];
    s[0] += a, s[1] += b, s[2] += a * a, s[3] += b * b, s[4] += a * b;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
    if (lane == 0) red[k][warp] = s[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[5] = {0, 0, 0, 0, 0};
    for (int k = 0; k < 5; ++k)
      for (int wv = 0; wv < 8; ++wv) t[k] += red[k][wv];
    const double nn = static_cast<double>(n);
    const double cov = t[4] - t[0] * t[1] / nn, va = t[2] - t[0] * t[0] / nn, vb = t[3] - t[1] * t[1] / nn;
    *out = static_cast<float>(cov / sqrt(va * vb));
  }
}

}  // namespace mca

using namespace mca;

extern "C" int mca_probe_epoch(const float* X, const float* Y, const int* order, int n, int batch_size, int n_out, int loss_kind,
                               int train, float* state, long long* step_dev, const mca_adamw_cfg* cfg_host, float* pred,
                               double* loss_sum, float* last_grad_norm, void* stream_) {
  if (n <= 0 || batch_size <= 0 || n_out <= 0 || n_out > PB_MAXL || loss_kind < 0 || loss_kind > 3) return MCA_ERR_SHAPE;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int P = n_out * PB_D + n_out, PX = P + 1;
  const int PA = (P + 3) & ~3, PXA = (PX + 3) & ~3;
  const size_t smem = (3 * static_cast<size_t>(PA) + PXA + static_cast<size_t>(PB_CL) * PXA + 16) * sizeof(float);
  static size_t attr = 0;
  if (smem > attr) {
    if (cudaFuncSetAttribute(probe_epoch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) !=
        cudaSuccess)
      return MCA_ERR_CUDA;
    attr = smem;
  }
  ProbeArgs a{X, Y, order, n, batch_size, n_out, loss_kind, train, state, step_dev, *cfg_host, pred, loss_sum, last_grad_norm};
  probe_epoch_kernel<<<PB_CL, PB_THREADS, smem, stream>>>(a);
  return check_launch();
}

extern "C" int mca_probe_pcc(const float* pred, const float* y, long long n, float* out, void* stream) {
  if (n <= 0) return MCA_ERR_SHAPE;
  probe_pcc_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(pred, y, n, out);
  return check_launch();
}
