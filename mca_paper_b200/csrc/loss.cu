// All-pairs temperature-scaled InfoNCE, forward and analytic backward, one CTA per (contrastive pair, direction).
// Replaces the Python loop MCAPretrainingLoss.forward (model.py:196-232) and, per pair,
// ContrastiveLossWithTemperature (utils/contrastive_loss_with_temperature.py:71-100,187): in-place clamp of
// logit_scale, T = exp(s), logits_a = a b_all^T T, logits_b = b a_all^T T, rows selected by the presence mask,
// labels = B*rank + i, (CE_a + CE_b)/2, NaN when no row is selected, then the NaN-aware mean of model.py:221-232 —
// with zero host synchronisations (the reference does one .item() per pair, model.py:225).
// `pooled_all` is the all-gathered [G*B, R, d] block; gradients are produced for every gathered row so the host can
// reduce-scatter them (the autograd of torch.distributed.nn.functional.all_gather, utils/distributed.py:45-46).
//
// Peer-memory exchange (data parallel inside one NVLink/NVSwitch domain), replacing the NCCL all_gather /
// reduce_scatter around these kernels: the gathered blocks live in P2P-mapped symmetric memory; every rank PUSHES its
// [B, R, d] pooled block into slot `rank` of every peer's gathered buffer (mca_p2p_push_rows: posted NVLink stores, no
// read latency), a flag barrier (mca_xgpu_barrier) publishes them, the loss kernels then run on local memory; the
// backward writes its [G*B, R, d] gradient block locally and, after a second barrier, mca_p2p_reduce_rows PULLS and
// sums every rank's slice of this rank's rows (the reduce-scatter).  (Reading the peers' rows from inside the loss
// kernels was measured slower: their 2 KB dot-product rows are latency-bound over NVLink.)  Everything is a plain
// kernel on one stream, so forward + loss + backward stay in one CUDA graph.
#include <math_constants.h>

#include "mca_b200.h"
#include "ptx.cuh"
#include "runtime.h"

namespace mca {

constexpr int LOSS_THREADS = 512;  // one thread per embedding column in the backward, 16 warps of dot products

struct LossArgs {
  const float* pooled_all;  // [GB, R, d]
  const uint8_t* present;   // [B, n_mod] local
  const mca_loss_pair* plan;
  float* logit_scale;
  int B, GB, R, d, n_mod, rank, n_pairs;
};

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// row r of gathered sample j
__device__ __forceinline__ const float* pooled_row(const LossArgs& a, int j, int r) {
  return a.pooled_all + (static_cast<long long>(j) * a.R + r) * a.d;
}

__device__ __forceinline__ bool row_selected(const LossArgs& a, const mca_loss_pair& p, int i) {
  unsigned bits = 0;
  for (int m = 0; m < a.n_mod; ++m) bits |= (a.present[i * a.n_mod + m] ? 1u : 0u) << m;
  if ((bits & p.all_mask) != p.all_mask) return false;
  if (p.any_mask != 0 && (bits & p.any_mask) == 0) return false;
  return true;
}

// One CTA = one (pair, direction): direction 0 scores the local a rows against every gathered b row, direction 1 the
// local b rows against every gathered a row.  The B local query rows are staged in shared memory once; each warp then
// reads a gathered key row ONCE (coalesced float4) and dots it with all B queries, so the global traffic is GB rows per
// CTA, not B*GB.  logits[i*GB + j] = T * q_i . k_j
constexpr int LOSS_MAXB = 32;  // local batch rows held in registers by the backward (configs/*_i.yaml use 32)

__device__ void stage_and_logits(const LossArgs& a, const mca_loss_pair& p, int dir, float T, float* qs, float* lg) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int rq = dir == 0 ? p.a_row : p.b_row, rk = dir == 0 ? p.b_row : p.a_row;
  const int d4 = a.d / 4;
  for (int t = threadIdx.x; t < a.B * d4; t += blockDim.x) {
    const int i = t / d4, c = t % d4;
    reinterpret_cast<float4*>(qs)[t] = reinterpret_cast<const float4*>(pooled_row(a, a.rank * a.B + i, rq))[c];
  }
  __syncthreads();
  for (int j = warp; j < a.GB; j += nwarps) {
    const float4* k4 = reinterpret_cast<const float4*>(pooled_row(a, j, rk));
    float acc[LOSS_MAXB];
#pragma unroll
    for (int i = 0; i < LOSS_MAXB; ++i) acc[i] = 0.f;
    for (int c = lane; c < d4; c += 32) {
      const float4 y = k4[c];
#pragma unroll
      for (int i = 0; i < LOSS_MAXB; ++i) {
        if (i < a.B) {
          const float4 x = reinterpret_cast<const float4*>(qs)[i * d4 + c];
          acc[i] += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < LOSS_MAXB; ++i) {
      if (i < a.B) {
        const float v = warp_sum_f(acc[i]);
        if (lane == 0) lg[i * a.GB + j] = v * T;
      }
    }
  }
  __syncthreads();
}

__global__ void clamp_scale_kernel(float* s, float lo, float hi) {
  if (threadIdx.x == 0) *s = fminf(fmaxf(*s, lo), hi);
}

// sum0[pair] (direction 0) / sum1[pair] (direction 1) = sum of the selected rows' cross entropies; the caller passes the
// `losses` and `w_default` outputs as the two scratch arrays, loss_reduce_kernel finishes them
__global__ void __launch_bounds__(LOSS_THREADS)
loss_fwd_kernel(LossArgs a, float* __restrict__ sum0, float* __restrict__ sum1) {
  extern __shared__ __align__(16) float sm[];
  float* qs = sm;
  float* lg = sm + a.B * a.d;
  __shared__ float s_sum;
  const int pair = blockIdx.x, dir = blockIdx.y;
  const mca_loss_pair p = a.plan[pair];
  const float T = expf(*a.logit_scale);
  if (threadIdx.x == 0) s_sum = 0.f;
  stage_and_logits(a, p, dir, T, qs, lg);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int i = warp; i < a.B; i += nwarps) {
    if (!row_selected(a, p, i)) continue;
    const float* row = lg + i * a.GB;
    float mx = -CUDART_INF_F;
    for (int j = lane; j < a.GB; j += 32) mx = fmaxf(mx, row[j]);
    mx = warp_max_f(mx);
    float se = 0.f;
    for (int j = lane; j < a.GB; j += 32) se += expf(row[j] - mx);
    se = warp_sum_f(se);
    if (lane == 0) atomicAdd(&s_sum, mx + logf(se) - row[a.rank * a.B + i]);
  }
  __syncthreads();
  if (threadIdx.x == 0) (dir == 0 ? sum0 : sum1)[pair] = s_sum;
}

// losses[p] = (CE_a + CE_b) / 2 averaged over the selected rows, NaN when no row is selected;
// summary[0] = loss (model.py:224-232), [1] = fcl_loss, [2] = no-fcl_loss (model.py:221-222), [3] = #non-NaN;
// w_default[p] = d loss / d loss_p
__global__ void __launch_bounds__(256)
loss_reduce_kernel(LossArgs a, float* __restrict__ losses, int P, float* __restrict__ summary,
                   float* __restrict__ w_default) {
  const mca_loss_pair* plan = a.plan;
  // thread = pair: selected-row count and the pair's loss from the two direction sums
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    int cnt = 0;
    for (int i = 0; i < a.B; ++i) cnt += row_selected(a, plan[p], i) ? 1 : 0;
    losses[p] = cnt == 0 ? CUDART_NAN_F : 0.5f * (losses[p] + w_default[p]) / static_cast<float>(cnt);
  }
  __syncthreads();
  __shared__ int s_nv;
  if (threadIdx.x == 0) {
    float tot = 0.f, fcl = 0.f, nofcl = 0.f;
    int nv = 0, nf = 0, nn = 0;
    for (int p = 0; p < P; ++p) {
      float v = losses[p];
      const bool isn = isnan(v);
      if (!isn) ++nv;
      if (isn) v = 0.f;
      else if (isinf(v)) v = v > 0 ? 3.402823466e+38f : -3.402823466e+38f;
      tot += v;
      if (plan[p].is_fcl) fcl += v, ++nf;
      else nofcl += v, ++nn;
    }
    summary[0] = nv == 0 ? tot : tot / static_cast<float>(nv);
    summary[1] = nf > 0 ? fcl / nf : 0.f;
    summary[2] = nn > 0 ? nofcl / nn : 0.f;
    summary[3] = static_cast<float>(nv);
    s_nv = nv;
  }
  __syncthreads();
  const int nv = s_nv;
  for (int p = threadIdx.x; p < P; p += blockDim.x)
    w_default[p] = (isnan(losses[p]) || nv == 0) ? 0.f : 1.0f / static_cast<float>(nv);
}

__global__ void __launch_bounds__(LOSS_THREADS)
loss_bwd_kernel(LossArgs a, const float* __restrict__ w, float* __restrict__ dpooled_all, float* __restrict__ dscale) {
  extern __shared__ __align__(16) float sm[];
  float* qs = sm;
  float* lg = sm + a.B * a.d;
  __shared__ int s_cnt;
  __shared__ float s_ds;
  const int pair = blockIdx.x, dir = blockIdx.y;
  const mca_loss_pair p = a.plan[pair];
  const float wp = w[pair];
  if (wp == 0.f) return;
  const float T = expf(*a.logit_scale);
  if (threadIdx.x == 0) {
    int c = 0;
    for (int i = 0; i < a.B; ++i) c += row_selected(a, p, i) ? 1 : 0;
    s_cnt = c;
    s_ds = 0.f;
  }
  stage_and_logits(a, p, dir, T, qs, lg);
  if (s_cnt == 0) return;
  const float gscale = wp * 0.5f / static_cast<float>(s_cnt);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  // logits -> dL/dlogits in place; accumulate dL/ds = sum dlogit * logit
  for (int i = warp; i < a.B; i += nwarps) {
    float* row = lg + i * a.GB;
    if (!row_selected(a, p, i)) {
      for (int j = lane; j < a.GB; j += 32) row[j] = 0.f;
      continue;
    }
    float mx = -CUDART_INF_F;
    for (int j = lane; j < a.GB; j += 32) mx = fmaxf(mx, row[j]);
    mx = warp_max_f(mx);
    float se = 0.f;
    for (int j = lane; j < a.GB; j += 32) se += expf(row[j] - mx);
    se = warp_sum_f(se);
    float ds = 0.f;
    const int label = a.rank * a.B + i;
    for (int j = lane; j < a.GB; j += 32) {
      const float l = row[j];
      const float g = gscale * (expf(l - mx) / se - (j == label ? 1.f : 0.f));
      ds += g * l;
      row[j] = g;
    }
    ds = warp_sum_f(ds);
    if (lane == 0) atomicAdd(&s_ds, ds);
  }
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.z == 0) atomicAdd(dscale, s_ds);
  // embedding gradients, thread = embedding column: d query_i += T sum_j dl[i][j] key_j ; d key_j += T sum_i dl[i][j] query_i.
  // Every gathered key row is read once per CTA (coalesced across the threads).  Under data parallelism the gathered rows
  // are dealt to gridDim.z CTAs per (pair, direction) — each recomputes the (cheap) logits and walks its own slice of keys:
  // the serial walk over all G*B rows was what grew with the world size (62 us at G = 8).
  const int rq = dir == 0 ? p.a_row : p.b_row, rk = dir == 0 ? p.b_row : p.a_row;
  const int jchunk = (a.GB + static_cast<int>(gridDim.z) - 1) / static_cast<int>(gridDim.z);
  const int j0 = static_cast<int>(blockIdx.z) * jchunk, j1 = min(a.GB, j0 + jchunk);
  for (int c = threadIdx.x; c < a.d; c += blockDim.x) {
    float qreg[LOSS_MAXB], dq[LOSS_MAXB];
#pragma unroll
    for (int i = 0; i < LOSS_MAXB; ++i) qreg[i] = i < a.B ? qs[i * a.d + c] : 0.f, dq[i] = 0.f;
    for (int j = j0; j < j1; ++j) {
      const float k = pooled_row(a, j, rk)[c];
      float dk = 0.f;
#pragma unroll
      for (int i = 0; i < LOSS_MAXB; ++i) {
        if (i < a.B) {
          const float g = lg[i * a.GB + j];
          dq[i] += g * k;
          dk += g * qreg[i];
        }
      }
      if (dk != 0.f) atomicAdd(dpooled_all + (static_cast<long long>(j) * a.R + rk) * a.d + c, T * dk);
    }
#pragma unroll
    for (int i = 0; i < LOSS_MAXB; ++i)
      if (i < a.B && dq[i] != 0.f)
        atomicAdd(dpooled_all + (static_cast<long long>(a.rank * a.B + i) * a.R + rq) * a.d + c, T * dq[i]);
  }
}

// ---- cross-GPU flag barrier over peer-mapped memory.  flags[g] (uint32[G], one array per rank, every rank can address
// all of them): rank r publishes the barrier's epoch in slot r of EVERY rank's array (release, system scope) and waits
// until every slot of its own array has reached the epoch (acquire).  The epoch is a device counter, so the same
// captured graph can be replayed.  A peer that never arrives is reported after ~10 s: err_flag (host-mapped) = 1 + its rank, then the kernel traps, so the
// step fails loudly instead of hanging the GPU or running on with stale data.
// Optional payload: one double of this rank (e.g. its partial gradient sum of squares) is stored into slot `rank` of
// every rank's payload array before the flag is raised, so it is visible to whoever passes the barrier.
__global__ void xgpu_barrier_kernel(uint32_t* const* __restrict__ flags_peers, int world, int rank,
                                    uint32_t* __restrict__ epoch, int* __restrict__ err_flag,
                                    const double* __restrict__ payload, double* const* __restrict__ payload_peers) {
  __shared__ uint32_t s_epoch;
  if (threadIdx.x == 0) s_epoch = ++(*epoch);
  __syncthreads();
  const uint32_t e = s_epoch;
  const int g = threadIdx.x;
  if (g >= world) return;
  if (payload != nullptr) payload_peers[g][rank] = *payload;
  __threadfence_system();  // everything this GPU wrote before the barrier is visible before the flag is
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flags_peers[g] + rank), "r"(e) : "memory");
  const uint32_t* mine = flags_peers[rank] + g;
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    if (static_cast<int32_t>(v - e) >= 0) break;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 10000000000ull) {
      // A peer never arrived: record it where the host can still read it (err_flag is pinned host memory) and stop
      // this GPU's step with a trap -- continuing would train on stale gathered rows / gradients / parameter shards.
      *reinterpret_cast<volatile int*>(err_flag) = 1 + g;
      __threadfence_system();
      __trap();
    }
  }
}

// dst_peers[g][off + i] = src[i] for every rank g (the push form of an all-gather; float4 lanes, posted P2P stores)
__global__ void __launch_bounds__(256)
p2p_push_rows_kernel(const float* __restrict__ src, float* const* __restrict__ dst_peers, long long off, long long n4,
                     int world) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    for (int g = 0; g < world; ++g) reinterpret_cast<float4*>(dst_peers[g] + off)[i] = v;
  }
}

// dst[i] = sum over ranks g of src_peers[g][off + i]  (the pull form of a reduce-scatter; float4 lanes, coalesced P2P reads)
__global__ void __launch_bounds__(256)
p2p_reduce_rows_kernel(const float* const* __restrict__ src_peers, long long off, float* __restrict__ dst, long long n4,
                       int world) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int g = 0; g < world; ++g) {
      const float4 v = reinterpret_cast<const float4*>(src_peers[g] + off)[i];
      acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
    reinterpret_cast<float4*>(dst)[i] = acc;
  }
}

}  // namespace mca

using namespace mca;

constexpr int LOSS_MAX_SMEM = 200 * 1024;  // B = 32 local rows x (d = 512 + GB = 256 gathered columns) x 4 B = 96 KB
static int loss_smem_bytes(int B, int GB, int d) { return (B * d + B * GB) * static_cast<int>(sizeof(float)); }

extern "C" int mca_contrastive_allpairs_fwd(const float* pooled_all, const uint8_t* present,
                                            const mca_loss_pair* plan_dev, int n_pairs, float* logit_scale, int B,
                                            int GB, int R, int d, int n_mod, int rank, float scale_min,
                                            float scale_max, float* losses, float* summary, float* w_default,
                                            void* stream_) {
  if (n_pairs <= 0 || B <= 0 || B > LOSS_MAXB || GB < B || (d % 4) != 0 || n_mod > MCA_MAX_MODALITIES) return MCA_ERR_SHAPE;
  const int smem = loss_smem_bytes(B, GB, d);
  if (smem > LOSS_MAX_SMEM) return MCA_ERR_SHAPE;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(loss_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LOSS_MAX_SMEM);
    cudaFuncSetAttribute(loss_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LOSS_MAX_SMEM);
    attr = true;
  }
  clamp_scale_kernel<<<1, 32, 0, stream>>>(logit_scale, scale_min, scale_max);
  LossArgs a{pooled_all, present, plan_dev, logit_scale, B, GB, R, d, n_mod, rank, n_pairs};
  loss_fwd_kernel<<<dim3(n_pairs, 2), LOSS_THREADS, smem, stream>>>(a, losses, w_default);
  loss_reduce_kernel<<<1, 256, 0, stream>>>(a, losses, n_pairs, summary, w_default);
  return check_launch();
}

extern "C" int mca_contrastive_allpairs_bwd(const float* pooled_all, const uint8_t* present,
                                            const mca_loss_pair* plan_dev, int n_pairs, float* logit_scale, int B,
                                            int GB, int R, int d, int n_mod, int rank, const float* w,
                                            float* dpooled_all, float* dscale, void* stream_) {
  if (n_pairs <= 0 || B <= 0 || B > LOSS_MAXB || GB < B || (d % 4) != 0) return MCA_ERR_SHAPE;
  const int smem = loss_smem_bytes(B, GB, d);
  if (smem > LOSS_MAX_SMEM) return MCA_ERR_SHAPE;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(loss_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LOSS_MAX_SMEM);
    attr = true;
  }
  LossArgs a{pooled_all, present, plan_dev, logit_scale, B, GB, R, d, n_mod, rank, n_pairs};
  const int zsplit = GB / B > 1 ? (GB / B > 8 ? 8 : GB / B) : 1;  // one slice of gathered rows per rank (at most 8)
  loss_bwd_kernel<<<dim3(n_pairs, 2, zsplit), LOSS_THREADS, smem, stream>>>(a, w, dpooled_all, dscale);
  return check_launch();
}

extern "C" int mca_xgpu_barrier(uint32_t* const* flags_peers_dev, int world, int rank, uint32_t* epoch_dev,
                                int* err_flag_dev, const double* payload, double* const* payload_peers_dev,
                                void* stream_) {
  if (world < 1 || world > 32 || rank < 0 || rank >= world) return MCA_ERR_ARG;
  if (payload != nullptr && payload_peers_dev == nullptr) return MCA_ERR_ARG;
  xgpu_barrier_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(flags_peers_dev, world, rank, epoch_dev,
                                                                             err_flag_dev, payload, payload_peers_dev);
  return check_launch();
}

extern "C" int mca_p2p_push_rows(const float* src, float* const* dst_peers_dev, long long off_elems, long long n, int world,
                                 void* stream_) {
  if (n <= 0 || (n % 4) != 0 || (off_elems % 4) != 0 || world < 1) return MCA_ERR_SHAPE;
  const long long n4 = n / 4;
  const unsigned grid = static_cast<unsigned>((n4 + 255) / 256 < 1184 ? (n4 + 255) / 256 : 1184);
  p2p_push_rows_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(src, dst_peers_dev, off_elems, n4, world);
  return check_launch();
}

extern "C" int mca_p2p_reduce_rows(const float* const* src_peers_dev, long long off_elems, float* dst, long long n,
                                   int world, void* stream_) {
  if (n <= 0 || (n % 4) != 0 || (off_elems % 4) != 0 || world < 1) return MCA_ERR_SHAPE;
  const long long n4 = n / 4;
  const unsigned grid = static_cast<unsigned>(n4 + 255) / 256 < 1184u ? static_cast<unsigned>((n4 + 255) / 256) : 1184u;
  p2p_reduce_rows_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(src_peers_dev, off_elems, dst, n4,
                                                                                   world);
  return check_launch();
}
