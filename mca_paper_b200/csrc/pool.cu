// Attention pooling with R learned queries (model.py:472-473 -> Attention.forward model.py:73-105 with
// context = final tokens, attn_mask = pool_mask, key_padding_mask = padding) and the tiny dense ops around it.
// The R x N score block is small (R <= 32), so this is SIMT code bound by reading K/V once per (sample, head,
// query) from L2; the expensive part of pooling (the K/V projection of all N tokens) runs on the tcgen05 GEMM.
// Probabilities are materialised ([B,H,R,N] fp32, ~10 MB) so the backward is exact, including the reference's
// fully-masked-row rule: -finfo.max fill makes such a row uniform 1/N over ALL keys, with no gradient to scores.
#include <math_constants.h>

#include "mca_b200.h"
#include "ptx.cuh"
#include "runtime.h"

namespace mca {

constexpr int DH = 64;

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, t) : v + t;
  }
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
  for (int w = 1; w < nw; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
  return r;
}

__device__ __forceinline__ void load_row64(const __nv_bfloat16* p, float (&v)[DH]) {
  const uint4* p4 = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint4 q = p4[i];
    v[8 * i + 0] = bf16_lo(q.x), v[8 * i + 1] = bf16_hi(q.x), v[8 * i + 2] = bf16_lo(q.y), v[8 * i + 3] = bf16_hi(q.y);
    v[8 * i + 4] = bf16_lo(q.z), v[8 * i + 5] = bf16_hi(q.z), v[8 * i + 6] = bf16_lo(q.w), v[8 * i + 7] = bf16_hi(q.w);
  }
}

__device__ __forceinline__ void load_row64(const float* p, float (&v)[DH]) {
  const float4* p4 = reinterpret_cast<const float4*>(p);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float4 q = p4[i];
    v[4 * i + 0] = q.x, v[4 * i + 1] = q.y, v[4 * i + 2] = q.z, v[4 * i + 3] = q.w;
  }
}
__device__ __forceinline__ float kv_to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float kv_to_float(float v) { return v; }

// grid = B*H*R, block 256.  qp [R, H*64] fp32 (already scaled), kv [B*N, 2*H*64] (K | V), bf16 or (fp32-parity mode) fp32
template <typename KV>
__global__ void __launch_bounds__(256)
pool_fwd_kernel(const float* __restrict__ qp, const KV* __restrict__ kv, const uint8_t* __restrict__ padding,
                const uint8_t* __restrict__ keygrp, const uint32_t* __restrict__ rowbits, float* __restrict__ probs,
                uint8_t* __restrict__ full_masked, float* __restrict__ out, int B, int H, int R, int N) {
  extern __shared__ float sc[];  // [N]
  __shared__ float q[DH];
  __shared__ float red[8];
  __shared__ float part[4][DH];
  const int r = blockIdx.x % R, h = (blockIdx.x / R) % H, b = blockIdx.x / (R * H);
  const int ld = 2 * H * DH;
  if (threadIdx.x < DH) q[threadIdx.x] = qp[r * H * DH + h * DH + threadIdx.x];
  __syncthreads();
  const uint32_t bits = rowbits[r];
  float mx = -CUDART_INF_F;
  int n_allowed = 0;
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    const bool ok = ((bits >> keygrp[j]) & 1u) && padding[static_cast<long long>(b) * N + j] == 0;
    float s = -CUDART_INF_F;
    if (ok) {
      float k[DH];
      load_row64(kv + (static_cast<long long>(b) * N + j) * ld + h * DH, k);
      s = 0.f;
#pragma unroll
      for (int c = 0; c < DH; ++c) s += q[c] * k[c];
      ++n_allowed;
    }
    sc[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = block_reduce(mx, red, true);
  const float tot_allowed = block_reduce(static_cast<float>(n_allowed), red, false);
  float* prow = probs + (static_cast<long long>(b * H + h) * R + r) * N;
  if (tot_allowed == 0.f) {
    // every key masked: softmax of a constant row = 1/N over all N keys (padded and disallowed ones included)
    const float u = 1.0f / static_cast<float>(N);
    for (int j = threadIdx.x; j < N; j += blockDim.x) sc[j] = u, prow[j] = u;
    if (threadIdx.x == 0 && h == 0) full_masked[b * R + r] = 1;
  } else {
    float se = 0.f;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
      const float e = sc[j] == -CUDART_INF_F ? 0.f : expf(sc[j] - mx);
      sc[j] = e;
      se += e;
    }
    se = block_reduce(se, red, false);
    const float inv = 1.0f / se;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
      const float p = sc[j] * inv;
      sc[j] = p;
      prow[j] = p;
    }
    if (threadIdx.x == 0 && h == 0) full_masked[b * R + r] = 0;
  }
  __syncthreads();
  // out[b, r, h*64 + c] = sum_j p[j] * V[b, j, h, c]
  const int c = threadIdx.x % DH, grp = threadIdx.x / DH;
  float acc = 0.f;
  for (int j = grp; j < N; j += 4) {
    const float p = sc[j];
    if (p != 0.f) acc += p * kv_to_float(kv[(static_cast<long long>(b) * N + j) * ld + H * DH + h * DH + c]);
  }
  part[grp][c] = acc;
  __syncthreads();
  if (threadIdx.x < DH)
    out[(static_cast<long long>(b) * R + r) * H * DH + h * DH + threadIdx.x] =
        part[0][threadIdx.x] + part[1][threadIdx.x] + part[2][threadIdx.x] + part[3][threadIdx.x];
}

// ------------------------------------------------------------------ cluster variants (R <= 16)
// One thread-block CLUSTER of PC CTAs = one (sample, head); CTA c owns the keys [c*NC, (c+1)*NC).  K and V rows are
// read once, the softmax statistics and the partial outputs are exchanged through distributed shared memory
// (mapa + st.shared::cluster between cluster barriers), so the whole pooling forward (and backward) is one launch of
// B*H*PC CTAs instead of B*H long-running blocks.  Scores / probabilities are kept [row][key] in shared memory so the
// per-row passes are bank-conflict free.
constexpr int PC = 8;
constexpr int PR = 16;  // pooled rows supported by the cluster kernels

__device__ __forceinline__ void cl_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cl_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// store into the same shared-memory variable of CTA `rank` of this cluster
__device__ __forceinline__ void st_peer_f32(float* local_ptr, uint32_t rank, float v) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local_ptr)), "r"(rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// shared layout of both kernels (floats): q[16*64] g[16*64] a[16*NCP] b[16*NCP] xs0[PC*16] xs1[PC*16] part[4*16*64]
// xpart[PC*16*64]
__host__ __device__ inline size_t pool_cl_smem_floats(int NCP) {
  return 2 * PR * DH + 2 * static_cast<size_t>(PR) * NCP + 2 * PC * PR + 4 * PR * DH + static_cast<size_t>(PC) * PR * DH;
}

__global__ void __cluster_dims__(PC, 1, 1) __launch_bounds__(256)
pool_fwd_cl_kernel(const float* __restrict__ qp, const __nv_bfloat16* __restrict__ kv, const uint8_t* __restrict__ padding,
                   const uint8_t* __restrict__ keygrp, const uint32_t* __restrict__ rowbits, float* __restrict__ probs,
                   uint8_t* __restrict__ full_masked, float* __restrict__ out, int B, int H, int R, int N, int NC, int NCP) {
  extern __shared__ __align__(16) float smc[];
  float* sm = smc;
  float* q = sm;                       // [16][64]
  float* sc = q + 2 * PR * DH;         // [16][NCP] scores -> probabilities
  float* xmax = sc + 2 * PR * NCP;     // [PC][16] row maxima of every CTA of the cluster
  float* xsum = xmax + PC * PR;        // [PC][16] row sums
  float* part = xsum + PC * PR;        // [4][16][64]
  float* xpart = part + 4 * PR * DH;   // [PC][16][64] partial outputs gathered by rank 0
  uint32_t* skey = reinterpret_cast<uint32_t*>(sc + PR * NCP);  // [NCP] bit r = pooled row r may see this key
  __shared__ uint32_t rmask[32];  // per key group: the pooled rows allowed to see it (the pool mask is block-structured:
  __shared__ uint32_t s_fm;       // a key is visible to ~2 of the 16 rows, so only those dot products are computed)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = cl_rank();
  const int cid = blockIdx.x / PC;
  const int h = cid % H, b = cid / H;
  const int ld = 2 * H * DH;
  const int j0 = static_cast<int>(rank) * NC;
  const int nloc = max(0, min(NC, N - j0));
  for (int i = tid; i < PR * DH; i += 256) q[i] = (i / DH) < R ? qp[(i / DH) * H * DH + h * DH + (i % DH)] : 0.f;
  if (tid < 32) {
    uint32_t m = 0;
    for (int r = 0; r < R; ++r) m |= ((rowbits[r] >> tid) & 1u) << r;
    rmask[tid] = m;
  }
  if (tid == 0) s_fm = 0u;
  __syncthreads();
  // ---- phase 1: masked scores of this CTA's keys, thread = key; only the rows that may see the key are scored
  for (int jl = tid; jl < nloc; jl += 256) {
    const int j = j0 + jl;
    float k[DH];
    load_row64(kv + (static_cast<long long>(b) * N + j) * ld + h * DH, k);
    const bool pad = padding[static_cast<long long>(b) * N + j] != 0;
    const uint32_t rs = pad ? 0u : rmask[keygrp[j] & 31u];
    skey[jl] = rs;
#pragma unroll
    for (int r = 0; r < PR; ++r) sc[r * NCP + jl] = -CUDART_INF_F;
    for (uint32_t m = rs; m != 0u; m &= m - 1u) {
      const int r = __ffs(m) - 1;
      const float4* q4 = reinterpret_cast<const float4*>(q + r * DH);
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < DH / 4; ++c) {
        const float4 w = q4[c];
        s += w.x * k[4 * c] + w.y * k[4 * c + 1] + w.z * k[4 * c + 2] + w.w * k[4 * c + 3];
      }
      sc[r * NCP + jl] = s;
    }
  }
  __syncthreads();
  // ---- phase 2: softmax statistics, warp w owns rows w and w + 8; exchanged across the cluster
  float lmax[2], gmax[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = warp + 8 * i;
    float mx = -CUDART_INF_F;
    for (int jl = lane; jl < nloc; jl += 32) mx = fmaxf(mx, sc[r * NCP + jl]);
    lmax[i] = warp_max(mx);
    if (lane < PC) st_peer_f32(&xmax[rank * PR + r], lane, lmax[i]);
  }
  cl_sync();
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = warp + 8 * i;
    float mx = -CUDART_INF_F;
#pragma unroll
    for (int p = 0; p < PC; ++p) mx = fmaxf(mx, xmax[p * PR + r]);
    gmax[i] = mx;
    float se = 0.f;
    if (mx != -CUDART_INF_F) {
      for (int jl = lane; jl < nloc; jl += 32) {
        const float s = sc[r * NCP + jl];
        const float e = s == -CUDART_INF_F ? 0.f : expf(s - mx);
        sc[r * NCP + jl] = e;
        se += e;
      }
    }
    se = warp_sum(se);
    if (lane < PC) st_peer_f32(&xsum[rank * PR + r], lane, se);
  }
  cl_sync();
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = warp + 8 * i;
    if (r >= R) {  // rows beyond R only pad the block to 16: keep them finite
      for (int jl = lane; jl < nloc; jl += 32) sc[r * NCP + jl] = 0.f;
      continue;
    }
    float* prow = probs + (static_cast<long long>(b * H + h) * R + r) * N + j0;
    if (gmax[i] == -CUDART_INF_F) {
      // every key masked: softmax of a constant row = 1/N over all N keys (padded and disallowed ones included)
      const float u = 1.0f / static_cast<float>(N);
      for (int jl = lane; jl < nloc; jl += 32) sc[r * NCP + jl] = u, prow[jl] = u;
      if (lane == 0) atomicOr(&s_fm, 1u << r);
    } else {
      float tot = 0.f;
#pragma unroll
      for (int p = 0; p < PC; ++p) tot += xsum[p * PR + r];
      const float inv = 1.0f / tot;
      for (int jl = lane; jl < nloc; jl += 32) {
        const float p = sc[r * NCP + jl] * inv;
        sc[r * NCP + jl] = p;
        prow[jl] = p;
      }
    }
    if (lane == 0 && h == 0 && rank == 0) full_masked[b * R + r] = gmax[i] == -CUDART_INF_F ? 1 : 0;
  }
  __syncthreads();
  // ---- phase 3: partial out[r, c] = sum over this CTA's keys p[r, j] * V[j, c]; thread = (dim c, key group g of 4)
  {
    const int c = tid % DH, g = tid / DH;
    float acc[PR];
#pragma unroll
    for (int r = 0; r < PR; ++r) acc[r] = 0.f;
    const __nv_bfloat16* vcol = kv + (static_cast<long long>(b) * N + j0) * ld + H * DH + h * DH + c;
    const uint32_t fmm = s_fm;  // fully masked rows weigh every key with 1/N
    for (int jb = g; jb < nloc; jb += 32) {  // eight independent loads in flight per thread
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = jb + 4 * u < nloc ? __bfloat162float(vcol[static_cast<long long>(jb + 4 * u) * ld]) : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int jl = jb + 4 * u < nloc ? jb + 4 * u : jb;  // v[u] = 0 beyond the chunk
        const uint32_t rs = skey[jl] | fmm;                  // warp-uniform (one key per warp and step): rows with p != 0
#pragma unroll
        for (int r = 0; r < PR; ++r)
          if ((rs >> r) & 1u) acc[r] += sc[r * NCP + jl] * v[u];
      }
    }
#pragma unroll
    for (int r = 0; r < PR; ++r) part[(g * PR + r) * DH + c] = acc[r];
  }
  __syncthreads();
  for (int i = tid; i < PR * DH; i += 256) {
    const float t = part[i] + part[PR * DH + i] + part[2 * PR * DH + i] + part[3 * PR * DH + i];
    st_peer_f32(&xpart[rank * PR * DH + i], 0, t);
  }
  cl_sync();
  if (rank == 0) {
    for (int i = tid; i < R * DH; i += 256) {
      float t = 0.f;
#pragma unroll
      for (int p = 0; p < PC; ++p) t += xpart[p * PR * DH + i];
      out[(static_cast<long long>(b) * R + i / DH) * H * DH + h * DH + (i % DH)] = t;
    }
  }
}

// Backward of the above in one launch: dP = g V^T, dS = P (dP - rowsum(P dP)) (zero for fully masked rows),
// dK = dS^T q, dV = P^T g (bf16 rows of dkv), dqp += dS K (cluster-reduced, then one atomicAdd per output and sample).
__global__ void __cluster_dims__(PC, 1, 1) __launch_bounds__(256)
pool_bwd_cl_kernel(const float* __restrict__ dout, const float* __restrict__ qp, const __nv_bfloat16* __restrict__ kv,
                   const float* __restrict__ probs, const uint8_t* __restrict__ full_masked,
                   __nv_bfloat16* __restrict__ dkv, float* __restrict__ dqp, int B, int H, int R, int N, int NC, int NCP) {
  extern __shared__ __align__(16) float smc[];
  float* sm = smc;
  float* q = sm;                        // [16][64]
  float* g = q + PR * DH;               // [16][64]
  float* ps = g + PR * DH;              // [16][NCP] probabilities
  float* ds = ps + PR * NCP;            // [16][NCP] dP, then dS
  float* xds = ds + PR * NCP;           // [PC][16] partial row sums of P dP
  float* part = xds + 2 * PC * PR;      // [4][16][64]
  float* xpart = part + 4 * PR * DH;    // [PC][16][64]
  uint32_t* skey = reinterpret_cast<uint32_t*>(xpart);  // [NCP] bit r = P[r, key] != 0 (xpart is only used at the end)
  __shared__ uint8_t fm[PR];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = cl_rank();
  const int cid = blockIdx.x / PC;
  const int h = cid % H, b = cid / H;
  const int ld = 2 * H * DH;
  const int j0 = static_cast<int>(rank) * NC;
  const int nloc = max(0, min(NC, N - j0));
  for (int i = tid; i < PR * DH; i += 256) {
    const int r = i / DH, c = i % DH;
    q[i] = r < R ? qp[r * H * DH + h * DH + c] : 0.f;
    g[i] = r < R ? dout[(static_cast<long long>(b) * R + r) * H * DH + h * DH + c] : 0.f;
  }
  if (tid < PR) fm[tid] = tid < R ? full_masked[b * R + tid] : 1;
  __syncthreads();
  // ---- dP[r, j] = g[r] . V[j], thread = key.  The pool mask is block-structured: a key carries probability in ~2 of
  // the 16 rows, and only those rows (minus the fully masked ones, whose scores get no gradient) need dP.
  for (int jl = tid; jl < nloc; jl += 256) {
    const int j = j0 + jl;
    float v[DH];
    load_row64(kv + (static_cast<long long>(b) * N + j) * ld + H * DH + h * DH, v);
    uint32_t rs = 0u, live = 0u;
#pragma unroll
    for (int r = 0; r < PR; ++r) {
      const float pv = r < R ? probs[(static_cast<long long>(b * H + h) * R + r) * N + j] : 0.f;
      ps[r * NCP + jl] = pv;
      ds[r * NCP + jl] = 0.f;
      if (pv != 0.f) {
        rs |= 1u << r;
        if (fm[r] == 0) live |= 1u << r;
      }
    }
    skey[jl] = rs;
    for (uint32_t m = live; m != 0u; m &= m - 1u) {
      const int r = __ffs(m) - 1;
      const float4* g4 = reinterpret_cast<const float4*>(g + r * DH);
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < DH / 4; ++c) {
        const float4 w = g4[c];
        s += w.x * v[4 * c] + w.y * v[4 * c + 1] + w.z * v[4 * c + 2] + w.w * v[4 * c + 3];
      }
      ds[r * NCP + jl] = s;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = warp + 8 * i;
    float s = 0.f;
    for (int jl = lane; jl < nloc; jl += 32) s += ps[r * NCP + jl] * ds[r * NCP + jl];
    s = warp_sum(s);
    if (lane < PC) st_peer_f32(&xds[rank * PR + r], lane, s);
  }
  cl_sync();
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = warp + 8 * i;
    float tot = 0.f;
#pragma unroll
    for (int p = 0; p < PC; ++p) tot += xds[p * PR + r];
    const bool dead = fm[r] != 0;
    for (int jl = lane; jl < nloc; jl += 32)
      ds[r * NCP + jl] = dead ? 0.f : ps[r * NCP + jl] * (ds[r * NCP + jl] - tot);
  }
  __syncthreads();
  // ---- dK / dV rows: thread = (key jl = t % 64 of a 64-key pass, 16-wide dim group t / 64)
  {
    const int c0 = (tid / 64) * 16;
    for (int jb = 0; jb < nloc; jb += 64) {
      const int jl = jb + (tid % 64);
      if (jl < nloc) {
        float ak[16], av[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) ak[i] = 0.f, av[i] = 0.f;
        const uint32_t rs = skey[jl];
#pragma unroll 2
        for (int r = 0; r < PR; ++r) {
          if (!((rs >> r) & 1u)) continue;  // P = dS = 0 for this (row, key)
          const float s = ds[r * NCP + jl], p = ps[r * NCP + jl];
          const float4* q4 = reinterpret_cast<const float4*>(q + r * DH + c0);
          const float4* g4 = reinterpret_cast<const float4*>(g + r * DH + c0);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 qq = q4[i], gg = g4[i];
            ak[4 * i] += s * qq.x, ak[4 * i + 1] += s * qq.y, ak[4 * i + 2] += s * qq.z, ak[4 * i + 3] += s * qq.w;
            av[4 * i] += p * gg.x, av[4 * i + 1] += p * gg.y, av[4 * i + 2] += p * gg.z, av[4 * i + 3] += p * gg.w;
          }
        }
        __nv_bfloat16* ok = dkv + (static_cast<long long>(b) * N + j0 + jl) * ld + h * DH + c0;
        __nv_bfloat16* ov = ok + H * DH;
        uint4 w0, w1;
        w0.x = pack_bf16x2(ak[0], ak[1]), w0.y = pack_bf16x2(ak[2], ak[3]), w0.z = pack_bf16x2(ak[4], ak[5]);
        w0.w = pack_bf16x2(ak[6], ak[7]);
        w1.x = pack_bf16x2(ak[8], ak[9]), w1.y = pack_bf16x2(ak[10], ak[11]), w1.z = pack_bf16x2(ak[12], ak[13]);
        w1.w = pack_bf16x2(ak[14], ak[15]);
        reinterpret_cast<uint4*>(ok)[0] = w0, reinterpret_cast<uint4*>(ok)[1] = w1;
        w0.x = pack_bf16x2(av[0], av[1]), w0.y = pack_bf16x2(av[2], av[3]), w0.z = pack_bf16x2(av[4], av[5]);
        w0.w = pack_bf16x2(av[6], av[7]);
        w1.x = pack_bf16x2(av[8], av[9]), w1.y = pack_bf16x2(av[10], av[11]), w1.z = pack_bf16x2(av[12], av[13]);
        w1.w = pack_bf16x2(av[14], av[15]);
        reinterpret_cast<uint4*>(ov)[0] = w0, reinterpret_cast<uint4*>(ov)[1] = w1;
      }
    }
  }
  // ---- dqp partial[r, c] = sum over this CTA's keys dS[r, j] * K[j, c]; thread = (dim c, key group of 4)
  {
    const int c = tid % DH, gq = tid / DH;
    float acc[PR];
#pragma unroll
    for (int r = 0; r < PR; ++r) acc[r] = 0.f;
    const __nv_bfloat16* kcol = kv + (static_cast<long long>(b) * N + j0) * ld + h * DH + c;
    for (int jb = gq; jb < nloc; jb += 32) {  // eight independent loads in flight per thread
      float kx[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) kx[u] = jb + 4 * u < nloc ? __bfloat162float(kcol[static_cast<long long>(jb + 4 * u) * ld]) : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int jl = jb + 4 * u < nloc ? jb + 4 * u : jb;  // kx[u] = 0 beyond the chunk
        const uint32_t rs = skey[jl];                        // warp-uniform
#pragma unroll
        for (int r = 0; r < PR; ++r)
          if ((rs >> r) & 1u) acc[r] += ds[r * NCP + jl] * kx[u];
      }
    }
#pragma unroll
    for (int r = 0; r < PR; ++r) part[(gq * PR + r) * DH + c] = acc[r];
  }
  __syncthreads();
  for (int i = tid; i < PR * DH; i += 256) {
    const float t = part[i] + part[PR * DH + i] + part[2 * PR * DH + i] + part[3 * PR * DH + i];
    st_peer_f32(&xpart[rank * PR * DH + i], 0, t);
  }
  cl_sync();
  if (rank == 0) {
    for (int i = tid; i < R * DH; i += 256) {
      float t = 0.f;
#pragma unroll
      for (int p = 0; p < PC; ++p) t += xpart[p * PR * DH + i];
      if (t != 0.f) atomicAdd(dqp + (i / DH) * H * DH + h * DH + (i % DH), t);
    }
  }
}

// dS (written over `probs_to_ds` copy): grid = B*H*R.  dout [B, R, H*64] fp32.
__global__ void __launch_bounds__(256)
pool_bwd_scores_kernel(const float* __restrict__ dout, const __nv_bfloat16* __restrict__ kv,
                       const float* __restrict__ probs, const uint8_t* __restrict__ full_masked,
                       float* __restrict__ ds, int B, int H, int R, int N) {
  extern __shared__ float dp[];  // [N]
  __shared__ float g[DH];
  __shared__ float red[8];
  const int r = blockIdx.x % R, h = (blockIdx.x / R) % H, b = blockIdx.x / (R * H);
  const int ld = 2 * H * DH;
  if (threadIdx.x < DH) g[threadIdx.x] = dout[(static_cast<long long>(b) * R + r) * H * DH + h * DH + threadIdx.x];
  __syncthreads();
  const float* prow = probs + (static_cast<long long>(b * H + h) * R + r) * N;
  float* drow = ds + (static_cast<long long>(b * H + h) * R + r) * N;
  if (full_masked[b * R + r]) {
    for (int j = threadIdx.x; j < N; j += blockDim.x) drow[j] = 0.f;
    return;
  }
  float dsum = 0.f;
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    const float p = prow[j];
    float d = 0.f;
    if (p != 0.f) {
      float v[DH];
      load_row64(kv + (static_cast<long long>(b) * N + j) * ld + H * DH + h * DH, v);
#pragma unroll
      for (int c = 0; c < DH; ++c) d += g[c] * v[c];
    }
    dp[j] = d;
    dsum += p * d;
  }
  dsum = block_reduce(dsum, red, false);
  for (int j = threadIdx.x; j < N; j += blockDim.x) drow[j] = prow[j] * (dp[j] - dsum);
}

// dK[b,j,h,:] = sum_r dS[b,h,r,j] qp[r,h,:] ; dV[b,j,h,:] = sum_r P[b,h,r,j] dout[b,r,h,:]   -> bf16 [B*N, 2*H*64]
// and, fused, this key chunk's share of dqp[r,h,:] = sum_b sum_j dS[b,h,r,j] K[b,j,h,:] (one atomicAdd per output
// per block; dqp must be zeroed by the launcher).
// grid = (ceil(N/64), H, B), block 256: thread = (key j_local = t%64, 16-wide dim group = t/64)
__global__ void __launch_bounds__(256)
pool_bwd_kv_kernel(const float* __restrict__ ds, const float* __restrict__ probs, const float* __restrict__ qp,
                   const float* __restrict__ dout, const __nv_bfloat16* __restrict__ kv, __nv_bfloat16* __restrict__ dkv,
                   float* __restrict__ dqp, int B, int H, int R, int N) {
  extern __shared__ float sm[];  // qs[R][64], gs[R][64], dsT[R][64], kt[64][65]
  float* qs = sm;
  float* gs = qs + R * DH;
  float* dss = gs + R * DH;
  float* kt = dss + R * 64;
  const int h = blockIdx.y, b = blockIdx.z;
  const int ld = 2 * H * DH;
  const int j0 = blockIdx.x * 64;
  for (int i = threadIdx.x; i < R * DH; i += blockDim.x) {
    const int r = i / DH, c = i % DH;
    qs[i] = qp[r * H * DH + h * DH + c];
    gs[i] = dout[(static_cast<long long>(b) * R + r) * H * DH + h * DH + c];
    const int jj = j0 + c;  // c doubles as the key index inside the chunk
    dss[i] = jj < N ? ds[(static_cast<long long>(b * H + h) * R + r) * N + jj] : 0.f;
  }
  for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) {  // K chunk: 64 keys x 64 dims, 16-byte loads
    const int jl = i / 8, q8 = i % 8;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (j0 + jl < N) {
      const uint4 q = *reinterpret_cast<const uint4*>(kv + (static_cast<long long>(b) * N + j0 + jl) * ld + h * DH + q8 * 8);
      v[0] = bf16_lo(q.x), v[1] = bf16_hi(q.x), v[2] = bf16_lo(q.y), v[3] = bf16_hi(q.y);
      v[4] = bf16_lo(q.z), v[5] = bf16_hi(q.z), v[6] = bf16_lo(q.w), v[7] = bf16_hi(q.w);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) kt[jl * 65 + q8 * 8 + e] = v[e];
  }
  __syncthreads();
  // ---- dqp partial: thread = (dim c = t%64, row group t/64 handles rows rg, rg+4, ...)
  {
    const int c = threadIdx.x % 64, rg = threadIdx.x / 64;
    for (int r = rg; r < R; r += 4) {
      float acc = 0.f;
#pragma unroll 8
      for (int jl = 0; jl < 64; ++jl) acc += dss[r * 64 + jl] * kt[jl * 65 + c];
      if (acc != 0.f) atomicAdd(dqp + r * H * DH + h * DH + c, acc);
    }
  }
  const int jl = threadIdx.x % 64;
  const int j = j0 + jl;
  const int c0 = (threadIdx.x / 64) * 16;
  if (j >= N) return;
  float ak[16], av[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) ak[i] = 0.f, av[i] = 0.f;
  for (int r = 0; r < R; ++r) {
    const float s = dss[r * 64 + jl];
    const float p = probs[(static_cast<long long>(b * H + h) * R + r) * N + j];
#pragma unroll
    for (int i = 0; i < 16; ++i) ak[i] += s * qs[r * DH + c0 + i], av[i] += p * gs[r * DH + c0 + i];
  }
  __nv_bfloat16* ok = dkv + (static_cast<long long>(b) * N + j) * ld + h * DH + c0;
  __nv_bfloat16* ov = ok + H * DH;
  uint4 q0, q1;
  q0.x = pack_bf16x2(ak[0], ak[1]), q0.y = pack_bf16x2(ak[2], ak[3]), q0.z = pack_bf16x2(ak[4], ak[5]),
  q0.w = pack_bf16x2(ak[6], ak[7]);
  q1.x = pack_bf16x2(ak[8], ak[9]), q1.y = pack_bf16x2(ak[10], ak[11]), q1.z = pack_bf16x2(ak[12], ak[13]),
  q1.w = pack_bf16x2(ak[14], ak[15]);
  reinterpret_cast<uint4*>(ok)[0] = q0, reinterpret_cast<uint4*>(ok)[1] = q1;
  q0.x = pack_bf16x2(av[0], av[1]), q0.y = pack_bf16x2(av[2], av[3]), q0.z = pack_bf16x2(av[4], av[5]),
  q0.w = pack_bf16x2(av[6], av[7]);
  q1.x = pack_bf16x2(av[8], av[9]), q1.y = pack_bf16x2(av[10], av[11]), q1.z = pack_bf16x2(av[12], av[13]),
  q1.w = pack_bf16x2(av[14], av[15]);
  reinterpret_cast<uint4*>(ov)[0] = q0, reinterpret_cast<uint4*>(ov)[1] = q1;
}

// ---- small fp32 GEMM: C[m,n] = alpha * sum_k A(m,k) B(n,k) (+ C if accumulate) (+ add[m % add_rows, n]); arbitrary
// strides.  32x32 output tile per block, 2x2 outputs per thread, 32-wide k steps staged through shared memory.  These
// problems have a few hundred rows at most and are pure latency, so the k range is split over a thread-block cluster
// (1,1,KS): every CTA reduces its k slice, ships its partial tile to CTA 0 through distributed shared memory and CTA 0
// adds the slices in a fixed order (deterministic) before the epilogue.
constexpr int SG_MAX_KS = 8;

__global__ void __launch_bounds__(256)
small_gemm_kernel(const float* __restrict__ A, long long sam, long long sak, const float* __restrict__ Bm,
                  long long sbn, long long sbk, float* __restrict__ C, long long ldc, const float* __restrict__ add,
                  long long ldadd, int add_rows, int M, int N, int K, float alpha, int accumulate, int k_per) {
  __shared__ float As[32][33], Bs[32][33];  // [k][row]
  __shared__ float xacc[SG_MAX_KS][32 * 32];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const int KS = gridDim.z;
  const uint32_t rank = blockIdx.z;
  const int kbeg = static_cast<int>(rank) * k_per, kend = min(K, kbeg + k_per);
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int k0 = kbeg; k0 < kend; k0 += 32) {
    // element e -> (row = e / 32, k = e % 32) when k is the fast stride, otherwise (k = e / 32, row = e % 32), so the
    // global reads stay coalesced for either layout
    for (int e = threadIdx.x; e < 32 * 32; e += 256) {
      int ra, ka, rb, kb;
      if (sak == 1) ra = e / 32, ka = e % 32; else ka = e / 32, ra = e % 32;
      if (sbk == 1) rb = e / 32, kb = e % 32; else kb = e / 32, rb = e % 32;
      As[ka][ra] = (m0 + ra < M && k0 + ka < kend) ? A[(m0 + ra) * sam + (k0 + ka) * sak] : 0.f;
      Bs[kb][rb] = (n0 + rb < N && k0 + kb < kend) ? Bm[(n0 + rb) * sbn + (k0 + kb) * sbk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float a0 = As[k][ty * 2], a1 = As[k][ty * 2 + 1], b0 = Bs[k][tx * 2], b1 = Bs[k][tx * 2 + 1];
      acc[0][0] += a0 * b0, acc[0][1] += a0 * b1, acc[1][0] += a1 * b0, acc[1][1] += a1 * b1;
    }
    __syncthreads();
  }
  if (KS > 1) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) st_peer_f32(&xacc[rank][(ty * 2 + i) * 32 + tx * 2 + j], 0, acc[i][j]);
    cl_sync();
    if (rank != 0) return;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float t = 0.f;
        for (int z = 0; z < KS; ++z) t += xacc[z][(ty * 2 + i) * 32 + tx * 2 + j];
        acc[i][j] = t;
      }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int m = m0 + ty * 2 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int n = n0 + tx * 2 + j;
      if (n >= N) continue;
      float v = alpha * acc[i][j];
      if (accumulate) v += C[m * ldc + n];
      if (add != nullptr) v += add[(add_rows > 0 ? m % add_rows : m) * ldadd + n];
      C[m * ldc + n] = v;
    }
  }
}

}  // namespace mca

using namespace mca;

// cluster kernels: R <= 16 and the [16][N/PC] score block of one CTA within shared memory
static bool pool_cluster_geometry(int R, int N, int* NC, int* NCP, size_t* bytes) {
  *NC = (N + PC - 1) / PC;
  *NCP = (*NC + 3) & ~3;
  *bytes = pool_cl_smem_floats(*NCP) * sizeof(float);
  return R <= PR && *bytes <= 200 * 1024;
}

extern "C" int mca_pool_attn_fwd(const float* qp, const void* kv, const uint8_t* padding, const uint8_t* keygrp,
                                 const uint32_t* rowbits, float* probs, uint8_t* full_masked, float* out, int B,
                                 int H, int R, int N, void* stream) {
  if (B <= 0 || R <= 0 || N <= 0 || N * 4 > 200 * 1024) return MCA_ERR_SHAPE;
  int NC, NCP;
  size_t bytes;
  if (pool_cluster_geometry(R, N, &NC, &NCP, &bytes)) {
    static bool attr2 = false;
    if (!attr2) {
      if (cudaFuncSetAttribute(pool_fwd_cl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
        return MCA_ERR_CUDA;
      attr2 = true;
    }
    pool_fwd_cl_kernel<<<B * H * PC, 256, bytes, reinterpret_cast<cudaStream_t>(stream)>>>(
        qp, reinterpret_cast<const __nv_bfloat16*>(kv), padding, keygrp, rowbits, probs, full_masked, out, B, H, R, N, NC,
        NCP);
    return check_launch();
  }
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(pool_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
  pool_fwd_kernel<__nv_bfloat16><<<B * H * R, 256, N * sizeof(float), reinterpret_cast<cudaStream_t>(stream)>>>(
      qp, reinterpret_cast<const __nv_bfloat16*>(kv), padding, keygrp, rowbits, probs, full_masked, out, B, H, R, N);
  return check_launch();
}

// fp32-parity mode (exact.cu): the same pooling on fp32 K | V rows [B*N, 2*H*64]
extern "C" int mca_x_pool_attn_fwd_f32(const float* qp, const float* kv32, const uint8_t* padding, const uint8_t* keygrp,
                                       const uint32_t* rowbits, float* probs, uint8_t* full_masked, float* out, int B,
                                       int H, int R, int N, void* stream) {
  if (B <= 0 || R <= 0 || N <= 0 || N * 4 > 200 * 1024) return MCA_ERR_SHAPE;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(pool_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
  pool_fwd_kernel<float><<<B * H * R, 256, N * sizeof(float), reinterpret_cast<cudaStream_t>(stream)>>>(
      qp, kv32, padding, keygrp, rowbits, probs, full_masked, out, B, H, R, N);
  return check_launch();
}

extern "C" int mca_pool_attn_bwd(const float* dout, const float* qp, const void* kv, const float* probs,
                                 const uint8_t* full_masked, float* ds_scratch, void* dkv, float* dqp, int B, int H,
                                 int R, int N, void* stream_) {
  if (B <= 0 || R <= 0 || N <= 0 || N * 4 > 200 * 1024) return MCA_ERR_SHAPE;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const __nv_bfloat16* kvb = reinterpret_cast<const __nv_bfloat16*>(kv);
  if (cudaMemsetAsync(dqp, 0, static_cast<size_t>(R) * H * DH * sizeof(float), stream) != cudaSuccess) return MCA_ERR_CUDA;
  int NC, NCP;
  size_t bytes;
  if (pool_cluster_geometry(R, N, &NC, &NCP, &bytes)) {
    static bool attr2 = false;
    if (!attr2) {
      if (cudaFuncSetAttribute(pool_bwd_cl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
        return MCA_ERR_CUDA;
      attr2 = true;
    }
    pool_bwd_cl_kernel<<<B * H * PC, 256, bytes, stream>>>(dout, qp, kvb, probs, full_masked,
                                                           reinterpret_cast<__nv_bfloat16*>(dkv), dqp, B, H, R, N, NC, NCP);
    return check_launch();
  }
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(pool_bwd_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
  pool_bwd_scores_kernel<<<B * H * R, 256, N * sizeof(float), stream>>>(dout, kvb, probs, full_masked, ds_scratch, B, H,
                                                                        R, N);
  dim3 g2((N + 63) / 64, H, B);
  const size_t sm2 = (3 * static_cast<size_t>(R) * DH + 64 * 65) * sizeof(float);
  if (sm2 > 48 * 1024) return MCA_ERR_SHAPE;  // R <= 42
  pool_bwd_kv_kernel<<<g2, 256, sm2, stream>>>(ds_scratch, probs, qp, dout, kvb, reinterpret_cast<__nv_bfloat16*>(dkv),
                                               dqp, B, H, R, N);
  return check_launch();
}

extern "C" int mca_small_gemm_f32(const float* A, long long sam, long long sak, const float* Bm, long long sbn,
                                  long long sbk, float* C, long long ldc, const float* add, long long ldadd,
                                  int add_rows, int M, int N, int K, float alpha, int accumulate, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0) return MCA_ERR_SHAPE;
  // split k over a cluster until the grid covers the machine (each CTA keeps at least one 32-wide k step)
  const int tiles = ((N + 31) / 32) * ((M + 31) / 32);
  int ks = 1;
  while (ks < SG_MAX_KS && tiles * ks < 2 * num_sms() && K / (2 * ks) >= 32) ks *= 2;
  const int k_per = ((K + ks - 1) / ks + 31) / 32 * 32;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((N + 31) / 32, (M + 31) / 32, ks);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = reinterpret_cast<cudaStream_t>(stream);
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 1, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = ks;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, small_gemm_kernel, A, sam, sak, Bm, sbn, sbk, C, ldc, add, ldadd, add_rows, M, N, K, alpha,
                         accumulate, k_per) != cudaSuccess)
    return MCA_ERR_CUDA;
  return check_launch();
}
