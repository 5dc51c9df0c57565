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

// grid = B*H*R, block 256.  qp [R, H*64] fp32 (already scaled), kv [B*N, 2*H*64] bf16 (K | V)
__global__ void __launch_bounds__(256)
pool_fwd_kernel(const float* __restrict__ qp, const __nv_bfloat16* __restrict__ kv, const uint8_t* __restrict__ padding,
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
    if (p != 0.f) acc += p * __bfloat162float(kv[(static_cast<long long>(b) * N + j) * ld + H * DH + h * DH + c]);
  }
  part[grp][c] = acc;
  __syncthreads();
  if (threadIdx.x < DH)
    out[(static_cast<long long>(b) * R + r) * H * DH + h * DH + threadIdx.x] =
        part[0][threadIdx.x] + part[1][threadIdx.x] + part[2][threadIdx.x] + part[3][threadIdx.x];
}

// Same computation with one block per (sample, head): K and V rows are read ONCE for all R pooled rows (the kernel
// above re-reads them per row).  Scores / probabilities of the whole [N, R] block live in shared memory.
// grid = B*H, block 512; requires R <= 16 and N * round_up(R,4) * 4 bytes of shared memory.
__global__ void __launch_bounds__(512)
pool_fwd_bh_kernel(const float* __restrict__ qp, const __nv_bfloat16* __restrict__ kv, const uint8_t* __restrict__ padding,
                   const uint8_t* __restrict__ keygrp, const uint32_t* __restrict__ rowbits, float* __restrict__ probs,
                   uint8_t* __restrict__ full_masked, float* __restrict__ out, int B, int H, int R, int N, int RS) {
  extern __shared__ float sm[];
  float* q = sm;              // [16][64]
  float* sc = sm + 16 * DH;   // [N][RS]
  const int h = blockIdx.x % H, b = blockIdx.x / H;
  const int ld = 2 * H * DH;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 16 * DH; i += blockDim.x) q[i] = (i / DH) < R ? qp[(i / DH) * H * DH + h * DH + (i % DH)] : 0.f;
  __syncthreads();
  // ---- phase 1: masked scores
  for (int j = tid; j < N; j += blockDim.x) {
    float k[DH];
    load_row64(kv + (static_cast<long long>(b) * N + j) * ld + h * DH, k);
    const bool pad = padding[static_cast<long long>(b) * N + j] != 0;
    const uint32_t kg = keygrp[j];
    for (int r = 0; r < R; ++r) {
      const float4* q4 = reinterpret_cast<const float4*>(q + r * DH);
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < DH / 4; ++c) {
        const float4 w = q4[c];
        s += w.x * k[4 * c] + w.y * k[4 * c + 1] + w.z * k[4 * c + 2] + w.w * k[4 * c + 3];
      }
      const bool ok = !pad && ((rowbits[r] >> kg) & 1u);
      sc[j * RS + r] = ok ? s : -CUDART_INF_F;
    }
    for (int r = R; r < RS; ++r) sc[j * RS + r] = 0.f;
  }
  __syncthreads();
  // ---- phase 2: softmax of row r by warp r (16 warps)
  for (int r = warp; r < R; r += blockDim.x / 32) {
    float mx = -CUDART_INF_F;
    for (int j = lane; j < N; j += 32) mx = fmaxf(mx, sc[j * RS + r]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float* prow = probs + (static_cast<long long>(b * H + h) * R + r) * N;
    if (mx == -CUDART_INF_F) {
      // every key masked: softmax of a constant row = 1/N over all N keys (padded and disallowed ones included)
      const float u = 1.0f / static_cast<float>(N);
      for (int j = lane; j < N; j += 32) sc[j * RS + r] = u, prow[j] = u;
      if (lane == 0 && h == 0) full_masked[b * R + r] = 1;
    } else {
      float se = 0.f;
      for (int j = lane; j < N; j += 32) {
        const float s = sc[j * RS + r];
        const float e = s == -CUDART_INF_F ? 0.f : expf(s - mx);
        sc[j * RS + r] = e;
        se += e;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
      const float inv = 1.0f / se;
      for (int j = lane; j < N; j += 32) {
        const float p = sc[j * RS + r] * inv;
        sc[j * RS + r] = p;
        prow[j] = p;
      }
      if (lane == 0 && h == 0) full_masked[b * R + r] = 0;
    }
  }
  __syncthreads();
  // ---- phase 3: out[r, c] = sum_j p[r, j] * V[j, c]; thread = (dim c, key group g of 8)
  const int c = tid % DH, g = tid / DH;
  float acc[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) acc[r] = 0.f;
  for (int j = g; j < N; j += 8) {
    const float v = __bfloat162float(kv[(static_cast<long long>(b) * N + j) * ld + H * DH + h * DH + c]);
    const float4* pj = reinterpret_cast<const float4*>(sc + j * RS);
#pragma unroll
    for (int r4 = 0; r4 < 4; ++r4)
      if (r4 * 4 < R) {  // rows beyond R inside the last group of four hold padding that is never written out
        const float4 p4 = pj[r4];
        acc[4 * r4] += p4.x * v, acc[4 * r4 + 1] += p4.y * v, acc[4 * r4 + 2] += p4.z * v, acc[4 * r4 + 3] += p4.w * v;
      }
  }
  __syncthreads();  // everyone is done reading sc: reuse it for the cross-group reduction [8][16][64]
  float* part = sc;
#pragma unroll
  for (int r = 0; r < 16; ++r) part[(g * 16 + r) * DH + c] = acc[r];
  __syncthreads();
  for (int i = tid; i < R * DH; i += blockDim.x) {
    const int r = i / DH, cc = i % DH;
    float t = 0.f;
#pragma unroll
    for (int gg = 0; gg < 8; ++gg) t += part[(gg * 16 + r) * DH + cc];
    out[(static_cast<long long>(b) * R + r) * H * DH + h * DH + cc] = t;
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
// strides.  32x32 output tile per block (these problems have at most a few hundred rows: small tiles keep every SM
// busy), 2x2 outputs per thread, 32-wide k steps staged through shared memory.
__global__ void __launch_bounds__(256)
small_gemm_kernel(const float* __restrict__ A, long long sam, long long sak, const float* __restrict__ Bm,
                  long long sbn, long long sbk, float* __restrict__ C, long long ldc, const float* __restrict__ add,
                  long long ldadd, int add_rows, int M, int N, int K, float alpha, int accumulate) {
  __shared__ float As[32][33], Bs[32][33];  // [k][row]
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int k0 = 0; k0 < K; k0 += 32) {
    // element e -> (row = e / 32, k = e % 32) when k is the fast stride, otherwise (k = e / 32, row = e % 32), so the
    // global reads stay coalesced for either layout
    for (int e = threadIdx.x; e < 32 * 32; e += 256) {
      int ra, ka, rb, kb;
      if (sak == 1) ra = e / 32, ka = e % 32; else ka = e / 32, ra = e % 32;
      if (sbk == 1) rb = e / 32, kb = e % 32; else kb = e / 32, rb = e % 32;
      As[ka][ra] = (m0 + ra < M && k0 + ka < K) ? A[(m0 + ra) * sam + (k0 + ka) * sak] : 0.f;
      Bs[kb][rb] = (n0 + rb < N && k0 + kb < K) ? Bm[(n0 + rb) * sbn + (k0 + kb) * sbk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float a0 = As[k][ty * 2], a1 = As[k][ty * 2 + 1], b0 = Bs[k][tx * 2], b1 = Bs[k][tx * 2 + 1];
      acc[0][0] += a0 * b0, acc[0][1] += a0 * b1, acc[1][0] += a1 * b0, acc[1][1] += a1 * b1;
    }
    __syncthreads();
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

extern "C" int mca_pool_attn_fwd(const float* qp, const void* kv, const uint8_t* padding, const uint8_t* keygrp,
                                 const uint32_t* rowbits, float* probs, uint8_t* full_masked, float* out, int B,
                                 int H, int R, int N, void* stream) {
  if (B <= 0 || R <= 0 || N <= 0 || N * 4 > 200 * 1024) return MCA_ERR_SHAPE;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(pool_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(pool_bwd_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
  const int RS = (R + 3) / 4 * 4;
  const size_t need = (16 * DH + static_cast<size_t>(N) * RS) * sizeof(float);
  if (R <= 16 && need <= 200 * 1024 && static_cast<size_t>(N) * RS >= 8 * 16 * DH) {
    static bool attr2 = false;
    if (!attr2) {
      cudaFuncSetAttribute(pool_fwd_bh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      attr2 = true;
    }
    pool_fwd_bh_kernel<<<B * H, 512, need, reinterpret_cast<cudaStream_t>(stream)>>>(
        qp, reinterpret_cast<const __nv_bfloat16*>(kv), padding, keygrp, rowbits, probs, full_masked, out, B, H, R, N, RS);
    return check_launch();
  }
  pool_fwd_kernel<<<B * H * R, 256, N * sizeof(float), reinterpret_cast<cudaStream_t>(stream)>>>(
      qp, reinterpret_cast<const __nv_bfloat16*>(kv), padding, keygrp, rowbits, probs, full_masked, out, B, H, R, N);
  return check_launch();
}

extern "C" int mca_pool_attn_bwd(const float* dout, const float* qp, const void* kv, const float* probs,
                                 const uint8_t* full_masked, float* ds_scratch, void* dkv, float* dqp, int B, int H,
                                 int R, int N, void* stream_) {
  if (B <= 0 || R <= 0 || N <= 0 || N * 4 > 200 * 1024) return MCA_ERR_SHAPE;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(pool_bwd_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
  const __nv_bfloat16* kvb = reinterpret_cast<const __nv_bfloat16*>(kv);
  pool_bwd_scores_kernel<<<B * H * R, 256, N * sizeof(float), stream>>>(dout, kvb, probs, full_masked, ds_scratch, B, H,
                                                                        R, N);
  dim3 g2((N + 63) / 64, H, B);
  if (cudaMemsetAsync(dqp, 0, static_cast<size_t>(R) * H * DH * sizeof(float), stream) != cudaSuccess) return MCA_ERR_CUDA;
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
  dim3 grid((N + 31) / 32, (M + 31) / 32);
  small_gemm_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(A, sam, sak, Bm, sbn, sbk, C, ldc, add,
                                                                             ldadd, add_rows, M, N, K, alpha, accumulate);
  return check_launch();
}
