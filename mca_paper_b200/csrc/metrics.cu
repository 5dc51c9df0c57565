// Embedding-space evaluation metrics on the device: alignment / uniformity of Wang & Isola as the
// reference computes them (utils/metrics.py:20-29: F.normalize, (x - y).norm(dim=1).pow(alpha).mean(),
// torch.pdist(x).pow(2).mul(-t).exp().mean().log()) and the retrieval ranks of utils/metrics.py:73-99 (cosine of every
// masked embedding against all targets, rank = number of targets scoring above the sample's own).  The O(M^2 D) pair
// work runs as 128x128 fp32 pair tiles on the CUDA cores (eval-only, exact fp32 differences rather than a Gram-matrix
// trick: uniformity of nearly collapsed embeddings is all cancellation); sums are reduced through per-block partials
// in a fixed order, so results are bit-reproducible.
#include "mca_b200.h"
#include "ptx.cuh"
#include "runtime.h"

namespace mca {

constexpr int MT_TILE = 128, MT_K = 16, MT_STRIDE = 132;

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// inv[i] = 1 / max(||x_i||_2, eps)   (F.normalize eps 1e-12, nn.CosineSimilarity eps 1e-8); one warp per row
__global__ void __launch_bounds__(256)
row_inv_norm_kernel(const float* __restrict__ x, long long M, int D, float eps, float* __restrict__ inv) {
  const int lane = threadIdx.x & 31;
  const long long r = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= M) return;
  const float* p = x + r * D;
  float ss = 0.f;
  for (int c = lane; c < D; c += 32) ss += p[c] * p[c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (lane == 0) inv[r] = 1.f / fmaxf(sqrtf(ss), eps);
}

__global__ void __launch_bounds__(256) fill_f32_kernel(float* __restrict__ p, long long n, float v) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i < n) p[i] = v;
}

// partial[block] = sum over the block's 8 rows of ||x_i' - y_i'||^alpha, x' = x / max(||x||, 1e-12) when norm
__global__ void __launch_bounds__(256)
alignment_kernel(const float* __restrict__ x, const float* __restrict__ y, long long M, int D, float alpha, int norm,
                 double* __restrict__ partial) {
  __shared__ double s_part[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long r = static_cast<long long>(blockIdx.x) * 8 + warp;
  double val = 0.0;
  if (r < M) {
    const float* px = x + r * D;
    const float* py = y + r * D;
    float sx = 1.f, sy = 1.f;
    if (norm) {
      float ax = 0.f, ay = 0.f;
      for (int c = lane; c < D; c += 32) ax += px[c] * px[c], ay += py[c] * py[c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ax += __shfl_xor_sync(0xffffffffu, ax, o), ay += __shfl_xor_sync(0xffffffffu, ay, o);
      sx = 1.f / fmaxf(sqrtf(ax), 1e-12f), sy = 1.f / fmaxf(sqrtf(ay), 1e-12f);
    }
    float ss = 0.f;
    for (int c = lane; c < D; c += 32) {
      const float d = __fmul_rn(px[c], sx) - __fmul_rn(py[c], sy);  // normalise, then subtract (no fma contraction)
      ss += d * d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float dist = sqrtf(ss);
    val = alpha == 2.f ? static_cast<double>(dist * dist) : static_cast<double>(powf(dist, alpha));
  }
  if (lane == 0) s_part[warp] = val;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += s_part[w];
    partial[blockIdx.x] = s;
  }
}

// out = take_log ? log(sum(partial) / denom) : sum(partial) / denom   (denom == 0 -> NaN, like mean() of an empty tensor)
__global__ void __launch_bounds__(256)
metric_finalize_kernel(const double* __restrict__ partial, long long n, double denom, int take_log, float* __restrict__ out) {
  __shared__ double s_w[8];
  double s = 0.0;
  for (long long i = threadIdx.x; i < n; i += 256) s += partial[i];
  s = warp_sum_f64(s);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += s_w[w];
    const double mean = denom > 0.0 ? tot / denom : __longlong_as_double(0x7ff8000000000000LL);
    out[0] = static_cast<float>(take_log ? log(mean) : mean);
  }
}

// own[i] = <e_i * inv_e[i], t_idx[i] * inv_t[idx[i]]>; one warp per row
__global__ void __launch_bounds__(256)
own_cosine_kernel(const float* __restrict__ emb, const float* __restrict__ inv_e, const float* __restrict__ tgt,
                  const float* __restrict__ inv_t, const long long* __restrict__ idx, long long M, long long T, int D,
                  float* __restrict__ own) {
  const int lane = threadIdx.x & 31;
  const long long r = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= M) return;
  const long long j = idx[r];
  float s = __int_as_float(0x7fc00000);
  if (j >= 0 && j < T) {
    const float* pe = emb + r * D;
    const float* pt = tgt + j * D;
    const float se = inv_e[r], st = inv_t[j];
    s = 0.f;
    for (int c = lane; c < D; c += 32) s += (pe[c] * se) * (pt[c] * st);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  }
  if (lane == 0) own[r] = s;
}

// One 128 x 128 tile of (row of A, row of B) pairs per block, rows pre-scaled by inv while staged; 256 threads, 8 x 8
// pairs each (two 4-row groups x two 4-column groups, so every shared-memory read is a 128-bit load), D walked in chunks
// of 16 through shared memory with the next chunk's global loads in flight during the FMAs.
//   MODE 0 (uniformity): A == B; partial[block] = sum over pairs i < j of exp(-t * ||a_i - a_j||^2); tiles below the
//                        diagonal exit at once.
//   MODE 1 (ranks):      ranks[i] += #{ j != idx[i] : <a_i, b_j> > own[i] }
template <int MODE>
__global__ void __launch_bounds__(256, 2)
pair_tile_kernel(const float* __restrict__ A, const float* __restrict__ invA, long long MA, const float* __restrict__ Bm,
                 const float* __restrict__ invB, long long MB, int D, float t, double* __restrict__ partial,
                 const float* __restrict__ own, const long long* __restrict__ idx, unsigned long long* __restrict__ ranks) {
  __shared__ __align__(16) float As[MT_K][MT_STRIDE], Bs[MT_K][MT_STRIDE];
  __shared__ double s_w[8];
  const int ti = blockIdx.y, tj = blockIdx.x;
  if (MODE == 0 && tj < ti) {
    if (threadIdx.x == 0) partial[static_cast<long long>(blockIdx.y) * gridDim.x + blockIdx.x] = 0.0;
    return;
  }
  const long long i0 = static_cast<long long>(ti) * MT_TILE, j0 = static_cast<long long>(tj) * MT_TILE;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int lk = threadIdx.x & 15, lr = threadIdx.x >> 4;   // staging: column lk of rows lr, lr + 16, ...
  float acc[8][8];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;

  float pa[8], pb[8];   // the staged chunk (scaled by the row's inv, an L1 hit after the first chunk)
  auto load_chunk = [&](int k0) {
    const int k = k0 + lk;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const long long ia = i0 + lr + 16 * e, ib = j0 + lr + 16 * e;
      pa[e] = (ia < MA && k < D) ? __fmul_rn(A[ia * D + k], invA[ia]) : 0.f;
      pb[e] = (ib < MB && k < D) ? __fmul_rn(Bm[ib * D + k], invB[ib]) : 0.f;
    }
  };
  auto store_chunk = [&]() {
#pragma unroll
    for (int e = 0; e < 8; ++e) As[lk][lr + 16 * e] = pa[e], Bs[lk][lr + 16 * e] = pb[e];
  };
  load_chunk(0);
  store_chunk();
  __syncthreads();
  for (int k0 = 0; k0 < D; k0 += MT_K) {
    const bool more = k0 + MT_K < D;
    if (more) load_chunk(k0 + MT_K);
#pragma unroll
    for (int kk = 0; kk < MT_K; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          if (MODE == 0) {
            const float d = a[r] - b[c];
            acc[r][c] = fmaf(d, d, acc[r][c]);
          } else {
            acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
          }
        }
    }
    __syncthreads();
    if (more) {
      store_chunk();
      __syncthreads();
    }
  }

  // pair (r, c) of this thread = rows i0 + (r < 4 ? 0 : 64) + ty*4 + (r & 3), columns j0 + (c < 4 ? 0 : 64) + tx*4 + (c & 3)
  if (MODE == 0) {
    double s = 0.0;
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const long long i = i0 + (r >> 2) * 64 + ty * 4 + (r & 3), j = j0 + (c >> 2) * 64 + tx * 4 + (c & 3);
        if (i < j && j < MA) s += static_cast<double>(expf(-t * acc[r][c]));
      }
    s = warp_sum_f64(s);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) tot += s_w[w];
      partial[static_cast<long long>(blockIdx.y) * gridDim.x + blockIdx.x] = tot;
    }
  } else {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const long long i = i0 + (r >> 2) * 64 + ty * 4 + (r & 3);
      int cnt = 0;
      if (i < MA) {
        const float o = own[i];
        const long long self = idx[i];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const long long j = j0 + (c >> 2) * 64 + tx * 4 + (c & 3);
          cnt += (j < MB && j != self && acc[r][c] > o) ? 1 : 0;
        }
      }
#pragma unroll
      for (int o2 = 8; o2 > 0; o2 >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o2);  // the 16 tx lanes of a row group
      if (tx == 0 && i < MA && cnt) atomicAdd(&ranks[i], static_cast<unsigned long long>(cnt));
    }
  }
}

}  // namespace mca

using namespace mca;

extern "C" int mca_row_inv_norms(const float* x, long long M, int D, float eps, float* inv, void* stream) {
  if (M <= 0 || D <= 0) return MCA_ERR_SHAPE;
  row_inv_norm_kernel<<<static_cast<unsigned>((M + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, M, D, eps, inv);
  return check_launch();
}

extern "C" long long mca_metric_scratch_doubles(long long M) {
  const long long t = (M + MT_TILE - 1) / MT_TILE, a = (M + 7) / 8;
  return t * t > a ? t * t : a;
}

extern "C" int mca_alignment(const float* x, const float* y, long long M, int D, float alpha, int norm, double* scratch,
                             float* out, void* stream_) {
  if (M < 0 || D <= 0) return MCA_ERR_SHAPE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  const long long nb = (M + 7) / 8;
  if (M > 0) alignment_kernel<<<static_cast<unsigned>(nb), 256, 0, st>>>(x, y, M, D, alpha, norm, scratch);
  metric_finalize_kernel<<<1, 256, 0, st>>>(scratch, nb, static_cast<double>(M), 0, out);
  return check_launch();
}

extern "C" int mca_uniformity(const float* x, long long M, int D, float t, int norm, float* inv_scratch, double* scratch,
                              float* out, void* stream_) {
  if (M < 0 || D <= 0) return MCA_ERR_SHAPE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  const long long tiles = (M + MT_TILE - 1) / MT_TILE;
  if (tiles > 65535) return MCA_ERR_SHAPE;
  if (M > 0) {
    if (norm) {
      row_inv_norm_kernel<<<static_cast<unsigned>((M + 7) / 8), 256, 0, st>>>(x, M, D, 1e-12f, inv_scratch);
    } else {
      fill_f32_kernel<<<static_cast<unsigned>((M + 255) / 256), 256, 0, st>>>(inv_scratch, M, 1.f);
    }
    pair_tile_kernel<0><<<dim3(static_cast<unsigned>(tiles), static_cast<unsigned>(tiles)), 256, 0, st>>>(
        x, inv_scratch, M, x, inv_scratch, M, D, t, scratch, nullptr, nullptr, nullptr);
  }
  metric_finalize_kernel<<<1, 256, 0, st>>>(scratch, tiles * tiles, 0.5 * static_cast<double>(M) * static_cast<double>(M - 1), 1, out);
  return check_launch();
}

extern "C" int mca_retrieval_ranks(const float* emb, const float* targets, const long long* idx, long long M, long long T,
                                   int D, float* inv_e, float* inv_t, float* own, long long* ranks, void* stream_) {
  if (M <= 0 || T <= 0 || D <= 0) return MCA_ERR_SHAPE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  const long long ti = (M + MT_TILE - 1) / MT_TILE, tj = (T + MT_TILE - 1) / MT_TILE;
  if (ti > 65535) return MCA_ERR_SHAPE;
  row_inv_norm_kernel<<<static_cast<unsigned>((M + 7) / 8), 256, 0, st>>>(emb, M, D, 1e-8f, inv_e);
  row_inv_norm_kernel<<<static_cast<unsigned>((T + 7) / 8), 256, 0, st>>>(targets, T, D, 1e-8f, inv_t);
  own_cosine_kernel<<<static_cast<unsigned>((M + 7) / 8), 256, 0, st>>>(emb, inv_e, targets, inv_t, idx, M, T, D, own);
  if (cudaMemsetAsync(ranks, 0, sizeof(long long) * M, st) != cudaSuccess) return MCA_ERR_CUDA;
  pair_tile_kernel<1><<<dim3(static_cast<unsigned>(tj), static_cast<unsigned>(ti)), 256, 0, st>>>(
      emb, inv_e, M, targets, inv_t, T, D, 0.f, nullptr, own, idx, reinterpret_cast<unsigned long long*>(ranks));
  return check_launch();
}
