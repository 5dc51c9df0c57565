// TabularEncoder pieces (TCGA configs): encoders.py:25-37 (nn.Embedding(max_norm=1.0) renormalises looked-up rows
// in place on every forward), encoders.py:55-72 (ContinuousValueEncoder: pad test on the raw value, clamp(max),
// Linear(1,d) -> ReLU; the d x d Linear runs on the tcgen05 GEMM, its LayerNorm + pad-zero + embedding add on
// ln512_fwd).  All bandwidth-bound elementwise work.
#include "mca_b200.h"
#include "ptx.cuh"
#include "runtime.h"

namespace mca {

// rows with ||row|| > max_norm are scaled by max_norm / (||row|| + 1e-7)  (torch embedding_renorm_)
__global__ void __launch_bounds__(256)
embedding_renorm_kernel(float* __restrict__ emb, int rows, int d, float max_norm) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  float* e = emb + static_cast<long long>(r) * d;
  float ss = 0.f;
  for (int c = lane; c < d; c += 32) ss += e[c] * e[c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float norm = sqrtf(ss);
  if (norm > max_norm) {
    const float s = max_norm / (norm + 1e-7f);
    for (int c = lane; c < d; c += 32) e[c] *= s;
  }
}

// h1[r, c] = relu(w1[c] * min(v[r], max_value) + b1[c]) (bf16 GEMM operand); vpad[r] = (v[r] == padding_value)
__global__ void __launch_bounds__(256)
tabular_fwd_kernel(const float* __restrict__ values, const float* __restrict__ w1, const float* __restrict__ b1,
                   __nv_bfloat16* __restrict__ h1, uint8_t* __restrict__ vpad, float max_value, float padding_value,
                   int d, long long rows) {
  const long long n = rows * (d / 2);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / (d / 2);
    const int c = static_cast<int>(i % (d / 2)) * 2;
    const float v = values[r];
    if (c == 0) vpad[r] = v == padding_value ? 1 : 0;
    const float vc = fminf(v, max_value);
    const float a = fmaxf(w1[c] * vc + b1[c], 0.f), b = fmaxf(w1[c + 1] * vc + b1[c + 1], 0.f);
    *reinterpret_cast<uint32_t*>(h1 + r * d + c) = pack_bf16x2(a, b);
  }
}

// dw1[c] += sum_r dh1[r,c] * [pre > 0] * vc ; db1[c] += sum_r dh1[r,c] * [pre > 0]
__global__ void __launch_bounds__(256)
tabular_bwd_kernel(const float* __restrict__ dh1, const float* __restrict__ values, const float* __restrict__ w1,
                   const float* __restrict__ b1, float* __restrict__ dw1, float* __restrict__ db1, float max_value,
                   int d, long long rows, int rows_per_block) {
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    const float w = w1[c], b = b1[c];
    float aw = 0.f, ab = 0.f;
    for (long long r = r0; r < r1; ++r) {
      const float vc = fminf(values[r], max_value);
      if (w * vc + b > 0.f) {
        const float g = dh1[r * d + c];
        aw += g * vc;
        ab += g;
      }
    }
    atomicAdd(dw1 + c, aw);
    atomicAdd(db1 + c, ab);
  }
}

// ---- index-driven embedding tables (SequenceEncoder encoders.py:145-166, SparseTabularEncoder :100-120)
// nn.Embedding(max_norm=1.0) renormalises only the rows that are LOOKED UP, once each: mark, then renormalise the
// marked rows (a row referenced by many tokens must not be scaled twice) and clear the marks.
__global__ void __launch_bounds__(256)
mark_rows_kernel(const long long* __restrict__ idx, long long n, int rows, uint8_t* __restrict__ flags, int* __restrict__ bad) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long v = idx[i];
    if (v < 0 || v >= rows) atomicOr(bad, 2);  // torch raises IndexError; reported through the step's error flag
    else flags[v] = 1;
  }
}
__global__ void __launch_bounds__(256)
renorm_marked_kernel(float* __restrict__ emb, int rows, int d, float max_norm, uint8_t* __restrict__ flags) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows || flags[r] == 0) return;
  float* e = emb + static_cast<long long>(r) * d;
  float ss = 0.f;
  for (int c = lane; c < d; c += 32) ss += e[c] * e[c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float norm = sqrtf(ss);
  if (norm > max_norm) {
    const float s = max_norm / (norm + 1e-7f);
    for (int c = lane; c < d; c += 32) e[c] *= s;
  }
  __syncwarp();
  if (lane == 0) flags[r] = 0;
}

// dst[b * dst_rows_per_b + dst_row_off + l, :] (+)= emb[idx[b, l], :] (+ pe[l, :]); one warp per token, float4 lanes
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ emb, const long long* __restrict__ idx, int rows_emb, int B, int L, int d,
                   const float* __restrict__ pe, float* __restrict__ dst, int dst_rows_per_b, int dst_row_off,
                   int accumulate) {
  const int lane = threadIdx.x & 31;
  const long long t = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (t >= static_cast<long long>(B) * L) return;
  const int b = static_cast<int>(t / L), l = static_cast<int>(t % L);
  long long v = idx[t];
  v = v < 0 ? 0 : (v >= rows_emb ? rows_emb - 1 : v);
  const float4* src = reinterpret_cast<const float4*>(emb + v * d);
  float4* out = reinterpret_cast<float4*>(dst + (static_cast<long long>(b) * dst_rows_per_b + dst_row_off + l) * d);
  for (int c = lane; c < d / 4; c += 32) {
    float4 q = src[c];
    if (pe != nullptr) {
      const float4 p = reinterpret_cast<const float4*>(pe + static_cast<long long>(l) * d)[c];
      q.x += p.x, q.y += p.y, q.z += p.z, q.w += p.w;
    }
    if (accumulate) {
      const float4 o = out[c];
      q.x += o.x, q.y += o.y, q.z += o.z, q.w += o.w;
    }
    out[c] = q;
  }
}

// demb[idx[b, l], :] += dsrc[b * src_rows_per_b + src_row_off + l, :] unless idx == skip_row (padding_idx: no gradient)
__global__ void __launch_bounds__(256)
scatter_add_rows_kernel(const float* __restrict__ dsrc, const long long* __restrict__ idx, int rows_emb, int B, int L,
                        int d, int src_rows_per_b, int src_row_off, int skip_row, float* __restrict__ demb) {
  const int lane = threadIdx.x & 31;
  const long long t = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (t >= static_cast<long long>(B) * L) return;
  const int b = static_cast<int>(t / L), l = static_cast<int>(t % L);
  const long long v = idx[t];
  if (v < 0 || v >= rows_emb || v == skip_row) return;
  const float* src = dsrc + (static_cast<long long>(b) * src_rows_per_b + src_row_off + l) * d;
  float* out = demb + v * d;
  for (int c = lane; c < d; c += 32) atomicAdd(out + c, src[c]);
}

// ---- PatchEncoder, "matrix" mode (encoders.py:217-274): 'b (h p1) (w p2) -> b (h w) (p1 p2)' and the pad mask
// all(patch == pad_token); one warp per patch
__global__ void __launch_bounds__(256)
patchify_kernel(const float* __restrict__ values, int B, int Hh, int Ww, int p1, int p2, float pad_token,
                float* __restrict__ tokens, uint8_t* __restrict__ mask) {
  const int lane = threadIdx.x & 31;
  const int nh = Hh / p1, nw = Ww / p2, in = p1 * p2;
  const long long t = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (t >= static_cast<long long>(B) * nh * nw) return;
  const int b = static_cast<int>(t / (nh * nw)), hw = static_cast<int>(t % (nh * nw));
  const int ph = hw / nw, pw = hw % nw;
  bool all_pad = true;
  for (int e = lane; e < in; e += 32) {
    const int i = e / p2, j = e % p2;
    const float v = values[(static_cast<long long>(b) * Hh + ph * p1 + i) * Ww + pw * p2 + j];
    tokens[t * in + e] = v;
    all_pad &= v == pad_token;
  }
  all_pad = __all_sync(0xffffffffu, all_pad);
  if (lane == 0) mask[t] = all_pad ? 1 : 0;
}

// nn.Dropout on token rows: keep with probability 1 - p, scale kept values by 1 / (1 - p).  The keep decision is a
// counter-based hash of (seed, step counter read from device memory, element), so the backward regenerates the very
// same mask from the same counter instead of storing it, and a captured graph draws a new mask at every replay.
__device__ __forceinline__ uint32_t mix32(uint64_t x) {
  x ^= x >> 33, x *= 0xff51afd7ed558ccdull, x ^= x >> 33, x *= 0xc4ceb9fe1a85ec53ull, x ^= x >> 33;
  return static_cast<uint32_t>(x);
}
__global__ void __launch_bounds__(256)
dropout_rows_kernel(float* __restrict__ x, int B, int L, int d, int rows_per_b, int row_off, float p,
                    unsigned long long seed, const long long* __restrict__ counter) {
  const unsigned long long ctr = static_cast<unsigned long long>(*counter);
  const float scale = 1.0f / (1.0f - p);
  const uint32_t thresh = static_cast<uint32_t>(static_cast<double>(p) * 4294967296.0);
  const long long n = static_cast<long long>(B) * L * d;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long tok = i / d;
    const int c = static_cast<int>(i % d);
    const long long row = (tok / L) * rows_per_b + row_off + (tok % L);
    const uint32_t h = mix32((seed * 0x9E3779B97F4A7C15ull) ^ (ctr * 0xD1B54A32D192ED03ull) ^ static_cast<uint64_t>(i));
    float* e = x + row * d + c;
    *e = h < thresh ? 0.f : *e * scale;
  }
}

}  // namespace mca

using namespace mca;

extern "C" int mca_embedding_renorm_indexed(float* emb, const long long* idx, long long n_idx, int rows, int d,
                                            float max_norm, uint8_t* flags_scratch, int* bad_index_flag, void* stream_) {
  if (rows <= 0 || d <= 0 || n_idx <= 0) return MCA_ERR_SHAPE;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const unsigned g1 = static_cast<unsigned>((n_idx + 255) / 256 < 1184 ? (n_idx + 255) / 256 : 1184);
  mark_rows_kernel<<<g1, 256, 0, stream>>>(idx, n_idx, rows, flags_scratch, bad_index_flag);
  renorm_marked_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(emb, rows, d, max_norm, flags_scratch);
  return check_launch();
}

extern "C" int mca_embedding_gather(const float* emb, const long long* idx, int rows_emb, int B, int L, int d,
                                    const float* pe, float* dst, int dst_rows_per_b, int dst_row_off, int accumulate,
                                    void* stream_) {
  if (rows_emb <= 0 || B <= 0 || L <= 0 || (d % 4) != 0) return MCA_ERR_SHAPE;
  const long long n = static_cast<long long>(B) * L;
  gather_rows_kernel<<<static_cast<unsigned>((n + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      emb, idx, rows_emb, B, L, d, pe, dst, dst_rows_per_b, dst_row_off, accumulate);
  return check_launch();
}

extern "C" int mca_embedding_scatter_add(const float* dsrc, const long long* idx, int rows_emb, int B, int L, int d,
                                         int src_rows_per_b, int src_row_off, int skip_row, float* demb, void* stream_) {
  if (rows_emb <= 0 || B <= 0 || L <= 0 || d <= 0) return MCA_ERR_SHAPE;
  const long long n = static_cast<long long>(B) * L;
  scatter_add_rows_kernel<<<static_cast<unsigned>((n + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      dsrc, idx, rows_emb, B, L, d, src_rows_per_b, src_row_off, skip_row, demb);
  return check_launch();
}

extern "C" int mca_patchify(const float* values, int B, int H, int W, int p1, int p2, float pad_token, float* tokens,
                            uint8_t* mask, void* stream_) {
  if (B <= 0 || p1 <= 0 || p2 <= 0 || H % p1 != 0 || W % p2 != 0) return MCA_ERR_SHAPE;
  const long long n = static_cast<long long>(B) * (H / p1) * (W / p2);
  patchify_kernel<<<static_cast<unsigned>((n + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      values, B, H, W, p1, p2, pad_token, tokens, mask);
  return check_launch();
}

extern "C" int mca_dropout_rows(float* x, int B, int L, int d, int rows_per_b, int row_off, float p,
                                unsigned long long seed, const long long* counter_dev, void* stream_) {
  if (B <= 0 || L <= 0 || d <= 0 || !(p >= 0.f && p < 1.f)) return MCA_ERR_SHAPE;
  if (p == 0.f) return MCA_OK;
  dropout_rows_kernel<<<148 * 4, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(x, B, L, d, rows_per_b, row_off, p,
                                                                                    seed, counter_dev);
  return check_launch();
}

extern "C" int mca_embedding_renorm(float* emb, int rows, int d, float max_norm, void* stream) {
  if (rows <= 0 || d <= 0) return MCA_ERR_SHAPE;
  embedding_renorm_kernel<<<(rows + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(emb, rows, d, max_norm);
  return check_launch();
}

extern "C" int mca_tabular_fwd(const float* values, const float* w1, const float* b1, void* h1_bf16, uint8_t* vpad,
                               float max_value, float padding_value, int d, long long rows, void* stream) {
  if (rows <= 0 || (d % 2) != 0) return MCA_ERR_SHAPE;
  tabular_fwd_kernel<<<148 * 4, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      values, w1, b1, reinterpret_cast<__nv_bfloat16*>(h1_bf16), vpad, max_value, padding_value, d, rows);
  return check_launch();
}

extern "C" int mca_tabular_bwd(const float* dh1, const float* values, const float* w1, const float* b1, float* dw1,
                               float* db1, const void* unused0, const void* unused1, float max_value,
                               float padding_value, int d, long long rows, void* stream) {
  (void)unused0, (void)unused1, (void)padding_value;
  if (rows <= 0) return MCA_ERR_SHAPE;
  const int rpb = 64;
  tabular_bwd_kernel<<<static_cast<unsigned>((rows + rpb - 1) / rpb), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      dh1, values, w1, b1, dw1, db1, max_value, d, rows, rpb);
  return check_launch();
}
