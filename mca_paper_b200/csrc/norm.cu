// Bandwidth-bound LayerNorm kernels (HBM roofline; 2*rows*d*4 algorithmic bytes per call).
//   ln512_fwd / ln512_bwd : d = 512, one warp per row, float4 loads (each lane owns 16 columns).
//       Replaces model.py:24-31 (LayerNorm with learnable gamma, constant beta) and, with the optional
//       pad / positional-table / row-scatter arguments, the tail of EmbeddedSequenceEncoder.forward
//       (encoders.py:192,205,208-209: LN(512) -> zero padded rows -> + sinusoidal PE, written straight into the
//       packed [B,N,512] token buffer so model.py:464 `pack` never copies).
//   lnw_fwd / lnw_param_bwd : arbitrary width (the encoders' input LayerNorm, encoders.py:189), emits the
//       zero-padded bf16 GEMM operand.
#include "mca_b200.h"
#include "ptx.cuh"
#include "runtime.h"

namespace mca {

constexpr int LN_D = 512;
constexpr float LN_EPS = 1e-5f;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// row r of the input maps to output row (r / seg_len) * out_stride_rows + out_row_off + (r % seg_len)
struct RowMap {
  int seg_len;          // rows per sample in the input (L); 0 = identity
  int out_rows_per_b;   // N
  int out_row_off;      // modality offset inside a sample
  __device__ __forceinline__ long long operator()(long long r) const {
    if (seg_len == 0) return r;
    return (r / seg_len) * out_rows_per_b + out_row_off + (r % seg_len);
  }
};

__global__ void __launch_bounds__(256)
ln512_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 float* __restrict__ y32, __nv_bfloat16* __restrict__ y16, float2* __restrict__ stats,
                 const uint8_t* __restrict__ pad, const float* __restrict__ pe, RowMap map, long long rows) {
  pdl_launch_dependents();
  pdl_wait();  // launched with programmatic serialization: resident early, starts when the predecessor has completed
  const int lane = threadIdx.x & 31;
  const long long r = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const long long ro = map(r);
  const bool padded = pad != nullptr && pad[r] != 0;
  float v[16];
  const float4* x4 = reinterpret_cast<const float4*>(x + r * LN_D);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 q = x4[lane + 32 * i];
    v[4 * i] = q.x, v[4 * i + 1] = q.y, v[4 * i + 2] = q.z, v[4 * i + 3] = q.w;
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  const float mean = warp_sum(s) * (1.0f / LN_D);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float d = v[i] - mean;
    ss += d * d;
  }
  const float rstd = rsqrtf(warp_sum(ss) * (1.0f / LN_D) + LN_EPS);
  if (stats != nullptr && lane == 0) stats[r] = make_float2(mean, rstd);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = (lane + 32 * i) * 4;
    const float4 g = *reinterpret_cast<const float4*>(gamma + c);
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (beta != nullptr) b = *reinterpret_cast<const float4*>(beta + c);
    float4 o;
    o.x = (v[4 * i] - mean) * rstd * g.x + b.x;
    o.y = (v[4 * i + 1] - mean) * rstd * g.y + b.y;
    o.z = (v[4 * i + 2] - mean) * rstd * g.z + b.z;
    o.w = (v[4 * i + 3] - mean) * rstd * g.w + b.w;
    if (padded) o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pe != nullptr) {
      const float4 p = *reinterpret_cast<const float4*>(pe + (r % map.seg_len) * LN_D + c);
      o.x += p.x, o.y += p.y, o.z += p.z, o.w += p.w;
    }
    if (y32 != nullptr) *reinterpret_cast<float4*>(y32 + ro * LN_D + c) = o;
    if (y16 != nullptr) {
      uint2 h;
      h.x = pack_bf16x2(o.x, o.y);
      h.y = pack_bf16x2(o.z, o.w);
      *reinterpret_cast<uint2*>(y16 + ro * LN_D + c) = h;
    }
  }
}

// Fused "residual add + LayerNorm" (model.py:118-122 wiring, reference quirk Q1: the residual branch starts from the
// NORMED tensor).  v = LN_prev(xprev) + y  is the new residual-stream value: it is written as fp32 (needed by the
// backward) and normalised again for the next GEMM operand.  LN_prev is recomputed from xprev and its saved
// statistics, so the normed tensor itself never goes to HBM in fp32.
__global__ void __launch_bounds__(256)
add_ln512_fwd_kernel(const float* __restrict__ xprev, const float2* __restrict__ stats_prev,
                     const float* __restrict__ gamma_prev, const float* __restrict__ beta_prev,
                     const __nv_bfloat16* __restrict__ y16, float* __restrict__ xnew, const float* __restrict__ gamma,
                     const float* __restrict__ beta, __nv_bfloat16* __restrict__ out16, float2* __restrict__ stats,
                     long long rows) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long r = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float2 sp = stats_prev[r];
  float v[16];
  const float4* x4 = reinterpret_cast<const float4*>(xprev + r * LN_D);
  const uint2* y2 = reinterpret_cast<const uint2*>(y16 + r * LN_D);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = (lane + 32 * i) * 4;
    const float4 q = x4[lane + 32 * i];
    const uint2 yy = y2[lane + 32 * i];
    const float4 g = *reinterpret_cast<const float4*>(gamma_prev + c);
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (beta_prev != nullptr) b = *reinterpret_cast<const float4*>(beta_prev + c);
    v[4 * i] = (q.x - sp.x) * sp.y * g.x + b.x + bf16_lo(yy.x);
    v[4 * i + 1] = (q.y - sp.x) * sp.y * g.y + b.y + bf16_hi(yy.x);
    v[4 * i + 2] = (q.z - sp.x) * sp.y * g.z + b.z + bf16_lo(yy.y);
    v[4 * i + 3] = (q.w - sp.x) * sp.y * g.w + b.w + bf16_hi(yy.y);
    if (xnew != nullptr)
      *reinterpret_cast<float4*>(xnew + r * LN_D + c) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  }
  if (out16 == nullptr) return;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  const float mean = warp_sum(s) * (1.0f / LN_D);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float d = v[i] - mean;
    ss += d * d;
  }
  const float rstd = rsqrtf(warp_sum(ss) * (1.0f / LN_D) + LN_EPS);
  if (stats != nullptr && lane == 0) stats[r] = make_float2(mean, rstd);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = (lane + 32 * i) * 4;
    const float4 g = *reinterpret_cast<const float4*>(gamma + c);
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (beta != nullptr) b = *reinterpret_cast<const float4*>(beta + c);
    uint2 h;
    h.x = pack_bf16x2((v[4 * i] - mean) * rstd * g.x + b.x, (v[4 * i + 1] - mean) * rstd * g.y + b.y);
    h.y = pack_bf16x2((v[4 * i + 2] - mean) * rstd * g.z + b.z, (v[4 * i + 3] - mean) * rstd * g.w + b.w);
    *reinterpret_cast<uint2*>(out16 + r * LN_D + c) = h;
  }
}

// dx = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat));  dgamma += sum_rows dy*xhat;  dbeta += sum_rows dy.
// dy is read at the mapped row (gathering the modality's rows out of the packed gradient); rows flagged in `pad`
// contribute nothing (encoders.py:205 masked_fill blocks their gradient).
__global__ void __launch_bounds__(256)
ln512_bwd_kernel(const float* __restrict__ dy, const __nv_bfloat16* __restrict__ dy_delta, const float* __restrict__ x,
                 const float2* __restrict__ stats, const float* __restrict__ gamma, float* __restrict__ dx32,
                 __nv_bfloat16* __restrict__ dx16, float* __restrict__ dgamma, float* __restrict__ dbeta,
                 const uint8_t* __restrict__ pad, RowMap map, long long rows) {
  __shared__ float sg[8][LN_D];
  __shared__ float sb[8][LN_D];
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float ag[16], ab[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) ag[i] = 0.f, ab[i] = 0.f;
  float g[16];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 q = *reinterpret_cast<const float4*>(gamma + (lane + 32 * i) * 4);
    g[4 * i] = q.x, g[4 * i + 1] = q.y, g[4 * i + 2] = q.z, g[4 * i + 3] = q.w;
  }
  for (long long r = static_cast<long long>(blockIdx.x) * 8 + warp; r < rows; r += static_cast<long long>(gridDim.x) * 8) {
    const long long rd = map(r);
    const bool padded = pad != nullptr && pad[r] != 0;
    float d[16], xh[16];
    const float2 st = stats[r];
    const float4* d4 = reinterpret_cast<const float4*>(dy + rd * LN_D);
    const float4* x4 = reinterpret_cast<const float4*>(x + r * LN_D);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 q = d4[lane + 32 * i];
      if (dy_delta != nullptr) {  // gradient that arrived through the branch GEMM (bf16), added to the residual path
        const uint2 dd = *reinterpret_cast<const uint2*>(dy_delta + rd * LN_D + (lane + 32 * i) * 4);
        q.x += bf16_lo(dd.x), q.y += bf16_hi(dd.x), q.z += bf16_lo(dd.y), q.w += bf16_hi(dd.y);
      }
      if (padded) q = make_float4(0.f, 0.f, 0.f, 0.f);
      d[4 * i] = q.x, d[4 * i + 1] = q.y, d[4 * i + 2] = q.z, d[4 * i + 3] = q.w;
      const float4 p = x4[lane + 32 * i];
      xh[4 * i] = (p.x - st.x) * st.y, xh[4 * i + 1] = (p.y - st.x) * st.y;
      xh[4 * i + 2] = (p.z - st.x) * st.y, xh[4 * i + 3] = (p.w - st.x) * st.y;
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      ag[i] += d[i] * xh[i];
      ab[i] += d[i];
      const float gd = g[i] * d[i];
      s1 += gd;
      s2 += gd * xh[i];
    }
    s1 = warp_sum(s1) * (1.0f / LN_D);
    s2 = warp_sum(s2) * (1.0f / LN_D);
    if (dx32 != nullptr || dx16 != nullptr) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = (lane + 32 * i) * 4;
        float4 o;
        o.x = st.y * (g[4 * i] * d[4 * i] - s1 - xh[4 * i] * s2);
        o.y = st.y * (g[4 * i + 1] * d[4 * i + 1] - s1 - xh[4 * i + 1] * s2);
        o.z = st.y * (g[4 * i + 2] * d[4 * i + 2] - s1 - xh[4 * i + 2] * s2);
        o.w = st.y * (g[4 * i + 3] * d[4 * i + 3] - s1 - xh[4 * i + 3] * s2);
        if (dx32 != nullptr) *reinterpret_cast<float4*>(dx32 + r * LN_D + c) = o;
        if (dx16 != nullptr) {
          uint2 h;
          h.x = pack_bf16x2(o.x, o.y);
          h.y = pack_bf16x2(o.z, o.w);
          *reinterpret_cast<uint2*>(dx16 + r * LN_D + c) = h;
        }
      }
    }
  }
  // block reduction of the per-warp column sums, then one atomic per column per block
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = (lane + 32 * i) * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) sg[warp][c + j] = ag[4 * i + j], sb[warp][c + j] = ab[4 * i + j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < LN_D; c += blockDim.x) {
    float tg = 0.f, tb = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tg += sg[w][c], tb += sb[w][c];
    if (dgamma != nullptr) atomicAdd(dgamma + c, tg);
    if (dbeta != nullptr) atomicAdd(dbeta + c, tb);
  }
}

// ---- arbitrary-width LayerNorm of encoder inputs: y(bf16, zero padded to kpad) = LN(x) ; padded rows -> 0
__global__ void __launch_bounds__(256)
lnw_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
               const uint8_t* __restrict__ pad, __nv_bfloat16* __restrict__ y, float2* __restrict__ stats, int width,
               int kpad, long long rows, int* __restrict__ nonfinite_flag) {
  const int lane = threadIdx.x & 31;
  const long long r = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const bool padded = pad != nullptr && pad[r] != 0;
  const float* xr = x + r * width;
  __nv_bfloat16* yr = y + r * kpad;
  float s = 0.f;
  bool bad = false;
  for (int c = lane; c < width; c += 32) {
    const float v = xr[c];
    bad |= !isfinite(v);
    s += padded ? 0.f : v;
  }
  if (bad) atomicOr(nonfinite_flag, 1);  // reference raises on non-finite tokens (encoders.py:197-198)
  const float mean = warp_sum(s) / width;
  float ss = 0.f;
  for (int c = lane; c < width; c += 32) {
    const float d = (padded ? 0.f : xr[c]) - mean;
    ss += d * d;
  }
  const float rstd = rsqrtf(warp_sum(ss) / width + LN_EPS);
  if (lane == 0) stats[r] = make_float2(mean, rstd);
  for (int c = lane; c < kpad; c += 32) {
    float o = 0.f;
    // a padded row is zero-filled first (encoders.py:199), so LN of it is just the bias; its output is discarded
    // downstream (zeroed again at encoders.py:205), so emit zeros and skip the work
    if (c < width && !padded) o = (xr[c] - mean) * rstd * w[c] + b[c];
    yr[c] = __float2bfloat16(o);
  }
}

// Column reductions over token rows.  Block = 128 rows x 128 columns: thread (cx = t % 32, ry = t / 32) walks rows
// ry, ry + 8, ... for four 32-column groups (every warp load is one coalesced 128-byte line), the eight row groups are
// combined in shared memory and one atomicAdd per column leaves the block.
constexpr int CR_ROWS = 128, CR_COLS = 128;

// parameter gradients of the input LayerNorm: dw[c] += sum_r dy[r,c]*xhat[r,c], db[c] += sum_r dy[r,c]
__global__ void __launch_bounds__(256)
lnw_param_bwd_kernel(const float* __restrict__ dy, int ld_dy, const float* __restrict__ x, const float2* __restrict__ stats,
                     const uint8_t* __restrict__ pad, float* __restrict__ dw, float* __restrict__ db, int width,
                     long long rows) {
  __shared__ float sw[8][CR_COLS], sb[8][CR_COLS];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const long long r0 = static_cast<long long>(blockIdx.x) * CR_ROWS;
  const long long r1 = min(rows, r0 + CR_ROWS);
  const int c0 = blockIdx.y * CR_COLS;
  float aw[4] = {0.f, 0.f, 0.f, 0.f}, ab[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long r = r0 + ry; r < r1; r += 8) {
    if (pad != nullptr && pad[r] != 0) continue;
    const float2 st = stats[r];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int c = c0 + g * 32 + cx;
      if (c < width) {
        const float d = dy[r * ld_dy + c];
        aw[g] += d * (x[r * width + c] - st.x) * st.y;
        ab[g] += d;
      }
    }
  }
#pragma unroll
  for (int g = 0; g < 4; ++g) sw[ry][g * 32 + cx] = aw[g], sb[ry][g * 32 + cx] = ab[g];
  __syncthreads();
  if (threadIdx.x < CR_COLS && c0 + threadIdx.x < width) {
    float tw = 0.f, tb = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tw += sw[i][threadIdx.x], tb += sb[i][threadIdx.x];
    atomicAdd(dw + c0 + threadIdx.x, tw);
    atomicAdd(db + c0 + threadIdx.x, tb);
  }
}

// column sums: out[c] += sum_r a[r, c]  (bias gradients)
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ a, int ld, float* __restrict__ out, int width, long long rows) {
  __shared__ float ss[8][CR_COLS];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const long long r0 = static_cast<long long>(blockIdx.x) * CR_ROWS;
  const long long r1 = min(rows, r0 + CR_ROWS);
  const int c0 = blockIdx.y * CR_COLS;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long r = r0 + ry; r < r1; r += 8) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int c = c0 + g * 32 + cx;
      if (c < width) acc[g] += a[r * ld + c];
    }
  }
#pragma unroll
  for (int g = 0; g < 4; ++g) ss[ry][g * 32 + cx] = acc[g];
  __syncthreads();
  if (threadIdx.x < CR_COLS && c0 + threadIdx.x < width) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += ss[i][threadIdx.x];
    atomicAdd(out + c0 + threadIdx.x, t);
  }
}

}  // namespace mca

using namespace mca;

extern "C" int mca_layernorm512_fwd(const float* x, const float* gamma, const float* beta, float* y32, void* y16,
                                    float* stats, const uint8_t* pad, const float* pe, int seg_len, int out_rows_per_b,
                                    int out_row_off, long long rows, void* stream) {
  if (rows <= 0) return MCA_ERR_SHAPE;
  if (pe != nullptr && seg_len <= 0) return MCA_ERR_ARG;
  RowMap map{seg_len, out_rows_per_b, out_row_off};
  const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
  if (launch_kernel(ln512_fwd_kernel, dim3(grid), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 2, x, gamma, beta, y32,
                    reinterpret_cast<__nv_bfloat16*>(y16), reinterpret_cast<float2*>(stats), pad, pe, map, rows) != cudaSuccess)
    return MCA_ERR_CUDA;
  return check_launch();
}

extern "C" int mca_layernorm512_bwd(const float* dy, const void* dy_delta_bf16, const float* x, const float* stats,
                                    const float* gamma, float* dx32, void* dx16, float* dgamma, float* dbeta,
                                    const uint8_t* pad, int seg_len, int out_rows_per_b, int out_row_off, long long rows,
                                    void* stream) {
  if (rows <= 0) return MCA_ERR_SHAPE;
  RowMap map{seg_len, out_rows_per_b, out_row_off};
  long long want = (rows + 7) / 8;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  const unsigned grid = static_cast<unsigned>(want < cap ? want : cap);
  if (launch_kernel(ln512_bwd_kernel, dim3(grid), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 2, dy,
                    reinterpret_cast<const __nv_bfloat16*>(dy_delta_bf16), x, reinterpret_cast<const float2*>(stats), gamma, dx32,
                    reinterpret_cast<__nv_bfloat16*>(dx16), dgamma, dbeta, pad, map, rows) != cudaSuccess)
    return MCA_ERR_CUDA;
  return check_launch();
}

extern "C" int mca_add_layernorm512_fwd(const float* xprev, const float* stats_prev, const float* gamma_prev,
                                        const float* beta_prev, const void* y_bf16, float* xnew, const float* gamma,
                                        const float* beta, void* out_bf16, float* stats, long long rows, void* stream) {
  if (rows <= 0) return MCA_ERR_SHAPE;
  if (xprev == nullptr || stats_prev == nullptr || gamma_prev == nullptr || y_bf16 == nullptr) return MCA_ERR_ARG;
  if (out_bf16 != nullptr && gamma == nullptr) return MCA_ERR_ARG;
  const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
  if (launch_kernel(add_ln512_fwd_kernel, dim3(grid), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 2, xprev,
                    reinterpret_cast<const float2*>(stats_prev), gamma_prev, beta_prev,
                    reinterpret_cast<const __nv_bfloat16*>(y_bf16), xnew, gamma, beta,
                    reinterpret_cast<__nv_bfloat16*>(out_bf16), reinterpret_cast<float2*>(stats), rows) != cudaSuccess)
    return MCA_ERR_CUDA;
  return check_launch();
}

extern "C" int mca_layernorm_in_fwd(const float* x, const float* w, const float* b, const uint8_t* pad, void* y_bf16,
                                    float* stats, int width, int kpad, long long rows, int* nonfinite_flag,
                                    void* stream) {
  if (rows <= 0 || width <= 0 || kpad < width || (kpad % 8) != 0) return MCA_ERR_SHAPE;
  const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
  lnw_fwd_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, w, b, pad, reinterpret_cast<__nv_bfloat16*>(y_bf16), reinterpret_cast<float2*>(stats), width, kpad, rows,
      nonfinite_flag);
  return check_launch();
}

extern "C" int mca_layernorm_in_param_bwd(const float* dy, int ld_dy, const float* x, const float* stats,
                                          const uint8_t* pad, float* dw, float* db, int width, long long rows,
                                          void* stream) {
  if (rows <= 0 || width <= 0) return MCA_ERR_SHAPE;
  dim3 grid(static_cast<unsigned>((rows + CR_ROWS - 1) / CR_ROWS), (width + CR_COLS - 1) / CR_COLS);
  lnw_param_bwd_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      dy, ld_dy, x, reinterpret_cast<const float2*>(stats), pad, dw, db, width, rows);
  return check_launch();
}

extern "C" int mca_colsum(const float* a, int ld, float* out, int width, long long rows, void* stream) {
  if (rows <= 0 || width <= 0) return MCA_ERR_SHAPE;
  dim3 grid(static_cast<unsigned>((rows + CR_ROWS - 1) / CR_ROWS), (width + CR_COLS - 1) / CR_COLS);
  colsum_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a, ld, out, width, rows);
  return check_launch();
}
