// fp32-parity forward mode ("exact" mode; north_star: fp32 loss and embeddings within 1e-3 of the reference, which runs
// fp32 end to end: train_accel_gpu.py:21 default Accelerator(), no autocast, model.py:73-105).
//
// The tensor cores only take bf16 here, so every dense contraction is computed as a 3-term split product on the SAME
// tcgen05 GEMM kernel: with x = hi + lo (hi = bf16(x), lo = bf16(x - hi): 16 significand bits),
//     A W^T  ~=  A_hi W_hi^T + A_hi W_lo^T + A_lo W_hi^T            (dropped: A_lo W_lo^T ~ 2^-16 relative)
// is ONE bf16 GEMM over K' = 3K whose operands are laid out [A_hi | A_hi | A_lo] and [W_hi | W_lo | W_hi] along K, with
// the usual fp32 accumulation in TMEM.  This file holds the operand producers of that layout (generic split, weight pack,
// the two encoder front ends), the fp32 GEGLU between FF1 and FF2, and an fp32 masked attention forward (SIMT, online
// softmax) that reads fp32 q/k/v and leaves fp32 + bf16 outputs and the log-sum-exp, so the regular (bf16) backward
// kernels consume what this forward saved.  Nothing here is on the default bf16 path.
#include <math_constants.h>

#include "mca_b200.h"
#include "ptx.cuh"
#include "runtime.h"

namespace mca {

__device__ __forceinline__ void split2(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// dst [rows, 3*kpad] <- src [rows, cols] (cols <= kpad, zero padded).  weight = 0: [hi | hi | lo], 1: [hi | lo | hi]
__global__ void __launch_bounds__(256)
split_kernel(const float* __restrict__ src, long long ld_src, __nv_bfloat16* __restrict__ dst, long long rows, int cols,
             int kpad, int weight) {
  const int c4 = kpad / 4;
  const long long n = rows * c4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / c4;
    const int c = static_cast<int>(i % c4) * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (c + 3 < cols) {
      const float4 q = *reinterpret_cast<const float4*>(src + r * ld_src + c);
      v[0] = q.x, v[1] = q.y, v[2] = q.z, v[3] = q.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c + j < cols) v[j] = src[r * ld_src + c + j];
    }
    __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split2(v[j], h[j], l[j]);
    __nv_bfloat16* d = dst + r * (3LL * kpad) + c;
    const uint2 hh = make_uint2(pack_bf16x2(__bfloat162float(h[0]), __bfloat162float(h[1])),
                                pack_bf16x2(__bfloat162float(h[2]), __bfloat162float(h[3])));
    const uint2 ll = make_uint2(pack_bf16x2(__bfloat162float(l[0]), __bfloat162float(l[1])),
                                pack_bf16x2(__bfloat162float(l[2]), __bfloat162float(l[3])));
    *reinterpret_cast<uint2*>(d) = hh;
    *reinterpret_cast<uint2*>(d + kpad) = weight ? ll : hh;
    *reinterpret_cast<uint2*>(d + 2 * kpad) = weight ? hh : ll;
  }
}

__device__ __forceinline__ int map_row_x(const mca_pack_desc& d, int r) {
  if (d.mode == 0) return d.dst_row0 + r;
  const bool gate = r >= d.half;
  const int v = gate ? r - d.half : r;
  return d.dst_row0 + (v / 64) * 128 + (gate ? 64 : 0) + (v % 64);
}

// the descriptors of mca_pack_weights, written as [hi | lo | hi] with row stride 3*dst_ld at 3*dst_off
__global__ void __launch_bounds__(256)
pack_weights_split_kernel(const float* __restrict__ params, __nv_bfloat16* __restrict__ arena3,
                          const mca_pack_desc* __restrict__ descs) {
  const mca_pack_desc d = descs[blockIdx.y];
  const int lane = threadIdx.x & 31;
  for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < d.rows; r += gridDim.x * 8) {
    const float* src = params + d.src_off + static_cast<long long>(r) * d.cols;
    __nv_bfloat16* dst = arena3 + 3 * d.dst_off + static_cast<long long>(map_row_x(d, r)) * (3LL * d.dst_ld);
    for (int c = lane; c < d.cols; c += 32) {
      __nv_bfloat16 h, l;
      split2(src[c] * d.scale, h, l);
      dst[c] = h, dst[d.dst_ld + c] = l, dst[2 * d.dst_ld + c] = h;
    }
  }
}

__device__ __forceinline__ float warp_sum_x(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// LayerNorm of the encoder inputs (encoders.py:187-190 first stage), written as the [hi | hi | lo] operand
__global__ void __launch_bounds__(256)
lnw_split_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                 const uint8_t* __restrict__ pad, __nv_bfloat16* __restrict__ y3, int width, int kpad, long long rows) {
  const int lane = threadIdx.x & 31;
  const long long r = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const bool padded = pad != nullptr && pad[r] != 0;
  const float* xr = x + r * width;
  float s = 0.f;
  for (int c = lane; c < width; c += 32) s += padded ? 0.f : xr[c];
  const float mean = warp_sum_x(s) / width;
  float ss = 0.f;
  for (int c = lane; c < width; c += 32) {
    const float d = (padded ? 0.f : xr[c]) - mean;
    ss += d * d;
  }
  const float rstd = rsqrtf(warp_sum_x(ss) / width + 1e-5f);
  __nv_bfloat16* yr = y3 + r * (3LL * kpad);
  for (int c = lane; c < kpad; c += 32) {
    float o = 0.f;
    if (c < width && !padded) o = (xr[c] - mean) * rstd * w[c] + b[c];
    __nv_bfloat16 h, l;
    split2(o, h, l);
    yr[c] = h, yr[kpad + c] = h, yr[2 * kpad + c] = l;
  }
}

// ContinuousValueEncoder first stage (encoders.py:60-75): relu(w1 * min(v, max) + b1), as the [hi | hi | lo] operand
__global__ void __launch_bounds__(256)
tabular_split_kernel(const float* __restrict__ values, const float* __restrict__ w1, const float* __restrict__ b1,
                     __nv_bfloat16* __restrict__ h3, float max_value, int d, long long rows) {
  const long long n = rows * d;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / d;
    const int c = static_cast<int>(i % d);
    const float vc = fminf(values[r], max_value);
    const float a = fmaxf(w1[c] * vc + b1[c], 0.f);
    __nv_bfloat16 h, l;
    split2(a, h, l);
    __nv_bfloat16* o = h3 + r * (3LL * d);
    o[c] = h, o[d + c] = h, o[2 * d + c] = l;
  }
}

// GEGLU (model.py:35-38, F.gelu exact) on the fp32 FF1 output u32 [M, 2*IP] laid out [64 value | 64 gate] per 128 columns.
// Writes the next operand h3 [M, 3*IP] = [hi | hi | lo] of h = x * gelu(g) and, for the regular backward, the bf16 copy
// of h and the factors a = gelu(g), bv = x * gelu'(g) in the layout MCA_EPI_GEGLU leaves in `u`.
__global__ void __launch_bounds__(256)
geglu_f32_kernel(const float* __restrict__ u32, __nv_bfloat16* __restrict__ h3, __nv_bfloat16* __restrict__ h16,
                 __nv_bfloat16* __restrict__ u16, long long M, int IP) {
  const int q4 = IP / 4;  // thread = four consecutive h columns (they never straddle a 64-column block)
  const long long n = M * q4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / q4;
    const int c = static_cast<int>(i % q4) * 4;
    const int blk = c >> 6, j = c & 63;
    const float4 xv = *reinterpret_cast<const float4*>(u32 + r * (2LL * IP) + blk * 128 + j);
    const float4 gv = *reinterpret_cast<const float4*>(u32 + r * (2LL * IP) + blk * 128 + 64 + j);
    const float x[4] = {xv.x, xv.y, xv.z, xv.w}, g[4] = {gv.x, gv.y, gv.z, gv.w};
    float hh[4], hl[4], av[4], bv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float cdf = 0.5f * (1.0f + erff(g[k] * 0.70710678118654752f));
      const float pdf = 0.3989422804014327f * expf(-0.5f * g[k] * g[k]);
      const float ge = g[k] * cdf;
      const float h = x[k] * ge;
      __nv_bfloat16 bh, bl;
      split2(h, bh, bl);
      hh[k] = __bfloat162float(bh), hl[k] = __bfloat162float(bl);
      av[k] = ge, bv[k] = x[k] * fmaf(g[k], pdf, cdf);
    }
    const uint2 whi = make_uint2(pack_bf16x2(hh[0], hh[1]), pack_bf16x2(hh[2], hh[3]));
    const uint2 wlo = make_uint2(pack_bf16x2(hl[0], hl[1]), pack_bf16x2(hl[2], hl[3]));
    __nv_bfloat16* o = h3 + r * (3LL * IP) + c;
    *reinterpret_cast<uint2*>(o) = whi;
    *reinterpret_cast<uint2*>(o + IP) = whi;
    *reinterpret_cast<uint2*>(o + 2 * IP) = wlo;
    *reinterpret_cast<uint2*>(h16 + r * IP + c) = whi;
    *reinterpret_cast<uint2*>(u16 + r * (2LL * IP) + blk * 128 + j) = make_uint2(pack_bf16x2(av[0], av[1]), pack_bf16x2(av[2], av[3]));
    *reinterpret_cast<uint2*>(u16 + r * (2LL * IP) + blk * 128 + 64 + j) =
        make_uint2(pack_bf16x2(bv[0], bv[1]), pack_bf16x2(bv[2], bv[3]));
  }
}

// vmean[b, c] = mean over ALL N rows of V[b, :, c] (the value of a fully masked query row, reference quirk Q4)
__global__ void __launch_bounds__(256)
vmean_f32_kernel(const float* __restrict__ qkv, int ld, int v_col0, int width, int N, float* __restrict__ vmean) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int part = threadIdx.x >> 5;  // 8 row groups
  __shared__ float red[8][32];
  float a = 0.f;
  if (c < width)
    for (int n = part; n < N; n += 8) a += qkv[(static_cast<long long>(b) * N + n) * ld + v_col0 + c];
  red[part][threadIdx.x & 31] = a;
  __syncthreads();
  if (part == 0 && c < width) {
    float t = 0.f;
#pragma unroll
    for (int p = 0; p < 8; ++p) t += red[p][threadIdx.x];
    vmean[static_cast<long long>(b) * width + c] = t / static_cast<float>(N);
  }
}

// fp32 masked attention forward (model.py:85-100).  CTA = (64-query tile, head, sample), 256 threads as a 16 x 16 grid of
// 4 x 4 register tiles; key blocks of 64; allowed(q, k) = rowbits[q] >> keygrp[k] & 1 and key k not padded.  Key blocks in
// which no query of the tile may see any live key are skipped.
constexpr int XA_BM = 64, XA_BN = 64, XA_DH = 64, XA_LD = 68;  // smem rows padded to 68 floats (16-byte aligned, conflict-free)
constexpr int XA_SMEM = 4 * XA_DH * XA_LD * 4 + XA_BN * 4;

__global__ void __launch_bounds__(256)
attn_fwd_f32_kernel(const float* __restrict__ qkv, const uint32_t* __restrict__ rowbits, const uint8_t* __restrict__ keygrp,
                    const uint8_t* __restrict__ padding, const float* __restrict__ vmean, float* __restrict__ out32,
                    __nv_bfloat16* __restrict__ out16, float* __restrict__ lse, int N, int H) {
  extern __shared__ __align__(16) float xs[];
  float* Qt = xs;                      // [d][i]
  float* Kt = Qt + XA_DH * XA_LD;      // [d][j]
  float* Vs = Kt + XA_DH * XA_LD;      // [j][c]
  float* Pt = Vs + XA_BN * XA_LD;      // [j][i]
  uint32_t* kinfo = reinterpret_cast<uint32_t*>(Pt + XA_BN * XA_LD);  // per key of the block: group | live << 8
  __shared__ int s_any;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * XA_BM;
  const int ld = 3 * H * XA_DH;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const long long row0 = static_cast<long long>(b) * N;
  // Q tile, transposed (rows past N are zero and never stored)
  for (int e = threadIdx.x; e < XA_BM * XA_DH; e += 256) {
    const int i = e >> 6, d = e & 63;
    Qt[d * XA_LD + i] = q0 + i < N ? qkv[(row0 + q0 + i) * ld + h * XA_DH + d] : 0.f;
  }
  uint32_t rb[4], tile_bits = 0;
#pragma unroll
  for (int a = 0; a < 4; ++a) rb[a] = q0 + ty * 4 + a < N ? rowbits[q0 + ty * 4 + a] : 0u;
  for (int i = 0; i < XA_BM; ++i)
    if (q0 + i < N) tile_bits |= rowbits[q0 + i];
  float m[4], l[4], o[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    m[a] = -CUDART_INF_F, l[a] = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) o[a][c] = 0.f;
  }
  for (int k0 = 0; k0 < N; k0 += XA_BN) {
    __syncthreads();  // previous block's Kt / Vs / Pt / kinfo are no longer read
    if (threadIdx.x == 0) s_any = 0;
    __syncthreads();
    if (threadIdx.x < XA_BN) {
      const int k = k0 + threadIdx.x;
      uint32_t info = 0xFFu;  // group 255, not live
      if (k < N) {
        const uint32_t g = keygrp[k];
        const uint32_t live = padding[row0 + k] == 0 ? 1u : 0u;
        info = g | (live << 8);
        if (live && ((tile_bits >> g) & 1u)) s_any = 1;
      }
      kinfo[threadIdx.x] = info;
    }
    __syncthreads();
    if (s_any == 0) continue;  // uniform: nobody in this tile may see a live key of this block
    for (int e = threadIdx.x; e < XA_BN * XA_DH; e += 256) {
      const int j = e >> 6, d = e & 63;
      const bool ok = k0 + j < N;
      const float* src = qkv + (row0 + k0 + j) * ld + h * XA_DH + d;
      Kt[d * XA_LD + j] = ok ? src[H * XA_DH] : 0.f;
      Vs[j * XA_LD + d] = ok ? src[2 * H * XA_DH] : 0.f;
    }
    __syncthreads();
    float s[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) s[a][c] = 0.f;
#pragma unroll 8
    for (int d = 0; d < XA_DH; ++d) {
      const float4 qv = *reinterpret_cast<const float4*>(Qt + d * XA_LD + ty * 4);
      const float4 kv = *reinterpret_cast<const float4*>(Kt + d * XA_LD + tx * 4);
      const float qa[4] = {qv.x, qv.y, qv.z, qv.w}, ka[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) s[a][c] = fmaf(qa[a], ka[c], s[a][c]);
    }
    // mask, block maximum per row (16 lanes share a row group)
    float mx[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      mx[a] = -CUDART_INF_F;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint32_t info = kinfo[tx * 4 + c];
        const bool ok = (info >> 8) && ((rb[a] >> (info & 255u)) & 1u);
        s[a][c] = ok ? s[a][c] : -CUDART_INF_F;
        mx[a] = fmaxf(mx[a], s[a][c]);
      }
#pragma unroll
      for (int of = 8; of > 0; of >>= 1) mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], of));
    }
    float alpha[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const float mn = fmaxf(m[a], mx[a]);
      alpha[a] = mn == -CUDART_INF_F ? 1.f : expf(m[a] - mn);  // m = -inf -> 0
      float ps = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float p = s[a][c] == -CUDART_INF_F ? 0.f : expf(s[a][c] - mn);
        ps += p;
        Pt[(tx * 4 + c) * XA_LD + ty * 4 + a] = p;
      }
#pragma unroll
      for (int of = 8; of > 0; of >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, of);
      l[a] = l[a] * alpha[a] + ps;
      m[a] = mn;
#pragma unroll
      for (int c = 0; c < 4; ++c) o[a][c] *= alpha[a];
    }
    __syncthreads();
#pragma unroll 8
    for (int j = 0; j < XA_BN; ++j) {
      const float4 pv = *reinterpret_cast<const float4*>(Pt + j * XA_LD + ty * 4);
      const float4 vv = *reinterpret_cast<const float4*>(Vs + j * XA_LD + tx * 4);
      const float pa[4] = {pv.x, pv.y, pv.z, pv.w}, va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) o[a][c] = fmaf(pa[a], va[c], o[a][c]);
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int qi = q0 + ty * 4 + a;
    if (qi >= N) continue;
    float v[4];
    float ls = CUDART_INF_F;
    if (l[a] != 0.f) {
      const float inv = 1.0f / l[a];
#pragma unroll
      for (int c = 0; c < 4; ++c) v[c] = o[a][c] * inv;
      ls = m[a] + logf(l[a]);
    } else {  // no live allowed key: uniform over all N keys
#pragma unroll
      for (int c = 0; c < 4; ++c) v[c] = vmean[static_cast<long long>(b) * H * XA_DH + h * XA_DH + tx * 4 + c];
    }
    const long long oidx = (row0 + qi) * (H * XA_DH) + h * XA_DH + tx * 4;
    *reinterpret_cast<float4*>(out32 + oidx) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<uint2*>(out16 + oidx) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
    if (tx == 0) lse[(static_cast<long long>(b) * H + h) * N + qi] = ls;
  }
}

}  // namespace mca

using namespace mca;

extern "C" int mca_x_split_f32(const float* src, long long ld_src, void* dst3, long long rows, int cols, int kpad,
                               int weight_layout, void* stream) {
  if (rows <= 0 || cols <= 0 || cols > kpad || (kpad % 4) != 0 || (ld_src % 4) != 0) return MCA_ERR_SHAPE;
  split_kernel<<<148 * 4, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      src, ld_src, reinterpret_cast<__nv_bfloat16*>(dst3), rows, cols, kpad, weight_layout);
  return check_launch();
}

extern "C" int mca_x_pack_weights_split(const float* params, void* arena3_bf16, const mca_pack_desc* descs_dev, int n_desc,
                                        void* stream) {
  if (n_desc <= 0) return MCA_ERR_SHAPE;
  dim3 grid(48, n_desc);
  pack_weights_split_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      params, reinterpret_cast<__nv_bfloat16*>(arena3_bf16), descs_dev);
  return check_launch();
}

extern "C" int mca_x_layernorm_in_split(const float* x, const float* w, const float* b, const uint8_t* pad, void* y3,
                                        int width, int kpad, long long rows, void* stream) {
  if (rows <= 0 || width <= 0 || width > kpad) return MCA_ERR_SHAPE;
  lnw_split_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, w, b, pad, reinterpret_cast<__nv_bfloat16*>(y3), width, kpad, rows);
  return check_launch();
}

extern "C" int mca_x_tabular_split(const float* values, const float* w1, const float* b1, void* h3, float max_value, int d,
                                   long long rows, void* stream) {
  if (rows <= 0 || d <= 0) return MCA_ERR_SHAPE;
  tabular_split_kernel<<<148 * 4, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      values, w1, b1, reinterpret_cast<__nv_bfloat16*>(h3), max_value, d, rows);
  return check_launch();
}

extern "C" int mca_x_geglu_f32(const float* u32, void* h3, void* h16, void* u16, long long M, int IP, void* stream) {
  if (M <= 0 || IP <= 0 || (IP % 64) != 0) return MCA_ERR_SHAPE;
  geglu_f32_kernel<<<148 * 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      u32, reinterpret_cast<__nv_bfloat16*>(h3), reinterpret_cast<__nv_bfloat16*>(h16), reinterpret_cast<__nv_bfloat16*>(u16),
      M, IP);
  return check_launch();
}

extern "C" int mca_x_attn_fwd_f32(const float* qkv32, const uint32_t* rowbits, const uint8_t* keygrp, const uint8_t* padding,
                                  float* vmean, float* out32, void* out16, float* lse, int B, int N, int H, void* stream_) {
  if (B <= 0 || N <= 0 || H <= 0) return MCA_ERR_SHAPE;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(attn_fwd_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, XA_SMEM) != cudaSuccess)
      return MCA_ERR_CUDA;
    attr = true;
  }
  const int width = H * XA_DH;
  vmean_f32_kernel<<<dim3((width + 31) / 32, B), 256, 0, stream>>>(qkv32, 3 * width, 2 * width, width, N, vmean);
  dim3 grid((N + XA_BM - 1) / XA_BM, H, B);
  attn_fwd_f32_kernel<<<grid, 256, XA_SMEM, stream>>>(qkv32, rowbits, keygrp, padding, vmean, out32,
                                                      reinterpret_cast<__nv_bfloat16*>(out16), lse, N, H);
  return check_launch();
}
