// Integer / byte plumbing around the dense kernels (all bit-exact, HBM- or latency-bound):
//   mca_build_offsets   — modality-dropout / pad-mask builder (model.py:455-466, encoders.py:199,205,307,339):
//                         packed key-padding bytes, modality presence, live counts, packed varlen index lists and
//                         cumulative offsets, per-(sample, key-tile) liveness classes for tile skipping.
//   mca_pack_weights    — fp32 state_dict-layout master weights -> bf16 kernel-layout operands (GEGLU row
//                         interleave for model.py:37, zero padding of the odd inner dim 1365 -> 1408, q-scale fold
//                         model.py:87 which is exact because dim_head**-0.5 = 2**-3).
//   mca_unpack_grads    — inverse map for weight gradients, summing split-K partial slabs.
//   small helpers       — fusion-token broadcast / batch-sum (model.py:460-461), strided fp32->bf16 cast, column mean.
#include "mca_b200.h"
#include "ptx.cuh"
#include "runtime.h"

namespace mca {

struct OffsetsArgs {
  const void* mask[MCA_MAX_MODALITIES];
  int elem_size[MCA_MAX_MODALITIES];  // 1 (bool / uint8) or 8 (int64)
  int len[MCA_MAX_MODALITIES];
  int off[MCA_MAX_MODALITIES];        // offset of the modality inside a sample
  int n_mod, B, N;
};

__device__ __forceinline__ bool mask_at(const void* p, int es, long long i) {
  return es == 1 ? (reinterpret_cast<const uint8_t*>(p)[i] != 0) : (reinterpret_cast<const long long*>(p)[i] != 0);
}

// one warp per (sample, modality): ordered compaction of the live positions
__global__ void __launch_bounds__(32)
offsets_kernel(OffsetsArgs a, uint8_t* __restrict__ padding, uint8_t* __restrict__ pad_mod,
               uint8_t* __restrict__ present, int* __restrict__ live_count, int* __restrict__ live_idx,
               int* __restrict__ any_absent) {
  const int b = blockIdx.x / a.n_mod, m = blockIdx.x % a.n_mod;
  const int lane = threadIdx.x;
  const int L = a.len[m], off = a.off[m];
  const long long base = static_cast<long long>(b) * L;
  uint8_t* pm = pad_mod + static_cast<long long>(off) * a.B + base;  // modality-major [B, L_m] bytes
  uint8_t* pk = padding + static_cast<long long>(b) * a.N + off;
  int* li = live_idx + static_cast<long long>(b) * a.N + off;
  int count = 0;
  for (int l0 = 0; l0 < L; l0 += 32) {
    const int l = l0 + lane;
    bool live = false;
    if (l < L) {
      const bool p = mask_at(a.mask[m], a.elem_size[m], base + l);
      pm[l] = p ? 1 : 0;
      pk[l] = p ? 1 : 0;
      live = !p;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, live);
    if (live) li[count + __popc(bal & ((1u << lane) - 1))] = off + l;
    count += __popc(bal);
  }
  for (int l = count + lane; l < L; l += 32) li[l] = -1;
  if (lane == 0) {
    live_count[b * a.n_mod + m] = count;
    present[b * a.n_mod + m] = count > 0 ? 1 : 0;  // model.py:458
    if (count == 0) atomicOr(any_absent, 1);
  }
  // fusion tokens are never padded (model.py:344-346,462): written by the first modality's warp
  if (m == 0) {
    int n_tok = 0;
    for (int i = 0; i < a.n_mod; ++i) n_tok += a.len[i];
    for (int l = n_tok + lane; l < a.N; l += 32) {
      padding[static_cast<long long>(b) * a.N + l] = 0;
      live_idx[static_cast<long long>(b) * a.N + l] = l;
    }
  }
}

// second phase: per (sample, key tile) class 0 = all live, 1 = mixed, 2 = all padded; exclusive cumsum of counts
__global__ void __launch_bounds__(128)
offsets_tiles_kernel(const uint8_t* __restrict__ padding, const int* __restrict__ kt_start,
                     const int* __restrict__ kt_len, int n_kt, uint8_t* __restrict__ kt_class,
                     uint32_t* __restrict__ kt_live, int B, int N, const int* __restrict__ live_count, int n_mod,
                     int* __restrict__ cu_live) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * n_kt) {
    const int b = i / n_kt, kt = i % n_kt;
    const uint8_t* p = padding + static_cast<long long>(b) * N + kt_start[kt];
    int npad = 0;
    const int len = kt_len[kt];
    uint32_t w[4] = {0, 0, 0, 0};
    for (int j = 0; j < len; ++j) {
      npad += p[j] != 0;
      if (p[j] == 0) w[j >> 5] |= 1u << (j & 31);
    }
    kt_class[i] = npad == 0 ? 0 : (npad == len ? 2 : 1);
    if (kt_live != nullptr) *reinterpret_cast<uint4*>(kt_live + static_cast<long long>(i) * 4) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  if (i == 0) {
    int acc = 0;
    for (int j = 0; j < B * n_mod; ++j) {
      cu_live[j] = acc;
      acc += live_count[j];
    }
    cu_live[B * n_mod] = acc;
  }
}

// ------------------------------------------------------------------ weight pack / grad unpack
__device__ __forceinline__ int map_row(const mca_pack_desc& d, int r) {
  if (d.mode == 0) return d.dst_row0 + r;
  const int gate = r >= d.half ? 1 : 0;
  const int v = r - gate * d.half;
  return d.dst_row0 + (v / 64) * 128 + gate * 64 + (v % 64);
}

// one warp per source row (no per-element division; loads coalesced along the row), grid.y = descriptor; rows whose width is
// a multiple of 4 move as float4 -> 4 x bf16 (every matrix but the 713- / 35- / 74-wide encoder projections)
__global__ void __launch_bounds__(256)
pack_weights_kernel(const float* __restrict__ params, __nv_bfloat16* __restrict__ arena,
                    const mca_pack_desc* __restrict__ descs) {
  const mca_pack_desc d = descs[blockIdx.y];
  const int lane = threadIdx.x & 31;
  const bool vec = (d.cols & 3) == 0 && (d.src_off & 3) == 0 && (d.dst_off & 3) == 0 && (d.dst_ld & 3) == 0;
  for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < d.rows; r += gridDim.x * 8) {
    const float* src = params + d.src_off + static_cast<long long>(r) * d.cols;
    __nv_bfloat16* dst = arena + d.dst_off + static_cast<long long>(map_row(d, r)) * d.dst_ld;
    if (vec) {
      for (int c = lane * 4; c < d.cols; c += 128) {
        const float4 q = *reinterpret_cast<const float4*>(src + c);
        *reinterpret_cast<uint2*>(dst + c) =
            make_uint2(pack_bf16x2(q.x * d.scale, q.y * d.scale), pack_bf16x2(q.z * d.scale, q.w * d.scale));
      }
    } else {
      for (int c = lane; c < d.cols; c += 32) dst[c] = __float2bfloat16(src[c] * d.scale);
    }
  }
}

// grads[row] = scale * sum over the n_splits slabs (warp per row; float4 lanes with four independent row segments in flight
// when the row width allows it)
__global__ void __launch_bounds__(256)
unpack_grads_kernel(float* __restrict__ grads, const float* __restrict__ partials,
                    const mca_pack_desc* __restrict__ descs) {
  const mca_pack_desc d = descs[blockIdx.y];
  const int lane = threadIdx.x & 31;
  const bool vec = (d.cols & 3) == 0 && (d.src_off & 3) == 0 && (d.dst_off & 3) == 0 && (d.dst_ld & 3) == 0 &&
                   (d.split_stride & 3) == 0;
  for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < d.rows; r += gridDim.x * 8) {
    float* dst = grads + d.src_off + static_cast<long long>(r) * d.cols;
    const float* p = partials + d.dst_off + static_cast<long long>(map_row(d, r)) * d.dst_ld;
    if (vec) {
      for (int c0 = lane * 4; c0 < d.cols; c0 += 512) {
        float4 acc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int z = 0; z < d.n_splits; ++z) {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (c0 + 128 * u < d.cols) {
              const float4 q = *reinterpret_cast<const float4*>(p + static_cast<long long>(z) * d.split_stride + c0 + 128 * u);
              acc[u].x += q.x, acc[u].y += q.y, acc[u].z += q.z, acc[u].w += q.w;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (c0 + 128 * u < d.cols)
            *reinterpret_cast<float4*>(dst + c0 + 128 * u) =
                make_float4(acc[u].x * d.scale, acc[u].y * d.scale, acc[u].z * d.scale, acc[u].w * d.scale);
      }
      continue;
    }
    for (int c0 = lane; c0 < d.cols; c0 += 128) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int z = 0; z < d.n_splits; ++z) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (c0 + 32 * u < d.cols) acc[u] += p[static_cast<long long>(z) * d.split_stride + c0 + 32 * u];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (c0 + 32 * u < d.cols) dst[c0 + 32 * u] = acc[u] * d.scale;
    }
  }
}

// ------------------------------------------------------------------ small helpers
// dst[b*rows_per_b + row_off + f, :] = src[f, :]
__global__ void broadcast_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int F, int d, int B,
                                      int rows_per_b, int row_off) {
  const long long n = static_cast<long long>(B) * F * d / 4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long e = i * 4;
    const int c = static_cast<int>(e % d);
    const int f = static_cast<int>((e / d) % F);
    const int b = static_cast<int>(e / (static_cast<long long>(d) * F));
    *reinterpret_cast<float4*>(dst + (static_cast<long long>(b) * rows_per_b + row_off + f) * d + c) =
        *reinterpret_cast<const float4*>(src + static_cast<long long>(f) * d + c);
  }
}
// out[f, :] (+)= sum_b src[b*rows_per_b + row_off + f, :]
__global__ void batchsum_rows_kernel(const float* __restrict__ src, float* __restrict__ out, int F, int d, int B,
                                     int rows_per_b, int row_off, int accumulate) {
  const int n = F * d;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int f = i / d, c = i % d;
    float acc = accumulate ? out[i] : 0.f;
    for (int b = 0; b < B; ++b) acc += src[(static_cast<long long>(b) * rows_per_b + row_off + f) * d + c];
    out[i] = acc;
  }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, long long ld_src, __nv_bfloat16* __restrict__ dst,
                                     long long ld_dst, long long rows, int cols) {
  pdl_launch_dependents();
  pdl_wait();
  const int c4 = cols / 4;
  const long long n = rows * c4;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  // four independent 16-byte loads in flight per thread (one per thread leaves the HBM pipe at ~60 %)
  for (; i + 3 * stride < n; i += 4 * stride) {
    float4 q[4];
    long long r[4];
    int c[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long e = i + u * stride;
      r[u] = e / c4, c[u] = static_cast<int>(e % c4) * 4;
      q[u] = *reinterpret_cast<const float4*>(src + r[u] * ld_src + c[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      uint2 h;
      h.x = pack_bf16x2(q[u].x, q[u].y);
      h.y = pack_bf16x2(q[u].z, q[u].w);
      *reinterpret_cast<uint2*>(dst + r[u] * ld_dst + c[u]) = h;
    }
  }
  for (; i < n; i += stride) {
    const long long r = i / c4;
    const int c = static_cast<int>(i % c4) * 4;
    const float4 q = *reinterpret_cast<const float4*>(src + r * ld_src + c);
    uint2 h;
    h.x = pack_bf16x2(q.x, q.y);
    h.y = pack_bf16x2(q.z, q.w);
    *reinterpret_cast<uint2*>(dst + r * ld_dst + c) = h;
  }
}

// ------------------------------------------------------------------ device-side collate
// The step in front of the hot path: MultimodalCollator (encoders.py:286-403).  The host stages only the LIVE rows of
// the present samples back to back (varlen: nothing is copied for absent modalities or for padding) with B+1 row
// offsets; these kernels expand them into the collators' dense layouts and build the masks.
// rows kernel: out[b, l, :] = l < len_b ? clean(src[off_b + l, :]) : fill; mask[b, l] = l >= len_b (truncation to L)
// one warp per output row (b, l): the row's length test and source offset are computed once, lanes stride over E
__global__ void __launch_bounds__(256)
collate_rows_kernel(const float* __restrict__ src, const int* __restrict__ row_off, int B, int L, int E, float fill,
                    int clean, float* __restrict__ out, uint8_t* __restrict__ mask) {
  const int lane = threadIdx.x & 31;
  const long long n_rows = static_cast<long long>(B) * L;
  for (long long t = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5); t < n_rows;
       t += static_cast<long long>(gridDim.x) * 8) {
    const int l = static_cast<int>(t % L), b = static_cast<int>(t / L);
    const int o = row_off[b], len = min(row_off[b + 1] - o, L);
    float* dst = out + t * E;
    if (l < len) {
      const float* s = src + (static_cast<long long>(o) + l) * E;
      for (int e = lane; e < E; e += 32) {
        float v = s[e];
        if (clean) {  // torch.nan_to_num defaults: nan -> 0, +-inf -> +-FLT_MAX
          if (isnan(v)) v = 0.f;
          else if (isinf(v)) v = v > 0.f ? 3.402823466e+38f : -3.402823466e+38f;
        }
        dst[e] = v;
      }
    } else {
      for (int e = lane; e < E; e += 32) dst[e] = fill;
    }
    if (mask != nullptr && lane == 0) mask[t] = l >= len ? 1 : 0;
  }
}

// values kernel (SequenceCollator): out[b, l] = l < len_b ? src[off_b + l] : pad_token; mask[b, l] = (out == pad_token)
// as int64 — a pad value INSIDE the data is masked too (encoders.py:307)
template <typename T>
__global__ void __launch_bounds__(256)
collate_values_kernel(const T* __restrict__ src, const int* __restrict__ off, int B, int L, T pad_token,
                      T* __restrict__ out, long long* __restrict__ mask) {
  const long long n = static_cast<long long>(B) * L;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int l = static_cast<int>(i % L), b = static_cast<int>(i / L);
    const int o = off[b], len = off[b + 1] - o;
    const T v = l < len ? src[o + l] : pad_token;
    out[i] = v;
    if (mask != nullptr) mask[i] = v == pad_token ? 1 : 0;
  }
}

}  // namespace mca

using namespace mca;

extern "C" int mca_collate_rows(const float* src, const int* row_off, int B, int L, int E, float fill, int clean, float* out,
                                uint8_t* mask, void* stream) {
  if (B <= 0 || L <= 0 || E <= 0) return MCA_ERR_SHAPE;
  const long long n_rows = static_cast<long long>(B) * L;
  const unsigned grid = static_cast<unsigned>((n_rows + 7) / 8 < 148 * 16 ? (n_rows + 7) / 8 : 148 * 16);
  collate_rows_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, row_off, B, L, E, fill, clean, out, mask);
  return check_launch();
}

extern "C" int mca_collate_values_f32(const float* src, const int* off, int B, int L, float pad_token, float* out,
                                      long long* mask, void* stream) {
  if (B <= 0 || L <= 0) return MCA_ERR_SHAPE;
  const long long n = static_cast<long long>(B) * L;
  collate_values_kernel<float><<<static_cast<unsigned>((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      src, off, B, L, pad_token, out, mask);
  return check_launch();
}

extern "C" int mca_collate_values_i64(const long long* src, const int* off, int B, int L, long long pad_token,
                                      long long* out, long long* mask, void* stream) {
  if (B <= 0 || L <= 0) return MCA_ERR_SHAPE;
  const long long n = static_cast<long long>(B) * L;
  collate_values_kernel<long long><<<static_cast<unsigned>((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      src, off, B, L, pad_token, out, mask);
  return check_launch();
}

extern "C" int mca_build_offsets(const void* const* masks_host, const int* elem_sizes_host, const int* lens_host,
                                 int n_mod, int B, int N, const int* kt_start, const int* kt_len, int n_kt,
                                 uint8_t* padding, uint8_t* pad_mod, uint8_t* present, int* live_count, int* live_idx,
                                 int* cu_live, uint8_t* kt_class, uint32_t* kt_live, int* any_absent, void* stream_) {
  if (n_mod <= 0 || n_mod > MCA_MAX_MODALITIES || B <= 0) return MCA_ERR_SHAPE;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  OffsetsArgs a;
  int off = 0;
  for (int m = 0; m < n_mod; ++m) {
    if (elem_sizes_host[m] != 1 && elem_sizes_host[m] != 8) return MCA_ERR_ARG;
    a.mask[m] = masks_host[m], a.elem_size[m] = elem_sizes_host[m], a.len[m] = lens_host[m], a.off[m] = off;
    off += lens_host[m];
  }
  if (off > N) return MCA_ERR_SHAPE;
  a.n_mod = n_mod, a.B = B, a.N = N;
  if (cudaMemsetAsync(any_absent, 0, sizeof(int), stream) != cudaSuccess) return MCA_ERR_CUDA;
  offsets_kernel<<<B * n_mod, 32, 0, stream>>>(a, padding, pad_mod, present, live_count, live_idx, any_absent);
  const int n = B * n_kt > 1 ? B * n_kt : 1;
  offsets_tiles_kernel<<<(n + 127) / 128, 128, 0, stream>>>(padding, kt_start, kt_len, n_kt, kt_class, kt_live, B, N,
                                                            live_count, n_mod, cu_live);
  return check_launch();
}

extern "C" int mca_pack_weights(const float* params, void* arena_bf16, const mca_pack_desc* descs_dev, int n_desc,
                                void* stream) {
  if (n_desc <= 0) return MCA_ERR_SHAPE;
  dim3 grid(48, n_desc);
  pack_weights_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      params, reinterpret_cast<__nv_bfloat16*>(arena_bf16), descs_dev);
  return check_launch();
}

extern "C" int mca_unpack_grads(float* grads, const float* partials, const mca_pack_desc* descs_dev, int n_desc,
                                void* stream) {
  if (n_desc <= 0) return MCA_ERR_SHAPE;
  dim3 grid(48, n_desc);
  unpack_grads_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(grads, partials, descs_dev);
  return check_launch();
}

extern "C" int mca_broadcast_rows(const float* src, float* dst, int F, int d, int B, int rows_per_b, int row_off,
                                  void* stream) {
  if (F <= 0 || (d % 4) != 0) return MCA_ERR_SHAPE;
  broadcast_rows_kernel<<<148, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, dst, F, d, B, rows_per_b, row_off);
  return check_launch();
}

extern "C" int mca_batchsum_rows(const float* src, float* out, int F, int d, int B, int rows_per_b, int row_off,
                                 int accumulate, void* stream) {
  if (F <= 0) return MCA_ERR_SHAPE;
  batchsum_rows_kernel<<<(F * d + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      src, out, F, d, B, rows_per_b, row_off, accumulate);
  return check_launch();
}

__global__ void query_skip_flags_kernel(const uint8_t* __restrict__ present, int n_blk, int n_mod, int B, int mode,
                                        uint8_t* __restrict__ skip_ok) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  uint8_t ok = mode == 2 ? 1 : 0;
  if (mode == 1) {
    ok = 1;
    for (int m = 0; m < n_mod; ++m) ok &= present[b * n_blk + m] != 0 ? 1 : 0;
  }
  skip_ok[b] = ok;
}

extern "C" int mca_query_skip_flags(const uint8_t* present, int n_blk, int n_mod, int B, int mode, uint8_t* skip_ok,
                                    void* stream) {
  if (B <= 0 || n_mod <= 0 || n_mod > n_blk || mode < 0 || mode > 2) return MCA_ERR_SHAPE;
  query_skip_flags_kernel<<<(B + 63) / 64, 64, 0, reinterpret_cast<cudaStream_t>(stream)>>>(present, n_blk, n_mod, B, mode,
                                                                                          skip_ok);
  return check_launch();
}

extern "C" int mca_cast_f32_bf16(const float* src, long long ld_src, void* dst, long long ld_dst, long long rows,
                                 int cols, void* stream) {
  if (rows <= 0 || (cols % 4) != 0) return MCA_ERR_SHAPE;
  if (launch_kernel(cast_f32_bf16_kernel, dim3(148 * 4), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 2, src, ld_src,
                    reinterpret_cast<__nv_bfloat16*>(dst), ld_dst, rows, cols) != cudaSuccess)
    return MCA_ERR_CUDA;
  return check_launch();
}
