// Epilogue of the tcgen05 GEMM kernels, shared by the single-CTA kernel (gemm.cu, 128 x 128 tiles) and the CTA-pair kernel
// (gemm2.cu, 256 x 256 tiles): one warp turns a [32 rows x 64 columns] slab of the fp32 accumulator (already in registers:
// r0 = columns 0..31, r1 = columns 32..63 of the slab, thread = row) into the output of the selected MCA_EPI_* mode, staged
// in the warp's 8 KB swizzled shared-memory area and moved by TMA store / reduce / load.  The kernels differ only in how a
// tile is cut into slabs, i.e. in the coordinates they pass.
#pragma once
#include "mca_b200.h"
#include "ptx.cuh"

namespace mca {

// ---- staging helpers: thread = one row of a [32 rows x 128 B] (or [32 x 64 B]) swizzled box
__device__ __forceinline__ uint32_t sw128_off(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }
__device__ __forceinline__ uint32_t sw64_off(int row, int chunk) { return row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4); }

__device__ __forceinline__ uint4 pack8(const float* v) {
  uint4 q;
  q.x = pack_bf16x2(v[0], v[1]);
  q.y = pack_bf16x2(v[2], v[3]);
  q.z = pack_bf16x2(v[4], v[5]);
  q.w = pack_bf16x2(v[6], v[7]);
  return q;
}
__device__ __forceinline__ void unpack8(const uint4 q, float* v) {
  v[0] = bf16_lo(q.x), v[1] = bf16_hi(q.x), v[2] = bf16_lo(q.y), v[3] = bf16_hi(q.y);
  v[4] = bf16_lo(q.z), v[5] = bf16_hi(q.z), v[6] = bf16_lo(q.w), v[7] = bf16_hi(q.w);
}

// Where a slab lives.  col0: first output column of the 64-column slab (every mode but GEGLU).  GEGLU tiles are made of
// [64 value | 64 gate] blocks: geglu_blk0 = first column of the slab's block, geglu_sub = which 32-pair group of it.
struct EpiSlab {
  int col0, geglu_blk0, geglu_sub, row0, z;
};

__device__ __forceinline__ bool epi_has_aux(int mode) { return mode == MCA_EPI_RESID || mode == MCA_EPI_GEGLU_BWD; }

// one elected lane: request the slab's aux operand (fp32 residual or the bf16 GEGLU backward factors) into the staging area
__device__ __forceinline__ void epi_request_aux(int mode, uint8_t* stg, const CUtensorMap* tmAux, uint64_t* xbar, const EpiSlab& s) {
  mbar_expect_tx(xbar, 8192);
  if (mode == MCA_EPI_RESID) {  // fp32 boxes [32 cols x 32 rows]
    tma_load_2d(stg, tmAux, xbar, s.col0, s.row0);
    tma_load_2d(stg + 4096, tmAux, xbar, s.col0 + 32, s.row0);
  } else {  // bf16 boxes [64 cols x 32 rows]: value-side and gate-side factors of this 64-column block
    const int blk = s.col0 / 64;
    tma_load_2d(stg, tmAux, xbar, blk * 128, s.row0);
    tma_load_2d(stg + 4096, tmAux, xbar, blk * 128 + 64, s.row0);
  }
}

// P: the kernel's parameter block (mode, N, bias, alpha, reduce)
template <typename P>
__device__ __forceinline__ void epi_store_slab(const P& p, const uint32_t (&r0)[32], const uint32_t (&r1)[32], uint8_t* stg, int lane,
                                               const EpiSlab& s, const CUtensorMap* tmO0, const CUtensorMap* tmO1, uint64_t* xbar,
                                               uint32_t& xph) {
  if (p.mode == MCA_EPI_BF16) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int j = c * 8 + i;
        v[i] = __uint_as_float(j < 32 ? r0[j] : r1[j - 32]) * p.alpha;
        if (p.bias != nullptr && s.col0 + j < p.N) v[i] += __ldg(p.bias + s.col0 + j);
      }
      *reinterpret_cast<uint4*>(stg + sw128_off(lane, c)) = pack8(v);
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (elect_one()) {
      tma_store_2d(tmO0, stg, s.col0, s.row0);
      bulk_commit_group();
    }
  } else if (p.mode == MCA_EPI_F32 || p.mode == MCA_EPI_RESID) {
    const bool has_aux = p.mode == MCA_EPI_RESID;
    if (has_aux) {
      mbar_wait(xbar, xph);
      xph ^= 1;
    }
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int ncol = s.col0 + b * 32;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float4 v;
        const uint32_t* r = b == 0 ? r0 : r1;
        v.x = __uint_as_float(r[4 * c + 0]) * p.alpha, v.y = __uint_as_float(r[4 * c + 1]) * p.alpha;
        v.z = __uint_as_float(r[4 * c + 2]) * p.alpha, v.w = __uint_as_float(r[4 * c + 3]) * p.alpha;
        if (p.bias != nullptr && ncol + 4 * c < p.N) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + ncol + 4 * c));
          v.x += bb.x, v.y += bb.y, v.z += bb.z, v.w += bb.w;
        }
        float4* dst = reinterpret_cast<float4*>(stg + b * 4096 + sw128_off(lane, c));
        if (has_aux) {
          const float4 a = *dst;
          v.x += a.x, v.y += a.y, v.z += a.z, v.w += a.w;
        }
        *dst = v;
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (elect_one()) {
      if (p.reduce) {
        tma_reduce_add_3d(tmO0, stg, s.col0, s.row0, 0);
        tma_reduce_add_3d(tmO0, stg + 4096, s.col0 + 32, s.row0, 0);
      } else {
        tma_store_3d(tmO0, stg, s.col0, s.row0, s.z);
        tma_store_3d(tmO0, stg + 4096, s.col0 + 32, s.row0, s.z);
      }
      bulk_commit_group();
    }
  } else if (p.mode == MCA_EPI_GEGLU) {
    // r0 = value x, r1 = gate g for 32 (x, g) pairs.  Stored for the backward: a = gelu(g), bv = x * gelu'(g);
    // forward output h = x * gelu(g).  Boxes of [32 cols x 32 rows] bf16 (64-byte rows, 64B swizzle).
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float av[8], bv[8], hv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float x = __uint_as_float(r0[c * 8 + i]), g = __uint_as_float(r1[c * 8 + i]);
        float cdf, pdf;
        gelu_cdf_pdf(g, cdf, pdf);
        const float ge = g * cdf;
        av[i] = ge;
        bv[i] = x * fmaf(g, pdf, cdf);
        hv[i] = x * ge;
      }
      const uint32_t o = sw64_off(lane, c);
      *reinterpret_cast<uint4*>(stg + o) = pack8(av);
      *reinterpret_cast<uint4*>(stg + 2048 + o) = pack8(bv);
      *reinterpret_cast<uint4*>(stg + 4096 + o) = pack8(hv);
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (elect_one()) {
      tma_store_2d(tmO1, stg, s.geglu_blk0 + s.geglu_sub * 32, s.row0);
      tma_store_2d(tmO1, stg + 2048, s.geglu_blk0 + 64 + s.geglu_sub * 32, s.row0);
      tma_store_2d(tmO0, stg + 4096, (s.geglu_blk0 / 128) * 64 + s.geglu_sub * 32, s.row0);
      bulk_commit_group();
    }
  } else {  // MCA_EPI_GEGLU_BWD: acc = dL/dh for 64 h-columns; staged (a | bv) are overwritten by (dL/dx | dL/dg)
    mbar_wait(xbar, xph);
    xph ^= 1;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float a[8], b[8];
      uint4* pa = reinterpret_cast<uint4*>(stg + sw128_off(lane, c));
      uint4* pb = reinterpret_cast<uint4*>(stg + 4096 + sw128_off(lane, c));
      unpack8(*pa, a);
      unpack8(*pb, b);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int j = c * 8 + i;
        const float d = __uint_as_float(j < 32 ? r0[j] : r1[j - 32]) * p.alpha;
        a[i] *= d;
        b[i] *= d;
      }
      *pa = pack8(a);
      *pb = pack8(b);
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (elect_one()) {
      const int blk = s.col0 / 64;
      tma_store_2d(tmO0, stg, blk * 128, s.row0);
      tma_store_2d(tmO0, stg + 4096, blk * 128 + 64, s.row0);
      bulk_commit_group();
    }
  }
}

}  // namespace mca
