// Block-sparse masked attention backward on tcgen05 (autograd of model.py:85-100).
//
// One CTA = one (sample, head, 128-key tile); it walks the query tiles that attend this key tile (the transpose of
// the forward schedule).  Per query tile five tensor-core products, all accumulating in TMEM:
//     S  = Q K^T          dP = dO V^T                       (128x128x64, K-major operands as TMA wrote them)
//     P  = exp2(S*log2e - lse*log2e)     dS = P * (dP - delta)      (registers; one thread per query row)
//     dV += P^T dO        dK += dS^T Q                      (P/dS tiles re-read MN-major: no transposes in smem)
//     dQ  = dS K                                            (fresh tile -> fp32 smem -> TMA reduce-add into dq_acc)
// dK/dV stay resident in TMEM across the whole loop and are written once.  Q/dO tiles are double buffered.
// Fully masked query rows carry lse = +inf, so their P is exactly 0 here; their uniform-1/N contribution to dV
// (reference quirk Q4) is the per-(sample, head) vector `ucorr`, added to every key row in the epilogue.
#include <math_constants.h>

#include "mca_b200.h"
#include "ptx.cuh"
#include "runtime.h"

namespace mca {

constexpr int AB_T = 128;
constexpr int AB_DH = 64;
constexpr int AB_THREADS = 192;
constexpr int AB_TILE = AB_T * AB_DH * 2;  // 16 KB bf16 [128, 64]
constexpr int AB_PT = AB_T * AB_T * 2;     // 32 KB bf16 [128, 128]
constexpr int AB_DQ = AB_T * AB_DH * 4;    // 32 KB fp32 [128, 64]
// sK, sV, 2x(sQ, sdO), sP, sdS, sdQ
constexpr int AB_SMEM = 2 * AB_TILE + 4 * AB_TILE + 2 * AB_PT + AB_DQ + 1024 /*align*/ + 1024 /*keybits + barriers*/;
constexpr float AB_LOG2E = 1.4426950408889634f;

struct AttnBwdArgs {
  const mca_attn_qtile* k_tiles_q;  // per key tile: start, len, slice of qt_list
  const mca_attn_ref* qt_list;
  const mca_attn_tile* q_tiles;
  const uint32_t* rowbits;
  const uint8_t* keygrp;
  const uint8_t* padding;
  const uint8_t* kt_class;
  const float* lse;     // [B,H,N]
  const float* delta;   // [B,H,N]
  const float* ucorr;   // [B, H*64]
  __nv_bfloat16* dqkv;  // [B*N, 3*H*64]
  int N, H, n_kt;
};

__device__ __forceinline__ void ab_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// 32 consecutive bf16 of row r into a [128 x 128] K-major swizzled tile (two 64-column halves of 16 KB)
__device__ __forceinline__ void store_tile_chunk(uint8_t* tile, int r, int cc, const float (&v)[32]) {
  uint8_t* half = tile + (cc >> 1) * (AB_PT / 2) + r * 128;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 w;
    w.x = pack_bf16x2(v[8 * q + 0], v[8 * q + 1]);
    w.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
    w.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]);
    w.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
    const int chunk = ((cc & 1) * 4 + q) ^ (r & 7);
    *reinterpret_cast<uint4*>(half + chunk * 16) = w;
  }
}

__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                const __grid_constant__ CUtensorMap tm_dq, const AttnBwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = base;
  uint8_t* sV = sK + AB_TILE;
  uint8_t* sQ = sV + AB_TILE;        // 2 stages
  uint8_t* sdO = sQ + 2 * AB_TILE;   // 2 stages
  uint8_t* sP = sdO + 2 * AB_TILE;
  uint8_t* sdS = sP + AB_PT;
  uint8_t* sdQ = sdS + AB_PT;
  uint32_t* keybit = reinterpret_cast<uint32_t*>(sdQ + AB_DQ);
  uint64_t* bars = reinterpret_cast<uint64_t*>(keybit + AB_T);
  uint64_t* kv_full = bars + 0;
  uint64_t* qdo_full = bars + 1;   // [2]
  uint64_t* qdo_empty = bars + 3;  // [2]
  uint64_t* sdp_full = bars + 5;
  uint64_t* sdp_empty = bars + 6;
  uint64_t* pds_full = bars + 7;
  uint64_t* dq_full = bars + 8;
  uint64_t* dq_empty = bars + 9;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const mca_attn_qtile KT = a.k_tiles_q[kt];
  const long long row0 = static_cast<long long>(b) * a.N;
  const int cls = a.kt_class[static_cast<long long>(b) * a.n_kt + kt];
  const int n_iter = cls == 2 ? 0 : KT.kt_cnt;
  const int HD = a.H * AB_DH;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_dq);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) mbar_init(&qdo_full[s], 1), mbar_init(&qdo_empty[s], 1);
    mbar_init(sdp_full, 1);
    mbar_init(sdp_empty, 128);
    mbar_init(pds_full, 128);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 128);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc(tmem_holder, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const uint32_t tS = tmem_base, tdP = tmem_base + 128, tdV = tmem_base + 256, tdK = tmem_base + 320,
                 tdQ = tmem_base + 384;

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (lane == 0 && n_iter > 0) {
      const int krow = static_cast<int>(row0 + KT.start);
      mbar_expect_tx(kv_full, 2 * AB_TILE);
      tma_load_2d(sK, &tm_qkv, kv_full, HD + h * AB_DH, krow);
      tma_load_2d(sV, &tm_qkv, kv_full, 2 * HD + h * AB_DH, krow);
      for (int it = 0; it < n_iter; ++it) {
        const int s = it & 1;
        const uint32_t sph = (it >> 1) & 1;
        const mca_attn_tile Q = a.q_tiles[a.qt_list[KT.kt_off + it].tile];
        const int qrow = static_cast<int>(row0 + Q.start);
        mbar_wait(&qdo_empty[s], sph ^ 1);
        mbar_expect_tx(&qdo_full[s], 2 * AB_TILE);
        tma_load_2d(sQ + s * AB_TILE, &tm_qkv, &qdo_full[s], h * AB_DH, qrow);
        tma_load_2d(sdO + s * AB_TILE, &tm_do, &qdo_full[s], h * AB_DH, qrow);
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    if (lane == 0 && n_iter > 0) {
      constexpr uint32_t id_s = make_idesc_bf16(AB_T, AB_T, false, false);    // S, dP: K-major x K-major
      constexpr uint32_t id_kv = make_idesc_bf16(AB_T, AB_DH, true, true);    // dV, dK: MN-major x MN-major
      constexpr uint32_t id_q = make_idesc_bf16(AB_T, AB_DH, false, true);    // dQ: K-major x MN-major
      const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV), p_addr = smem_u32(sP), ds_addr = smem_u32(sdS);
      mbar_wait(kv_full, 0);
      for (int it = 0; it < n_iter; ++it) {
        const int s = it & 1;
        const uint32_t sph = (it >> 1) & 1, ph = it & 1;
        const uint32_t q_addr = smem_u32(sQ + s * AB_TILE), do_addr = smem_u32(sdO + s * AB_TILE);
        mbar_wait(&qdo_full[s], sph);
        mbar_wait(sdp_empty, ph ^ 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < AB_DH / 16; ++k)
          umma_bf16(tS, make_smem_desc_sw128(q_addr + k * 32, 16, 1024), make_smem_desc_sw128(k_addr + k * 32, 16, 1024),
                    id_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < AB_DH / 16; ++k)
          umma_bf16(tdP, make_smem_desc_sw128(do_addr + k * 32, 16, 1024), make_smem_desc_sw128(v_addr + k * 32, 16, 1024),
                    id_s, k > 0 ? 1u : 0u);
        umma_commit(sdp_full);
        mbar_wait(pds_full, ph);
        mbar_wait(dq_empty, ph ^ 1);
        tc_fence_after();
        // contraction over the 128 query rows: 8 steps of 16 rows (2 KB); the key halves are 16 KB apart (LBO)
#pragma unroll
        for (int k = 0; k < AB_T / 16; ++k)
          umma_bf16(tdV, make_smem_desc_sw128(p_addr + k * 2048, AB_PT / 2, 1024),
                    make_smem_desc_sw128(do_addr + k * 2048, 8192, 1024), id_kv, (it > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < AB_T / 16; ++k)
          umma_bf16(tdK, make_smem_desc_sw128(ds_addr + k * 2048, AB_PT / 2, 1024),
                    make_smem_desc_sw128(q_addr + k * 2048, 8192, 1024), id_kv, (it > 0 || k > 0) ? 1u : 0u);
        // dQ = dS K: contraction over the 128 keys
#pragma unroll
        for (int k = 0; k < AB_T / 16; ++k)
          umma_bf16(tdQ, make_smem_desc_sw128(ds_addr + (k >> 2) * (AB_PT / 2) + (k & 3) * 32, 16, 1024),
                    make_smem_desc_sw128(k_addr + k * 2048, 8192, 1024), id_q, k > 0 ? 1u : 0u);
        umma_commit(&qdo_empty[s]);
        umma_commit(dq_full);
      }
    }
  } else {
    // ===================== compute warps: thread = row =====================
    const int r = warp * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
    {  // key bitmask of this key tile (fixed for the whole CTA)
      const int kj = KT.start + r;
      uint32_t bit = 0;
      if (r < KT.len && a.padding[row0 + kj] == 0) bit = 1u << a.keygrp[kj];
      keybit[r] = bit;
    }
    ab_bar_sync(1, 128);
    for (int it = 0; it < n_iter; ++it) {
      const uint32_t ph = it & 1;
      const mca_attn_ref ref = a.qt_list[KT.kt_off + it];
      const mca_attn_tile Q = a.q_tiles[ref.tile];
      const int qi = Q.start + r;
      const bool valid = r < Q.len;
      const long long sidx = (static_cast<long long>(b) * a.H + h) * a.N + qi;
      const float lse2 = valid ? a.lse[sidx] * AB_LOG2E : CUDART_INF_F;
      const float dlt = valid ? a.delta[sidx] : 0.f;
      const uint32_t rb = a.rowbits[min(qi, a.N - 1)];
      const bool masked = (ref.flags & 1) || cls == 1 || KT.len < AB_T;
      mbar_wait(sdp_full, ph);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < AB_T / 32; ++cc) {
        uint32_t sv[32], dv[32];
        tmem_ld32(tS + lane_sel + cc * 32, sv);
        tmem_ld32(tdP + lane_sel + cc * 32, dv);
        tmem_ld_wait();
        float p[32], ds[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float pi = exp2f(__uint_as_float(sv[i]) * AB_LOG2E - lse2);
          if (masked && !(rb & keybit[cc * 32 + i])) pi = 0.f;
          p[i] = pi;
          ds[i] = pi * (__uint_as_float(dv[i]) - dlt);
        }
        store_tile_chunk(sP, r, cc, p);
        store_tile_chunk(sdS, r, cc, ds);
      }
      tc_fence_before();
      mbar_arrive(sdp_empty);
      fence_proxy_async_smem();
      mbar_arrive(pds_full);
      // ---- dQ tile: TMEM -> fp32 swizzled smem -> TMA reduce-add
      mbar_wait(dq_full, ph);
      tc_fence_after();
      if (r == 0) bulk_wait_group_read0();  // the previous reduce has finished reading sdQ
      ab_bar_sync(1, 128);
#pragma unroll
      for (int cc = 0; cc < AB_DH / 32; ++cc) {
        uint32_t v[32];
        tmem_ld32(tdQ + lane_sel + cc * 32, v);
        tmem_ld_wait();
        uint8_t* rowp = sdQ + cc * (AB_DQ / 2) + r * 128;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<uint4*>(rowp + ((q ^ (r & 7)) * 16)) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
      tc_fence_before();
      mbar_arrive(dq_empty);
      fence_proxy_async_smem();
      ab_bar_sync(1, 128);
      if (r == 0) {
        const int qrow = static_cast<int>(row0 + Q.start);
        tma_reduce_add_2d(&tm_dq, sdQ, h * AB_DH, qrow);
        tma_reduce_add_2d(&tm_dq, sdQ + AB_DQ / 2, h * AB_DH + 32, qrow);
        bulk_commit_group();
      }
    }
    // ---- epilogue: dK, dV of this key tile (thread = key row).  tcgen05.ld is warp-collective: every lane loads,
    // only rows inside the tile store.
    {
      const bool store = r < KT.len;
      __nv_bfloat16* drow = a.dqkv + (row0 + KT.start + (store ? r : 0)) * (3 * HD) + h * AB_DH;
      const float* uc = a.ucorr + static_cast<long long>(b) * HD + h * AB_DH;
#pragma unroll
      for (int which = 0; which < 2; ++which) {  // 0: dK -> column block 1, 1: dV -> column block 2
        __nv_bfloat16* dst = drow + (which + 1) * HD;
#pragma unroll
        for (int cc = 0; cc < AB_DH / 32; ++cc) {
          float v[32];
          if (n_iter > 0) {
            uint32_t t[32];
            tmem_ld32((which == 0 ? tdK : tdV) + lane_sel + cc * 32, t);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(t[i]);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.f;
          }
          if (which == 1) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += uc[cc * 32 + i];
          }
          if (store) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 w;
              w.x = pack_bf16x2(v[8 * q + 0], v[8 * q + 1]);
              w.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
              w.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]);
              w.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
              reinterpret_cast<uint4*>(dst + cc * 32)[q] = w;
            }
          }
        }
      }
    }
    if (r == 0) bulk_wait_group0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// delta[b,h,n] = sum_c dO*O ; ucorr[b, h*64+c] += dO[row, h*64+c] / N for rows whose lse is +inf (fully masked)
__global__ void __launch_bounds__(256)
attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout,
                     const float* __restrict__ lse, float* __restrict__ delta, float* __restrict__ ucorr, int B, int N,
                     int H) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= static_cast<long long>(B) * N) return;
  const int b = static_cast<int>(row / N), n = static_cast<int>(row % N);
  const int HD = H * AB_DH;  // 512: each lane owns 16 consecutive columns, 4 lanes per head
  const int c0 = lane * (HD / 32);
  const uint4* o4 = reinterpret_cast<const uint4*>(out + row * HD + c0);
  const uint4* d4 = reinterpret_cast<const uint4*>(dout + row * HD + c0);
  float dv[16];
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const uint4 o = o4[i], d = d4[i];
    const uint32_t ow[4] = {o.x, o.y, o.z, o.w}, dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      dv[8 * i + 2 * j] = bf16_lo(dw[j]), dv[8 * i + 2 * j + 1] = bf16_hi(dw[j]);
      acc += bf16_lo(ow[j]) * dv[8 * i + 2 * j] + bf16_hi(ow[j]) * dv[8 * i + 2 * j + 1];
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  const int h = c0 / AB_DH;
  const long long sidx = (static_cast<long long>(b) * H + h) * N + n;
  if ((lane & 3) == 0) delta[sidx] = acc;
  if (lse[sidx] == CUDART_INF_F) {
    const float inv = 1.0f / static_cast<float>(N);
#pragma unroll
    for (int i = 0; i < 16; ++i) atomicAdd(ucorr + static_cast<long long>(b) * HD + c0 + i, dv[i] * inv);
  }
}

}  // namespace mca

using namespace mca;

extern "C" int mca_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse,
                            const mca_attn_qtile* k_tiles_q, int n_kt, const mca_attn_ref* qt_list,
                            const mca_attn_tile* q_tiles, int n_qt, const uint32_t* rowbits, const uint8_t* keygrp,
                            const uint8_t* padding, const uint8_t* kt_class, float* delta, float* ucorr,
                            float* dq_accum, void* dqkv, int B, int N, int H, void* stream_) {
  (void)n_qt;
  if (B <= 0 || N <= 0 || H * AB_DH != 512) return MCA_ERR_SHAPE;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int HD = H * AB_DH;
  const long long M = static_cast<long long>(B) * N;
  CUtensorMap tm_qkv, tm_do, tm_dq;
  int rc = make_tmap_2d_bf16(&tm_qkv, qkv, 3 * HD, M, 3 * HD, AB_DH, AB_T);
  if (rc != MCA_OK) return rc;
  rc = make_tmap_2d_bf16(&tm_do, dout, HD, M, HD, AB_DH, AB_T);
  if (rc != MCA_OK) return rc;
  rc = make_tmap_2d_f32(&tm_dq, dq_accum, HD, M, HD, 32, AB_T);
  if (rc != MCA_OK) return rc;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM) != cudaSuccess)
      return MCA_ERR_CUDA;
    attr = true;
  }
  if (cudaMemsetAsync(dq_accum, 0, M * HD * sizeof(float), stream) != cudaSuccess) return MCA_ERR_CUDA;
  if (cudaMemsetAsync(ucorr, 0, static_cast<size_t>(B) * HD * sizeof(float), stream) != cudaSuccess) return MCA_ERR_CUDA;
  attn_bwd_prep_kernel<<<static_cast<unsigned>((M + 7) / 8), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(out), reinterpret_cast<const __nv_bfloat16*>(dout), lse, delta, ucorr, B, N, H);
  AttnBwdArgs a{k_tiles_q, qt_list, q_tiles, rowbits, keygrp, padding, kt_class, lse, delta, ucorr,
                reinterpret_cast<__nv_bfloat16*>(dqkv), N, H, n_kt};
  dim3 grid(n_kt, H, B);
  attn_bwd_kernel<<<grid, AB_THREADS, AB_SMEM, stream>>>(tm_qkv, tm_do, tm_dq, a);
  if (cudaGetLastError() != cudaSuccess) return MCA_ERR_CUDA;
  // dQ: fp32 accumulator -> bf16 first column block of dqkv
  return mca_cast_f32_bf16(dq_accum, HD, dqkv, 3 * HD, M, HD, stream_);
}
