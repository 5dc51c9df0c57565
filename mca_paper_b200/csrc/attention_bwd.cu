// Block-sparse masked attention backward on tcgen05 (autograd of model.py:85-100).
//
// One CTA = one (sample, head, 128-key tile); it walks the query tiles that attend this key tile (the transpose of
// the forward schedule).  Scores are computed TRANSPOSED (keys on the TMEM lanes, queries along the columns) so
// that P^T and dS^T can be fed back to the tensor core straight from TMEM as the A operands of the dV and dK
// products; only dS also goes to shared memory (for dQ).  Each query tile is processed as two halves of 64 queries
// that ping-pong between two compute warpgroups, so the tensor pipe always has the other half's products to run
// while one half is in the exp / multiply stage:
//     X(t,h): S^T  = K Q_h^T          dP^T = V dO_h^T            (128 x 64 x 64, smem x smem -> TMEM region h)
//     C(t,h): P^T  = exp2(S^T*log2e - lse_q*log2e)   dS^T = P^T * (dP^T - delta_q)     (thread = key row)
//     Y(t,h): dV  += P^T dO_h         dK  += dS^T Q_h            (A from TMEM, B = the TMA tiles read MN-major)
//     Z(t)  : dQ   = dS K             (128 x 64 x 128; fresh tile -> fp32 smem -> TMA reduce-add into dq_acc)
// dK/dV stay resident in TMEM across the whole loop and are written once.  Q/dO tiles are double buffered.
// Masking costs nothing per element for the common tiles: a query that may not see this tile's key group gets
// lse = +inf (P = 0) when its lse is staged; only tiles with padded / missing keys or mixed key groups (the fusion
// sub-blocks) take a per-element select.  Fully masked query rows carry lse = +inf from the forward; their
// uniform-1/N contribution to dV (reference quirk Q4) is the per-(sample, head) vector `ucorr`, added in the epilogue.
#include <math_constants.h>

#include "mca_b200.h"
#include "ptx.cuh"
#include "runtime.h"

namespace mca {

constexpr int AB_T = 128;
constexpr int AB_DH = 64;
constexpr int AB_THREADS = 320;            // 2 compute warpgroups + TMA warp + MMA warp
constexpr int AB_TILE = AB_T * AB_DH * 2;  // 16 KB bf16 [128, 64]
constexpr int AB_DS = AB_T * AB_T * 2;     // 32 KB bf16 dS tile, stored [key][query] in two 64-query halves
constexpr int AB_DQ = AB_T * AB_DH * 4;    // 32 KB fp32 [128, 64]
// sK, sV, 2x(sQ, sdO), 2x sdS, sdQ, per-query staging + barriers
constexpr int AB_SMEM = 2 * AB_TILE + 4 * AB_TILE + 2 * AB_DS + AB_DQ + 1024 /*align*/ + 4096;
constexpr float AB_LOG2E = 1.4426950408889634f;

struct AttnBwdArgs {
  const mca_attn_qtile* k_tiles_q;  // per key tile: start, len, slice of qt_list
  const mca_attn_ref* qt_list;
  const mca_attn_tile* q_tiles;
  const uint32_t* rowbits;
  const uint8_t* keygrp;
  const uint8_t* tile_grp;
  const uint8_t* padding;
  const uint8_t* kt_class;
  const float* lse;     // [B,H,N]
  const float* delta;   // [B,H,N]
  const float* ucorr;   // [B, H*64]
  __nv_bfloat16* dqkv;  // [B*N, 3*H*64]
  int N, H, n_kt;
};

__device__ __forceinline__ void ab_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                const __grid_constant__ CUtensorMap tm_dq, const AttnBwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = base;
  uint8_t* sV = sK + AB_TILE;
  uint8_t* sQ = sV + AB_TILE;        // 2 stages
  uint8_t* sdO = sQ + 2 * AB_TILE;   // 2 stages
  uint8_t* sdS = sdO + 2 * AB_TILE;  // 2 buffers
  uint8_t* sdQ = sdS + 2 * AB_DS;
  float* s_lse = reinterpret_cast<float*>(sdQ + AB_DQ);  // [2 halves][2 buffers][64]
  float* s_dl = s_lse + 256;
  uint32_t* s_rb = reinterpret_cast<uint32_t*>(s_dl + 256);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_rb + 256);
  uint64_t* kv_full = bars + 0;
  uint64_t* qdo_full = bars + 1;   // [2]
  uint64_t* qdo_empty = bars + 3;  // [2]
  uint64_t* x_full = bars + 5;     // [2] per half
  uint64_t* c_done = bars + 7;     // [2] per half
  uint64_t* z_full = bars + 9;     // [2] per dQ buffer
  uint64_t* dq_free = bars + 11;   // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x % a.H, b = blockIdx.x / a.H;
  const int kt = blockIdx.y;
  const mca_attn_qtile KT = a.k_tiles_q[kt];
  const long long row0 = static_cast<long long>(b) * a.N;
  const int cls = a.kt_class[static_cast<long long>(b) * a.n_kt + kt];
  const int n_iter = cls == 2 ? 0 : KT.kt_cnt;
  const int HD = a.H * AB_DH;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_dq);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&qdo_full[s], 1), mbar_init(&qdo_empty[s], 1);
      mbar_init(&x_full[s], 1), mbar_init(&c_done[s], 128);
      mbar_init(&z_full[s], 1), mbar_init(&dq_free[s], 128);
    }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(tmem_holder, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  // region h: [S^T_h | dP^T_h] (64 + 64 fp32 columns), later [P^T_h (32) .. | dS^T_h (32) ..]
  const uint32_t tdV = tmem_base + 256, tdK = tmem_base + 320, tdQ = tmem_base + 384;  // tdQ: 2 x 64 columns

  if (warp == 8) {
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
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    if (lane == 0 && n_iter > 0) {
      constexpr uint32_t id_x = make_idesc_bf16(AB_T, 64, false, false);    // S^T, dP^T: K-major x K-major, N = 64
      constexpr uint32_t id_y = make_idesc_bf16(AB_T, AB_DH, false, true);  // dV, dK: A from TMEM, B MN-major
      constexpr uint32_t id_z = make_idesc_bf16(AB_T, AB_DH, true, true);   // dQ: MN-major x MN-major
      const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV);
      auto issue_x = [&](int t, int hf) {
        const int s = t & 1;
        const uint32_t q_addr = smem_u32(sQ + s * AB_TILE) + hf * 8192, do_addr = smem_u32(sdO + s * AB_TILE) + hf * 8192;
        const uint32_t reg = tmem_base + hf * 128;
#pragma unroll
        for (int k = 0; k < AB_DH / 16; ++k)
          umma_bf16(reg, make_smem_desc_sw128(k_addr + k * 32, 16, 1024), make_smem_desc_sw128(q_addr + k * 32, 16, 1024),
                    id_x, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < AB_DH / 16; ++k)
          umma_bf16(reg + 64, make_smem_desc_sw128(v_addr + k * 32, 16, 1024),
                    make_smem_desc_sw128(do_addr + k * 32, 16, 1024), id_x, k > 0 ? 1u : 0u);
        umma_commit(&x_full[hf]);
      };
      auto issue_y = [&](int t, int hf) {
        const int s = t & 1;
        const uint32_t q_addr = smem_u32(sQ + s * AB_TILE) + hf * 8192, do_addr = smem_u32(sdO + s * AB_TILE) + hf * 8192;
        const uint32_t reg = tmem_base + hf * 128;
        // contraction over the 64 queries of this half: 4 steps of 16 query rows (2 KB of the MN-major B tile)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ts(tdV, reg + k * 8, make_smem_desc_sw128(do_addr + k * 2048, 8192, 1024), id_y,
                       (t > 0 || hf > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ts(tdK, reg + 64 + k * 8, make_smem_desc_sw128(q_addr + k * 2048, 8192, 1024), id_y,
                       (t > 0 || hf > 0 || k > 0) ? 1u : 0u);
      };
      mbar_wait(kv_full, 0);
      mbar_wait(&qdo_full[0], 0);
      tc_fence_after();
      issue_x(0, 0);
      issue_x(0, 1);
      for (int t = 0; t < n_iter; ++t) {
        const uint32_t ph = t & 1;
        mbar_wait(&c_done[0], ph);
        tc_fence_after();
        issue_y(t, 0);
        if (t + 1 < n_iter) {
          mbar_wait(&qdo_full[(t + 1) & 1], ((t + 1) >> 1) & 1);
          tc_fence_after();
          issue_x(t + 1, 0);
        }
        mbar_wait(&c_done[1], ph);
        tc_fence_after();
        issue_y(t, 1);
        umma_commit(&qdo_empty[t & 1]);  // Q / dO of this tile are no longer read once these retire
        mbar_wait(&dq_free[t & 1], ((t >> 1) & 1) ^ 1);
        tc_fence_after();
        {  // dQ = dS K: contraction over the 128 keys (8 steps of 16 key rows)
          const uint32_t ds_addr = smem_u32(sdS + (t & 1) * AB_DS);
#pragma unroll
          for (int k = 0; k < AB_T / 16; ++k)
            umma_bf16(tdQ + (t & 1) * 64, make_smem_desc_sw128(ds_addr + k * 2048, AB_DS / 2, 1024),
                      make_smem_desc_sw128(k_addr + k * 2048, 8192, 1024), id_z, k > 0 ? 1u : 0u);
          umma_commit(&z_full[t & 1]);
        }
        if (t + 1 < n_iter) issue_x(t + 1, 1);
      }
    }
  } else {
    // ===================== compute warpgroups: thread = key row, warpgroup = query half =====================
    const int hf = warp >> 2;             // which 64-query half of every tile
    const int r = (warp & 3) * 32 + lane;  // key row = TMEM lane
    const int wt = threadIdx.x & 127;      // thread index inside the warpgroup
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t reg = tmem_base + hf * 128 + lane_sel;
    const int kgrp = a.tile_grp[kt];
    const bool has_dead = cls == 1 || KT.len < AB_T;
    bool live = false;
    uint32_t mygrp = 0;
    {
      const int kj = KT.start + r;
      if (r < KT.len) {
        live = a.padding[row0 + kj] == 0;
        mygrp = a.keygrp[kj];
      }
    }
    for (int t = 0; t < n_iter; ++t) {
      const uint32_t ph = t & 1;
      const mca_attn_ref ref = a.qt_list[KT.kt_off + t];
      const mca_attn_tile Q = a.q_tiles[ref.tile];
      float* my_lse = s_lse + (hf * 2 + (t & 1)) * 64;
      float* my_dl = s_dl + (hf * 2 + (t & 1)) * 64;
      uint32_t* my_rb = s_rb + (hf * 2 + (t & 1)) * 64;
      if (wt < 64) {  // stage this half's per-query lse (with the group mask folded in), delta and row bits
        const int qr = hf * 64 + wt;
        const int qi = Q.start + qr;
        float l2 = CUDART_INF_F, dl = 0.f;
        uint32_t rb = 0;
        if (qr < Q.len) {
          const long long sidx = (static_cast<long long>(b) * a.H + h) * a.N + qi;
          rb = a.rowbits[qi];
          if (kgrp == 255 || ((rb >> kgrp) & 1u)) l2 = a.lse[sidx] * AB_LOG2E;
          dl = a.delta[sidx];
        }
        my_lse[wt] = l2, my_dl[wt] = dl, my_rb[wt] = rb;
      }
      ab_bar_sync(1 + hf, 128);
      mbar_wait(&x_full[hf], ph);
      tc_fence_after();
      uint32_t sv[2][32], dv[2][32];
      tmem_ld32(reg, sv[0]);
      tmem_ld32(reg + 32, sv[1]);
      tmem_ld32(reg + 64, dv[0]);
      tmem_ld32(reg + 96, dv[1]);
      tmem_ld_wait();
      if (t >= 2) mbar_wait(&z_full[t & 1], ((t - 2) >> 1) & 1);  // dS buffer (t&1) has been consumed by dQ(t-2)
      uint8_t* ds_row = sdS + (t & 1) * AB_DS + hf * (AB_DS / 2) + r * 128;
      uint32_t pp[32], dd[32];
#pragma unroll
      for (int c = 0; c < 8; ++c) {  // 8 queries per step
        float p[8], ds[8];
#pragma unroll
        for (int g4 = 0; g4 < 2; ++g4) {
          const float4 l4 = *reinterpret_cast<const float4*>(my_lse + c * 8 + g4 * 4);
          const float4 d4 = *reinterpret_cast<const float4*>(my_dl + c * 8 + g4 * 4);
          const float lq[4] = {l4.x, l4.y, l4.z, l4.w}, dq[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int e = c * 8 + g4 * 4 + i;
            float pi = fast_ex2(fmaf(__uint_as_float(sv[e >> 5][e & 31]), AB_LOG2E, -lq[i]));
            if (kgrp == 255 && !((my_rb[e] >> mygrp) & 1u)) pi = 0.f;
            p[g4 * 4 + i] = pi;
            ds[g4 * 4 + i] = pi * (__uint_as_float(dv[e >> 5][e & 31]) - dq[i]);
          }
        }
        if (has_dead && !live) {
#pragma unroll
          for (int i = 0; i < 8; ++i) p[i] = 0.f, ds[i] = 0.f;
        }
        uint4 w;
        w.x = pack_bf16x2(ds[0], ds[1]), w.y = pack_bf16x2(ds[2], ds[3]);
        w.z = pack_bf16x2(ds[4], ds[5]), w.w = pack_bf16x2(ds[6], ds[7]);
        dd[4 * c] = w.x, dd[4 * c + 1] = w.y, dd[4 * c + 2] = w.z, dd[4 * c + 3] = w.w;
        pp[4 * c] = pack_bf16x2(p[0], p[1]), pp[4 * c + 1] = pack_bf16x2(p[2], p[3]);
        pp[4 * c + 2] = pack_bf16x2(p[4], p[5]), pp[4 * c + 3] = pack_bf16x2(p[6], p[7]);
        *reinterpret_cast<uint4*>(ds_row + ((c ^ (r & 7)) << 4)) = w;
      }
      tmem_st32(reg, pp);        // P^T_h over the first 32 columns of S^T_h
      tmem_st32(reg + 64, dd);   // dS^T_h over the first 32 columns of dP^T_h
      tmem_st_wait();
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&c_done[hf]);
      // ---- warpgroup 1 drains dQ of the previous tile: TMEM -> fp32 swizzled smem -> TMA reduce-add
      if (hf == 1 && t > 0) {
        const int tp = t - 1;
        const mca_attn_tile Qp = a.q_tiles[a.qt_list[KT.kt_off + tp].tile];
        mbar_wait(&z_full[tp & 1], (tp >> 1) & 1);
        tc_fence_after();
        if (wt == 0) bulk_wait_group_read0();  // the previous reduce has finished reading sdQ
        ab_bar_sync(2, 128);
#pragma unroll
        for (int cc = 0; cc < AB_DH / 32; ++cc) {
          uint32_t v[32];
          tmem_ld32(tdQ + (tp & 1) * 64 + lane_sel + cc * 32, v);
          tmem_ld_wait();
          uint8_t* rowp = sdQ + cc * (AB_DQ / 2) + r * 128;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<uint4*>(rowp + ((q ^ (r & 7)) * 16)) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
        tc_fence_before();
        mbar_arrive(&dq_free[tp & 1]);
        fence_proxy_async_smem();
        ab_bar_sync(2, 128);
        if (wt == 0) {
          const int qrow = static_cast<int>(row0 + Qp.start);
          tma_reduce_add_2d(&tm_dq, sdQ, h * AB_DH, qrow);
          tma_reduce_add_2d(&tm_dq, sdQ + AB_DQ / 2, h * AB_DH + 32, qrow);
          bulk_commit_group();
        }
      }
    }
    if (n_iter > 0) {
      const int tp = n_iter - 1;
      mbar_wait(&z_full[tp & 1], (tp >> 1) & 1);  // the last dQ product retired => every MMA of this CTA retired
      tc_fence_after();
      if (hf == 1) {
        const mca_attn_tile Qp = a.q_tiles[a.qt_list[KT.kt_off + tp].tile];
        if (wt == 0) bulk_wait_group_read0();
        ab_bar_sync(2, 128);
#pragma unroll
        for (int cc = 0; cc < AB_DH / 32; ++cc) {
          uint32_t v[32];
          tmem_ld32(tdQ + (tp & 1) * 64 + lane_sel + cc * 32, v);
          tmem_ld_wait();
          uint8_t* rowp = sdQ + cc * (AB_DQ / 2) + r * 128;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<uint4*>(rowp + ((q ^ (r & 7)) * 16)) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
        fence_proxy_async_smem();
        ab_bar_sync(2, 128);
        if (wt == 0) {
          const int qrow = static_cast<int>(row0 + Qp.start);
          tma_reduce_add_2d(&tm_dq, sdQ, h * AB_DH, qrow);
          tma_reduce_add_2d(&tm_dq, sdQ + AB_DQ / 2, h * AB_DH + 32, qrow);
          bulk_commit_group();
        }
      }
    }
    // ---- epilogue: warpgroup 0 writes dK, warpgroup 1 writes dV (+ the uniform-row correction); thread = key row.
    // tcgen05.ld is warp-collective: every lane loads, only rows inside the tile store.
    {
      const bool store = r < KT.len;
      const int which = hf;  // 0: dK -> column block 1, 1: dV -> column block 2
      __nv_bfloat16* dst = a.dqkv + (row0 + KT.start + (store ? r : 0)) * (3 * HD) + h * AB_DH + (which + 1) * HD;
      const float* uc = a.ucorr + static_cast<long long>(b) * HD + h * AB_DH;
#pragma unroll
      for (int cc = 0; cc < AB_DH / 32; ++cc) {
        float v[32];
        if (n_iter > 0) {
          uint32_t tt[32];
          tmem_ld32((which == 0 ? tdK : tdV) + lane_sel + cc * 32, tt);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(tt[i]);
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
    if (hf == 1 && wt == 0) bulk_wait_group0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
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
                            const uint8_t* tile_grp, const uint8_t* padding, const uint8_t* kt_class, float* delta,
                            float* ucorr, float* dq_accum, void* dqkv, int B, int N, int H, void* stream_) {
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
  AttnBwdArgs a{k_tiles_q, qt_list, q_tiles, rowbits, keygrp, tile_grp, padding, kt_class, lse, delta, ucorr,
                reinterpret_cast<__nv_bfloat16*>(dqkv), N, H, n_kt};
  dim3 grid(B * H, n_kt);
  attn_bwd_kernel<<<grid, AB_THREADS, AB_SMEM, stream>>>(tm_qkv, tm_do, tm_dq, a);
  if (cudaGetLastError() != cudaSuccess) return MCA_ERR_CUDA;
  // dQ: fp32 accumulator -> bf16 first column block of dqkv
  return mca_cast_f32_bf16(dq_accum, HD, dqkv, 3 * HD, M, HD, stream_);
}
