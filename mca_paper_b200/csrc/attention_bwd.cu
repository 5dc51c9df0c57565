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
constexpr int AB_THREADS = 352;            // 2 compute warpgroups + TMA warp + MMA warp + dQ-reduce warp
constexpr int AB_TILE = AB_T * AB_DH * 2;  // 16 KB bf16 [128, 64]
constexpr int AB_DS = AB_T * AB_T * 2;     // 32 KB bf16 dS tile, stored [key][query] in two 64-query halves
constexpr int AB_DQ = AB_T * AB_DH * 4;    // 32 KB fp32 [128, 64]
constexpr int AB_QSTAGES = 3;
constexpr int AB_MAX_ITERS = 64;  // query tiles attending one key tile
// sK, sV, 3x(sQ, sdO), sdS (half 0 double buffered, half 1 single), sdQ; per-query staging and barriers are static
constexpr int AB_SMEM = 2 * AB_TILE + 2 * AB_QSTAGES * AB_TILE + 3 * (AB_DS / 2) + AB_DQ;
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

#ifdef MCA_TRACE
// debug-only timeline (built with MCA_NVCC_EXTRA=-DMCA_TRACE): clock64 stamps of one CTA, read by mca_debug_read_trace
__device__ long long g_trace[4 * 16 * 16 + 8];
#define TR(role, t, e) do { if (blockIdx.x == 5 && blockIdx.y == 3 && (t) < 16) g_trace[((role) * 16 + (t)) * 16 + (e)] = clock64(); } while (0)
#define TRG(e) do { if (blockIdx.x == 5 && blockIdx.y == 3) g_trace[4 * 16 * 16 + (e)] = clock64(); } while (0)
#else
#define TR(role, t, e) do { } while (0)
#define TRG(e) do { } while (0)
#endif

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                const __grid_constant__ CUtensorMap tm_dq, const AttnBwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];  // kept in the shared address space: no generic-pointer casts
  uint8_t* sK = smem;
  uint8_t* sV = sK + AB_TILE;
  uint8_t* sQ = sV + AB_TILE;                 // AB_QSTAGES stages
  uint8_t* sdO = sQ + AB_QSTAGES * AB_TILE;   // AB_QSTAGES stages
  uint8_t* sdS = sdO + AB_QSTAGES * AB_TILE;  // [half 0, buffer 0][half 0, buffer 1][half 1], 16 KB each
  uint8_t* sdQ = sdS + 3 * (AB_DS / 2);       // two 32-column fp32 boxes
  __shared__ int2 s_qt[AB_MAX_ITERS];                 // (start, len) of every query tile this CTA visits
  __shared__ __align__(16) float s_lse[2][2][64];     // [half][buffer][query]: -lse*log2e, -inf = masked
  __shared__ __align__(16) float s_dl[2][2][64];      // -delta
  __shared__ __align__(16) uint32_t s_rb[2][2][64];   // allowed-key-group bits (mixed-group tiles only)
  __shared__ uint64_t bars[24];
  __shared__ uint32_t tmem_holder_s;
  uint64_t* kv_full = bars + 0;
  uint64_t* qdo_full = bars + 1;   // [AB_QSTAGES]
  uint64_t* qdo_empty = bars + 4;  // [AB_QSTAGES]
  uint64_t* x_full = bars + 7;     // [2] per half: S^T / dP^T of the half are in TMEM
  uint64_t* x_free = bars + 9;     // [2] per half: the compute warpgroup has copied them to registers
  uint64_t* c_done = bars + 11;    // [2] per half: P^T / dS^T written (TMEM + smem)
  uint64_t* y_done = bars + 13;    // dV / dK products of one half retired: the P^T / dS^T columns are free
  uint64_t* z_full = bars + 14;    // [2] by tile parity: dQ product retired (dS consumed, dQ accumulator complete)
  uint64_t* dq_free = bars + 16;   // dQ accumulator copied to registers
  uint64_t* sdq_full = bars + 17;  // dQ tile staged in shared memory
  uint64_t* sdq_free = bars + 18;  // the TMA reduce has finished reading the staged tile
  uint32_t* tmem_holder = &tmem_holder_s;

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
    for (int s = 0; s < AB_QSTAGES; ++s) mbar_init(&qdo_full[s], 1), mbar_init(&qdo_empty[s], 1);
    for (int s = 0; s < 2; ++s) mbar_init(&x_full[s], 1), mbar_init(&x_free[s], 128), mbar_init(&c_done[s], 128);
    mbar_init(y_done, 1);
    mbar_init(&z_full[0], 1), mbar_init(&z_full[1], 1);
    mbar_init(dq_free, 128);
    mbar_init(sdq_full, 128);
    mbar_init(sdq_free, 1);
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(tmem_holder, 512);
  for (int i = threadIdx.x; i < n_iter; i += AB_THREADS) {
    const mca_attn_tile Q = a.q_tiles[a.qt_list[KT.kt_off + i].tile];
    s_qt[i] = make_int2(Q.start, Q.len);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  // columns: [0,128) S^T_0 | dP^T_0, [128,256) S^T_1 | dP^T_1, [256,320) P^T (32) | dS^T (32) of the half in flight,
  //          [320,384) dV, [384,448) dK, [448,512) dQ
  const uint32_t tP = tmem_base + 256, tdV = tmem_base + 320, tdK = tmem_base + 384, tdQ = tmem_base + 448;
  if (threadIdx.x == 0) TRG(0);

  if (warp == 8) {
    // ===================== TMA producer (whole warp loops, one elected lane issues) =====================
    if (n_iter > 0) {
      const int krow = static_cast<int>(row0 + KT.start);
      if (elect_one()) {
        mbar_expect_tx(kv_full, 2 * AB_TILE);
        tma_load_2d(sK, &tm_qkv, kv_full, HD + h * AB_DH, krow);
        tma_load_2d(sV, &tm_qkv, kv_full, 2 * HD + h * AB_DH, krow);
      }
      __syncwarp();
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % AB_QSTAGES;
        const uint32_t sph = (it / AB_QSTAGES) & 1;
        const int qrow = static_cast<int>(row0 + s_qt[it].x);
        mbar_wait(&qdo_empty[s], sph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&qdo_full[s], 2 * AB_TILE);
          tma_load_2d(sQ + s * AB_TILE, &tm_qkv, &qdo_full[s], h * AB_DH, qrow);
          tma_load_2d(sdO + s * AB_TILE, &tm_do, &qdo_full[s], h * AB_DH, qrow);
        }
        __syncwarp();
      }
    }
  } else if (warp == 10) {
    // ===================== dQ reduce issuer: staged fp32 tile -> TMA reduce-add into dq_acc =====================
    for (int t = 0; t < n_iter; ++t) {
      const int qrow = static_cast<int>(row0 + s_qt[t].x);
      mbar_wait(sdq_full, t & 1);
      if (elect_one()) {  // same membermask every time -> same leader, which owns the bulk async-groups
        tma_reduce_add_2d(&tm_dq, sdQ, h * AB_DH, qrow);
        tma_reduce_add_2d(&tm_dq, sdQ + AB_DQ / 2, h * AB_DH + 32, qrow);
        bulk_commit_group();
        bulk_wait_group_read0();
        mbar_arrive(sdq_free);
      }
      __syncwarp();
    }
    if (elect_one()) bulk_wait_group0();
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    // The whole warp runs the loop and the waits (convergent code keeps the descriptors in uniform registers, so a
    // tcgen05.mma costs one uniform add); one elected lane issues the MMAs and commits.  Independent accumulation
    // chains are interleaved (S^T with dP^T, dV with dK).
    if (n_iter > 0) {
      constexpr uint32_t id_x = make_idesc_bf16(AB_T, 64, false, false);    // S^T, dP^T: K-major x K-major, N = 64
      constexpr uint32_t id_y = make_idesc_bf16(AB_T, AB_DH, false, true);  // dV, dK: A from TMEM, B MN-major
      constexpr uint32_t id_z = make_idesc_bf16(AB_T, AB_DH, true, true);   // dQ: MN-major x MN-major
      const uint64_t dk_k = make_smem_desc_sw128(smem_u32(sK), 16, 1024);     // K as the K-major A operand of S^T
      const uint64_t dv_k = make_smem_desc_sw128(smem_u32(sV), 16, 1024);     // V as the K-major A operand of dP^T
      const uint64_t dk_mn = make_smem_desc_sw128(smem_u32(sK), 8192, 1024);  // K as the MN-major B operand of dQ
      const uint64_t dq_k0 = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
      const uint64_t do_k0 = make_smem_desc_sw128(smem_u32(sdO), 16, 1024);
      const uint64_t dq_mn0 = make_smem_desc_sw128(smem_u32(sQ), 8192, 1024);
      const uint64_t do_mn0 = make_smem_desc_sw128(smem_u32(sdO), 8192, 1024);
      auto issue_x = [&](int t, int hf) {
        const uint64_t off = static_cast<uint64_t>(((t % AB_QSTAGES) * AB_TILE + hf * 8192) >> 4);
        const uint32_t reg = tmem_base + hf * 128;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < AB_DH / 16; ++k) {
            umma_bf16(reg, dk_k + k * 2, dq_k0 + off + k * 2, id_x, k > 0 ? 1u : 0u);
            umma_bf16(reg + 64, dv_k + k * 2, do_k0 + off + k * 2, id_x, k > 0 ? 1u : 0u);
          }
          umma_commit(&x_full[hf]);
        }
        __syncwarp();
      };
      auto issue_y = [&](int t, int hf, bool last_of_tile) {
        const uint64_t off = static_cast<uint64_t>(((t % AB_QSTAGES) * AB_TILE + hf * 8192) >> 4);
        const uint32_t acc = (t > 0 || hf > 0) ? 1u : 0u;
        if (elect_one()) {
          // contraction over the 64 queries of this half: 4 steps of 16 query rows (2 KB of the MN-major B tile)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_bf16_ts(tdV, tP + k * 8, do_mn0 + off + k * 128, id_y, k > 0 ? 1u : acc);
            umma_bf16_ts(tdK, tP + 32 + k * 8, dq_mn0 + off + k * 128, id_y, k > 0 ? 1u : acc);
          }
          umma_commit(y_done);
          if (last_of_tile) umma_commit(&qdo_empty[t % AB_QSTAGES]);  // Q / dO of this tile are no longer read
        }
        __syncwarp();
      };
      mbar_wait(kv_full, 0);
      if (lane == 0) TRG(1);
      mbar_wait(&qdo_full[0], 0);
      if (lane == 0) TRG(2);
      tc_fence_after();
      issue_x(0, 0);
      issue_x(0, 1);
      for (int t = 0; t < n_iter; ++t) {
        const uint32_t ph = t & 1;
        const bool more = t + 1 < n_iter;
        if (lane == 0) TR(2, t, 0);
        if (more) {
          mbar_wait(&qdo_full[(t + 1) % AB_QSTAGES], ((t + 1) / AB_QSTAGES) & 1);
          mbar_wait(&x_free[0], ph);  // S^T_0 / dP^T_0 of tile t are in registers: overwrite them right away
          tc_fence_after();
          issue_x(t + 1, 0);
        }
        if (lane == 0) TR(2, t, 1);
        mbar_wait(&c_done[0], ph);
        if (lane == 0) TR(2, t, 2);
        tc_fence_after();
        issue_y(t, 0, false);
        if (lane == 0) TR(2, t, 3);
        if (more) {
          mbar_wait(&x_free[1], ph);
          tc_fence_after();
          issue_x(t + 1, 1);
        }
        if (lane == 0) TR(2, t, 4);
        mbar_wait(&c_done[1], ph);
        if (lane == 0) TR(2, t, 5);
        tc_fence_after();
        issue_y(t, 1, true);
        if (t > 0) {
          mbar_wait(dq_free, (t - 1) & 1);
          tc_fence_after();
        }
        if (lane == 0) TR(2, t, 6);
        {  // dQ = dS K: contraction over the 128 keys (8 steps of 16 key rows); the two 64-query halves of dS
           // live in separate buffers: half 0 at buffer (t & 1), half 1 behind both (always a positive offset)
          const uint64_t dds = make_smem_desc_sw128(smem_u32(sdS) + (t & 1) * (AB_DS / 2), (2 - (t & 1)) * (AB_DS / 2), 1024);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < AB_T / 16; ++k) umma_bf16(tdQ, dds + k * 128, dk_mn + k * 128, id_z, k > 0 ? 1u : 0u);
            umma_commit(&z_full[t & 1]);
          }
          __syncwarp();
        }
        if (lane == 0) TR(2, t, 7);
      }
    }
  } else if (warp < 8) {
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
    const bool dead = has_dead && !live;
    const long long sbase = (static_cast<long long>(b) * a.H + h) * a.N;
    // raw per-query values of tile t go global -> shared with cp.async (no registers held across an iteration)
    auto stage_async = [&](int t) {
      if (wt < 64 && t < n_iter) {
        const int qi = min(s_qt[t].x + hf * 64 + wt, a.N - 1);
        cp_async4(&s_lse[hf][t & 1][wt], a.lse + sbase + qi);
        cp_async4(&s_dl[hf][t & 1][wt], a.delta + sbase + qi);
        cp_async4(&s_rb[hf][t & 1][wt], a.rowbits + qi);
      }
      cp_async_commit();
    };
    auto drain_dq = [&](int tp) {  // dQ of tile tp: TMEM -> registers -> fp32 swizzled smem (the reduce warp ships it)
      mbar_wait(&z_full[tp & 1], (tp >> 1) & 1);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld32(tdQ + lane_sel, v0);
      tmem_ld32(tdQ + lane_sel + 32, v1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(dq_free);
      if (tp > 0) mbar_wait(sdq_free, (tp - 1) & 1);
      uint8_t* rowp = sdQ + r * 128;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        *reinterpret_cast<uint4*>(rowp + ((q ^ (r & 7)) << 4)) = make_uint4(v0[4 * q], v0[4 * q + 1], v0[4 * q + 2], v0[4 * q + 3]);
        *reinterpret_cast<uint4*>(rowp + AB_DQ / 2 + ((q ^ (r & 7)) << 4)) = make_uint4(v1[4 * q], v1[4 * q + 1], v1[4 * q + 2], v1[4 * q + 3]);
      }
      fence_proxy_async_smem();
      mbar_arrive(sdq_full);
    };
    stage_async(0);
    for (int t = 0; t < n_iter; ++t) {
      const uint32_t ph = t & 1;
      if (wt == 0) TR(hf, t, 0);
      cp_async_wait_all();
      if (wt < 64) {  // fix up the staged values in place: fold the group mask into lse, pre-negate for the FMAs
        const int qr = hf * 64 + wt;
        const bool valid = qr < s_qt[t].y;
        const uint32_t rb = valid ? s_rb[hf][t & 1][wt] : 0u;
        const bool sees = valid && (kgrp == 255 || ((rb >> kgrp) & 1u));
        s_lse[hf][t & 1][wt] = sees ? -s_lse[hf][t & 1][wt] * AB_LOG2E : -CUDART_INF_F;
        s_dl[hf][t & 1][wt] = valid ? -s_dl[hf][t & 1][wt] : 0.f;
        s_rb[hf][t & 1][wt] = rb;
      }
      ab_bar_sync(1 + hf, 128);
      stage_async(t + 1);  // after the barrier: every thread is done with the other buffer (tile t-1)
      if (wt == 0) TR(hf, t, 1);
      mbar_wait(&x_full[hf], ph);
      if (wt == 0) TR(hf, t, 2);
      tc_fence_after();
      uint32_t sv[2][32], dv[2][32];
      tmem_ld32(reg, sv[0]);
      tmem_ld32(reg + 32, sv[1]);
      tmem_ld32(reg + 64, dv[0]);
      tmem_ld32(reg + 96, dv[1]);
      tmem_ld_wait();
      if (wt == 0) TR(hf, t, 3);
      tc_fence_before();
      mbar_arrive(&x_free[hf]);  // the next tile's S^T / dP^T of this half may be issued now
      const float4* lse4 = reinterpret_cast<const float4*>(s_lse[hf][t & 1]);
      const float4* dl4 = reinterpret_cast<const float4*>(s_dl[hf][t & 1]);
      const uint32_t* rbq = s_rb[hf][t & 1];
      // P^T = exp2(S^T*log2e - lse2[q]),  dS^T = P^T * (dP^T - delta[q])   (in place in sv / dv)
#pragma unroll
      for (int g = 0; g < 16; ++g) {
        const float4 l4 = lse4[g], d4 = dl4[g];
        const float lq[4] = {l4.x, l4.y, l4.z, l4.w}, dq[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int e = g * 4 + i;
          const float pi = fast_ex2(fmaf(__uint_as_float(sv[e >> 5][e & 31]), AB_LOG2E, lq[i]));
          const float di = pi * (__uint_as_float(dv[e >> 5][e & 31]) + dq[i]);
          sv[e >> 5][e & 31] = __float_as_uint(pi);
          dv[e >> 5][e & 31] = __float_as_uint(di);
        }
      }
      if (kgrp == 255) {  // mixed key groups (fusion sub-blocks): per-(query, key) visibility
#pragma unroll
        for (int e = 0; e < 64; ++e)
          if (!((rbq[e] >> mygrp) & 1u)) sv[e >> 5][e & 31] = 0u, dv[e >> 5][e & 31] = 0u;
      }
      if (dead) {
#pragma unroll
        for (int e = 0; e < 64; ++e) sv[e >> 5][e & 31] = 0u, dv[e >> 5][e & 31] = 0u;
      }
      uint32_t pp[32], dd[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int e = 2 * j;
        pp[j] = pack_bf16x2(__uint_as_float(sv[e >> 5][e & 31]), __uint_as_float(sv[(e + 1) >> 5][(e + 1) & 31]));
        dd[j] = pack_bf16x2(__uint_as_float(dv[e >> 5][e & 31]), __uint_as_float(dv[(e + 1) >> 5][(e + 1) & 31]));
      }
      if (wt == 0) TR(hf, t, 8);
      // dS^T -> shared memory for the dQ product.  Half 0 is double buffered by tile parity; half 1 has one buffer
      // that dQ(t-1) must have finished reading.
      uint8_t* ds_row;
      if (hf == 0) {
        if (t >= 2) mbar_wait(&z_full[t & 1], ((t - 2) >> 1) & 1);
        ds_row = sdS + (t & 1) * (AB_DS / 2) + r * 128;
      } else {
        if (t >= 1) mbar_wait(&z_full[(t - 1) & 1], ((t - 1) >> 1) & 1);
        ds_row = sdS + 2 * (AB_DS / 2) + r * 128;
      }
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(ds_row + ((c ^ (r & 7)) << 4)) = make_uint4(dd[4 * c], dd[4 * c + 1], dd[4 * c + 2], dd[4 * c + 3]);
      // the P^T / dS^T columns are shared by both halves: wait until the previous half's dV / dK products retired
      const int yk = 2 * t + hf - 1;  // index of that commit on y_done
      if (wt == 0) TR(hf, t, 4);
      if (yk >= 0) {
        mbar_wait(y_done, yk & 1);
        tc_fence_after();
      }
      if (wt == 0) TR(hf, t, 5);
      tmem_st32(tP + lane_sel, pp);
      tmem_st32(tP + 32 + lane_sel, dd);
      tmem_st_wait();
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&c_done[hf]);
      if (wt == 0) TR(hf, t, 6);
      // the two warpgroups take turns draining dQ: tile tp is handled by warpgroup tp & 1 one iteration later
      if (t > 0 && ((t - 1) & 1) == hf) drain_dq(t - 1);
      if (wt == 0) TR(hf, t, 7);
    }
    if (n_iter > 0) {
      const int tp = n_iter - 1;
      if ((tp & 1) == hf) drain_dq(tp);
      mbar_wait(&z_full[tp & 1], (tp >> 1) & 1);  // the last dQ product retired => every MMA of this CTA retired
      tc_fence_after();
    }
    // ---- epilogue: warpgroup 0 writes dK, warpgroup 1 writes dV (+ the uniform-row correction); thread = key row.
    // tcgen05.ld is warp-collective: every lane loads, only rows inside the tile store.
    if (wt == 0) TRG(3 + hf);
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
    if (wt == 0) TRG(5 + hf);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) TRG(7);
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
  if (B <= 0 || N <= 0 || H * AB_DH != 512 || n_qt <= 0 || n_qt > AB_MAX_ITERS) return MCA_ERR_SHAPE;
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

#ifdef MCA_TRACE
extern "C" int mca_debug_read_trace(long long* host_dst, int n) {
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(host_dst, mca::g_trace, sizeof(long long) * n) == cudaSuccess ? 0 : 3;
}
#endif
