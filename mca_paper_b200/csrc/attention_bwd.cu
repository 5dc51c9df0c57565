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
// The per-query terms ride on the tensor core: both score products get one extra k-step whose B rows hold
// (-lse, -delta) of the query, each split into three bf16 terms (exact to 2^-25), against constant 0/1 A rows, so
// the accumulators arrive as S^T - lse_q and dP^T - delta_q and the compute warps never load a per-query value
// (those broadcast shared-memory loads were a third of the kernel's shared-memory traffic, which is what bounds it).
// Masking costs nothing per element for the common tiles: a query that may not see this tile's key group gets
// -lse = -60000 (P = 0) in that row; only tiles with padded / missing keys or mixed key groups (the fusion
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
constexpr int AB_MAX_ITERS = 128;  // upper bound of query tiles attending one key tile (the launcher checks the total)
// sK, sV, 3x(sQ, sdO), sdS (half 0 double buffered, half 1 single), sdQ, extra k-step operands; barriers are static
// extra k-step operands (K-major, no swizzle: 8-row x 16-byte core matrices).  Only the first 8-element k chunk of a
// row carries data; the second chunk of EVERY operand is the same all-zero block (the descriptor's leading-dimension
// offset points each tile at it), so a [128 x 16] operand costs 2 KB: A rows for S^T, A rows for dP^T, B rows of the
// 128 queries of a tile (double buffered by tile parity), zero block.
constexpr int AB_EXT_TILE = AB_T * 16;
constexpr int AB_EXT = 5 * AB_EXT_TILE;
constexpr int AB_SMEM = 2 * AB_TILE + 2 * AB_QSTAGES * AB_TILE + 3 * (AB_DS / 2) + AB_DQ + AB_EXT;
constexpr float AB_MASKED = -60000.f;  // -lse of a query that must not see this key tile: exp2 underflows to 0
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
__device__ long long g_bcta[4096 * 4];  // (globaltimer start, end, smid, n_iter) of every CTA
__device__ __forceinline__ long long gtimer_b() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ int smid_b() { int s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); return s; }
#define TR(role, t, e) do { if (blockIdx.x == 5 && blockIdx.y == 3 && (t) < 16) g_trace[((role) * 16 + (t)) * 16 + (e)] = clock64(); } while (0)
#define TRG(e) do { if (blockIdx.x == 5 && blockIdx.y == 3) g_trace[4 * 16 * 16 + (e)] = clock64(); } while (0)
#else
#define TR(role, t, e) do { } while (0)
#define TRG(e) do { } while (0)
#endif

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
  uint8_t* sAS = sdQ + AB_DQ;                 // extra k-step: A rows of the S^T product (ones in k = 0..2)
  uint8_t* sAD = sAS + AB_EXT_TILE;           //               A rows of the dP^T product (ones in k = 3..5)
  uint8_t* sBX = sAD + AB_EXT_TILE;           //               B rows [tile parity][query]: -lse (k 0..2), -delta (k 3..5)
  uint8_t* sZ = sBX + 2 * AB_EXT_TILE;        //               the shared all-zero second k chunk
  __shared__ int2 s_qt[AB_MAX_ITERS];                 // (start, len) of every query tile this CTA visits
  __shared__ __align__(16) uint32_t s_rb[2][3][64];   // allowed-key-group bits (mixed-group tiles only), by tile % 3
  __shared__ float s_uc[AB_DH];  // this (sample, head)'s uniform-row correction of dV, fetched up front
  __shared__ uint64_t bars[24];
  __shared__ uint32_t tmem_holder_s;
  uint64_t* kv_full = bars + 0;
  uint64_t* qdo_full = bars + 1;   // [AB_QSTAGES]
  uint64_t* qdo_empty = bars + 4;  // [AB_QSTAGES]
  uint64_t* x_full = bars + 7;     // S^T / dP^T of the tile (both query halves, one N = 128 product each) are in TMEM
  uint64_t* x_free = bars + 9;     // both compute warpgroups have copied them to registers and written the next B rows
  uint64_t* c_done = bars + 11;    // [2] per half: P^T / dS^T written (TMEM + smem)
  uint64_t* y_done = bars + 13;    // [2] per half: its dV / dK products retired, the shared P^T / dS^T columns are free
                                   // (one barrier per half: a waiter can then never be a whole phase ahead of it)
  uint64_t* z_full = bars + 15;    // [2] by tile parity: dQ product retired (dS consumed, dQ accumulator complete)
  uint64_t* dq_free = bars + 17;   // dQ accumulator copied to registers
  uint64_t* sdq_full = bars + 18;  // dQ tile staged in shared memory
  uint64_t* sdq_free = bars + 19;  // the TMA reduce has finished reading the staged tile
  uint32_t* tmem_holder = &tmem_holder_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x % a.H, b = blockIdx.x / a.H;
  const int kt = blockIdx.y;
  const mca_attn_qtile KT = a.k_tiles_q[kt];
  const long long row0 = static_cast<long long>(b) * a.N;
  const int cls = a.kt_class[static_cast<long long>(b) * a.n_kt + kt];
  const int n_iter = cls == 2 ? 0 : KT.kt_cnt;
  const int HD = a.H * AB_DH;
#ifdef MCA_TRACE
  const int cta_lin = blockIdx.y * gridDim.x + blockIdx.x;
  if (threadIdx.x == 0 && cta_lin < 4096) { g_bcta[cta_lin * 4] = gtimer_b(); g_bcta[cta_lin * 4 + 2] = smid_b(); g_bcta[cta_lin * 4 + 3] = n_iter; }
#endif

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_dq);
    mbar_init(kv_full, 1);
    for (int s = 0; s < AB_QSTAGES; ++s) mbar_init(&qdo_full[s], 1), mbar_init(&qdo_empty[s], 1);
    mbar_init(x_full, 1), mbar_init(x_free, 256);
    for (int s = 0; s < 2; ++s) mbar_init(&c_done[s], 128);
    mbar_init(&y_done[0], 1), mbar_init(&y_done[1], 1);
    mbar_init(&z_full[0], 1), mbar_init(&z_full[1], 1);
    mbar_init(dq_free, 128);
    mbar_init(sdq_full, 128);
    mbar_init(sdq_free, 1);
    fence_mbar_init();
    // K / V and the first Q / dO tile are requested right away: their DRAM latency overlaps the TMEM allocation, the
    // schedule staging and the constant-tile setup below instead of following the CTA-wide barrier
    if (n_iter > 0) {
      const int krow = static_cast<int>(row0 + KT.start);
      mbar_expect_tx(kv_full, 2 * AB_TILE);
      tma_load_2d(sK, &tm_qkv, kv_full, HD + h * AB_DH, krow);
      tma_load_2d(sV, &tm_qkv, kv_full, 2 * HD + h * AB_DH, krow);
      const int qrow = static_cast<int>(row0 + a.q_tiles[a.qt_list[KT.kt_off].tile].start);
      mbar_expect_tx(&qdo_full[0], 2 * AB_TILE);
      tma_load_2d(sQ, &tm_qkv, &qdo_full[0], h * AB_DH, qrow);
      tma_load_2d(sdO, &tm_do, &qdo_full[0], h * AB_DH, qrow);
    }
  }
  if (warp == 9) tmem_alloc(tmem_holder, 512);
  for (int i = threadIdx.x; i < n_iter; i += AB_THREADS) {
    const mca_attn_tile Q = a.q_tiles[a.qt_list[KT.kt_off + i].tile];
    s_qt[i] = make_int2(Q.start, Q.len);
  }
  if (threadIdx.x >= 256 && threadIdx.x < 256 + AB_DH)
    s_uc[threadIdx.x - 256] = a.ucorr[static_cast<long long>(b) * HD + h * AB_DH + threadIdx.x - 256];
  // constant rows of the extra k-step (bf16 1.0 = 0x3F80) and the zero block
  for (int i = threadIdx.x; i < 3 * AB_T; i += AB_THREADS) {
    const int row = i & 127, which = i >> 7;
    uint8_t* dst = (which == 0 ? sAS : (which == 1 ? sAD : sZ)) + row * 16;
    *reinterpret_cast<uint4*>(dst) = which == 0 ? make_uint4(0x3F803F80u, 0x00003F80u, 0u, 0u)
                                   : (which == 1 ? make_uint4(0u, 0x3F800000u, 0x3F803F80u, 0u) : make_uint4(0u, 0u, 0u, 0u));
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  // columns: [0,128) S^T (queries 0..127), [128,256) dP^T, [256,288) P^T (16) | dS^T (16) of half 0's quarter in flight,
  //          [288,320) the same for half 1,
  //          [320,384) dV, [384,448) dK, [448,512) dQ
  const uint32_t tP = tmem_base + 256, tdV = tmem_base + 320, tdK = tmem_base + 384, tdQ = tmem_base + 448;
  if (threadIdx.x == 0) TRG(0);

  if (warp == 8) {
    // ===================== TMA producer (whole warp loops, one elected lane issues) =====================
    if (n_iter > 0) {
      for (int it = 1; it < n_iter; ++it) {  // K / V and tile 0 were requested before the CTA-wide barrier
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
      constexpr uint32_t id_x = make_idesc_bf16(AB_T, AB_T, false, false);  // S^T, dP^T: K-major x K-major, N = 128
      constexpr uint32_t id_y = make_idesc_bf16(AB_T, AB_DH, false, true);  // dV, dK: A from TMEM, B MN-major
      constexpr uint32_t id_z = make_idesc_bf16(AB_T, AB_DH, true, true);   // dQ: MN-major x MN-major
      const uint64_t dk_k = make_smem_desc_sw128(smem_u32(sK), 16, 1024);     // K as the K-major A operand of S^T
      const uint64_t dv_k = make_smem_desc_sw128(smem_u32(sV), 16, 1024);     // V as the K-major A operand of dP^T
      const uint64_t dk_mn = make_smem_desc_sw128(smem_u32(sK), 8192, 1024);  // K as the MN-major B operand of dQ
      const uint64_t dq_k0 = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
      const uint64_t do_k0 = make_smem_desc_sw128(smem_u32(sdO), 16, 1024);
      const uint64_t dq_mn0 = make_smem_desc_sw128(smem_u32(sQ), 8192, 1024);
      const uint64_t do_mn0 = make_smem_desc_sw128(smem_u32(sdO), 8192, 1024);
      // leading-dimension offset = distance to the shared zero chunk, stride 128 B between 8-row groups
      const uint64_t d_as = make_smem_desc_nosw(smem_u32(sAS), 4 * AB_EXT_TILE, 128);
      const uint64_t d_ad = make_smem_desc_nosw(smem_u32(sAD), 3 * AB_EXT_TILE, 128);
      const uint64_t d_bx0 = make_smem_desc_nosw(smem_u32(sBX), 2 * AB_EXT_TILE, 128);
      const uint64_t d_bx1 = make_smem_desc_nosw(smem_u32(sBX) + AB_EXT_TILE, AB_EXT_TILE, 128);
      // S^T = K Q^T and dP^T = V dO^T for all 128 queries of the tile at once: the K / V rows (the A operands) are
      // read from shared memory once per k-step instead of once per query half
      auto issue_x = [&](int t) {
        const uint64_t off = static_cast<uint64_t>(((t % AB_QSTAGES) * AB_TILE) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < AB_DH / 16; ++k) {
            umma_bf16(tmem_base, dk_k + k * 2, dq_k0 + off + k * 2, id_x, k > 0 ? 1u : 0u);
            umma_bf16(tmem_base + 128, dv_k + k * 2, do_k0 + off + k * 2, id_x, k > 0 ? 1u : 0u);
          }
          const uint64_t bx = (t & 1) ? d_bx1 : d_bx0;
          umma_bf16(tmem_base, d_as, bx, id_x, 1u);        // S^T  -= lse_q
          umma_bf16(tmem_base + 128, d_ad, bx, id_x, 1u);  // dP^T -= delta_q
          umma_commit(x_full);
        }
        __syncwarp();
      };
      // dV += P^T dO, dK += dS^T Q for one QUARTER (32 queries) of the tile: every warpgroup owns a private 32-column
      // P^T | dS^T buffer (tP + 32 * half), so it never waits for the other warpgroup's products, only for its own
      // previous quarter's (which run under its next quarter's exp / multiply work)
      auto issue_y = [&](int t, int hf, int qd, bool last_of_tile) {
        const uint64_t off = static_cast<uint64_t>(((t % AB_QSTAGES) * AB_TILE + hf * 8192 + qd * 4096) >> 4);
        const uint32_t acc = (t > 0 || hf > 0 || qd > 0) ? 1u : 0u;
        const uint32_t tp = tP + 32 * hf;
        if (elect_one()) {
          // contraction over the 32 queries of this quarter: 2 steps of 16 query rows (2 KB of the MN-major B tile)
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            umma_bf16_ts(tdV, tp + k * 8, do_mn0 + off + k * 128, id_y, k > 0 ? 1u : acc);
            umma_bf16_ts(tdK, tp + 16 + k * 8, dq_mn0 + off + k * 128, id_y, k > 0 ? 1u : acc);
          }
          umma_commit(&y_done[hf]);
          if (last_of_tile) umma_commit(&qdo_empty[t % AB_QSTAGES]);  // Q / dO of this tile are no longer read
        }
        __syncwarp();
      };
      mbar_wait(kv_full, 0);
      if (lane == 0) TRG(1);
      mbar_wait(&qdo_full[0], 0);
      if (lane == 0) TRG(2);
      // x_free phase n = "the B rows of tiles <= n + 1 are written and S^T / dP^T of tile n - 1 have been read"
      mbar_wait(x_free, 0);
      tc_fence_after();
      issue_x(0);
      for (int t = 0; t < n_iter; ++t) {
        const uint32_t ph = t & 1;
        const bool more = t + 1 < n_iter;
        if (lane == 0) TR(2, t, 0);
        if (more) {
          mbar_wait(&qdo_full[(t + 1) % AB_QSTAGES], ((t + 1) / AB_QSTAGES) & 1);
          mbar_wait(x_free, ph ^ 1);  // S^T / dP^T of tile t are in registers: overwrite them right away
          tc_fence_after();
          issue_x(t + 1);
        }
        if (lane == 0) TR(2, t, 1);
        // quarters in the order the warpgroups produce them: (half 0, q 0), (half 1, q 0), (half 0, q 1), (half 1, q 1);
        // c_done[h] completes twice per tile: phase parity = quarter index
#pragma unroll
        for (int qd = 0; qd < 2; ++qd) {
          mbar_wait(&c_done[0], static_cast<uint32_t>(qd));
          tc_fence_after();
          issue_y(t, 0, qd, false);
          mbar_wait(&c_done[1], static_cast<uint32_t>(qd));
          tc_fence_after();
          issue_y(t, 1, qd, qd == 1);
        }
        if (lane == 0) TR(2, t, 5);
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
    const uint32_t reg = tmem_base + hf * 64 + lane_sel;  // S^T columns of this half; dP^T is 128 columns further
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
    // Per-query values: the first 64 threads of each warpgroup own one query of the half.  They load (lse, delta,
    // allowed-group bits) one tile ahead into registers and turn them into the B row of the extra k-step.
    float r_lse = CUDART_INF_F, r_dl = 0.f;
    uint32_t r_rb = 0u;
    auto load_q = [&](int t) {
      if (wt < 64 && t < n_iter) {
        const int qi = min(s_qt[t].x + hf * 64 + wt, a.N - 1);
        r_lse = a.lse[sbase + qi];
        r_dl = a.delta[sbase + qi];
        r_rb = a.rowbits[qi];
      }
    };
    // v = h + m + l exactly: three 8-bit slices of the fp32 significand, each a bf16 (truncation, no cvt instructions)
    auto split3 = [](float v, uint32_t& h, uint32_t& m, uint32_t& l) {
      const uint32_t vh = __float_as_uint(v) & 0xFFFF0000u;
      const float r1 = v - __uint_as_float(vh);
      const uint32_t vm = __float_as_uint(r1) & 0xFFFF0000u;
      const float r2 = r1 - __uint_as_float(vm);
      h = vh >> 16, m = vm >> 16, l = __float_as_uint(r2) >> 16;
    };
    auto write_ext = [&](int t) {  // B row of this thread's query for tile t (from the registers loaded for t)
      if (wt < 64 && t < n_iter) {
        const bool valid = hf * 64 + wt < s_qt[t].y;
        const uint32_t rb = valid ? r_rb : 0u;
        const bool sees = valid && (kgrp == 255 || ((rb >> kgrp) & 1u)) && r_lse != CUDART_INF_F;
        uint32_t lh, lm, ll, dh, dm, dl;
        split3(sees ? -r_lse : AB_MASKED, lh, lm, ll);
        split3(valid ? -r_dl : 0.f, dh, dm, dl);
        const int q = hf * 64 + wt;  // row of the [128 x 16] B tile of this tile parity
        *reinterpret_cast<uint4*>(sBX + (t & 1) * AB_EXT_TILE + q * 16) =
            make_uint4(lh | (lm << 16), ll | (dh << 16), dm | (dl << 16), 0u);
        if (kgrp == 255) s_rb[hf][t % 3][wt] = rb;
        fence_proxy_async_smem();
      }
    };
    auto drain_dq = [&](int tp) {  // dQ of tile tp: TMEM -> registers -> fp32 swizzled smem (the reduce warp ships it)
      mbar_wait(&z_full[tp & 1], (tp >> 1) & 1);
      if (wt == 0) TR(hf, tp, 11);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld32(tdQ + lane_sel, v0);
      tmem_ld32(tdQ + lane_sel + 32, v1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(dq_free);
      if (wt == 0) TR(hf, tp, 12);
      if (tp > 0) mbar_wait(sdq_free, (tp - 1) & 1);
      if (wt == 0) TR(hf, tp, 13);
      uint8_t* rowp = sdQ + r * 128;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        *reinterpret_cast<uint4*>(rowp + ((q ^ (r & 7)) << 4)) = make_uint4(v0[4 * q], v0[4 * q + 1], v0[4 * q + 2], v0[4 * q + 3]);
        *reinterpret_cast<uint4*>(rowp + AB_DQ / 2 + ((q ^ (r & 7)) << 4)) = make_uint4(v1[4 * q], v1[4 * q + 1], v1[4 * q + 2], v1[4 * q + 3]);
      }
      fence_proxy_async_smem();
      mbar_arrive(sdq_full);
    };
    {  // tiles 0 and 1: both sets of per-query loads in flight before either is consumed (one DRAM latency, not two)
      load_q(0);
      const float l0 = r_lse, d0 = r_dl;
      const uint32_t b0 = r_rb;
      load_q(1);
      const float l1 = r_lse, d1 = r_dl;
      const uint32_t b1 = r_rb;
      r_lse = l0, r_dl = d0, r_rb = b0;
      write_ext(0);
      r_lse = l1, r_dl = d1, r_rb = b1;
      write_ext(1);
      load_q(2);
    }
    mbar_arrive(x_free);  // phase 0: the B rows of tiles 0 and 1 are in place
    if (kgrp == 255) ab_bar_sync(1 + hf, 128);  // ... and so are their group-bit slots, for every thread of the warpgroup
    for (int t = 0; t < n_iter; ++t) {
      const uint32_t ph = t & 1;
      if (wt == 0) TR(hf, t, 0);
      mbar_wait(x_full, ph);
      if (wt == 0) TR(hf, t, 2);
      tc_fence_after();
      uint32_t sv[2][32], dv[2][32];
      tmem_ld32(reg, sv[0]);
      tmem_ld32(reg + 32, sv[1]);
      tmem_ld32(reg + 128, dv[0]);
      tmem_ld32(reg + 160, dv[1]);
      tmem_ld_wait();
      if (wt == 0) TR(hf, t, 3);
      tc_fence_before();
      mbar_arrive(x_free);  // the next tile's S^T / dP^T may be issued once both warpgroups got here
      if (wt == 0) TR(hf, t, 9);
      const uint32_t* rbq = s_rb[hf][t % 3];
      // dS^T also goes to shared memory for the dQ product.  Half 0 is double buffered by tile parity; half 1 has one
      // buffer that dQ(t-1) must have finished reading.
      uint8_t* ds_row;
      if (hf == 0) {
        if (t >= 2) mbar_wait(&z_full[t & 1], ((t - 2) >> 1) & 1);
        ds_row = sdS + (t & 1) * (AB_DS / 2) + r * 128;
      } else {
        if (t >= 1) mbar_wait(&z_full[(t - 1) & 1], ((t - 1) >> 1) & 1);
        ds_row = sdS + 2 * (AB_DS / 2) + r * 128;
      }
      const uint32_t tp_mine = tP + 32 * hf + lane_sel;  // this warpgroup's private P^T (16 columns) | dS^T (16 columns)
#pragma unroll
      for (int qd = 0; qd < 2; ++qd) {  // the half's 64 queries as two quarters of 32: sv[qd], dv[qd]
        // accumulators already hold S^T - lse[q] and dP^T - delta[q]:  P^T = exp2(log2e * .),  dS^T = P^T * (.)
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const float pi = fast_ex2(__uint_as_float(sv[qd][e]) * AB_LOG2E);
          sv[qd][e] = __float_as_uint(pi);
          dv[qd][e] = __float_as_uint(pi * __uint_as_float(dv[qd][e]));
        }
        if (kgrp == 255) {  // mixed key groups (fusion sub-blocks): per-(query, key) visibility
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (!((rbq[qd * 32 + e] >> mygrp) & 1u)) sv[qd][e] = 0u, dv[qd][e] = 0u;
        }
        if (dead) {
#pragma unroll
          for (int e = 0; e < 32; ++e) sv[qd][e] = 0u, dv[qd][e] = 0u;
        }
        uint32_t pp[16], dd[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          pp[j] = pack_bf16x2(__uint_as_float(sv[qd][2 * j]), __uint_as_float(sv[qd][2 * j + 1]));
          dd[j] = pack_bf16x2(__uint_as_float(dv[qd][2 * j]), __uint_as_float(dv[qd][2 * j + 1]));
        }
        if (wt == 0 && qd == 1) TR(hf, t, 8);
        // this quarter's 32 queries = bytes [64 qd, 64 qd + 64) of the key's dS row (128B-swizzled 16-byte chunks)
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(ds_row + (((4 * qd + c) ^ (r & 7)) << 4)) = make_uint4(dd[4 * c], dd[4 * c + 1], dd[4 * c + 2], dd[4 * c + 3]);
        if (wt == 0 && qd == 1) TR(hf, t, 4);
        // the private buffer is free once this warpgroup's previous quarter's dV / dK products retired
        if (t > 0 || qd > 0) {
          mbar_wait(&y_done[hf], static_cast<uint32_t>(qd ^ 1));  // commit index 2t + qd - 1
          tc_fence_after();
        }
        if (wt == 0 && qd == 1) TR(hf, t, 5);
        tmem_st16(tp_mine, pp);
        tmem_st16(tp_mine + 16, dd);
        tmem_st_wait();
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&c_done[hf]);
      }
      if (wt == 0) TR(hf, t, 6);
      // B rows of tile t + 2 go into the buffer X(t) has finished with (x_full was observed above); X(t + 2) is only
      // issued after this thread's x_free arrival of iteration t + 1, so the write is off the critical path.  Mixed-group
      // key tiles also keep the queries' group bits in shared memory (slot tile % 3): the barrier makes sure nobody still
      // reads the slot being replaced (tile t - 1) and publishes the slots written in earlier iterations.
      if (kgrp == 255) ab_bar_sync(1 + hf, 128);
      write_ext(t + 2);
      load_q(t + 3);
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
      uint32_t t0[32], t1[32];
      if (n_iter > 0) {  // both 32-column halves of the accumulator in flight before the wait
        tmem_ld32((which == 0 ? tdK : tdV) + lane_sel, t0);
        tmem_ld32((which == 0 ? tdK : tdV) + lane_sel + 32, t1);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) t0[i] = 0u, t1[i] = 0u;
      }
#pragma unroll
      for (int cc = 0; cc < AB_DH / 32; ++cc) {
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(cc == 0 ? t0[i] : t1[i]);
        if (which == 1) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += s_uc[cc * 32 + i];
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
#ifdef MCA_TRACE
  if (threadIdx.x == 0 && cta_lin < 4096) g_bcta[cta_lin * 4 + 1] = gtimer_b();
#endif
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// delta[b,h,n] = sum_c dO*O ; ucorr[b, h*64+c] += dO[row, h*64+c] / N for rows whose lse is +inf (fully masked).
// Block (x, b) walks 64 rows of sample b: the masked rows' dO are summed in registers, then across the 8 warps in shared
// memory, and only then added to ucorr — one atomic per column and block instead of one per column and row (with 40 %
// modality dropout a third of all rows are fully masked and the per-row atomics on 4096 addresses dominated the step).
constexpr int PREP_ROWS = 64;
__global__ void __launch_bounds__(256)
attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout,
                     const float* __restrict__ lse, float* __restrict__ delta, float* __restrict__ ucorr, int B, int N,
                     int H) {
  __shared__ float s_part[8][512];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int HD = H * AB_DH;  // 512: each lane owns 16 consecutive columns, 4 lanes per head
  const int c0 = lane * (HD / 32);
  const int h = c0 / AB_DH;
  float usum[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) usum[i] = 0.f;
  bool any = false;
  const int n_end = min(N, (static_cast<int>(blockIdx.x) + 1) * PREP_ROWS);
  for (int n = blockIdx.x * PREP_ROWS + warp; n < n_end; n += 8) {
    const long long row = static_cast<long long>(b) * N + n;
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
    const long long sidx = (static_cast<long long>(b) * H + h) * N + n;
    if ((lane & 3) == 0) delta[sidx] = acc;
    if (lse[sidx] == CUDART_INF_F) {
      any = true;
#pragma unroll
      for (int i = 0; i < 16; ++i) usum[i] += dv[i];
    }
  }
  const bool block_any = __syncthreads_or(any);
  if (!block_any) return;
#pragma unroll
  for (int i = 0; i < 16; ++i) s_part[warp][c0 + i] = usum[i];
  __syncthreads();
  const float inv = 1.0f / static_cast<float>(N);
  for (int c = threadIdx.x; c < HD; c += 256) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_part[w][c];
    if (t != 0.f) atomicAdd(ucorr + static_cast<long long>(b) * HD + c, t * inv);
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
  attn_bwd_prep_kernel<<<dim3((N + PREP_ROWS - 1) / PREP_ROWS, B), 256, 0, stream>>>(
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
extern "C" int mca_debug_read_cta_bwd(long long* host_dst, int n) {
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(host_dst, mca::g_bcta, sizeof(long long) * n) == cudaSuccess ? 0 : 3;
}
#endif
