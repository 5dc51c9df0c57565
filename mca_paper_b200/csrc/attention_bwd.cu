// Block-sparse masked attention backward on tcgen05 (autograd of model.py:85-100).
//
// Work item = one (sample, head, 128-key tile); the item walks the query tiles that attend this key tile (the transpose
// of the forward schedule).  The kernel is PERSISTENT: one CTA per SM pulls items from a self-resetting work queue and
// carries its pipeline across items (the next item's Q / dO tiles, per-query terms and schedule are staged while the
// current item is computed; TMEM, barriers and the constant operand tiles are set up once).  A CTA per item spent ~30 %
// of its life outside its tile loop: launch latency, set-up, the exposed head of the pipeline and a serial epilogue
// (profiles/r2_attn_notes.md).
//
// Scores are computed TRANSPOSED (keys on the TMEM lanes, queries along the columns) so that P^T and dS^T can be fed
// back to the tensor core straight from TMEM as the A operands of the dV and dK products; only dS also goes to shared
// memory (for dQ).  Per query tile g (flat index over the CTA's items):
//     X(g): S^T  = K Q^T            dP^T = V dO^T            (128 x 128 x 64, smem x smem -> TMEM, both query halves)
//     C(g): P^T  = exp2(S^T*log2e - lse_q*log2e)   dS^T = P^T * (dP^T - delta_q)       (thread = key row)
//     Y(g): dV  += P^T dO           dK  += dS^T Q            (A from TMEM, B = the TMA tiles read MN-major), per quarter
//     Z(g): dQ   = dS K             (128 x 64 x 128; fresh tile -> fp32 smem -> TMA reduce-add into dq_acc)
// dK/dV stay resident in TMEM across an item and are written once.  The per-query terms ride on the tensor core: both
// score products get one extra k-step whose B rows hold (-lse, -delta) of the query, each split into three bf16 terms
// (exact to 2^-25), against constant 0/1 A rows, so the accumulators arrive as S^T - lse_q and dP^T - delta_q.
// Masking costs nothing per element for the common tiles: a query that may not see this tile's key group gets
// -lse = -60000 (P = 0) in that row; only tiles with padded / missing keys or mixed key groups (the fusion
// sub-blocks) take a per-element select.  Fully masked query rows carry lse = +inf from the forward; their
// uniform-1/N contribution to dV (reference quirk Q4) is the per-(sample, head) vector `ucorr`, added in the epilogue.
//
// Warp roles (512 threads, registers redistributed with setmaxnreg):
//   warpgroups 0, 1 : compute, one per 64-query half of every tile (thread = key row)
//   warpgroup 2     : per-query helper (thread = query): writes the extra k-step B rows two tiles ahead, drains the dQ
//                     accumulator TMEM -> fp32 shared memory and issues its TMA reduce-add (the first two used to sit on
//                     the compute warps' critical path)
//   warp 12 : work queue + TMA producer      warps 13 / 14 / 15 : MMA issuers of (S^T, dP^T) / dQ / (dV, dK): one
//             issuing warp spent ~40 % of every tile polling its barriers one after another while the tensor pipe idled
#include <math_constants.h>

#include "mca_b200.h"
#include "ptx.cuh"
#include "runtime.h"

namespace mca {

constexpr int AB_T = 128;
constexpr int AB_DH = 64;
constexpr int AB_THREADS = 512;
constexpr int AB_TILE = AB_T * AB_DH * 2;  // 16 KB bf16 [128, 64]
constexpr int AB_DS = AB_T * AB_T * 2;     // 32 KB bf16 dS tile, stored [key][query] in two 64-query halves
constexpr int AB_DQ = AB_T * AB_DH * 4;    // 32 KB fp32 [128, 64]
constexpr int AB_QSTAGES = 3;
constexpr int AB_RING = 4;         // staged work items (the helper looks two query tiles ahead: at most three items)
constexpr int AB_MAX_ITERS = 128;  // upper bound of query tiles attending one key tile (the launcher checks the total)
// sK, sV, 3x(sQ, sdO), sdS (half 0 double buffered, half 1 single), sdQ, extra k-step operands
// extra k-step operands (K-major, no swizzle: 8-row x 16-byte core matrices).  Only the first 8-element k chunk of a
// row carries data; the second chunk of EVERY operand is the same all-zero block (the descriptor's leading-dimension
// offset points each tile at it), so a [128 x 16] operand costs 2 KB: A rows for S^T, A rows for dP^T, B rows of the
// 128 queries of a tile (double buffered by tile parity), zero block.
constexpr int AB_EXT_TILE = AB_T * 16;
constexpr int AB_EXT = 5 * AB_EXT_TILE;
constexpr int AB_SMEM = 2 * AB_TILE + 2 * AB_QSTAGES * AB_TILE + 3 * (AB_DS / 2) + AB_DQ + AB_EXT;
constexpr float AB_MASKED = -60000.f;  // -lse of a query that must not see this key tile: exp2 underflows to 0
constexpr float AB_LOG2E = 1.4426950408889634f;
// register budget per thread after setmaxnreg (x 128 threads per warpgroup): 2 x 176 + 96 + 56 = 504 <= 512

struct AttnBwdArgs {
  const mca_attn_qtile* k_tiles_q;  // per key tile: start, len, slice of qt_list
  const mca_attn_ref* qt_list;
  const mca_attn_tile* q_tiles;
  const uint32_t* rowbits;
  const uint8_t* keygrp;
  const uint8_t* tile_grp;
  const uint8_t* padding;
  const uint8_t* kt_class;
  const uint8_t* skip_ok;  // [B] or nullptr: all-padded query tiles of this sample are left out (mca_query_skip_flags)
  const float* lse;     // [B,H,N]
  const float* delta;   // [B,H,N]
  const float* ucorr;   // [B, H*64]
  __nv_bfloat16* dqkv;  // [B*N, 3*H*64]
  int N, H, n_kt;
  int n_items, BH;      // work items = (key tile, sample, head), key tile slow
};

struct AbItem { int n_iter, b, h, kt, kstart, klen, kgrp, cls; };  // n_iter == 0: no more work

// Work queue: [0] next item, [1] CTAs that have left.  The last CTA out resets both for the next launch on the stream.
__device__ unsigned int g_ab_ctr[2];

__device__ __forceinline__ void ab_setmaxnreg_inc_compute() { asm volatile("setmaxnreg.inc.sync.aligned.u32 176;"); }
__device__ __forceinline__ void ab_setmaxnreg_dec_helper() { asm volatile("setmaxnreg.dec.sync.aligned.u32 96;"); }
__device__ __forceinline__ void ab_setmaxnreg_dec_misc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 56;"); }

__device__ __forceinline__ void ab_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

#ifdef MCA_TRACE
// debug-only timeline (built with -DMCA_TRACE): clock64 stamps of one CTA, read by mca_debug_read_trace
__device__ long long g_trace[4 * 16 * 16 + 8];
__device__ long long g_bcta[4096 * 4];  // (globaltimer start, end, smid, n_iter) of every work item
__device__ __forceinline__ long long gtimer_b() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ int smid_b() { int s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); return s; }
#define TR(role, t, e) do { if (blockIdx.x == 5 && (t) >= 13 && (t) < 29) g_trace[((role) * 16 + (t) - 13) * 16 + (e)] = clock64(); } while (0)
#define TRG(e) do { if (blockIdx.x == 5) g_trace[4 * 16 * 16 + (e)] = clock64(); } while (0)
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
  __shared__ int2 s_qt[AB_RING][AB_MAX_ITERS];        // (start, len) of every query tile of the staged items
  __shared__ AbItem s_item[AB_RING];
  __shared__ __align__(16) uint32_t s_rb[2][3][64];   // allowed-key-group bits (mixed-group tiles only), by tile % 3
  __shared__ uint64_t bars[36];
  __shared__ uint32_t tmem_holder_s;
  uint64_t* item_full = bars + 0;   // [AB_RING] item header + query-tile list staged
  uint64_t* item_empty = bars + 4;  // [AB_RING] every consumer is done with the slot
  uint64_t* kv_full = bars + 8;
  uint64_t* kv_empty = bars + 9;    // every product reading this item's K / V has retired
  uint64_t* qdo_full = bars + 10;   // [AB_QSTAGES]
  uint64_t* qdo_empty = bars + 13;  // [AB_QSTAGES]
  uint64_t* ext_full = bars + 16;   // [2] by tile parity: B rows of the extra k-step written
  uint64_t* ext_empty = bars + 18;  // [2] the score products that read them have retired
  uint64_t* x_full = bars + 20;     // S^T / dP^T of the tile (both query halves, one N = 128 product each) are in TMEM
  uint64_t* x_free = bars + 21;     // both compute warpgroups have copied them to registers
  uint64_t* c_done = bars + 22;     // [2] per half: P^T / dS^T written (TMEM + smem)
  uint64_t* y_done = bars + 24;     // [2] per half: its dV / dK products retired, the private P^T / dS^T columns are free
  uint64_t* z_full = bars + 26;     // [2] by tile parity: dQ product retired (dS consumed, dQ accumulator complete)
  uint64_t* dq_free = bars + 28;    // dQ accumulator copied to registers
  uint64_t* ds_full = bars + 29;    // both halves of the tile's dS are in shared memory (what the dQ product waits for)
  uint64_t* acc_full = bars + 30;   // every dV / dK product of the item has retired
  uint64_t* acc_free = bars + 31;   // the item's dK / dV accumulators have been copied to registers
  uint32_t* tmem_holder = &tmem_holder_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int HD = a.H * AB_DH;
  pdl_launch_dependents();

  if (warp == 12 && lane == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_dq);
    for (int s = 0; s < AB_RING; ++s) mbar_init(&item_full[s], 1), mbar_init(&item_empty[s], 15);
    mbar_init(kv_full, 1), mbar_init(kv_empty, 2);
    for (int s = 0; s < AB_QSTAGES; ++s) mbar_init(&qdo_full[s], 1), mbar_init(&qdo_empty[s], 2);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ext_full[s], 128), mbar_init(&ext_empty[s], 1);
      mbar_init(&c_done[s], 128), mbar_init(&y_done[s], 1), mbar_init(&z_full[s], 1);
    }
    mbar_init(x_full, 1), mbar_init(x_free, 256);
    mbar_init(dq_free, 128);
    mbar_init(ds_full, 256);
    mbar_init(acc_full, 1);
    mbar_init(acc_free, 256);
    fence_mbar_init();
  }
  if (warp == 13) tmem_alloc(tmem_holder, 512);
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
  pdl_wait();  // set-up ran under the predecessor's tail (PDL); nothing above touches global memory
  // columns: [0,128) S^T (queries 0..127), [128,256) dP^T, [256,288) P^T (16) | dS^T (16) of half 0's quarter in flight,
  //          [288,320) the same for half 1,
  //          [320,384) dV, [384,448) dK, [448,512) dQ
  const uint32_t tP = tmem_base + 256, tdV = tmem_base + 320, tdK = tmem_base + 384, tdQ = tmem_base + 448;
  if (threadIdx.x == 0) TRG(0);

  if (warp >= 12) {
    ab_setmaxnreg_dec_misc();
    if (warp == 12) {
      // ===================== work queue + TMA producer (whole warp loops, one elected lane issues) =================
      // stage(n): pull the next non-empty item from the queue into ring slot n % AB_RING.  Items whose key tile holds no
      // live key for the sample have no products at all: their dK rows are 0 and their dV rows the uniform-row
      // correction, written here directly.
      auto stage = [&](int n) -> bool {
        const int slot = n & (AB_RING - 1);
        mbar_wait(&item_empty[slot], ((n / AB_RING) & 1) ^ 1);
        for (;;) {
          unsigned int item = 0;
          if (lane == 0) item = atomicAdd(&g_ab_ctr[0], 1u);
          item = __shfl_sync(0xffffffffu, item, 0);
          if (item >= static_cast<unsigned int>(a.n_items)) {
            if (lane == 0) {
              s_item[slot].n_iter = 0;
              mbar_arrive(&item_full[slot]);
            }
            __syncwarp();
            return false;
          }
          const int kt = static_cast<int>(item) / a.BH, bh = static_cast<int>(item) % a.BH;
          const int b = bh / a.H, h = bh % a.H;
          const mca_attn_qtile KT = a.k_tiles_q[kt];
          const int cls = a.kt_class[static_cast<long long>(b) * a.n_kt + kt];
          // the query tiles that attend this key tile; varlen: tiles whose rows are all padded are left out when the
          // sample's flag allows it (their dQ rows stay zero, they add nothing to dK / dV)
          int cnt = 0;
          if (cls != 2) {
            const bool skip = a.skip_ok != nullptr && a.skip_ok[b] != 0;
            for (int i0 = 0; i0 < KT.kt_cnt; i0 += 32) {
              const int i = i0 + lane;
              bool keep = false;
              int2 v = make_int2(0, 0);
              if (i < KT.kt_cnt) {
                const int tid = a.qt_list[KT.kt_off + i].tile;
                const mca_attn_tile Q = a.q_tiles[tid];
                v = make_int2(Q.start, Q.len);
                keep = !(skip && a.kt_class[static_cast<long long>(b) * a.n_kt + tid] == 2);
              }
              const uint32_t m = __ballot_sync(0xffffffffu, keep);
              if (keep) s_qt[slot][cnt + __popc(m & ((1u << lane) - 1u))] = v;
              cnt += __popc(m);
            }
          }
          if (cnt == 0) {
            const float* uc = a.ucorr + static_cast<long long>(b) * HD + h * AB_DH;
            for (int i = lane; i < KT.len * 16; i += 32) {  // 16 chunks of 8 bf16 per row: 8 of dK, 8 of dV
              const int row = i >> 4, ch = i & 15;
              __nv_bfloat16* dst = a.dqkv + (static_cast<long long>(b) * a.N + KT.start + row) * (3 * HD) + h * AB_DH +
                                   (ch < 8 ? HD : 2 * HD) + (ch & 7) * 8;
              uint4 w = make_uint4(0u, 0u, 0u, 0u);
              if (ch >= 8) {
                const float* u = uc + (ch & 7) * 8;
                w.x = pack_bf16x2(u[0], u[1]), w.y = pack_bf16x2(u[2], u[3]);
                w.z = pack_bf16x2(u[4], u[5]), w.w = pack_bf16x2(u[6], u[7]);
              }
              *reinterpret_cast<uint4*>(dst) = w;
            }
            continue;
          }
          __syncwarp();
          if (lane == 0) {
            s_item[slot] = AbItem{cnt, b, h, kt, KT.start, KT.len, a.tile_grp[kt], cls};
            mbar_arrive(&item_full[slot]);
          }
          __syncwarp();
          return true;
        }
      };
      bool more = stage(0);
      if (more) more = stage(1);
      int g = 0;
      for (int n = 0;; ++n) {
        const int slot = n & (AB_RING - 1);
        const AbItem I = s_item[slot];
        if (I.n_iter == 0) break;
        const long long row0 = static_cast<long long>(I.b) * a.N;
        const int krow = static_cast<int>(row0 + I.kstart);
        mbar_wait(kv_empty, (n & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(kv_full, 2 * AB_TILE);
          tma_load_2d(sK, &tm_qkv, kv_full, HD + I.h * AB_DH, krow);
          tma_load_2d(sV, &tm_qkv, kv_full, 2 * HD + I.h * AB_DH, krow);
        }
        __syncwarp();
        for (int t = 0; t < I.n_iter; ++t, ++g) {
          const int s = g % AB_QSTAGES;
          const int qrow = static_cast<int>(row0 + s_qt[slot][t].x);
          mbar_wait(&qdo_empty[s], ((g / AB_QSTAGES) & 1) ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&qdo_full[s], 2 * AB_TILE);
            tma_load_2d(sQ + s * AB_TILE, &tm_qkv, &qdo_full[s], I.h * AB_DH, qrow);
            tma_load_2d(sdO + s * AB_TILE, &tm_do, &qdo_full[s], I.h * AB_DH, qrow);
          }
          __syncwarp();
        }
        if (more) more = stage(n + 2);
      }
    } else if (warp == 14) {
      // ===================== third MMA issuer: the dQ product =====================
      constexpr uint32_t id_z = make_idesc_bf16(AB_T, AB_DH, true, true);    // dQ: MN-major x MN-major
      const uint64_t dk_mn = make_smem_desc_sw128(smem_u32(sK), 8192, 1024);  // K as the MN-major B operand of dQ
      int g = 0;
      for (int n = 0;; ++n) {
        const int slot = n & (AB_RING - 1);
        mbar_wait(&item_full[slot], (n / AB_RING) & 1);
        const int n_iter = s_item[slot].n_iter;
        if (n_iter == 0) break;
        for (int t = 0; t < n_iter; ++t, ++g) {
          mbar_wait(ds_full, g & 1);  // both halves of dS(g) are in shared memory
          if (g > 0) mbar_wait(dq_free, (g - 1) & 1);  // the previous dQ accumulator has been drained
          tc_fence_after();
          // dQ = dS K: contraction over the 128 keys (8 steps of 16 key rows); the two 64-query halves of dS live in
          // separate buffers: half 0 at buffer (g & 1), half 1 behind both (always a positive offset)
          const uint64_t dds = make_smem_desc_sw128(smem_u32(sdS) + (g & 1) * (AB_DS / 2), (2 - (g & 1)) * (AB_DS / 2), 1024);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < AB_T / 16; ++k) umma_bf16(tdQ, dds + k * 128, dk_mn + k * 128, id_z, k > 0 ? 1u : 0u);
            umma_commit(&z_full[g & 1]);
            if (t == n_iter - 1) umma_commit(kv_empty);  // this issuer's last read of the item's K / V
          }
          __syncwarp();
          if (lane == 0) TR(3, g, 7);
        }
        if (lane == 0) mbar_arrive(&item_empty[slot]);
      }
    } else if (warp == 13) {
      // ===================== MMA issuer =====================
      // The whole warp runs the loop and the waits (convergent code keeps the descriptors in uniform registers, so a
      // tcgen05.mma costs one uniform add); one elected lane issues the MMAs and commits.  Independent accumulation
      // chains are interleaved (S^T with dP^T, dV with dK).
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
      // S^T = K Q^T and dP^T = V dO^T for all 128 queries of tile gx at once (after its operands, the previous tile's
      // read-out and its extra k-step rows are in place)
      auto issue_x = [&](int gx) {
        mbar_wait(&qdo_full[gx % AB_QSTAGES], (gx / AB_QSTAGES) & 1);
        if (lane == 0) TR(2, gx - 1, 2);
        if (gx > 0) mbar_wait(x_free, (gx - 1) & 1);
        if (lane == 0) TR(2, gx - 1, 3);
        mbar_wait(&ext_full[gx & 1], (gx >> 1) & 1);
        if (lane == 0) TR(2, gx - 1, 4);
        tc_fence_after();
        const uint64_t off = static_cast<uint64_t>(((gx % AB_QSTAGES) * AB_TILE) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < AB_DH / 16; ++k) {
            umma_bf16(tmem_base, dk_k + k * 2, dq_k0 + off + k * 2, id_x, k > 0 ? 1u : 0u);
            umma_bf16(tmem_base + 128, dv_k + k * 2, do_k0 + off + k * 2, id_x, k > 0 ? 1u : 0u);
          }
          const uint64_t bx = (gx & 1) ? d_bx1 : d_bx0;
          umma_bf16(tmem_base, d_as, bx, id_x, 1u);        // S^T  -= lse_q
          umma_bf16(tmem_base + 128, d_ad, bx, id_x, 1u);  // dP^T -= delta_q
          umma_commit(x_full);
          umma_commit(&ext_empty[gx & 1]);
          umma_commit(&qdo_empty[gx % AB_QSTAGES]);  // this issuer's only read of the tile's Q / dO
        }
        __syncwarp();
      };
      int g = 0;
      for (int n = 0;; ++n) {
        const int slot = n & (AB_RING - 1);
        mbar_wait(&item_full[slot], (n / AB_RING) & 1);
        const int n_iter = s_item[slot].n_iter;
        if (n_iter == 0) break;
        mbar_wait(kv_full, n & 1);
        if (lane == 0) TRG(1);
        for (int t = 0; t < n_iter; ++t, ++g) {
          if (lane == 0) TR(2, g, 0);
          if (t == 0) issue_x(g);  // K / V are single-buffered: an item's first score products cannot be issued ahead
          if (t + 1 < n_iter) issue_x(g + 1);  // S^T / dP^T of tile g are in registers: overwrite them right away
          if (lane == 0) TR(2, g, 1);
          if (t == n_iter - 1 && elect_one()) umma_commit(kv_empty);  // this issuer's last read of the item's K / V
          __syncwarp();
          if (lane == 0) TR(2, g, 7);
        }
        if (lane == 0) mbar_arrive(&item_empty[slot]);
      }
    } else if (warp == 15) {
      // ===================== second MMA issuer: the dV / dK products =====================
      // A single issuing warp spent ~40 % of every tile polling barriers one after another (an mbarrier wait costs
      // ~100 cycles even when it is already complete) while the tensor pipe sat idle; the quarter products, whose
      // barriers arrive one by one from the compute warpgroups, have their own issuer (profiles/r2_attn_notes.md).
      constexpr uint32_t id_y = make_idesc_bf16(AB_T, AB_DH, false, true);  // dV, dK: A from TMEM, B MN-major
      const uint64_t dq_mn0 = make_smem_desc_sw128(smem_u32(sQ), 8192, 1024);
      const uint64_t do_mn0 = make_smem_desc_sw128(smem_u32(sdO), 8192, 1024);
      // dV += P^T dO, dK += dS^T Q for one QUARTER (32 queries) of the tile: every warpgroup owns a private 32-column
      // P^T | dS^T buffer (tP + 32 * half), so it never waits for the other warpgroup's products, only for its own
      // previous quarter's (which run under its next quarter's exp / multiply work)
      auto issue_y = [&](int gy, int t, int hf, int qd, bool last_of_tile) {
        const uint64_t off = static_cast<uint64_t>(((gy % AB_QSTAGES) * AB_TILE + hf * 8192 + qd * 4096) >> 4);
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
          if (last_of_tile) umma_commit(&qdo_empty[gy % AB_QSTAGES]);  // this issuer's last read of the tile's Q / dO
        }
        __syncwarp();
      };
      int g = 0;
      for (int n = 0;; ++n) {
        const int slot = n & (AB_RING - 1);
        mbar_wait(&item_full[slot], (n / AB_RING) & 1);
        const int n_iter = s_item[slot].n_iter;
        if (n_iter == 0) break;
        if (n > 0) mbar_wait(acc_free, (n - 1) & 1);  // the previous item's dK / dV have been read out of TMEM
        for (int t = 0; t < n_iter; ++t, ++g) {
          // quarters in the order the warpgroups produce them: (half 0, q 0), (half 1, q 0), (half 0, q 1), (half 1, q 1);
          // c_done[h] completes twice per tile: phase parity = quarter index
#pragma unroll
          for (int qd = 0; qd < 2; ++qd) {
            mbar_wait(&c_done[0], static_cast<uint32_t>(qd));
            tc_fence_after();
            issue_y(g, t, 0, qd, false);
            mbar_wait(&c_done[1], static_cast<uint32_t>(qd));
            tc_fence_after();
            issue_y(g, t, 1, qd, qd == 1);
          }
        }
        if (elect_one()) umma_commit(acc_full);  // arrives once every dV / dK product of the item has retired
        __syncwarp();
        if (lane == 0) mbar_arrive(&item_empty[slot]);
      }
    }
  } else if (warp >= 8) {
    // ===================== per-query helper warpgroup: thread = query of the tile =====================
    ab_setmaxnreg_dec_helper();
    const int q = threadIdx.x - 256;        // query row of every tile = TMEM lane of the dQ accumulator
    const int hq = q >> 6;                   // its 64-query half
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    // v = h + m + l exactly: three 8-bit slices of the fp32 significand, each a bf16 (truncation, no cvt instructions)
    auto split3 = [](float v, uint32_t& h, uint32_t& m, uint32_t& l) {
      const uint32_t vh = __float_as_uint(v) & 0xFFFF0000u;
      const float r1 = v - __uint_as_float(vh);
      const uint32_t vm = __float_as_uint(r1) & 0xFFFF0000u;
      const float r2 = r1 - __uint_as_float(vm);
      h = vh >> 16, m = vm >> 16, l = __float_as_uint(r2) >> 16;
    };
    // "ext" cursor: walks the flat tile sequence ahead of the drain; e_n / e_t = item / local tile of flat tile ge
    int e_n = 0, e_t = -1, e_iter = 0, ge = -1;
    bool e_end = false;
    float r_lse = CUDART_INF_F, r_dl = 0.f;  // values loaded for the NEXT tile to be written
    uint32_t r_rb = 0u;
    int r_kgrp = 0;
    bool r_valid = false, r_have = false;
    auto ext_advance_and_load = [&]() {  // move to the next flat tile and issue its per-query loads
      r_have = false;
      if (e_end) return;
      if (e_t < 0) {  // very first call
        mbar_wait(&item_full[0], 0);
        e_iter = s_item[0].n_iter;
        e_t = 0;
      } else if (++e_t >= e_iter) {
        ++e_n;
        const int slot = e_n & (AB_RING - 1);
        mbar_wait(&item_full[slot], (e_n / AB_RING) & 1);
        e_iter = s_item[slot].n_iter;
        e_t = 0;
      }
      if (e_iter == 0) {
        e_end = true;
        return;
      }
      const int slot = e_n & (AB_RING - 1);
      const AbItem I = s_item[slot];
      const int2 Q = s_qt[slot][e_t];
      const int qi = min(Q.x + q, a.N - 1);
      const long long sidx = (static_cast<long long>(I.b) * a.H + I.h) * a.N + qi;
      r_lse = a.lse[sidx];
      r_dl = a.delta[sidx];
      r_rb = a.rowbits[qi];
      r_valid = q < Q.y;
      r_kgrp = I.kgrp;
      r_have = true;
      ++ge;
    };
    auto ext_write = [&]() {  // B row of this thread's query for flat tile ge (from the registers loaded for it)
      if (!r_have) return;
      if (ge >= 2) mbar_wait(&ext_empty[ge & 1], ((ge >> 1) & 1) ^ 1);  // the products that read this buffer have retired
      const uint32_t rb = r_valid ? r_rb : 0u;
      const bool sees = r_valid && (r_kgrp == 255 || ((rb >> r_kgrp) & 1u)) && r_lse != CUDART_INF_F;
      uint32_t lh, lm, ll, dh, dm, dl;
      split3(sees ? -r_lse : AB_MASKED, lh, lm, ll);
      split3(r_valid ? -r_dl : 0.f, dh, dm, dl);
      *reinterpret_cast<uint4*>(sBX + (ge & 1) * AB_EXT_TILE + q * 16) = make_uint4(lh | (lm << 16), ll | (dh << 16), dm | (dl << 16), 0u);
      if (r_kgrp == 255) s_rb[hq][ge % 3][q & 63] = rb;
      fence_proxy_async_smem();
      mbar_arrive(&ext_full[ge & 1]);
    };
    // tiles 0 and 1 up front, tile 2 loaded
    ext_advance_and_load();
    ext_write();
    ext_advance_and_load();
    ext_write();
    ext_advance_and_load();
    int g = 0;
    for (int n = 0;; ++n) {
      const int slot = n & (AB_RING - 1);
      mbar_wait(&item_full[slot], (n / AB_RING) & 1);
      const int n_iter = s_item[slot].n_iter;
      if (n_iter == 0) break;
      const long long row0 = static_cast<long long>(s_item[slot].b) * a.N;
      const int h_item = s_item[slot].h;
      for (int t = 0; t < n_iter; ++t, ++g) {
        // extra k-step rows of tile g + 2 (values loaded one iteration ago), then the loads of tile g + 3
        ext_write();
        ext_advance_and_load();
        // dQ of tile g: TMEM -> registers -> fp32 swizzled smem -> TMA reduce-add into dq_acc, issued by this
        // warpgroup's thread 0 (which owns the bulk async-groups)
        mbar_wait(&z_full[g & 1], (g >> 1) & 1);
        tc_fence_after();
        uint32_t v0[32], v1[32];
        tmem_ld32(tdQ + lane_sel, v0);
        tmem_ld32(tdQ + lane_sel + 32, v1);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(dq_free);
        if (q == 0) bulk_wait_group_read0();  // the previous tile's reduce has finished reading the staging buffer
        ab_bar_sync(1, 128);
        uint8_t* rowp = sdQ + q * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          *reinterpret_cast<uint4*>(rowp + ((c ^ (q & 7)) << 4)) = make_uint4(v0[4 * c], v0[4 * c + 1], v0[4 * c + 2], v0[4 * c + 3]);
          *reinterpret_cast<uint4*>(rowp + AB_DQ / 2 + ((c ^ (q & 7)) << 4)) = make_uint4(v1[4 * c], v1[4 * c + 1], v1[4 * c + 2], v1[4 * c + 3]);
        }
        fence_proxy_async_smem();
        ab_bar_sync(1, 128);
        if (q == 0) {
          const int qrow = static_cast<int>(row0 + s_qt[slot][t].x);
          tma_reduce_add_2d(&tm_dq, sdQ, h_item * AB_DH, qrow);
          tma_reduce_add_2d(&tm_dq, sdQ + AB_DQ / 2, h_item * AB_DH + 32, qrow);
          bulk_commit_group();
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&item_empty[slot]);
    }
    if (q == 0) bulk_wait_group0();
  } else {
    // ===================== compute warpgroups: thread = key row, warpgroup = query half =====================
    ab_setmaxnreg_inc_compute();
    const int hf = warp >> 2;             // which 64-query half of every tile
    const int r = (warp & 3) * 32 + lane;  // key row = TMEM lane
    const int wt = threadIdx.x & 127;      // thread index inside the warpgroup
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t reg = tmem_base + hf * 64 + lane_sel;  // S^T columns of this half; dP^T is 128 columns further
    const uint32_t tp_mine = tP + 32 * hf + lane_sel;     // this warpgroup's private P^T (16 columns) | dS^T (16 columns)
    int g = 0;
    for (int n = 0;; ++n) {
      const int slot = n & (AB_RING - 1);
      mbar_wait(&item_full[slot], (n / AB_RING) & 1);
      const AbItem I = s_item[slot];
      const int n_iter = I.n_iter;
      if (n_iter == 0) break;
#ifdef MCA_TRACE
      const int item_lin = I.kt * a.BH + I.b * a.H + I.h;
      if (threadIdx.x == 0 && item_lin < 4096) { g_bcta[item_lin * 4] = gtimer_b(); g_bcta[item_lin * 4 + 2] = smid_b(); g_bcta[item_lin * 4 + 3] = n_iter; }
#endif
      const int kgrp = I.kgrp;
      const long long row0 = static_cast<long long>(I.b) * a.N;
      const bool has_dead = I.cls == 1 || I.klen < AB_T;
      bool live = false;
      uint32_t mygrp = 0;
      if (r < I.klen) {
        live = a.padding[row0 + I.kstart + r] == 0;
        mygrp = a.keygrp[I.kstart + r];
      }
      const bool dead = has_dead && !live;
      for (int t = 0; t < n_iter; ++t, ++g) {
        const uint32_t ph = g & 1;
        if (wt == 0) TR(hf, g, 0);
        mbar_wait(x_full, ph);
        if (wt == 0) TR(hf, g, 2);
        tc_fence_after();
        uint32_t sv[2][32], dv[2][32];
        tmem_ld32(reg, sv[0]);
        tmem_ld32(reg + 32, sv[1]);
        tmem_ld32(reg + 128, dv[0]);
        tmem_ld32(reg + 160, dv[1]);
        tmem_ld_wait();
        if (wt == 0) TR(hf, g, 3);
        tc_fence_before();
        mbar_arrive(x_free);  // the next tile's S^T / dP^T may be issued once both warpgroups got here
        if (wt == 0) TR(hf, g, 9);
        // (mixed-group tiles) the queries' group bits were published before ext_full, which the score products waited for
        const uint32_t* rbq = s_rb[hf][g % 3];
        // dS^T also goes to shared memory for the dQ product.  Half 0 is double buffered by tile parity; half 1 has one
        // buffer that dQ(g-1) must have finished reading.
        uint8_t* ds_row;
        if (hf == 0) {
          if (g >= 2) mbar_wait(&z_full[g & 1], ((g - 2) >> 1) & 1);
          ds_row = sdS + (g & 1) * (AB_DS / 2) + r * 128;
        } else {
          if (g >= 1) mbar_wait(&z_full[(g - 1) & 1], ((g - 1) >> 1) & 1);
          ds_row = sdS + 2 * (AB_DS / 2) + r * 128;
        }
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
          if (wt == 0 && qd == 1) TR(hf, g, 8);
          // this quarter's 32 queries = bytes [64 qd, 64 qd + 64) of the key's dS row (128B-swizzled 16-byte chunks)
#pragma unroll
          for (int c = 0; c < 4; ++c)
            *reinterpret_cast<uint4*>(ds_row + (((4 * qd + c) ^ (r & 7)) << 4)) = make_uint4(dd[4 * c], dd[4 * c + 1], dd[4 * c + 2], dd[4 * c + 3]);
          if (wt == 0 && qd == 1) TR(hf, g, 4);
          // the private buffer is free once this warpgroup's previous quarter's dV / dK products retired
          if (g > 0 || qd > 0) {
            mbar_wait(&y_done[hf], static_cast<uint32_t>(qd ^ 1));  // commit index 2g + qd - 1
            tc_fence_after();
          }
          if (wt == 0 && qd == 1) TR(hf, g, 5);
          tmem_st16(tp_mine, pp);
          tmem_st16(tp_mine + 16, dd);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&c_done[hf]);
          if (qd == 1) {
            // dS in shared memory is only read by the dQ product, which waits for ds_full: one generic -> async proxy
            // fence per tile (after both quarters' stores) instead of one per quarter
            fence_proxy_async_smem();
            mbar_arrive(ds_full);
          }
        }
        if (wt == 0) TR(hf, g, 6);
      }
      // ---- item epilogue: warpgroup 0 writes dK, warpgroup 1 writes dV (+ the uniform-row correction); thread = key row.
      mbar_wait(acc_full, n & 1);  // every dV / dK product of the item has retired
      tc_fence_after();
      if (wt == 0) TRG(3 + hf);
      {
        const bool store = r < I.klen;
        const int which = hf;  // 0: dK -> column block 1, 1: dV -> column block 2
        __nv_bfloat16* dst = a.dqkv + (row0 + I.kstart + (store ? r : 0)) * (3 * HD) + I.h * AB_DH + (which + 1) * HD;
        uint32_t t0[32], t1[32];
        // tcgen05.ld is warp-collective: every lane loads, only rows inside the tile store
        tmem_ld32((which == 0 ? tdK : tdV) + lane_sel, t0);
        tmem_ld32((which == 0 ? tdK : tdV) + lane_sel + 32, t1);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(acc_free);  // the next item's first dV / dK products may overwrite the accumulators
        const float* uc = a.ucorr + static_cast<long long>(I.b) * HD + I.h * AB_DH;
#pragma unroll
        for (int cc = 0; cc < AB_DH / 32; ++cc) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(cc == 0 ? t0[i] : t1[i]);
          if (which == 1) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 u = *reinterpret_cast<const float4*>(uc + cc * 32 + i);
              v[i] += u.x, v[i + 1] += u.y, v[i + 2] += u.z, v[i + 3] += u.w;
            }
          }
          if (store) {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              uint4 w;
              w.x = pack_bf16x2(v[8 * q4 + 0], v[8 * q4 + 1]);
              w.y = pack_bf16x2(v[8 * q4 + 2], v[8 * q4 + 3]);
              w.z = pack_bf16x2(v[8 * q4 + 4], v[8 * q4 + 5]);
              w.w = pack_bf16x2(v[8 * q4 + 6], v[8 * q4 + 7]);
              reinterpret_cast<uint4*>(dst + cc * 32)[q4] = w;
            }
          }
        }
      }
      if (wt == 0) TRG(5 + hf);
#ifdef MCA_TRACE
      if (threadIdx.x == 0 && item_lin < 4096) g_bcta[item_lin * 4 + 1] = gtimer_b();
#endif
      __syncwarp();
      if (lane == 0) mbar_arrive(&item_empty[slot]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) TRG(7);
  if (warp == 13) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  // last CTA out resets the work queue for the next launch
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&g_ab_ctr[1], 1u) == gridDim.x - 1) {
      g_ab_ctr[0] = 0u;
      g_ab_ctr[1] = 0u;
      __threadfence();
    }
  }
}

// delta[b,h,n] = sum_c dO*O ; ucorr[b, h*64+c] += dO[row, h*64+c] / N for rows whose lse is +inf (fully masked); dq_zero[row] = 0.
// Block (x, b) walks 16 rows of sample b (two per warp: 320 blocks of 64 rows left the HBM pipe half empty); the masked rows' dO are summed in registers, then across the 8 warps in shared
// memory, and only then added to ucorr — one atomic per column and block instead of one per column and row (with 40 %
// modality dropout a third of all rows are fully masked and the per-row atomics on 4096 addresses dominated the step).
constexpr int PREP_ROWS = 16;
__global__ void __launch_bounds__(256)
attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout,
                     const float* __restrict__ lse, float* __restrict__ delta, float* __restrict__ ucorr,
                     float* __restrict__ dq_zero, int B, int N, int H) {
  __shared__ float s_part[8][512];
  pdl_launch_dependents();
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
    // the fp32 dQ accumulator of this row is cleared here (the main kernel reduce-adds into it): one pass over the rows
    // instead of a separate 42 MB memset in front of it
    float4* z4 = reinterpret_cast<float4*>(dq_zero + row * HD + c0);
#pragma unroll
    for (int i = 0; i < 4; ++i) z4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
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
                            const uint8_t* tile_grp, const uint8_t* padding, const uint8_t* kt_class,
                            const uint8_t* skip_ok, float* delta, float* ucorr, float* dq_accum, void* dqkv, int B, int N,
                            int H, void* stream_) {
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
  if (cudaMemsetAsync(ucorr, 0, static_cast<size_t>(B) * HD * sizeof(float), stream) != cudaSuccess) return MCA_ERR_CUDA;
  attn_bwd_prep_kernel<<<dim3((N + PREP_ROWS - 1) / PREP_ROWS, B), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(out), reinterpret_cast<const __nv_bfloat16*>(dout), lse, delta, ucorr, dq_accum, B,
      N, H);
  const int n_items = B * H * n_kt;  // item index = key tile (slow) x (sample, head)
  AttnBwdArgs a{k_tiles_q, qt_list, q_tiles, rowbits, keygrp, tile_grp, padding, kt_class, skip_ok, lse, delta, ucorr,
                reinterpret_cast<__nv_bfloat16*>(dqkv), N, H, n_kt, n_items, B * H};
  const int grid = n_items < num_sms() ? n_items : num_sms();  // persistent: one CTA per SM pulls from the queue
  if (launch_kernel(attn_bwd_kernel, dim3(grid), dim3(AB_THREADS), AB_SMEM, stream, 1, tm_qkv, tm_do, tm_dq, a) != cudaSuccess)
    return MCA_ERR_CUDA;
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
