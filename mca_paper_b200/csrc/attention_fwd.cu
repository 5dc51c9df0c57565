// Block-sparse masked multi-head attention forward on tcgen05 (replaces model.py:85-100: q*scale, QK^T, static
// attn_mask fill, key_padding_mask fill, softmax, PV — without ever materialising the [B,h,N,N] score tensor).
//
// One CTA = one (sample, head, 128-row query tile); two CTAs share an SM so one CTA's tensor work hides the other's
// softmax.  Warps 0-3: softmax, one thread per query row (TMEM lane); warp 4: TMA producer (K and V double
// buffered); warp 5: tcgen05.mma issuer.  Per visited key tile t:
//     S_t = Q K_t^T  (128x128x64, smem x smem -> TMEM)   issued as soon as the softmax warps have copied S_{t-1} to
//                                                          registers, so it overlaps the softmax of tile t-1
//     P_t = exp2(S_t*log2e - m)  -> bf16 pairs written back to TMEM (tcgen05.st), never to shared memory
//     O  += P_t V_t  (A operand from TMEM, V consumed MN-major as loaded), accumulated in TMEM over all tiles
// The running maximum m is only raised when the tile maximum exceeds it by more than 2^8 (softmax is shift
// invariant, P stays below 256), so the O accumulator in TMEM is rescaled on a small minority of tiles.
// Key tiles come from a static, host-built schedule (only tiles holding at least one statically allowed
// (query,key) pair — 43 % of all pairs for MCA at CMU shape), heaviest query tiles first, and are skipped at run time
// when every key in them is padded for this sample.  Partially allowed / partially padded / ragged tiles apply a
// 128-bit per-row mask: live-key bits built by mca_build_offsets, ANDed with the row's allowed key groups.
// Reference quirk: masks are filled with -finfo.max, so a query row with no live allowed key is
// exactly uniform over ALL N keys; those rows get the per-(sample,head) mean of V and LSE = +inf (which makes the
// backward treat their P as 0; their 1/N contribution to dV is added separately).
#include <math_constants.h>

#include "mca_b200.h"
#include "ptx.cuh"
#include "runtime.h"

namespace mca {

constexpr int AT_BM = 128;    // query rows per tile
constexpr int AT_BN = 128;    // keys per tile
constexpr int AT_DH = 64;
constexpr int AT_SOFT_WARPS = 4;                       // one thread per query row (TMEM lane)
constexpr int AT_THREADS = (AT_SOFT_WARPS + 2) * 32;   // + TMA warp + MMA warp
constexpr int AT_TILE_BYTES = AT_BM * AT_DH * 2;       // 16 KB
constexpr int AT_MAX_KT = 128;                         // key tiles one query tile may visit (the launcher checks)
constexpr int AT_SCHED_BYTES = AT_MAX_KT * 32;
constexpr int AT_SMEM = 6 * AT_TILE_BYTES + 2 * AT_SCHED_BYTES;  // 2 x Q, 2 x K, 2 x V, 2 schedules
constexpr float LOG2E = 1.4426950408889634f;
constexpr float AT_RESCALE_THRESHOLD = 8.0f;  // log2 units

#ifdef MCA_TRACE
// debug-only timeline (-DMCA_TRACE): clock64 stamps of one CTA [role][tile][event] + (start, end, smid) of every work item
__device__ long long g_ftrace[10 * 24 * 8 + 8];
__device__ long long g_fcta[4096 * 4];
#define FTR(role, t, e) do { if (blockIdx.x == 5 && (t) >= 24 && (t) < 48) g_ftrace[((role) * 24 + (t) - 24) * 8 + (e)] = clock64(); } while (0)
__device__ __forceinline__ long long gtimer() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ int smid() { int s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); return s; }
#else
#define FTR(role, t, e) do { } while (0)
#endif

struct AttnFwdArgs {
  const mca_attn_qtile* q_tiles;  // sorted by descending work
  const mca_attn_ref* kt_list;
  const mca_attn_tile* k_tiles;
  const uint32_t* rowbits;   // [N]
  const uint8_t* keygrp;     // [N]
  const uint8_t* tile_grp;   // [n_kt] key group of the tile, 255 = mixed
  const uint8_t* kt_class;   // [B, n_kt]
  const uint32_t* kt_live;   // [B, n_kt, 4] live-key bits
  const uint8_t* padding;    // [B, N] (only read when skip_ok != nullptr)
  const uint8_t* skip_ok;    // [B] or nullptr: query tiles of this sample whose rows are ALL padded need no scores
  const float* vmean;        // [B, H*64]
  __nv_bfloat16* out;        // [B*N, H*64]
  float* lse;                // [B, H, N]
  int N, H, n_kt;
  int n_items, BH;           // work items = (query tile, sample, head), query tile slow
};

// One visited key tile of a work item, staged in shared memory by the producer warp so that no role chases global
// pointers inside the loop (the kt_list -> k_tiles / kt_class / tile_grp / kt_live chain cost ~430 cycles per tile).
struct __align__(16) AtSched {
  int kstart;        // first key position inside the sample
  int klen;          // valid keys of the tile
  int flags;         // bit 0: per-element masking needed, bits 8..15: key group (255 = mixed)
  int pad;
  uint32_t live[4];  // live-key bits of this (sample, tile)
};
struct AtItem { int n_it, b, h, qstart, qlen, item, pad0, pad1; };  // n_it < 0: no more work

// Work queue of the persistent kernel: [0] next item, [1] CTAs that have left.  The last CTA out resets both, so the
// counters are zero again when the next launch (same stream) starts; the library never allocates device memory.
__device__ unsigned int g_at_ctr[2];

// The softmax role.  Every shared-memory object it touches is addressed through the static shared window (constant
// offsets): generic 64-bit pointers for the seven barriers and the schedule cost ~20 registers and, next to the 128
// registers of the S tile, ~1 KB of spills per thread.
__device__ __forceinline__ void at_softmax_role(const AttnFwdArgs& a, AtSched* sched_all, const AtItem* s_item, uint64_t* bars,
                                                uint32_t tmem_base) {
  uint64_t* item_full = bars + 0;
  uint64_t* item_empty = bars + 2;
  uint64_t* s_full = bars + 16;
  uint64_t* s_empty = bars + 17;
  uint64_t* p_full = bars + 18;
  uint64_t* pv_done = bars + 19;
  uint64_t* o_free = bars + 20;
  const uint32_t tS = tmem_base, tP = tmem_base + 128, tO = tmem_base + 192;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    // ===================== softmax / epilogue: thread = query row (TMEM lane) =====================
    // (Two other layouts were measured at the same 2 CTAs per SM: two threads per row in different warps, 112 us, and a
    // quad of lanes per row with the 16x256b fragment, 106 us against 111 us for this one; with the work queue added
    // both no longer fit the 96 registers that 10 warps per CTA leave and spilled: profiles/r2_attn_notes.md.)
    const int r = warp * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
    int g = 0;
    for (int n = 0;; ++n) {
      const int slot = n & 1;
      mbar_wait(&item_full[slot], (n >> 1) & 1);
      const int n_it = s_item[slot].n_it;
      if (n_it < 0) break;
#ifdef MCA_TRACE
      if (warp == 0 && lane == 0 && s_item[slot].item < 4096) { const int ii = s_item[slot].item; g_fcta[ii * 4] = gtimer(); g_fcta[ii * 4 + 2] = smid(); g_fcta[ii * 4 + 3] = n_it; }
#endif
      const AtSched* sched = sched_all + slot * AT_MAX_KT;
      // allowed key groups of this row (the row may run past the tile's valid rows: it is then never stored)
      const uint32_t rb = a.rowbits[min(s_item[slot].qstart + r, a.N - 1)];
      float m2 = -CUDART_INF_F, l_run = 0.f;  // reference maximum (log2 units) and running sum
      for (int it = 0; it < n_it; ++it, ++g) {
        const bool masked = sched[it].flags & 1;
        if (lane == 0) FTR(warp, g, 0);
        mbar_wait(s_full, g & 1);
        if (lane == 0) FTR(warp, g, 1);
        tc_fence_after();
        uint32_t sv[4][32];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) tmem_ld32(tS + lane_sel + cc * 32, sv[cc]);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(s_empty);  // S is in registers: the next QK^T may overwrite it
        if (lane == 0) FTR(warp, g, 2);
        if (masked) {
          // allowed-key bits: the tile's live-key words, ANDed with the row's visibility of the tile's key group (or of
          // every key's group for the mixed fusion sub-block tiles)
          const AtSched e = sched[it];
          const int grp = (e.flags >> 8) & 255;
          uint32_t aw[4] = {e.live[0], e.live[1], e.live[2], e.live[3]};
          if (grp != 255) {
            if (((rb >> grp) & 1u) == 0) aw[0] = aw[1] = aw[2] = aw[3] = 0;
          } else {
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              uint32_t bits = 0;
#pragma unroll 1
              for (int j = 0; j < 32; ++j) {
                const int kj = e.kstart + w * 32 + j;
                if (w * 32 + j < e.klen && ((rb >> a.keygrp[kj]) & 1u)) bits |= 1u << j;
              }
              aw[w] &= bits;
            }
          }
#pragma unroll
          for (int cc = 0; cc < 4; ++cc)
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (!((aw[cc] >> i) & 1u)) sv[cc][i] = __float_as_uint(-CUDART_INF_F);
        }
        float mx0 = -CUDART_INF_F, mx1 = -CUDART_INF_F, mx2 = -CUDART_INF_F, mx3 = -CUDART_INF_F;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          mx0 = fmaxf(mx0, __uint_as_float(sv[0][i]));
          mx1 = fmaxf(mx1, __uint_as_float(sv[1][i]));
          mx2 = fmaxf(mx2, __uint_as_float(sv[2][i]));
          mx3 = fmaxf(mx3, __uint_as_float(sv[3][i]));
        }
        const float t2 = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * LOG2E;
        const bool grow = t2 > m2 + AT_RESCALE_THRESHOLD;
        const float m2n = grow ? t2 : m2;
        const float alpha = grow ? fast_ex2(m2 - m2n) : 1.0f;  // m2 = -inf -> 0
        if (lane == 0) FTR(warp, g, 3);
        if (g > 0) {
          // P of the previous tile has been consumed, O is quiescent (the previous tile may belong to the previous item,
          // whose epilogue already waited for it: the barrier is then simply found complete)
          mbar_wait(pv_done, (g - 1) & 1);
          tc_fence_after();
          if (it > 0 && __any_sync(0xffffffffu, grow)) {
#pragma unroll
            for (int cc = 0; cc < AT_DH / 16; ++cc) {
              uint32_t ov[16];
              tmem_ld16(tO + lane_sel + cc * 16, ov);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
              tmem_st16(tO + lane_sel + cc * 16, ov);
            }
          }
        }
        if (lane == 0) FTR(warp, g, 4);
        l_run *= alpha;
        m2 = m2n;
        const float moff = (m2 == -CUDART_INF_F) ? 0.f : m2;
        uint64_t sum2 = f2_pack(0.f, 0.f);
        const uint64_t log2e2 = f2_pack(LOG2E, LOG2E), nmoff2 = f2_pack(-moff, -moff);
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) {  // 32 keys -> 16 packed columns
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float t0, t1;
            f2_unpack(f2_fma(f2_pack(__uint_as_float(sv[hh][2 * j]), __uint_as_float(sv[hh][2 * j + 1])), log2e2, nmoff2), t0, t1);
            const float p0 = fast_ex2(t0), p1 = fast_ex2(t1);
            sum2 = f2_add(sum2, f2_pack(p0, p1));
            pk[j] = pack_bf16x2(p0, p1);
          }
          tmem_st16(tP + lane_sel + hh * 16, pk);
        }
        {
          float sum0, sum1;
          f2_unpack(sum2, sum0, sum1);
          l_run += sum0 + sum1;
        }
        if (lane == 0) FTR(warp, g, 5);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(p_full);
        if (lane == 0) FTR(warp, g, 6);
      }
      // ---- item epilogue
      uint32_t ov[2][32];
      float lse = CUDART_INF_F;
      if (n_it > 0) {  // uniform across the CTA
        mbar_wait(pv_done, (g - 1) & 1);
        tc_fence_after();
        tmem_ld32(tO + lane_sel, ov[0]);
        tmem_ld32(tO + lane_sel + 32, ov[1]);
        tmem_ld_wait();
        tc_fence_before();
      } else {
        // defined on every path: an undefined accumulator would be live across the whole tile loop above (64 registers
        // next to the 128 of the S tile = 1 KB of spills per thread)
#pragma unroll
        for (int i = 0; i < AT_DH; ++i) ov[i >> 5][i & 31] = 0u;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free);  // the next item's first PV may overwrite the accumulator
      const AtItem I = s_item[slot];
      const int b = I.b, h = I.h;
      const long long row0 = static_cast<long long>(b) * a.N;
      const int qi = I.qstart + r;
      const bool row_valid = r < I.qlen;
      if (l_run != 0.f) {
        const float inv = 1.0f / l_run;
#pragma unroll
        for (int i = 0; i < AT_DH; ++i) ov[i >> 5][i & 31] = __float_as_uint(__uint_as_float(ov[i >> 5][i & 31]) * inv);
        lse = (m2 + log2f(l_run)) * 0.6931471805599453f;
      } else {  // reference quirk Q4: a row with no live allowed key is uniform over all N keys
        const float* vm = a.vmean + static_cast<long long>(b) * a.H * AT_DH + h * AT_DH;
#pragma unroll
        for (int i = 0; i < AT_DH; ++i) ov[i >> 5][i & 31] = __float_as_uint(vm[i]);
      }
      if (row_valid) {
        __nv_bfloat16* orow = a.out + (row0 + qi) * (a.H * AT_DH) + h * AT_DH;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint32_t* v = &ov[q >> 2][(q & 3) * 8];
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(v[0]), __uint_as_float(v[1]));
          w.y = pack_bf16x2(__uint_as_float(v[2]), __uint_as_float(v[3]));
          w.z = pack_bf16x2(__uint_as_float(v[4]), __uint_as_float(v[5]));
          w.w = pack_bf16x2(__uint_as_float(v[6]), __uint_as_float(v[7]));
          reinterpret_cast<uint4*>(orow)[q] = w;
        }
        a.lse[(static_cast<long long>(b) * a.H + h) * a.N + qi] = lse;
      }
#ifdef MCA_TRACE
      if (warp == 0 && lane == 0 && I.item < 4096) g_fcta[I.item * 4 + 1] = gtimer();
#endif
      __syncwarp();
      if (lane == 0) mbar_arrive(&item_empty[slot]);
    }
    }
}

// Persistent kernel: 2 CTAs per SM, each pulls (query tile, sample, head) items from the queue, heaviest first.  The
// pipeline is carried ACROSS items: the producer warp already loads the next item's Q / K / V and stages its schedule
// while the current item is computed, the MMA warp issues the next item's first QK^T before the current item's last
// PV, TMEM / barriers / descriptors are set up once per CTA.  (A CTA per item spent ~27 % of its life in launch
// latency, set-up and the exposed head / tail of its pipeline: profiles/r2_attn_notes.md.)
__global__ void __launch_bounds__(AT_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const AttnFwdArgs a) {
  extern __shared__ __align__(1024) uint8_t base[];  // kept in the shared address space: no generic-pointer casts
  uint8_t* sQ = base;                      // 2 stages (by item parity)
  uint8_t* sK = base + 2 * AT_TILE_BYTES;  // 2 stages (by tile parity)
  uint8_t* sV = base + 4 * AT_TILE_BYTES;  // 2 stages
  AtSched* sched_all = reinterpret_cast<AtSched*>(base + 6 * AT_TILE_BYTES);  // 2 x AT_MAX_KT entries
  __shared__ uint64_t bars[24];
  __shared__ uint32_t tmem_holder_s;
  __shared__ AtItem s_item[2];
  uint64_t* item_full = bars + 0;   // [2] schedule + header of the item staged
  uint64_t* item_empty = bars + 2;  // [2] every consumer is done with the slot
  uint64_t* q_full = bars + 4;      // [2]
  uint64_t* q_empty = bars + 6;     // [2]
  uint64_t* k_full = bars + 8;      // [2]
  uint64_t* k_empty = bars + 10;    // [2]
  uint64_t* v_full = bars + 12;     // [2]
  uint64_t* v_empty = bars + 14;    // [2]
  uint64_t* s_full = bars + 16;
  uint64_t* s_empty = bars + 17;
  uint64_t* p_full = bars + 18;
  uint64_t* pv_done = bars + 19;
  uint64_t* o_free = bars + 20;     // the softmax warps have copied the item's O accumulator to registers
  uint32_t* tmem_holder = &tmem_holder_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();

  if (warp == AT_SOFT_WARPS && lane == 0) {
    tma_prefetch_desc(&tm_qkv);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&item_full[i], 1);
      mbar_init(&item_empty[i], 1 + AT_SOFT_WARPS);
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, AT_SOFT_WARPS * 32);
    mbar_init(p_full, AT_SOFT_WARPS * 32);
    mbar_init(pv_done, 1);
    mbar_init(o_free, AT_SOFT_WARPS);
    fence_mbar_init();
  }
  if (warp == AT_SOFT_WARPS + 1) tmem_alloc(tmem_holder, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const uint32_t tS = tmem_base, tP = tmem_base + 128, tO = tmem_base + 192;
  pdl_wait();  // set-up ran under the predecessor's tail (PDL); nothing above touches global memory

  if (warp == AT_SOFT_WARPS) {
    // ===================== producer: work queue, schedule staging, TMA loads (whole warp loops) =====================
    int g = 0;  // running tile counter of this CTA (K / V ring position)
    for (int n = 0;; ++n) {
      const int slot = n & 1;
      const uint32_t iph = (n >> 1) & 1;
      mbar_wait(&item_empty[slot], iph ^ 1);
      unsigned int item = 0;
      if (lane == 0) item = atomicAdd(&g_at_ctr[0], 1u);
      item = __shfl_sync(0xffffffffu, item, 0);
      if (item >= static_cast<unsigned int>(a.n_items)) {
        if (lane == 0) {
          s_item[slot].n_it = -1;
          mbar_arrive(&item_full[slot]);
        }
        break;
      }
      const int y = static_cast<int>(item) / a.BH, bh = static_cast<int>(item) % a.BH;
      const int b = bh / a.H, h = bh % a.H;
      const mca_attn_qtile Q = a.q_tiles[y];
      const long long row0 = static_cast<long long>(b) * a.N;
      // the query tile is requested before the schedule is staged
      mbar_wait(&q_empty[slot], iph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&q_full[slot], AT_TILE_BYTES);
        tma_load_2d(sQ + slot * AT_TILE_BYTES, &tm_qkv, &q_full[slot], h * AT_DH, static_cast<int>(row0 + Q.start));
      }
      __syncwarp();
      // varlen: a query tile whose rows are all padded produces rows nobody reads when the sample's flag allows it
      // (mca_query_skip_flags): it is staged with an empty schedule (its rows get the fully-masked value, LSE = +inf)
      bool dead_q = false;
      if (a.skip_ok != nullptr && a.skip_ok[b] != 0) {
        const uint8_t* pr = a.padding + row0 + Q.start;
        bool live = false;
        for (int i = lane; i < Q.len; i += 32) live |= pr[i] == 0;
        dead_q = !__any_sync(0xffffffffu, live);
      }
      // stage the item's schedule: the visited key tiles that hold at least one live key for this sample, in order
      AtSched* sched = sched_all + slot * AT_MAX_KT;
      const uint8_t* cls = a.kt_class + static_cast<long long>(b) * a.n_kt;
      int n_it = 0;
      for (int t0 = 0; t0 < Q.kt_cnt; t0 += 32) {
        const int t = t0 + lane;
        bool keep = false;
        AtSched e;
        if (t < Q.kt_cnt) {
          const mca_attn_ref ref = a.kt_list[Q.kt_off + t];
          const int c = cls[ref.tile];
          keep = c != 2 && !dead_q;
          const mca_attn_tile K = a.k_tiles[ref.tile];
          const int grp = a.tile_grp[ref.tile];
          const bool masked = (ref.flags & 1) || c == 1 || K.len < AT_BN || grp == 255;
          const uint4 lw = *reinterpret_cast<const uint4*>(a.kt_live + (static_cast<long long>(b) * a.n_kt + ref.tile) * 4);
          e.kstart = K.start, e.klen = K.len, e.flags = (masked ? 1 : 0) | (grp << 8), e.pad = 0;
          e.live[0] = lw.x, e.live[1] = lw.y, e.live[2] = lw.z, e.live[3] = lw.w;
        }
        const uint32_t m = __ballot_sync(0xffffffffu, keep);
        if (keep) sched[n_it + __popc(m & ((1u << lane) - 1u))] = e;
        n_it += __popc(m);
      }
      __syncwarp();
      if (lane == 0) {
        s_item[slot] = AtItem{n_it, b, h, Q.start, Q.len, static_cast<int>(item), 0, 0};
        mbar_arrive(&item_full[slot]);
      }
      __syncwarp();
      for (int it = 0; it < n_it; ++it, ++g) {
        const int krow = static_cast<int>(row0 + sched[it].kstart);
        const int st = g & 1;
        const uint32_t ph = (g >> 1) & 1;
        if (lane == 0) FTR(9, g, 0);
        mbar_wait(&k_empty[st], ph ^ 1);
        if (lane == 0) FTR(9, g, 1);
        if (elect_one()) {
          mbar_expect_tx(&k_full[st], AT_TILE_BYTES);
          tma_load_2d(sK + st * AT_TILE_BYTES, &tm_qkv, &k_full[st], a.H * AT_DH + h * AT_DH, krow);
        }
        __syncwarp();
        mbar_wait(&v_empty[st], ph ^ 1);
        if (lane == 0) FTR(9, g, 2);
        if (elect_one()) {
          mbar_expect_tx(&v_full[st], AT_TILE_BYTES);
          tma_load_2d(sV + st * AT_TILE_BYTES, &tm_qkv, &v_full[st], 2 * a.H * AT_DH + h * AT_DH, krow);
        }
        __syncwarp();
      }
    }
  } else if (warp == AT_SOFT_WARPS + 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the loop and the waits (convergent code keeps the descriptors in uniform registers, so a
    // tcgen05.mma costs one uniform add); one elected lane issues the MMAs and commits.  The loop is flat over the tiles
    // of successive items: QK^T of tile g is issued before PV of tile g - 1, also across an item boundary.
    constexpr uint32_t idesc_s = make_idesc_bf16(AT_BM, AT_BN, false, false);
    constexpr uint32_t idesc_o = make_idesc_bf16(AT_BM, AT_DH, false, true);
    const uint64_t dq0 = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
    const uint64_t dk0 = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
    const uint64_t dv0 = make_smem_desc_sw128(smem_u32(sV), 8192, 1024);
    int pend_g = -1, pend_n = 0;   // tile whose PV is still to be issued, and its item
    bool pend_first = false;
    auto issue_pv = [&]() {
      const int j = pend_g, st = j & 1;
      mbar_wait(&v_full[st], (j >> 1) & 1);
      if (lane == 0) FTR(8, j, 4);
      mbar_wait(p_full, j & 1);
      if (lane == 0) FTR(8, j, 5);
      // the first PV of an item overwrites the accumulator: the previous item's O must have been read out
      if (pend_first && pend_n > 0) mbar_wait(o_free, (pend_n - 1) & 1);
      tc_fence_after();
      const uint64_t dv = dv0 + static_cast<uint64_t>((st * AT_TILE_BYTES) >> 4);
      const uint32_t acc0 = pend_first ? 0u : 1u;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < AT_BN / 16; ++k) umma_bf16_ts(tO, tP + k * 8, dv + k * (2048 >> 4), idesc_o, k > 0 ? 1u : acc0);
        umma_commit(&v_empty[st]);
        umma_commit(pv_done);
      }
      __syncwarp();
      if (lane == 0) FTR(8, j, 6);
    };
    int g = 0;
    for (int n = 0;; ++n) {
      const int slot = n & 1;
      const uint32_t iph = (n >> 1) & 1;
      mbar_wait(&item_full[slot], iph);
      const int n_it = s_item[slot].n_it;
      if (n_it < 0) break;
      if (n_it > 0) mbar_wait(&q_full[slot], iph);
      const uint64_t dq = dq0 + static_cast<uint64_t>((slot * AT_TILE_BYTES) >> 4);
      for (int it = 0; it < n_it; ++it, ++g) {
        const int st = g & 1;
        if (lane == 0) FTR(8, g, 0);
        mbar_wait(&k_full[st], (g >> 1) & 1);
        if (lane == 0) FTR(8, g, 1);
        mbar_wait(s_empty, (g & 1) ^ 1);
        if (lane == 0) FTR(8, g, 2);
        tc_fence_after();
        const uint64_t dk = dk0 + static_cast<uint64_t>((st * AT_TILE_BYTES) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < AT_DH / 16; ++k) umma_bf16(tS, dq + k * 2, dk + k * 2, idesc_s, k > 0 ? 1u : 0u);
          umma_commit(&k_empty[st]);
          if (it == n_it - 1) umma_commit(&q_empty[slot]);  // the item's last read of its Q tile
          umma_commit(s_full);
        }
        __syncwarp();
        if (lane == 0) FTR(8, g, 3);
        if (pend_g >= 0) issue_pv();
        pend_g = g, pend_n = n, pend_first = it == 0;
      }
      if (n_it == 0 && elect_one()) mbar_arrive(&q_empty[slot]);  // nothing read the (loaded) Q tile
      __syncwarp();
      if (lane == 0) mbar_arrive(&item_empty[slot]);
      // The last PV of the item is normally issued behind the next item's first QK^T.  It may only be deferred when
      // that item is already staged and has tiles: the softmax warps wait for this PV before they release the item
      // slot, and the producer needs a free slot to stage anything further (a run of empty items would deadlock).
      if (pend_g >= 0) {
        const int ns = (n + 1) & 1;
        bool defer = mbar_try_wait(&item_full[ns], ((n + 1) >> 1) & 1);
        defer = __shfl_sync(0xffffffffu, defer ? 1 : 0, 0) != 0;
        if (defer) defer = s_item[ns].n_it > 0;
        if (!defer) {
          issue_pv();
          pend_g = -1;
        }
      }
    }
    if (pend_g >= 0) issue_pv();
  } else {
    at_softmax_role(a, sched_all, s_item, bars, tmem_base);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == AT_SOFT_WARPS + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
  // last CTA out resets the work queue for the next launch
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&g_at_ctr[1], 1u) == gridDim.x - 1) {
      g_at_ctr[0] = 0u;
      g_at_ctr[1] = 0u;
      __threadfence();
    }
  }
}

// vmean[b, c] = mean over all N rows of V[b, :, c]; only needed when some modality is absent (device flag).
// Block (x, b) sums 64 rows of sample b (thread = two adjacent bf16 columns, coalesced 1 KB per row) and adds its partial
// to vmean (zeroed by the launcher): 40 x B blocks instead of one long-running block per (sample, 128 columns).
constexpr int VM_ROWS = 64;
__global__ void __launch_bounds__(256)
vmean_kernel(const __nv_bfloat16* __restrict__ qkv, int ld, int v_col0, int width, int N, const int* __restrict__ any_absent,
             float* __restrict__ vmean) {
  pdl_launch_dependents();
  if (*any_absent == 0) return;
  const int b = blockIdx.y;
  const int n0 = blockIdx.x * VM_ROWS, n1 = min(N, n0 + VM_ROWS);
  const float inv = 1.0f / static_cast<float>(N);
  for (int c = 2 * threadIdx.x; c < width; c += 2 * blockDim.x) {
    float a0 = 0.f, a1 = 0.f;
    const __nv_bfloat16* col = qkv + static_cast<long long>(b) * N * ld + v_col0 + c;
#pragma unroll 8
    for (int n = n0; n < n1; ++n) {
      const uint32_t w = *reinterpret_cast<const uint32_t*>(col + static_cast<long long>(n) * ld);
      a0 += bf16_lo(w), a1 += bf16_hi(w);
    }
    atomicAdd(vmean + static_cast<long long>(b) * width + c, a0 * inv);
    atomicAdd(vmean + static_cast<long long>(b) * width + c + 1, a1 * inv);
  }
}

}  // namespace mca

using namespace mca;

extern "C" int mca_attn_fwd(const void* qkv, const mca_attn_qtile* q_tiles, int n_qt, const mca_attn_ref* kt_list,
                            const mca_attn_tile* k_tiles, int n_kt, const uint32_t* rowbits, const uint8_t* keygrp,
                            const uint8_t* tile_grp, const uint8_t* kt_class, const uint32_t* kt_live,
                            const uint8_t* padding, const uint8_t* skip_ok, const int* any_absent, float* vmean, void* out,
                            float* lse, int B, int N, int H, void* stream_) {
  if (B <= 0 || N <= 0 || H <= 0 || n_qt <= 0 || n_kt > AT_MAX_KT) return MCA_ERR_SHAPE;
  if (skip_ok != nullptr && padding == nullptr) return MCA_ERR_ARG;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int ld = 3 * H * AT_DH;
  CUtensorMap tm;
  int rc = make_tmap_2d_bf16(&tm, qkv, static_cast<uint64_t>(ld), static_cast<uint64_t>(B) * N, static_cast<uint64_t>(ld),
                             AT_DH, AT_BM);
  if (rc != MCA_OK) return rc;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM) != cudaSuccess)
      return MCA_ERR_CUDA;
    attr = true;
  }
  if (cudaMemsetAsync(vmean, 0, static_cast<size_t>(B) * H * AT_DH * sizeof(float), stream) != cudaSuccess) return MCA_ERR_CUDA;
  dim3 gv((N + VM_ROWS - 1) / VM_ROWS, B);
  vmean_kernel<<<gv, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(qkv), ld, 2 * H * AT_DH, H * AT_DH, N,
                                       any_absent, vmean);
  const int n_items = B * H * n_qt;  // item index = query tile (slow, heaviest first) x (sample, head)
  AttnFwdArgs a{q_tiles, kt_list, k_tiles, rowbits, keygrp, tile_grp, kt_class, kt_live, padding, skip_ok, vmean,
                reinterpret_cast<__nv_bfloat16*>(out), lse, N, H, n_kt, n_items, B * H};
  const int grid = n_items < 2 * num_sms() ? n_items : 2 * num_sms();  // persistent: two CTAs per SM pull from the queue
  if (launch_kernel(attn_fwd_kernel, dim3(grid), dim3(AT_THREADS), AT_SMEM, stream, 1, tm, a) != cudaSuccess) return MCA_ERR_CUDA;
  return check_launch();
}

#ifdef MCA_TRACE
extern "C" int mca_debug_read_trace_fwd(long long* tr, int n_tr, long long* cta, int n_cta) {
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(tr, mca::g_ftrace, sizeof(long long) * n_tr) != cudaSuccess) return 3;
  return cudaMemcpyFromSymbol(cta, mca::g_fcta, sizeof(long long) * n_cta) == cudaSuccess ? 0 : 3;
}
#endif
