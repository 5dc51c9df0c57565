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
constexpr int AT_SOFT_WARPS = 8;                       // 16 query rows per softmax warp (a quad of lanes per row)
constexpr int AT_THREADS = (AT_SOFT_WARPS + 2) * 32;   // + TMA warp + MMA warp
constexpr int AT_TILE_BYTES = AT_BM * AT_DH * 2;       // 16 KB
constexpr int AT_MAX_KT = 128;                         // key tiles one query tile may visit (the launcher checks)
constexpr int AT_SCHED_BYTES = AT_MAX_KT * 32;
constexpr int AT_SMEM = 5 * AT_TILE_BYTES + AT_SCHED_BYTES + 1024;  // Q, 2 x K, 2 x V, schedule (+ alignment slack)
constexpr float LOG2E = 1.4426950408889634f;
constexpr float AT_RESCALE_THRESHOLD = 8.0f;  // log2 units

#ifdef MCA_TRACE
// debug-only timeline (-DMCA_TRACE): clock64 stamps of one CTA [role][tile][event] + (start, end, smid) of every CTA
__device__ long long g_ftrace[10 * 24 * 8 + 8];
__device__ long long g_fcta[4096 * 4];
#define FTR(role, t, e) do { if (blockIdx.x == 5 && blockIdx.y == 3 && (t) < 24) g_ftrace[((role) * 24 + (t)) * 8 + (e)] = clock64(); } while (0)
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
  const float* vmean;        // [B, H*64]
  __nv_bfloat16* out;        // [B*N, H*64]
  float* lse;                // [B, H, N]
  int N, H, n_kt;
};

// One visited key tile of this CTA, staged in shared memory by the prologue so that no role chases global pointers
// inside the loop (the kt_list -> k_tiles / kt_class / tile_grp / kt_live chain cost ~430 cycles per tile).
struct __align__(16) AtSched {
  int kstart;        // first key position inside the sample
  int klen;          // valid keys of the tile
  int flags;         // bit 0: per-element masking needed, bits 8..15: key group (255 = mixed)
  int pad;
  uint32_t live[4];  // live-key bits of this (sample, tile)
};

__global__ void __launch_bounds__(AT_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const AttnFwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = base;
  uint8_t* sK = base + AT_TILE_BYTES;      // 2 stages
  uint8_t* sV = base + 3 * AT_TILE_BYTES;  // 2 stages
  AtSched* sched = reinterpret_cast<AtSched*>(base + 5 * AT_TILE_BYTES);
  __shared__ uint64_t bars[13];
  __shared__ uint32_t tmem_holder_s;
  __shared__ int s_nit;
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;   // [2]
  uint64_t* k_empty = bars + 3;  // [2]
  uint64_t* v_full = bars + 5;   // [2]
  uint64_t* v_empty = bars + 7;  // [2]
  uint64_t* s_full = bars + 9;
  uint64_t* s_empty = bars + 10;
  uint64_t* p_full = bars + 11;
  uint64_t* pv_done = bars + 12;
  uint32_t* tmem_holder = &tmem_holder_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x % a.H, b = blockIdx.x / a.H;
#ifdef MCA_TRACE
  const int cta_lin = blockIdx.y * gridDim.x + blockIdx.x;
  if (threadIdx.x == 0 && cta_lin < 4096) { g_fcta[cta_lin * 4] = gtimer(); g_fcta[cta_lin * 4 + 2] = smid(); g_fcta[cta_lin * 4 + 3] = clock64(); }
#endif
  const mca_attn_qtile Q = a.q_tiles[blockIdx.y];
  const long long row0 = static_cast<long long>(b) * a.N;

  if (warp == AT_SOFT_WARPS) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_qkv);
      mbar_init(q_full, 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&k_full[i], 1);
        mbar_init(&k_empty[i], 1);
        mbar_init(&v_full[i], 1);
        mbar_init(&v_empty[i], 1);
      }
      mbar_init(s_full, 1);
      mbar_init(s_empty, AT_SOFT_WARPS * 32);
      mbar_init(p_full, AT_SOFT_WARPS * 32);
      mbar_init(pv_done, 1);
      fence_mbar_init();
      // the query tile is requested before anything else of the prologue
      mbar_expect_tx(q_full, AT_TILE_BYTES);
      tma_load_2d(sQ, &tm_qkv, q_full, h * AT_DH, static_cast<int>(row0 + Q.start));
    }
    __syncwarp();
    // stage this CTA's schedule: the visited key tiles that hold at least one live key for this sample, in order
    const uint8_t* cls = a.kt_class + static_cast<long long>(b) * a.n_kt;
    int n_it = 0;
    for (int t0 = 0; t0 < Q.kt_cnt; t0 += 32) {
      const int t = t0 + lane;
      bool keep = false;
      AtSched e;
      if (t < Q.kt_cnt) {
        const mca_attn_ref ref = a.kt_list[Q.kt_off + t];
        const int c = cls[ref.tile];
        keep = c != 2;
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
    if (lane == 0) s_nit = n_it;
  }
  if (warp == AT_SOFT_WARPS + 1) tmem_alloc(tmem_holder, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const uint32_t tS = tmem_base, tP = tmem_base + 128, tO = tmem_base + 192;
  const int n_it = s_nit;

  if (warp == AT_SOFT_WARPS) {
    // ===================== TMA producer (whole warp loops, one elected lane issues) =====================
    for (int it = 0; it < n_it; ++it) {
      const int krow = static_cast<int>(row0 + sched[it].kstart);
      const int st = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      if (lane == 0) FTR(9, it, 0);
      mbar_wait(&k_empty[st], ph ^ 1);
      if (lane == 0) FTR(9, it, 1);
      if (elect_one()) {
        mbar_expect_tx(&k_full[st], AT_TILE_BYTES);
        tma_load_2d(sK + st * AT_TILE_BYTES, &tm_qkv, &k_full[st], a.H * AT_DH + h * AT_DH, krow);
      }
      __syncwarp();
      mbar_wait(&v_empty[st], ph ^ 1);
      if (lane == 0) FTR(9, it, 2);
      if (elect_one()) {
        mbar_expect_tx(&v_full[st], AT_TILE_BYTES);
        tma_load_2d(sV + st * AT_TILE_BYTES, &tm_qkv, &v_full[st], 2 * a.H * AT_DH + h * AT_DH, krow);
      }
      __syncwarp();
    }
  } else if (warp == AT_SOFT_WARPS + 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the loop and the waits (convergent code keeps the descriptors in uniform registers, so a
    // tcgen05.mma costs one uniform add); one elected lane issues the MMAs and commits.
    constexpr uint32_t idesc_s = make_idesc_bf16(AT_BM, AT_BN, false, false);
    constexpr uint32_t idesc_o = make_idesc_bf16(AT_BM, AT_DH, false, true);
    const uint64_t dq0 = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
    const uint64_t dk0 = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
    const uint64_t dv0 = make_smem_desc_sw128(smem_u32(sV), 8192, 1024);
    auto issue_pv = [&](int j) {
      const int st = j & 1;
      mbar_wait(&v_full[st], (j >> 1) & 1);
      if (lane == 0) FTR(8, j, 4);
      mbar_wait(p_full, j & 1);
      if (lane == 0) FTR(8, j, 5);
      tc_fence_after();
      const uint64_t dv = dv0 + static_cast<uint64_t>((st * AT_TILE_BYTES) >> 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < AT_BN / 16; ++k)
          umma_bf16_ts(tO, tP + k * 8, dv + k * (2048 >> 4), idesc_o, (j > 0 || k > 0) ? 1u : 0u);
        umma_commit(&v_empty[st]);
        umma_commit(pv_done);
      }
      __syncwarp();
      if (lane == 0) FTR(8, j, 6);
    };
    mbar_wait(q_full, 0);
    for (int it = 0; it < n_it; ++it) {
      const int st = it & 1;
      if (lane == 0) FTR(8, it, 0);
      mbar_wait(&k_full[st], (it >> 1) & 1);
      if (lane == 0) FTR(8, it, 1);
      mbar_wait(s_empty, (it & 1) ^ 1);
      if (lane == 0) FTR(8, it, 2);
      tc_fence_after();
      const uint64_t dk = dk0 + static_cast<uint64_t>((st * AT_TILE_BYTES) >> 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < AT_DH / 16; ++k) umma_bf16(tS, dq0 + k * 2, dk + k * 2, idesc_s, k > 0 ? 1u : 0u);
        umma_commit(&k_empty[st]);
        umma_commit(s_full);
      }
      __syncwarp();
      if (lane == 0) FTR(8, it, 3);
      if (it > 0) issue_pv(it - 1);
    }
    if (n_it > 0) issue_pv(n_it - 1);
  } else {
    // ===================== softmax / epilogue: warp = 16 query rows, a quad of lanes per row ======================
    // Fragment of tcgen05.ld.16x256b (the m16n8 accumulator layout): lane i holds rows i/4 and i/4 + 8 of the warp's
    // 16 TMEM lanes and columns 8j + 2(i%4) + {0,1}: row maxima / sums are two quad shuffles, no shared memory and no
    // barrier between warps.  Eight softmax warps per CTA (two CTAs per SM): four independent softmax warps per
    // scheduler; one warp per scheduler reaches ~8 of the SM's 16 exp2 per clock, four reach ~14 (ubench/mufu_rate.cu).
    const int wq = warp & 3;                 // TMEM lane quarter this warp may access
    const int lbase = wq * 32 + (warp >> 2) * 16;  // first of the warp's 16 lanes (= query rows of the tile)
    const int qd = lane & 3;                 // position inside the quad
    const int rA = lbase + (lane >> 2), rB = rA + 8;
    const int qiA = Q.start + rA, qiB = Q.start + rB;  // rows inside the sample (may run past the tile: never stored)
    const uint32_t rbA = a.rowbits[min(qiA, a.N - 1)], rbB = a.rowbits[min(qiB, a.N - 1)];
    const uint32_t lsel = static_cast<uint32_t>(lbase) << 16;
    float m2A = -CUDART_INF_F, m2B = -CUDART_INF_F;  // reference maxima (log2 units)
    float lA = 0.f, lB = 0.f;                        // partial row sums of this lane's columns
    for (int it = 0; it < n_it; ++it) {
      const AtSched e = sched[it];
      const bool masked = e.flags & 1;
      uint32_t awA[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
      uint32_t awB[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
      if (masked) {
        const int grp = (e.flags >> 8) & 255;
#pragma unroll
        for (int w = 0; w < 4; ++w) awA[w] = awB[w] = e.live[w];
        if (grp != 255) {
          if (((rbA >> grp) & 1u) == 0) awA[0] = awA[1] = awA[2] = awA[3] = 0;
          if (((rbB >> grp) & 1u) == 0) awB[0] = awB[1] = awB[2] = awB[3] = 0;
        } else {
          // mixed key groups (the fusion sub-blocks): per-key visibility; only this lane's columns are needed
#pragma unroll 1
          for (int j = 0; j < 16; ++j) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const int kk = 8 * j + 2 * qd + c;
              const uint32_t g = kk < e.klen ? a.keygrp[e.kstart + kk] : 31u;
              const uint32_t bit = 1u << (kk & 31);
              if (kk >= e.klen || !((rbA >> g) & 1u)) awA[kk >> 5] &= ~bit;
              if (kk >= e.klen || !((rbB >> g) & 1u)) awB[kk >> 5] &= ~bit;
            }
          }
        }
#pragma unroll
        for (int w = 0; w < 4; ++w) awA[w] >>= 2 * qd, awB[w] >>= 2 * qd;  // bit 8(j&3) + c = column 8j + 2 qd + c
      }
      if (lane == 0) FTR(warp, it, 0);
      mbar_wait(s_full, it & 1);
      if (lane == 0) FTR(warp, it, 1);
      tc_fence_after();
      uint32_t sv[2][32];  // sv[hh][4 jj + {0,1}] = row A, [4 jj + {2,3}] = row B, columns 64 hh + 8 jj + 2 qd + {0,1}
      tmem_ld16x256b_x8(tS + lsel, sv[0]);
      tmem_ld16x256b_x8(tS + lsel + 64, sv[1]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(s_empty);  // S is in registers: the next QK^T may overwrite it
      if (lane == 0) FTR(warp, it, 2);
      if (masked) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const int w = hh * 2 + (jj >> 2), sh = 8 * (jj & 3);
            if (!((awA[w] >> sh) & 1u)) sv[hh][4 * jj + 0] = __float_as_uint(-CUDART_INF_F);
            if (!((awA[w] >> (sh + 1)) & 1u)) sv[hh][4 * jj + 1] = __float_as_uint(-CUDART_INF_F);
            if (!((awB[w] >> sh) & 1u)) sv[hh][4 * jj + 2] = __float_as_uint(-CUDART_INF_F);
            if (!((awB[w] >> (sh + 1)) & 1u)) sv[hh][4 * jj + 3] = __float_as_uint(-CUDART_INF_F);
          }
      }
      float mxA = -CUDART_INF_F, mxB = -CUDART_INF_F;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          mxA = fmaxf(mxA, fmaxf(__uint_as_float(sv[hh][4 * jj]), __uint_as_float(sv[hh][4 * jj + 1])));
          mxB = fmaxf(mxB, fmaxf(__uint_as_float(sv[hh][4 * jj + 2]), __uint_as_float(sv[hh][4 * jj + 3])));
        }
      mxA = fmaxf(mxA, __shfl_xor_sync(0xffffffffu, mxA, 1));
      mxB = fmaxf(mxB, __shfl_xor_sync(0xffffffffu, mxB, 1));
      mxA = fmaxf(mxA, __shfl_xor_sync(0xffffffffu, mxA, 2));
      mxB = fmaxf(mxB, __shfl_xor_sync(0xffffffffu, mxB, 2));
      const float tA = mxA * LOG2E, tB = mxB * LOG2E;
      const bool growA = tA > m2A + AT_RESCALE_THRESHOLD, growB = tB > m2B + AT_RESCALE_THRESHOLD;
      const float nA = growA ? tA : m2A, nB = growB ? tB : m2B;
      const float alphaA = growA ? fast_ex2(m2A - nA) : 1.0f;  // m2 = -inf -> 0
      const float alphaB = growB ? fast_ex2(m2B - nB) : 1.0f;
      if (lane == 0) FTR(warp, it, 3);
      lA *= alphaA, lB *= alphaB;
      m2A = nA, m2B = nB;
      const float offA = (m2A == -CUDART_INF_F) ? 0.f : m2A, offB = (m2B == -CUDART_INF_F) ? 0.f : m2B;
      const uint64_t log2e2 = f2_pack(LOG2E, LOG2E), noffA = f2_pack(-offA, -offA), noffB = f2_pack(-offB, -offB);
      uint64_t sumA = f2_pack(0.f, 0.f), sumB = f2_pack(0.f, 0.f);
      uint32_t pk[32];  // 16x128b fragment of the bf16-packed P tile: pk[2 j] = row A, pk[2 j + 1] = row B, column 4 j + qd
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          float a0, a1, b0, b1;
          f2_unpack(f2_fma(f2_pack(__uint_as_float(sv[hh][4 * jj]), __uint_as_float(sv[hh][4 * jj + 1])), log2e2, noffA), a0, a1);
          f2_unpack(f2_fma(f2_pack(__uint_as_float(sv[hh][4 * jj + 2]), __uint_as_float(sv[hh][4 * jj + 3])), log2e2, noffB), b0, b1);
          const float pa0 = fast_ex2(a0), pa1 = fast_ex2(a1), pb0 = fast_ex2(b0), pb1 = fast_ex2(b1);
          sumA = f2_add(sumA, f2_pack(pa0, pa1));
          sumB = f2_add(sumB, f2_pack(pb0, pb1));
          pk[2 * (hh * 8 + jj)] = pack_bf16x2(pa0, pa1);
          pk[2 * (hh * 8 + jj) + 1] = pack_bf16x2(pb0, pb1);
        }
      {
        float s0, s1;
        f2_unpack(sumA, s0, s1);
        lA += s0 + s1;
        f2_unpack(sumB, s0, s1);
        lB += s0 + s1;
      }
      if (lane == 0) FTR(warp, it, 4);
      if (it > 0) {
        mbar_wait(pv_done, (it - 1) & 1);  // P of the previous tile has been consumed, O is quiescent
        tc_fence_after();
        if (__any_sync(0xffffffffu, growA || growB)) {
          uint32_t ov[32];
          tmem_ld16x256b_x8(tO + lsel, ov);
          tmem_ld_wait();
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            ov[4 * jj] = __float_as_uint(__uint_as_float(ov[4 * jj]) * alphaA);
            ov[4 * jj + 1] = __float_as_uint(__uint_as_float(ov[4 * jj + 1]) * alphaA);
            ov[4 * jj + 2] = __float_as_uint(__uint_as_float(ov[4 * jj + 2]) * alphaB);
            ov[4 * jj + 3] = __float_as_uint(__uint_as_float(ov[4 * jj + 3]) * alphaB);
          }
          tmem_st16x256b_x8(tO + lsel, ov);
        }
      }
      if (lane == 0) FTR(warp, it, 5);
      tmem_st16x128b_x16(tP + lsel, pk);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full);
      if (lane == 0) FTR(warp, it, 6);
    }
    // ---- epilogue: normalise and store the warp's 16 rows (each lane: 2 rows x 16 columns)
    uint32_t ov[32];
    if (n_it > 0) {  // uniform across the CTA
      mbar_wait(pv_done, (n_it - 1) & 1);
      tc_fence_after();
      tmem_ld16x256b_x8(tO + lsel, ov);
      tmem_ld_wait();
    }
    lA += __shfl_xor_sync(0xffffffffu, lA, 1);
    lB += __shfl_xor_sync(0xffffffffu, lB, 1);
    lA += __shfl_xor_sync(0xffffffffu, lA, 2);
    lB += __shfl_xor_sync(0xffffffffu, lB, 2);
    const float* vm = a.vmean + static_cast<long long>(b) * a.H * AT_DH + h * AT_DH;
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const float l_tot = rr ? lB : lA;
      const float m2 = rr ? m2B : m2A;
      const int r = rr ? rB : rA, qi = rr ? qiB : qiA;
      float lse = CUDART_INF_F;
      float inv = 0.f;
      if (l_tot != 0.f) {
        inv = 1.0f / l_tot;
        lse = (m2 + log2f(l_tot)) * 0.6931471805599453f;
      }
      if (r < Q.len) {
        __nv_bfloat16* orow = a.out + (row0 + qi) * (a.H * AT_DH) + h * AT_DH;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int c = 8 * jj + 2 * qd;
          float o0, o1;
          if (l_tot != 0.f) {
            o0 = __uint_as_float(ov[4 * jj + 2 * rr]) * inv, o1 = __uint_as_float(ov[4 * jj + 2 * rr + 1]) * inv;
          } else {  // reference quirk Q4: a row with no live allowed key is uniform over all N keys
            o0 = vm[c], o1 = vm[c + 1];
          }
          *reinterpret_cast<uint32_t*>(orow + c) = pack_bf16x2(o0, o1);
        }
        if (qd == 0) a.lse[(static_cast<long long>(b) * a.H + h) * a.N + qi] = lse;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
#ifdef MCA_TRACE
  if (threadIdx.x == 0 && cta_lin < 4096) g_fcta[cta_lin * 4 + 1] = gtimer();
  if (threadIdx.x == 0 && blockIdx.x == 5 && blockIdx.y == 3) { g_ftrace[10 * 24 * 8] = g_fcta[cta_lin * 4 + 3]; g_ftrace[10 * 24 * 8 + 1] = clock64(); }
#endif
  if (warp == AT_SOFT_WARPS + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// vmean[b, c] = mean over all N rows of V[b, :, c]; only needed when some modality is absent (device flag).
// Block (x, b) sums 64 rows of sample b (thread = two adjacent bf16 columns, coalesced 1 KB per row) and adds its partial
// to vmean (zeroed by the launcher): 40 x B blocks instead of one long-running block per (sample, 128 columns).
constexpr int VM_ROWS = 64;
__global__ void __launch_bounds__(256)
vmean_kernel(const __nv_bfloat16* __restrict__ qkv, int ld, int v_col0, int width, int N, const int* __restrict__ any_absent,
             float* __restrict__ vmean) {
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
                            const int* any_absent, float* vmean, void* out, float* lse, int B, int N, int H,
                            void* stream_) {
  if (B <= 0 || N <= 0 || H <= 0 || n_qt <= 0 || n_kt > AT_MAX_KT) return MCA_ERR_SHAPE;
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
  AttnFwdArgs a{q_tiles, kt_list, k_tiles, rowbits, keygrp, tile_grp, kt_class, kt_live, vmean,
                reinterpret_cast<__nv_bfloat16*>(out), lse, N, H, n_kt};
  dim3 grid(B * H, n_qt);  // x fastest: every (sample, head) of the heaviest query tile is dispatched first
  attn_fwd_kernel<<<grid, AT_THREADS, AT_SMEM, stream>>>(tm, a);
  return check_launch();
}

#ifdef MCA_TRACE
extern "C" int mca_debug_read_trace_fwd(long long* tr, int n_tr, long long* cta, int n_cta) {
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(tr, mca::g_ftrace, sizeof(long long) * n_tr) != cudaSuccess) return 3;
  return cudaMemcpyFromSymbol(cta, mca::g_fcta, sizeof(long long) * n_cta) == cudaSuccess ? 0 : 3;
}
#endif
