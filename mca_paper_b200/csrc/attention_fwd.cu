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
constexpr int AT_THREADS = 192;
constexpr int AT_TILE_BYTES = AT_BM * AT_DH * 2;  // 16 KB
constexpr int AT_SMEM = 5 * AT_TILE_BYTES + 1024 + 256;  // Q, 2 x K, 2 x V
constexpr float LOG2E = 1.4426950408889634f;
constexpr float AT_RESCALE_THRESHOLD = 8.0f;  // log2 units

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

__global__ void __launch_bounds__(AT_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const AttnFwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = base;
  uint8_t* sK = base + AT_TILE_BYTES;      // 2 stages
  uint8_t* sV = base + 3 * AT_TILE_BYTES;  // 2 stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + 5 * AT_TILE_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;   // [2]
  uint64_t* k_empty = bars + 3;  // [2]
  uint64_t* v_full = bars + 5;   // [2]
  uint64_t* v_empty = bars + 7;  // [2]
  uint64_t* s_full = bars + 9;
  uint64_t* s_empty = bars + 10;
  uint64_t* p_full = bars + 11;
  uint64_t* pv_done = bars + 12;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x % a.H, b = blockIdx.x / a.H;
  const mca_attn_qtile Q = a.q_tiles[blockIdx.y];
  const long long row0 = static_cast<long long>(b) * a.N;
  const uint8_t* cls = a.kt_class + static_cast<long long>(b) * a.n_kt;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tm_qkv);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 128);
    mbar_init(p_full, 128);
    mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc(tmem_holder, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const uint32_t tS = tmem_base, tP = tmem_base + 128, tO = tmem_base + 192;

  if (warp == 4) {
    // ===================== TMA producer (whole warp loops, one elected lane issues) =====================
    if (elect_one()) {
      mbar_expect_tx(q_full, AT_TILE_BYTES);
      tma_load_2d(sQ, &tm_qkv, q_full, h * AT_DH, static_cast<int>(row0 + Q.start));
    }
    __syncwarp();
    int it = 0;
    for (int t = 0; t < Q.kt_cnt; ++t) {
      const mca_attn_ref ref = a.kt_list[Q.kt_off + t];
      if (cls[ref.tile] == 2) continue;
      const int krow = static_cast<int>(row0 + a.k_tiles[ref.tile].start);
      const int st = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      mbar_wait(&k_empty[st], ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&k_full[st], AT_TILE_BYTES);
        tma_load_2d(sK + st * AT_TILE_BYTES, &tm_qkv, &k_full[st], a.H * AT_DH + h * AT_DH, krow);
      }
      __syncwarp();
      mbar_wait(&v_empty[st], ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&v_full[st], AT_TILE_BYTES);
        tma_load_2d(sV + st * AT_TILE_BYTES, &tm_qkv, &v_full[st], 2 * a.H * AT_DH + h * AT_DH, krow);
      }
      __syncwarp();
      ++it;
    }
  } else if (warp == 5) {
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
      mbar_wait(p_full, j & 1);
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
    };
    mbar_wait(q_full, 0);
    int it = 0;
    for (int t = 0; t < Q.kt_cnt; ++t) {
      const mca_attn_ref ref = a.kt_list[Q.kt_off + t];
      if (cls[ref.tile] == 2) continue;
      const int st = it & 1;
      mbar_wait(&k_full[st], (it >> 1) & 1);
      mbar_wait(s_empty, (it & 1) ^ 1);
      tc_fence_after();
      const uint64_t dk = dk0 + static_cast<uint64_t>((st * AT_TILE_BYTES) >> 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < AT_DH / 16; ++k) umma_bf16(tS, dq0 + k * 2, dk + k * 2, idesc_s, k > 0 ? 1u : 0u);
        umma_commit(&k_empty[st]);
        umma_commit(s_full);
      }
      __syncwarp();
      if (it > 0) issue_pv(it - 1);
      ++it;
    }
    if (it > 0) issue_pv(it - 1);
  } else {
    // ===================== softmax / epilogue: thread = query row =====================
    const int r = warp * 32 + lane;
    const int qi = Q.start + r;  // row inside the sample (may run past the tile's valid rows: never stored)
    const bool row_valid = r < Q.len;
    const uint32_t rb = a.rowbits[min(qi, a.N - 1)];
    const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
    float m2 = -CUDART_INF_F, l_run = 0.f;  // reference maximum (log2 units) and running sum
    int it = 0;
    for (int t = 0; t < Q.kt_cnt; ++t) {
      const mca_attn_ref ref = a.kt_list[Q.kt_off + t];
      const int c = cls[ref.tile];
      if (c == 2) continue;
      const mca_attn_tile K = a.k_tiles[ref.tile];
      const int grp = a.tile_grp[ref.tile];
      const bool masked = (ref.flags & 1) || c == 1 || K.len < AT_BN || grp == 255;
      uint32_t aw[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
      if (masked) {
        const uint4 lw = *reinterpret_cast<const uint4*>(a.kt_live + (static_cast<long long>(b) * a.n_kt + ref.tile) * 4);
        aw[0] = lw.x, aw[1] = lw.y, aw[2] = lw.z, aw[3] = lw.w;
        if (grp != 255) {
          if (((rb >> grp) & 1u) == 0) aw[0] = aw[1] = aw[2] = aw[3] = 0;
        } else {
#pragma unroll 1
          for (int w = 0; w < 4; ++w) {
            uint32_t bits = 0;
            for (int j = 0; j < 32; ++j) {
              const int kj = K.start + w * 32 + j;
              if (w * 32 + j < K.len && ((rb >> a.keygrp[kj]) & 1u)) bits |= 1u << j;
            }
            aw[w] &= bits;
          }
        }
      }
      mbar_wait(s_full, it & 1);
      tc_fence_after();
      uint32_t sv[4][32];
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) tmem_ld32(tS + lane_sel + cc * 32, sv[cc]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(s_empty);  // S is in registers: the next QK^T may overwrite it
      if (masked) {
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
      if (it > 0) {
        mbar_wait(pv_done, (it - 1) & 1);  // P of the previous tile has been consumed, O is quiescent
        tc_fence_after();
        if (__any_sync(0xffffffffu, grow)) {
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
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full);
      ++it;
    }
    uint32_t ov[2][32];
    float lse = CUDART_INF_F;
    if (it > 0) {  // uniform across the CTA
      mbar_wait(pv_done, (it - 1) & 1);
      tc_fence_after();
      tmem_ld32(tO + lane_sel, ov[0]);
      tmem_ld32(tO + lane_sel + 32, ov[1]);
      tmem_ld_wait();
    }
    if (l_run != 0.f) {
      const float inv = 1.0f / l_run;
#pragma unroll
      for (int i = 0; i < AT_DH; ++i) ov[i >> 5][i & 31] = __float_as_uint(__uint_as_float(ov[i >> 5][i & 31]) * inv);
      lse = (m2 + log2f(l_run)) * 0.6931471805599453f;
    } else {
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
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
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
  if (B <= 0 || N <= 0 || H <= 0 || n_qt <= 0) return MCA_ERR_SHAPE;
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
