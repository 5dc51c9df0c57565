// Block-sparse masked multi-head attention forward on tcgen05 (replaces model.py:85-100: q*scale, QK^T, static
// attn_mask fill, key_padding_mask fill, softmax, PV — without ever materialising the [B,h,N,N] score tensor).
//
// One CTA = one (sample, head, 128-row query tile).  Warps 0-3: softmax, one thread per query row (TMEM lane);
// warp 4: TMA producer; warp 5: tcgen05.mma issuer.  Per visited key tile:
//     S = Q K^T  (128x128x64, smem x smem -> TMEM)    ->  row max / exp2 / row sum in registers
//     P (bf16) -> 128B-swizzled smem                  ->  O_j = P V (128x64x128, V consumed MN-major as loaded)
//     o_acc = o_acc * alpha + O_j                      (running rescale kept in registers; TMEM tile is fresh)
// Key tiles come from a static, host-built schedule (only tiles holding at least one statically allowed
// (query,key) pair — 43 % of all pairs for MCA at CMU shape) and are skipped at run time when every key in them is
// padded for this sample.  Partially allowed / partially padded / ragged tiles evaluate a per-key bitmask:
// allowed(i,j) = rowbits[i] & keybit[j], keybit = 1 << group(j) or 0 when padded.
// Reference quirk: masks are filled with -finfo.max, so a query row with no live allowed key is
// exactly uniform over ALL N keys; those rows get the per-(sample,head) mean of V and LSE = +inf (which makes the
// backward treat their P as 0; their 1/N contribution to dV is added separately).
// Two CTAs fit per SM (80 KB smem, 256 TMEM columns each) so one CTA's tensor work hides the other's softmax.
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
constexpr int AT_P_BYTES = AT_BM * AT_BN * 2;     // 32 KB
constexpr int AT_SMEM = 3 * AT_TILE_BYTES + AT_P_BYTES + 1024 + 1024;
constexpr float LOG2E = 1.4426950408889634f;

struct AttnFwdArgs {
  const mca_attn_qtile* q_tiles;
  const mca_attn_ref* kt_list;
  const mca_attn_tile* k_tiles;
  const uint32_t* rowbits;   // [N]
  const uint8_t* keygrp;     // [N]
  const uint8_t* padding;    // [B, N]
  const uint8_t* kt_class;   // [B, n_kt]
  const float* vmean;        // [B, H*64]
  __nv_bfloat16* out;        // [B*N, H*64]
  float* lse;                // [B, H, N]
  int N, H, n_kt;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(AT_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const AttnFwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = base;
  uint8_t* sK = base + AT_TILE_BYTES;
  uint8_t* sV = base + 2 * AT_TILE_BYTES;
  uint8_t* sP = base + 3 * AT_TILE_BYTES;
  uint32_t* keybit = reinterpret_cast<uint32_t*>(sP + AT_P_BYTES);  // [128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(keybit + AT_BN);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = bars + 2;
  uint64_t* v_full = bars + 3;
  uint64_t* v_empty = bars + 4;
  uint64_t* s_full = bars + 5;
  uint64_t* s_empty = bars + 6;
  uint64_t* p_full = bars + 7;
  uint64_t* o_full = bars + 8;
  uint64_t* o_empty = bars + 9;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const mca_attn_qtile Q = a.q_tiles[qt];
  const long long row0 = static_cast<long long>(b) * a.N;
  const uint8_t* cls = a.kt_class + static_cast<long long>(b) * a.n_kt;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tm_qkv);
    mbar_init(q_full, 1);
    mbar_init(k_full, 1);
    mbar_init(k_empty, 1);
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(s_empty, 128);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    mbar_init(o_empty, 128);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc(tmem_holder, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const uint32_t tS = tmem_base, tO = tmem_base + 128;

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(q_full, AT_TILE_BYTES);
      tma_load_2d(sQ, &tm_qkv, q_full, h * AT_DH, static_cast<int>(row0 + Q.start));
      uint32_t ph = 0;
      for (int t = 0; t < Q.kt_cnt; ++t) {
        const mca_attn_ref ref = a.kt_list[Q.kt_off + t];
        if (cls[ref.tile] == 2) continue;
        const int krow = static_cast<int>(row0 + a.k_tiles[ref.tile].start);
        mbar_wait(k_empty, ph ^ 1);
        mbar_expect_tx(k_full, AT_TILE_BYTES);
        tma_load_2d(sK, &tm_qkv, k_full, a.H * AT_DH + h * AT_DH, krow);
        mbar_wait(v_empty, ph ^ 1);
        mbar_expect_tx(v_full, AT_TILE_BYTES);
        tma_load_2d(sV, &tm_qkv, v_full, 2 * a.H * AT_DH + h * AT_DH, krow);
        ph ^= 1;
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(AT_BM, AT_BN, false, false);
      constexpr uint32_t idesc_o = make_idesc_bf16(AT_BM, AT_DH, false, true);
      const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV), p_addr = smem_u32(sP);
      mbar_wait(q_full, 0);
      uint32_t ph = 0;
      for (int t = 0; t < Q.kt_cnt; ++t) {
        const mca_attn_ref ref = a.kt_list[Q.kt_off + t];
        if (cls[ref.tile] == 2) continue;
        mbar_wait(k_full, ph);
        mbar_wait(s_empty, ph ^ 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < AT_DH / 16; ++k)
          umma_bf16(tS, make_smem_desc_sw128(q_addr + k * 32, 16, 1024), make_smem_desc_sw128(k_addr + k * 32, 16, 1024),
                    idesc_s, k > 0 ? 1u : 0u);
        umma_commit(k_empty);
        umma_commit(s_full);
        mbar_wait(v_full, ph);
        mbar_wait(p_full, ph);
        mbar_wait(o_empty, ph ^ 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < AT_BN / 16; ++k)
          umma_bf16(tO, make_smem_desc_sw128(p_addr + (k >> 2) * (AT_P_BYTES / 2) + (k & 3) * 32, 16, 1024),
                    make_smem_desc_sw128(v_addr + k * 2048, 8192, 1024), idesc_o, k > 0 ? 1u : 0u);
        umma_commit(v_empty);
        umma_commit(o_full);
        ph ^= 1;
      }
    }
  } else {
    // ===================== softmax / accumulate / epilogue: thread = query row =====================
    const int r = warp * 32 + lane;
    const int qi = Q.start + r;  // row inside the sample (may run past the tile's valid rows: never stored)
    const bool row_valid = r < Q.len;
    const uint32_t rb = a.rowbits[min(qi, a.N - 1)];
    const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
    float m_run = -CUDART_INF_F, l_run = 0.f;
    float o_acc[AT_DH];
#pragma unroll
    for (int i = 0; i < AT_DH; ++i) o_acc[i] = 0.f;
    uint32_t ph = 0;
    for (int t = 0; t < Q.kt_cnt; ++t) {
      const mca_attn_ref ref = a.kt_list[Q.kt_off + t];
      const int c = cls[ref.tile];
      if (c == 2) continue;
      const mca_attn_tile K = a.k_tiles[ref.tile];
      const bool masked = (ref.flags & 1) || c == 1 || K.len < AT_BN;
      if (masked) {
        named_bar_sync(1, 128);
        const int kj = K.start + r;
        uint32_t bit = 0;
        if (r < K.len && a.padding[row0 + kj] == 0) bit = 1u << a.keygrp[kj];
        keybit[r] = bit;
        named_bar_sync(1, 128);
      }
      mbar_wait(s_full, ph);
      tc_fence_after();
      // ---- pass 1: row max over allowed keys
      float mx = m_run;
#pragma unroll 1
      for (int cc = 0; cc < AT_BN / 32; ++cc) {
        uint32_t v[32];
        tmem_ld32(tS + lane_sel + cc * 32, v);
        tmem_ld_wait();
        if (masked) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (rb & keybit[cc * 32 + i]) mx = fmaxf(mx, __uint_as_float(v[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
      }
      const float alpha = (m_run == -CUDART_INF_F) ? 0.f : exp2f((m_run - mx) * LOG2E);
      const float moff = (mx == -CUDART_INF_F) ? 0.f : mx * LOG2E;
      // ---- pass 2: p = exp2(s*log2e - m*log2e), write bf16 P into the swizzled K-major tile
      float rowsum = 0.f;
      uint8_t* prow = sP + r * 128;
#pragma unroll 1
      for (int cc = 0; cc < AT_BN / 32; ++cc) {
        uint32_t v[32];
        tmem_ld32(tS + lane_sel + cc * 32, v);
        tmem_ld_wait();
        float p[32];
        if (masked) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            p[i] = (rb & keybit[cc * 32 + i]) ? exp2f(__uint_as_float(v[i]) * LOG2E - moff) : 0.f;
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) p[i] = exp2f(__uint_as_float(v[i]) * LOG2E - moff);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) rowsum += p[i];
        // 32 keys = 64 B = 4 chunks of 16 B; chunk index within the 128 B row = (cc&1)*4 + q, half = cc>>1
        uint8_t* half = prow + (cc >> 1) * (AT_P_BYTES / 2);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 w;
          w.x = pack_bf16x2(p[8 * q + 0], p[8 * q + 1]);
          w.y = pack_bf16x2(p[8 * q + 2], p[8 * q + 3]);
          w.z = pack_bf16x2(p[8 * q + 4], p[8 * q + 5]);
          w.w = pack_bf16x2(p[8 * q + 6], p[8 * q + 7]);
          const int chunk = ((cc & 1) * 4 + q) ^ (r & 7);
          *reinterpret_cast<uint4*>(half + chunk * 16) = w;
        }
      }
      tc_fence_before();
      mbar_arrive(s_empty);
      fence_proxy_async_smem();
      mbar_arrive(p_full);
      l_run = l_run * alpha + rowsum;
      m_run = mx;
      // ---- O_j from TMEM, running rescale in registers
      mbar_wait(o_full, ph);
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < AT_DH / 32; ++cc) {
        uint32_t v[32];
        tmem_ld32(tO + lane_sel + cc * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o_acc[cc * 32 + i] = o_acc[cc * 32 + i] * alpha + __uint_as_float(v[i]);
      }
      tc_fence_before();
      mbar_arrive(o_empty);
      ph ^= 1;
    }
    if (row_valid) {
      __nv_bfloat16* orow = a.out + (row0 + qi) * (a.H * AT_DH) + h * AT_DH;
      float lse;
      if (l_run == 0.f) {
        const float* vm = a.vmean + static_cast<long long>(b) * a.H * AT_DH + h * AT_DH;
#pragma unroll
        for (int i = 0; i < AT_DH; ++i) o_acc[i] = vm[i];
        lse = CUDART_INF_F;
      } else {
        const float inv = 1.0f / l_run;
#pragma unroll
        for (int i = 0; i < AT_DH; ++i) o_acc[i] *= inv;
        lse = m_run + logf(l_run);
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        uint4 w;
        w.x = pack_bf16x2(o_acc[8 * q + 0], o_acc[8 * q + 1]);
        w.y = pack_bf16x2(o_acc[8 * q + 2], o_acc[8 * q + 3]);
        w.z = pack_bf16x2(o_acc[8 * q + 4], o_acc[8 * q + 5]);
        w.w = pack_bf16x2(o_acc[8 * q + 6], o_acc[8 * q + 7]);
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

// vmean[b, c] = mean over all N rows of V[b, :, c]; only needed when some modality is absent (device flag)
__global__ void __launch_bounds__(256)
vmean_kernel(const __nv_bfloat16* __restrict__ qkv, int ld, int v_col0, int width, int N, const int* __restrict__ any_absent,
             float* __restrict__ vmean) {
  if (*any_absent == 0) return;
  __shared__ float part[2][128];
  const int b = blockIdx.x;
  const int c = blockIdx.y * 128 + (threadIdx.x & 127), g = threadIdx.x >> 7;
  float acc = 0.f;
  if (c < width)
    for (int n = g; n < N; n += 2) acc += __bfloat162float(qkv[(static_cast<long long>(b) * N + n) * ld + v_col0 + c]);
  part[g][threadIdx.x & 127] = acc;
  __syncthreads();
  if (g == 0 && c < width) vmean[static_cast<long long>(b) * width + c] = (part[0][threadIdx.x] + part[1][threadIdx.x]) / N;
}

}  // namespace mca

using namespace mca;

extern "C" int mca_attn_fwd(const void* qkv, const mca_attn_qtile* q_tiles, int n_qt, const mca_attn_ref* kt_list,
                            const mca_attn_tile* k_tiles, int n_kt, const uint32_t* rowbits, const uint8_t* keygrp,
                            const uint8_t* padding, const uint8_t* kt_class, const int* any_absent, float* vmean,
                            void* out, float* lse, int B, int N, int H, void* stream_) {
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
  dim3 gv(B, (H * AT_DH + 127) / 128);
  vmean_kernel<<<gv, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(qkv), ld, 2 * H * AT_DH, H * AT_DH, N,
                                       any_absent, vmean);
  AttnFwdArgs a{q_tiles, kt_list, k_tiles, rowbits, keygrp, padding, kt_class, vmean,
                reinterpret_cast<__nv_bfloat16*>(out), lse, N, H, n_kt};
  dim3 grid(n_qt, H, B);
  attn_fwd_kernel<<<grid, AT_THREADS, AT_SMEM, stream>>>(tm, a);
  return check_launch();
}
