// Persistent warp-specialised bf16 GEMM for sm_100a: TMA -> 128B-swizzled smem ring -> tcgen05.mma (fp32
// accumulators in TMEM, double buffered) -> fused epilogues read back with tcgen05.ld.
//
//   out[M,N] = A[M,K] * B[N,K]^T           (both operands may independently be K-major or MN-major in memory)
//
// This one kernel serves every dense contraction of the MCA training step (reference: model.py:83,105
// to_q/to_kv/to_out, model.py:49-51 GEGLU feed-forward, encoders.py:190 token projection, and their autograd
// transposes):  forward uses (A K-major, B K-major), dX uses (A K-major, B MN-major = the weight as stored),
// dW uses (A MN-major, B MN-major) with split-K over the token dimension.
#include "mca_b200.h"
#include "ptx.cuh"
#include "runtime.h"

namespace mca {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle-128B row
constexpr int GEMM_THREADS = 256;

struct GemmParams {
  int M, N, K;
  int k_splits;
  int mode;
  void* out0;
  long long ld0;
  void* out1;
  long long ld1;
  const void* aux0;
  long long ldaux;
  const float* bias;
  float alpha;
};

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int SMEM = STAGES * (A_BYTES + B_BYTES) + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = 2 * BN;
};

// ---- epilogue: one thread = one output row, processes 32 consecutive columns held in r[]
__device__ __forceinline__ void store_bf16_32(__nv_bfloat16* dst, const float (&v)[32]) {
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 q;
    q.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
    q.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
    q.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
    q.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
    d4[i] = q;
  }
}
__device__ __forceinline__ void store_f32_32(float* dst, const float (&v)[32]) {
  float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
  for (int i = 0; i < 8; ++i) d4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
__device__ __forceinline__ void load_f32_32(const float* src, float (&v)[32]) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float4 q = s4[i];
    v[4 * i] = q.x, v[4 * i + 1] = q.y, v[4 * i + 2] = q.z, v[4 * i + 3] = q.w;
  }
}
__device__ __forceinline__ void load_bf16_32(const __nv_bfloat16* src, float (&v)[32]) {
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 q = s4[i];
    v[8 * i + 0] = bf16_lo(q.x), v[8 * i + 1] = bf16_hi(q.x);
    v[8 * i + 2] = bf16_lo(q.y), v[8 * i + 3] = bf16_hi(q.y);
    v[8 * i + 4] = bf16_lo(q.z), v[8 * i + 5] = bf16_hi(q.z);
    v[8 * i + 6] = bf16_lo(q.w), v[8 * i + 7] = bf16_hi(q.w);
  }
}

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = base;
  uint8_t* sB = base + STAGES * Cfg::A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + STAGES * Cfg::B_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_m = (p.M + BM - 1) / BM;
  const int tiles_n = (p.N + BN - 1) / BN;
  const int kb_total = (p.K + BK - 1) / BK;
  const int kb_per_split = (kb_total + p.k_splits - 1) / p.k_splits;
  const int num_tiles = tiles_m * tiles_n * p.k_splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 128);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_holder, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int nt = tile % tiles_n;
        const int mt = (tile / tiles_n) % tiles_m;
        const int z = tile / (tiles_n * tiles_m);
        const int m0 = mt * BM, n0 = nt * BN;
        const int kb0 = z * kb_per_split;
        const int kb1 = min(kb_total, kb0 + kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_expect_tx(&full_bar[s], Cfg::A_BYTES + Cfg::B_BYTES);
          uint8_t* a_dst = sA + s * Cfg::A_BYTES;
          uint8_t* b_dst = sB + s * Cfg::B_BYTES;
          if constexpr (!A_MN) {
            tma_load_2d(a_dst, &tmA, &full_bar[s], kb * BK, m0);
          } else {
#pragma unroll
            for (int h = 0; h < BM / 64; ++h) tma_load_2d(a_dst + h * 8192, &tmA, &full_bar[s], m0 + 64 * h, kb * BK);
          }
          if constexpr (!B_MN) {
            tma_load_2d(b_dst, &tmB, &full_bar[s], kb * BK, n0);
          } else {
#pragma unroll
            for (int h = 0; h < BN / 64; ++h) tma_load_2d(b_dst + h * 8192, &tmB, &full_bar[s], n0 + 64 * h, kb * BK);
          }
          if (++s == STAGES) s = 0, ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (single thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN, B_MN);
      int s = 0, as = 0;
      uint32_t ph = 0, aph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int z = tile / (tiles_n * tiles_m);
        const int kb0 = z * kb_per_split;
        const int kb1 = min(kb_total, kb0 + kb_per_split);
        mbar_wait(&tempty_bar[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + s * Cfg::A_BYTES);
          const uint32_t b_addr = smem_u32(sB + s * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major: 16 bf16 = 32 B along the swizzled row; rows of 8 are 1024 B apart (SBO).
            // MN-major: 16 k-rows = 2 KB; the next 64 MN elements live one TMA box (8 KB) further (LBO).
            const uint64_t da = A_MN ? make_smem_desc_sw128(a_addr + k * 2048, 8192, 1024)
                                     : make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t db = B_MN ? make_smem_desc_sw128(b_addr + k * 2048, 8192, 1024)
                                     : make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            umma_bf16(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs retire
          if (++s == STAGES) s = 0, ph ^= 1;
        }
        umma_commit(&tfull_bar[as]);  // accumulator complete -> epilogue
        if (++as == 2) as = 0, aph ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps: TMEM -> registers -> global =====================
    const int ew = warp & 3;  // TMEM sub-partition this warp may read
    int as = 0;
    uint32_t aph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int nt = tile % tiles_n;
      const int mt = (tile / tiles_n) % tiles_m;
      const int z = tile / (tiles_n * tiles_m);
      const int n0 = nt * BN;
      const long long row = static_cast<long long>(mt) * BM + ew * 32 + lane;
      const bool row_ok = row < p.M;
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + as * BN;

      if (p.mode == MCA_EPI_GEGLU) {
        // columns of each 128-wide block are [64 value | 64 gate] (W1 rows interleaved on the host side)
#pragma unroll 1
        for (int blk = 0; blk < BN / 128; ++blk) {
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t rv[32], rg[32];
            tmem_ld32(t_row + blk * 128 + c * 32, rv);
            tmem_ld32(t_row + blk * 128 + 64 + c * 32, rg);
            tmem_ld_wait();
            const int ncol = n0 + blk * 128 + c * 32;  // column of the value chunk in u
            if (row_ok && ncol < p.N) {
              float xv[32], gv[32], hv[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                // round to bf16 first so forward h and the backward recompute see identical u
                xv[i] = __bfloat162float(__float2bfloat16(__uint_as_float(rv[i])));
                gv[i] = __bfloat162float(__float2bfloat16(__uint_as_float(rg[i])));
                hv[i] = gelu_exact(gv[i]) * xv[i];
              }
              __nv_bfloat16* u = reinterpret_cast<__nv_bfloat16*>(p.out1) + row * p.ld1;
              store_bf16_32(u + ncol, xv);
              store_bf16_32(u + ncol + 64, gv);
              __nv_bfloat16* h = reinterpret_cast<__nv_bfloat16*>(p.out0) + row * p.ld0;
              store_bf16_32(h + (ncol / 128) * 64 + c * 32, hv);
            }
          }
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(t_row + c * 32, r);
          tmem_ld_wait();
          const int ncol = n0 + c * 32;
          if (!row_ok || ncol >= p.N) continue;
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * p.alpha;
          if (p.bias != nullptr) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += __ldg(p.bias + ncol + i);
          }
          if (p.mode == MCA_EPI_BF16) {
            store_bf16_32(reinterpret_cast<__nv_bfloat16*>(p.out0) + row * p.ld0 + ncol, v);
          } else if (p.mode == MCA_EPI_F32) {
            float* o = reinterpret_cast<float*>(p.out0) + static_cast<long long>(z) * p.M * p.ld0;
            store_f32_32(o + row * p.ld0 + ncol, v);
          } else if (p.mode == MCA_EPI_RESID) {
            float a[32];
            load_f32_32(reinterpret_cast<const float*>(p.aux0) + row * p.ldaux + ncol, a);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += a[i];
            store_f32_32(reinterpret_cast<float*>(p.out0) + row * p.ld0 + ncol, v);
            if (p.out1 != nullptr) store_bf16_32(reinterpret_cast<__nv_bfloat16*>(p.out1) + row * p.ld1 + ncol, v);
          } else if (p.mode == MCA_EPI_GEGLU_BWD) {
            // v = dL/dh for h columns [ncol, ncol+32); u holds (value, gate) in the interleaved layout
            const int blk = ncol / 64, off = ncol % 64;
            const __nv_bfloat16* u = reinterpret_cast<const __nv_bfloat16*>(p.aux0) + row * p.ldaux + blk * 128 + off;
            float xv[32], gv[32];
            load_bf16_32(u, xv);
            load_bf16_32(u + 64, gv);
            float dxv[32], dgv[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              dxv[i] = v[i] * gelu_exact(gv[i]);
              dgv[i] = v[i] * xv[i] * gelu_exact_grad(gv[i]);
            }
            __nv_bfloat16* du = reinterpret_cast<__nv_bfloat16*>(p.out0) + row * p.ld0 + blk * 128 + off;
            store_bf16_32(du, dxv);
            store_bf16_32(du + 64, dgv);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[as]);
      if (++as == 2) as = 0, aph ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN, bool A_MN, bool B_MN>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM) != cudaSuccess)
      return MCA_ERR_CUDA;
    attr_set = true;
  }
  const int tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN) * p.k_splits;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  kern<<<grid, GEMM_THREADS, Cfg::SMEM, stream>>>(tmA, tmB, p);
  return cudaGetLastError() == cudaSuccess ? MCA_OK : MCA_ERR_CUDA;
}

}  // namespace mca

using namespace mca;

extern "C" int mca_gemm_bf16(const void* A, int a_mn_major, long long lda, const void* B, int b_mn_major,
                             long long ldb, int M, int N, int K, int k_splits, int mode, void* out0, long long ld0,
                             void* out1, long long ld1, const void* aux0, long long ldaux, const float* bias,
                             float alpha, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (M <= 0 || N <= 0 || K <= 0 || k_splits < 1 || (N % 32) != 0) return MCA_ERR_SHAPE;
  if (mode != MCA_EPI_F32 && k_splits != 1) return MCA_ERR_SHAPE;
  const int kb_total = (K + BK - 1) / BK;
  if (k_splits > kb_total) k_splits = kb_total;
  {  // every split must own at least one k-block
    const int per = (kb_total + k_splits - 1) / k_splits;
    k_splits = (kb_total + per - 1) / per;
  }
  const int BN = 128;
  CUtensorMap tmA, tmB;
  int rc;
  // K-major: global [rows, K] (K contiguous), box {64 k, rows}.  MN-major: global [K, rows] (rows contiguous), box {64 rows, 64 k}.
  rc = a_mn_major ? make_tmap_2d_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, BK)
                  : make_tmap_2d_bf16(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM);
  if (rc != MCA_OK) return rc;
  rc = b_mn_major ? make_tmap_2d_bf16(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, BK)
                  : make_tmap_2d_bf16(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, BN);
  if (rc != MCA_OK) return rc;
  GemmParams p;
  p.M = M, p.N = N, p.K = K, p.k_splits = k_splits, p.mode = mode;
  p.out0 = out0, p.ld0 = ld0, p.out1 = out1, p.ld1 = ld1, p.aux0 = aux0, p.ldaux = ldaux, p.bias = bias;
  p.alpha = alpha;
  if (!a_mn_major && !b_mn_major) return launch_gemm<128, false, false>(tmA, tmB, p, stream);
  if (!a_mn_major && b_mn_major) return launch_gemm<128, false, true>(tmA, tmB, p, stream);
  if (a_mn_major && b_mn_major) return launch_gemm<128, true, true>(tmA, tmB, p, stream);
  return launch_gemm<128, true, false>(tmA, tmB, p, stream);
}
