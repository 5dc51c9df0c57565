// Persistent warp-specialised bf16 GEMM for sm_100a: TMA -> 128B-swizzled smem ring -> tcgen05.mma (fp32
// accumulators in TMEM, double buffered) -> fused epilogues read back with tcgen05.ld, staged in swizzled shared
// memory and written (and, for the in-place epilogues, pre-loaded) by TMA so every global access is a full line.
//
//   out[M,N] = A[M,K] * B[N,K]^T           (both operands may independently be K-major or MN-major in memory)
//
// This one kernel serves every dense contraction of the MCA training step (reference: model.py:83,105
// to_q/to_kv/to_out, model.py:49-51 GEGLU feed-forward, encoders.py:190 token projection, and their autograd
// transposes):  forward uses (A K-major, B K-major), dX uses (A K-major, B MN-major = the weight as stored),
// dW uses (A MN-major, B MN-major) with split-K over the token dimension.
//
// Warp roles (384 threads, one CTA per SM): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator,
// warps 4-11 epilogue.  Epilogue warp w owns TMEM lane quarter w%4 (32 accumulator rows) and column half (w-4)/4
// (64 of the tile's 128 columns); it has a private 8 KB staging area, so the epilogue needs no block-wide barrier.
#include <cstdlib>

#include "mca_b200.h"
#include "ptx.cuh"
#include "runtime.h"
#include "gemm_epilogue.cuh"

namespace mca {

constexpr int BM = 128;
constexpr int BN = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle-128B row
constexpr int GEMM_THREADS = 384;
constexpr int STAGES = 5;
constexpr int A_BYTES = BM * BK * 2;
constexpr int B_BYTES = BN * BK * 2;
constexpr int EPI_WARPS = 8;
constexpr int EPI_BYTES_PER_WARP = 8192;
constexpr int GEMM_SMEM = STAGES * (A_BYTES + B_BYTES) + EPI_WARPS * EPI_BYTES_PER_WARP + 1024 /*align slack*/ + 256;
constexpr int TMEM_COLS = 2 * BN;

struct GemmParams {
  int M, N, K;
  int k_splits;
  int mode;
  int reduce;  // MCA_EPI_F32 only: TMA reduce-add into slab 0 instead of a store into slab z
  const float* bias;
  float alpha;
};

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmO0, const __grid_constant__ CUtensorMap tmO1,
               const __grid_constant__ CUtensorMap tmAux, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  pdl_launch_dependents();
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = base;
  uint8_t* sB = base + STAGES * A_BYTES;
  uint8_t* sEpi = sB + STAGES * B_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sEpi + EPI_WARPS * EPI_BYTES_PER_WARP);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* aux_bar = tempty_bar + 2;  // [EPI_WARPS]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(aux_bar + EPI_WARPS);

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
    tma_prefetch_desc(&tmO0);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], EPI_WARPS);
    }
    for (int i = 0; i < EPI_WARPS; ++i) mbar_init(&aux_bar[i], 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_holder, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_wait();  // set-up done under the predecessor's tail; from here on global memory is touched

  if (warp == 0) {
    // ===================== TMA producer (whole warp loops, one elected lane issues) =====================
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
        uint8_t* a_dst = sA + s * A_BYTES;
        uint8_t* b_dst = sB + s * B_BYTES;
        if (elect_one()) {
          mbar_expect_tx(&full_bar[s], A_BYTES + B_BYTES);
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
        }
        __syncwarp();
        if (++s == STAGES) s = 0, ph ^= 1;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the loop (convergent code keeps the shared-memory descriptors in uniform registers, so
    // one tcgen05.mma costs a couple of uniform adds); a single elected lane issues the MMAs and the commits.
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN, B_MN);
    // k-step inside a 64-wide k-block: K-major advances 32 B along the swizzled row, MN-major 16 k-rows = 2 KB
    constexpr uint64_t a_step = (A_MN ? 2048u : 32u) >> 4, b_step = (B_MN ? 2048u : 32u) >> 4;
    const uint64_t da0 = A_MN ? make_smem_desc_sw128(smem_u32(sA), 8192, 1024) : make_smem_desc_sw128(smem_u32(sA), 16, 1024);
    const uint64_t db0 = B_MN ? make_smem_desc_sw128(smem_u32(sB), 8192, 1024) : make_smem_desc_sw128(smem_u32(sB), 16, 1024);
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
        const uint64_t da = da0 + static_cast<uint64_t>((s * A_BYTES) >> 4);
        const uint64_t db = db0 + static_cast<uint64_t>((s * B_BYTES) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16(d_tmem, da + k * a_step, db + k * b_step, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs retire
          if (kb + 1 == kb1) umma_commit(&tfull_bar[as]);  // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++s == STAGES) s = 0, ph ^= 1;
      }
      if (++as == 2) as = 0, aph ^= 1;
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps: TMEM -> registers -> swizzled smem -> TMA store =====================
    const int ew = warp - 4;
    const int q = warp & 3;   // TMEM lane quarter this warp may read
    const int hf = ew >> 2;   // column half of the tile
    uint8_t* stg = sEpi + ew * EPI_BYTES_PER_WARP;
    uint64_t* xbar = &aux_bar[ew];
    const bool has_aux = p.mode == MCA_EPI_RESID || p.mode == MCA_EPI_GEGLU_BWD;
    int as = 0;
    uint32_t aph = 0, xph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int nt = tile % tiles_n;
      const int mt = (tile / tiles_n) % tiles_m;
      const int z = tile / (tiles_n * tiles_m);
      const int n0 = nt * BN;
      const int row0 = mt * BM + q * 32;
      // the previous tile's TMA stores must have finished reading this warp's staging area
      if (elect_one()) bulk_wait_group_read0();  // same membermask -> same leader as the lane that issued the stores
      __syncwarp();
      const EpiSlab slab{n0 + hf * 64, n0, hf, row0, z};
      if (has_aux && elect_one()) epi_request_aux(p.mode, stg, &tmAux, xbar, slab);
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
      uint32_t r0[32], r1[32];
      if (p.mode == MCA_EPI_GEGLU) {  // tile columns are [64 value | 64 gate]
        tmem_ld32(t_row + hf * 32, r0);
        tmem_ld32(t_row + 64 + hf * 32, r1);
      } else {
        tmem_ld32(t_row + hf * 64, r0);
        tmem_ld32(t_row + hf * 64 + 32, r1);
      }
      tmem_ld_wait();
      // accumulator is in registers: hand the TMEM buffer back to the MMA warp right away
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      if (++as == 2) as = 0, aph ^= 1;

      epi_store_slab(p, r0, r1, stg, lane, slab, &tmO0, &tmO1, xbar, xph);
    }
    if (elect_one()) bulk_wait_group0();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <bool A_MN, bool B_MN>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO0, const CUtensorMap& tmO1,
                       const CUtensorMap& tmAux, const GemmParams& p, cudaStream_t stream) {
  auto kern = gemm_tc_kernel<A_MN, B_MN>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM) != cudaSuccess)
      return MCA_ERR_CUDA;
    attr_set = true;
  }
  const int tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN) * p.k_splits;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  if (launch_kernel(kern, dim3(grid), dim3(GEMM_THREADS), GEMM_SMEM, stream, 1, tmA, tmB, tmO0, tmO1, tmAux, p) != cudaSuccess)
    return MCA_ERR_CUDA;
  return cudaGetLastError() == cudaSuccess ? MCA_OK : MCA_ERR_CUDA;
}

}  // namespace mca

using namespace mca;

namespace mca {
int gemm2_dispatch(const void* A, int a_mn_major, long long lda, const void* B, int b_mn_major, long long ldb, int M, int N,
                   int K, int k_splits, int mode, void* out0, long long ld0, void* out1, long long ld1, const void* aux0,
                   long long ldaux, const float* bias, float alpha, cudaStream_t stream);
// CTA-pair kernel (gemm2.cu) unless MCA_GEMM_2CTA=0 or the problem has fewer than 256 rows
static bool use_cta_pairs() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MCA_GEMM_2CTA");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}
}  // namespace mca

extern "C" int mca_gemm_bf16(const void* A, int a_mn_major, long long lda, const void* B, int b_mn_major,
                             long long ldb, int M, int N, int K, int k_splits, int mode, void* out0, long long ld0,
                             void* out1, long long ld1, const void* aux0, long long ldaux, const float* bias,
                             float alpha, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (M <= 0 || N <= 0 || K <= 0 || k_splits < 1 || (N % 32) != 0) return MCA_ERR_SHAPE;
  if (mode < MCA_EPI_BF16 || mode > MCA_EPI_F32_ACC || out0 == nullptr) return MCA_ERR_ARG;
  if (mode != MCA_EPI_F32 && mode != MCA_EPI_F32_ACC && k_splits != 1) return MCA_ERR_SHAPE;
  if ((mode == MCA_EPI_GEGLU && (out1 == nullptr || (N % 128) != 0)) ||
      ((mode == MCA_EPI_RESID || mode == MCA_EPI_GEGLU_BWD) && aux0 == nullptr) ||
      (mode == MCA_EPI_GEGLU_BWD && (N % 64) != 0))
    return MCA_ERR_ARG;
  k_splits = gemm_effective_splits(K, k_splits);  // every split owns at least one k-block
  if (M >= 256 && use_cta_pairs())
    return gemm2_dispatch(A, a_mn_major, lda, B, b_mn_major, ldb, M, N, K, k_splits, mode, out0, ld0, out1, ld1, aux0, ldaux,
                          bias, alpha, stream);
  CUtensorMap tmA, tmB, tmO0, tmO1, tmAux;
  int rc;
  // K-major: global [rows, K] (K contiguous), box {64 k, rows}.  MN-major: global [K, rows] (rows contiguous), box {64 rows, 64 k}.
  rc = a_mn_major ? make_tmap_2d_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, BK)
                  : make_tmap_2d_bf16(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM);
  if (rc != MCA_OK) return rc;
  rc = b_mn_major ? make_tmap_2d_bf16(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, BK)
                  : make_tmap_2d_bf16(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, BN);
  if (rc != MCA_OK) return rc;
  // epilogue maps: one box = what one epilogue warp stages (32 rows x 128 B, or x 64 B for the GEGLU outputs)
  const uint64_t uM = (uint64_t)M, uN = (uint64_t)N;
  if (mode == MCA_EPI_BF16) {
    const uint64_t dims[2] = {uN, uM}, st[1] = {(uint64_t)ld0};
    const uint32_t box[2] = {64, 32};
    rc = make_tmap(&tmO0, 2, out0, 2, dims, st, box, 128);
  } else if (mode == MCA_EPI_F32 || mode == MCA_EPI_RESID || mode == MCA_EPI_F32_ACC) {
    const uint64_t dims[3] = {uN, uM, (uint64_t)(mode == MCA_EPI_F32_ACC ? 1 : k_splits)}, st[2] = {(uint64_t)ld0, uM * (uint64_t)ld0};
    const uint32_t box[3] = {32, 32, 1};
    rc = make_tmap(&tmO0, 4, out0, 3, dims, st, box, 128);
  } else if (mode == MCA_EPI_GEGLU) {
    const uint64_t dims0[2] = {uN / 2, uM}, st0[1] = {(uint64_t)ld0};
    const uint64_t dims1[2] = {uN, uM}, st1[1] = {(uint64_t)ld1};
    const uint32_t box[2] = {32, 32};
    rc = make_tmap(&tmO0, 2, out0, 2, dims0, st0, box, 64);
    if (rc == MCA_OK) rc = make_tmap(&tmO1, 2, out1, 2, dims1, st1, box, 64);
  } else {  // GEGLU_BWD: du has 2N columns
    const uint64_t dims[2] = {2 * uN, uM}, st[1] = {(uint64_t)ld0};
    const uint32_t box[2] = {64, 32};
    rc = make_tmap(&tmO0, 2, out0, 2, dims, st, box, 128);
  }
  if (rc != MCA_OK) return rc;
  if (mode != MCA_EPI_GEGLU) tmO1 = tmO0;
  if (mode == MCA_EPI_RESID) {
    const uint64_t dims[2] = {uN, uM}, st[1] = {(uint64_t)ldaux};
    const uint32_t box[2] = {32, 32};
    rc = make_tmap(&tmAux, 4, aux0, 2, dims, st, box, 128);
  } else if (mode == MCA_EPI_GEGLU_BWD) {
    const uint64_t dims[2] = {2 * uN, uM}, st[1] = {(uint64_t)ldaux};
    const uint32_t box[2] = {64, 32};
    rc = make_tmap(&tmAux, 2, aux0, 2, dims, st, box, 128);
  } else {
    tmAux = tmO0;
  }
  if (rc != MCA_OK) return rc;
  GemmParams p;
  p.M = M, p.N = N, p.K = K, p.k_splits = k_splits, p.bias = bias, p.alpha = alpha;
  p.reduce = mode == MCA_EPI_F32_ACC ? 1 : 0;
  p.mode = mode == MCA_EPI_F32_ACC ? static_cast<int>(MCA_EPI_F32) : mode;
  if (!a_mn_major && !b_mn_major) return launch_gemm<false, false>(tmA, tmB, tmO0, tmO1, tmAux, p, stream);
  if (!a_mn_major && b_mn_major) return launch_gemm<false, true>(tmA, tmB, tmO0, tmO1, tmAux, p, stream);
  if (a_mn_major && b_mn_major) return launch_gemm<true, true>(tmA, tmB, tmO0, tmO1, tmAux, p, stream);
  return launch_gemm<true, false>(tmA, tmB, tmO0, tmO1, tmAux, p, stream);
}
