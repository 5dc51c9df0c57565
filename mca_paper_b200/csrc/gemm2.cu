// CTA-pair (tcgen05 cta_group::2) variant of the persistent bf16 GEMM: two CTAs on neighbouring SMs form a cluster
// and cooperate on one 256 x 256 output tile.  Each CTA stages its own 128 rows of A and HALF of the 256 B rows per
// k-block, the leader CTA issues tcgen05.mma.cta_group::2 (M = 256, N = 256) which reads the operand halves from both
// CTAs' shared memory and writes each CTA's 128 x 256 accumulator slice into that CTA's TMEM.  Per CTA and k-block that
// is 32 KB of TMA fill + 32 KB of operand reads for 512 tensor cycles — half the shared-memory traffic per FLOP of the
// single-CTA 128 x 128 kernel in gemm.cu, whose main loop ncu shows pinned at ~50 % tensor activity by exactly that
// traffic (profiles/r1_ncu_gemm_*).  Same operand layouts, split-K and fused epilogues as gemm.cu.
//
// Barrier protocol (s = smem stage, a = accumulator stage):
//   full[s]   (leader's copy, count 2): each CTA's producer arrives with expect_tx for its own 32 KB; both CTAs' TMA
//             loads complete_tx on the leader's barrier (cta_group::2 form, peer bit cleared in the address)
//   empty[s]  (one per CTA, count 1): the leader's tcgen05.commit multicasts the arrival to both CTAs
//   tfull[a]  (one per CTA, count 1): multicast commit after the last k-block of a tile
//   tempty[a] (leader's copy, count 16): the 8 epilogue warps of BOTH CTAs arrive (remote arrive from the peer)
#include "mca_b200.h"
#include "ptx.cuh"
#include "runtime.h"
#include "gemm_epilogue.cuh"

namespace mca {

namespace g2 {

constexpr int BM = 128;        // rows per CTA (256 per cluster)
constexpr int BN = 256;        // tile columns
constexpr int BNH = BN / 2;    // B rows staged per CTA
constexpr int BK = 64;
constexpr int THREADS = 384;
constexpr int STAGES = 5;
constexpr int A_BYTES = BM * BK * 2;
constexpr int B_BYTES = BNH * BK * 2;
constexpr int EPI_WARPS = 8;
constexpr int EPI_BYTES_PER_WARP = 8192;
constexpr int SMEM = STAGES * (A_BYTES + B_BYTES) + EPI_WARPS * EPI_BYTES_PER_WARP + 1024;
constexpr int TMEM_COLS = 2 * BN;  // 512: double-buffered 128 x 256 fp32 accumulators
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> CTA 0 of the pair

struct Params {
  int M, N, K;
  int k_splits;
  int mode;
  int reduce;  // MCA_EPI_F32 only: TMA reduce-add into slab 0 instead of a store into slab z
  const float* bias;
  float alpha;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* holder, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(holder)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[256 x 16, halves in both CTAs' smem] * B[256 x 16, halves in both CTAs' smem]
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all previously issued MMAs retire) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
      : "memory");
}
// TMA load into this CTA's smem, completion bytes credited to the LEADER CTA's barrier
__device__ __forceinline__ void tma2_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_leader(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(smem_u32(bar) & PEER_MASK), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_MASK) : "memory");
}

template <bool A_MN, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm2_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmO0, const __grid_constant__ CUtensorMap tmO1,
                const __grid_constant__ CUtensorMap tmAux, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint8_t* sEpi = sB + STAGES * B_BYTES;
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], tfull_bar[2], tempty_bar[2], aux_bar[EPI_WARPS];
  __shared__ uint32_t tmem_holder;

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int tiles_m = (p.M + 2 * BM - 1) / (2 * BM);
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
      mbar_init(&full_bar[i], 2);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 2 * EPI_WARPS);
    }
    for (int i = 0; i < EPI_WARPS; ++i) mbar_init(&aux_bar[i], 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc2(&tmem_holder, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anyone arrives on them remotely
  tc_fence_after();
  const uint32_t tmem_base = tmem_holder;
  pdl_wait();  // set-up done under the predecessor's tail; from here on global memory is touched

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; whole warp loops, one elected lane issues) =====================
    int s = 0;
    uint32_t ph = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
      const int nt = tile % tiles_n;
      const int mt = (tile / tiles_n) % tiles_m;
      const int z = tile / (tiles_n * tiles_m);
      const int m0 = mt * 2 * BM + static_cast<int>(rank) * BM;
      const int n0 = nt * BN + static_cast<int>(rank) * BNH;
      const int kb0 = z * kb_per_split;
      const int kb1 = min(kb_total, kb0 + kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* a_dst = sA + s * A_BYTES;
        uint8_t* b_dst = sB + s * B_BYTES;
        if (elect_one()) {
          mbar_expect_tx_leader(&full_bar[s], A_BYTES + B_BYTES);
          if constexpr (!A_MN) {
            tma2_load_2d(a_dst, &tmA, &full_bar[s], kb * BK, m0);
          } else {
#pragma unroll
            for (int h = 0; h < BM / 64; ++h) tma2_load_2d(a_dst + h * 8192, &tmA, &full_bar[s], m0 + 64 * h, kb * BK);
          }
          if constexpr (!B_MN) {
            tma2_load_2d(b_dst, &tmB, &full_bar[s], kb * BK, n0);
          } else {
#pragma unroll
            for (int h = 0; h < BNH / 64; ++h) tma2_load_2d(b_dst + h * 8192, &tmB, &full_bar[s], n0 + 64 * h, kb * BK);
          }
        }
        __syncwarp();
        if (++s == STAGES) s = 0, ph ^= 1;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only; convergent warp, one elected lane issues) =====================
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN, A_MN, B_MN);
      constexpr uint64_t a_step = (A_MN ? 2048u : 32u) >> 4, b_step = (B_MN ? 2048u : 32u) >> 4;
      const uint64_t da0 = A_MN ? make_smem_desc_sw128(smem_u32(sA), 8192, 1024) : make_smem_desc_sw128(smem_u32(sA), 16, 1024);
      const uint64_t db0 = B_MN ? make_smem_desc_sw128(smem_u32(sB), 8192, 1024) : make_smem_desc_sw128(smem_u32(sB), 16, 1024);
      int s = 0, as = 0;
      uint32_t ph = 0, aph = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
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
              umma2_bf16(d_tmem, da + k * a_step, db + k * b_step, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            umma2_commit_mc(&empty_bar[s]);                       // frees the smem slot in both CTAs
            if (kb + 1 == kb1) umma2_commit_mc(&tfull_bar[as]);   // accumulator complete -> both epilogues
          }
          __syncwarp();
          if (++s == STAGES) s = 0, ph ^= 1;
        }
        if (++as == 2) as = 0, aph ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps (both CTAs): TMEM -> registers -> swizzled smem -> TMA store =====================
    const int ew = warp - 4;
    const int q = warp & 3;   // TMEM lane quarter this warp may read
    const int hf = ew >> 2;   // which 128 of the tile's 256 columns; each warp works through them in two 64-column slabs
    uint8_t* stg = sEpi + ew * EPI_BYTES_PER_WARP;
    uint64_t* xbar = &aux_bar[ew];
    const bool has_aux = p.mode == MCA_EPI_RESID || p.mode == MCA_EPI_GEGLU_BWD;
    int as = 0;
    uint32_t aph = 0, xph = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
      const int nt = tile % tiles_n;
      const int mt = (tile / tiles_n) % tiles_m;
      const int z = tile / (tiles_n * tiles_m);
      const int n0 = nt * BN;
      const int row0 = mt * 2 * BM + static_cast<int>(rank) * BM + q * 32;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
#pragma unroll 1
      for (int sl = 0; sl < 2; ++sl) {
        const int sc = hf * 128 + sl * 64;  // first tile column of this slab (all modes but GEGLU)
        // the previous slab's TMA stores must have finished reading this warp's staging area
        if (elect_one()) bulk_wait_group_read0();
        __syncwarp();
        const EpiSlab slab{n0 + sc, n0 + hf * 128, sl, row0, z};
        if (has_aux && elect_one()) epi_request_aux(p.mode, stg, &tmAux, xbar, slab);
        if (sl == 0) {
          mbar_wait(&tfull_bar[as], aph);
          tc_fence_after();
        }
        uint32_t r0[32], r1[32];
        if (p.mode == MCA_EPI_GEGLU) {  // each 128-column block is [64 value | 64 gate]; block hf, 32-pair group sl
          tmem_ld32(t_row + hf * 128 + sl * 32, r0);
          tmem_ld32(t_row + hf * 128 + 64 + sl * 32, r1);
        } else {
          tmem_ld32(t_row + sc, r0);
          tmem_ld32(t_row + sc + 32, r1);
        }
        tmem_ld_wait();
        if (sl == 1) {  // the whole accumulator slice of this warp is in registers: hand the TMEM stage back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(&tempty_bar[as]);
        }

        epi_store_slab(p, r0, r1, stg, lane, slab, &tmO0, &tmO1, xbar, xph);
      }
      if (++as == 2) as = 0, aph ^= 1;
    }
    if (elect_one()) bulk_wait_group0();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA of the pair may free TMEM or exit while its peer can still address it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, TMEM_COLS);
  }
}

template <bool A_MN, bool B_MN>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO0, const CUtensorMap& tmO1,
                  const CUtensorMap& tmAux, const Params& p, cudaStream_t stream) {
  auto kern = gemm2_tc_kernel<A_MN, B_MN>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess) return MCA_ERR_CUDA;
    attr_set = true;
  }
  const int tiles = ((p.M + 2 * BM - 1) / (2 * BM)) * ((p.N + BN - 1) / BN) * p.k_splits;
  const int max_clusters = num_sms() / 2;
  const int clusters = tiles < max_clusters ? tiles : max_clusters;
  if (launch_kernel(kern, dim3(2 * clusters), dim3(THREADS), SMEM, stream, 1, tmA, tmB, tmO0, tmO1, tmAux, p) != cudaSuccess)
    return MCA_ERR_CUDA;
  return cudaGetLastError() == cudaSuccess ? MCA_OK : MCA_ERR_CUDA;
}

}  // namespace g2

// Same contract as mca_gemm_bf16 (arguments already validated by the caller).
int gemm2_dispatch(const void* A, int a_mn_major, long long lda, const void* B, int b_mn_major, long long ldb, int M, int N,
                   int K, int k_splits, int mode, void* out0, long long ld0, void* out1, long long ld1, const void* aux0,
                   long long ldaux, const float* bias, float alpha, cudaStream_t stream) {
  using namespace g2;
  CUtensorMap tmA, tmB, tmO0, tmO1, tmAux;
  int rc;
  rc = a_mn_major ? make_tmap_2d_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, BK)
                  : make_tmap_2d_bf16(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM);
  if (rc != MCA_OK) return rc;
  rc = b_mn_major ? make_tmap_2d_bf16(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, BK)
                  : make_tmap_2d_bf16(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, BNH);
  if (rc != MCA_OK) return rc;
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
  } else {
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
  Params p;
  p.M = M, p.N = N, p.K = K, p.k_splits = k_splits, p.bias = bias, p.alpha = alpha;
  p.reduce = mode == MCA_EPI_F32_ACC ? 1 : 0;
  p.mode = mode == MCA_EPI_F32_ACC ? static_cast<int>(MCA_EPI_F32) : mode;
  if (!a_mn_major && !b_mn_major) return launch<false, false>(tmA, tmB, tmO0, tmO1, tmAux, p, stream);
  if (!a_mn_major && b_mn_major) return launch<false, true>(tmA, tmB, tmO0, tmO1, tmAux, p, stream);
  if (a_mn_major && b_mn_major) return launch<true, true>(tmA, tmB, tmO0, tmO1, tmAux, p, stream);
  return launch<true, false>(tmA, tmB, tmO0, tmO1, tmAux, p, stream);
}

}  // namespace mca
