// Microbenchmark: issue rate of tcgen05.mma (kind::f16, bf16 -> fp32) for the shapes the MCA kernels use.
// One CTA per SM, one thread issues `iters` MMAs back to back on static smem / TMEM operands, one commit at the end.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../include -I../../mca_paper_b200/csrc mma_rate.cu -o mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace mca;

struct Cfg { int n; int a_mode; int b_mn; int other_warps_lds; };  // a_mode: 0 smem K-major, 1 smem MN-major, 2 TMEM

__global__ void __launch_bounds__(160, 1) mma_rate_kernel(int n, int a_mode, int b_mn, int iters, int lds_traffic, long long* out, int n_acc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&holder, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = holder;
  if (warp == 0 && lane == 0) {
    const uint32_t idesc = make_idesc_bf16(128, n, a_mode == 1, b_mn != 0);
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 32768);
    uint64_t da[4], db[4];
    for (int k = 0; k < 4; ++k) {
      da[k] = a_mode == 1 ? make_smem_desc_sw128(a_addr + k * 2048, 8192, 1024) : make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
      db[k] = b_mn ? make_smem_desc_sw128(b_addr + k * 2048, 8192, 1024) : make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
    }
    const uint32_t d0 = tm, d1 = tm + (n_acc > 1 ? n : 0);
    long long t0 = clock64();
    if (a_mode == 2) {
      for (int i = 0; i < iters; i += 8) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          umma_bf16_ts(d0, tm + 480 + k * 8, db[k], idesc, 1u);
          umma_bf16_ts(d1, tm + 480 + k * 8, db[k], idesc, 1u);
        }
      }
    } else {
      for (int i = 0; i < iters; i += 8) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          umma_bf16(d0, da[k], db[k], idesc, 1u);
          umma_bf16(d1, da[k], db[k], idesc, 1u);
        }
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  } else if (warp >= 1 && lds_traffic) {
    // competing shared-memory traffic from 4 warps (128-bit loads + stores on a private region)
    uint4* p = reinterpret_cast<uint4*>(smem + 65536) + (threadIdx.x - 32);
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (int i = 0; i < lds_traffic; ++i) {
      uint4 v = p[(i & 7) * 128];
      acc.x += v.x; acc.y ^= v.y;
      if (lds_traffic > 1 && (i & 1)) p[(i & 7) * 128] = acc;
    }
    if (acc.x == 0x12345678u) out[1] = acc.y;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  const int iters = 2048;
  const char* am[] = {"A smem K-major", "A smem MN-major", "A TMEM"};
  for (int n_acc : {1, 2})
    for (int a_mode : {0, 1, 2})
      for (int b_mn : {0, 1})
        for (int n : {64, 128, 256}) {
          if (n_acc * n > 448) continue;
          const int grid = 148;
          double res[3];
          int li = 0;
          for (int lds : {0, 3000, 6001}) {  // 0: none; even: 128-bit loads only; odd: loads + stores (4 warps)
            mma_rate_kernel<<<grid, 160, 96 * 1024>>>(n, a_mode, b_mn, 64, 0, d, n_acc);  // warm
            mma_rate_kernel<<<grid, 160, 96 * 1024>>>(n, a_mode, b_mn, iters, lds, d, n_acc);
            long long c = 0;
            cudaError_t e = cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
            res[li++] = double(c) / iters;
          }
          printf("accumulators %d  M=128 N=%3d K=16  %-16s B %-8s : %7.1f cycles/MMA  (ideal %d)   with LDS traffic %7.1f   with LDS+STS traffic %7.1f\n",
                 n_acc, n, am[a_mode], b_mn ? "MN-major" : "K-major", res[0], 128 * n / 256, res[1], res[2]);
        }
  return 0;
}
