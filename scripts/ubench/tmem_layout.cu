// Probe of the tcgen05.ld / tcgen05.st fragment layouts used by the attention kernels: fill TMEM through 32x32b stores
// (thread = lane, register = column) with value = lane * 1000 + column, read it back through 16x256b / 16x128b loads
// and check which (lane, column) every register of every thread received; then the reverse for the 16x128b / 16x256b
// stores.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../include -I../../mca_paper_b200/csrc tmem_layout.cu -o tmem_layout
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace mca;

__global__ void __launch_bounds__(128, 1) probe(int* out) {
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc(&holder, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = holder;
  const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
  // fill columns [0, 64): value = row * 1000 + col
  uint32_t v[32];
  for (int c0 = 0; c0 < 64; c0 += 32) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = (warp * 32 + lane) * 1000 + c0 + i;
    tmem_st32(tm + lane_sel + c0, v);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  int bad = 0;
  // 16x256b.x8 : 16 lanes x 64 columns, rows [base, base + 16)
  for (int hf = 0; hf < 2; ++hf) {
    uint32_t r[32];
    tmem_ld16x256b_x8(tm + (static_cast<uint32_t>(warp * 32 + hf * 16) << 16), r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int row = warp * 32 + hf * 16 + lane / 4 + (e >= 2 ? 8 : 0);
        const int col = 8 * j + 2 * (lane % 4) + (e & 1);
        if (r[4 * j + e] != static_cast<uint32_t>(row * 1000 + col)) ++bad;
      }
  }
  atomicAdd(&out[0], bad);
  bad = 0;
  // 16x128b.x16 : 16 lanes x 64 columns
  for (int hf = 0; hf < 2; ++hf) {
    uint32_t r[32];
    tmem_ld16x128b_x16(tm + (static_cast<uint32_t>(warp * 32 + hf * 16) << 16), r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int row = warp * 32 + hf * 16 + lane / 4 + (e ? 8 : 0);
        const int col = 4 * j + (lane % 4);
        if (r[2 * j + e] != static_cast<uint32_t>(row * 1000 + col)) ++bad;
      }
  }
  atomicAdd(&out[1], bad);
  __syncthreads();
  // stores: write columns [64, 128) through 16x128b.x16 and [128, 192) through 16x256b.x8, read back with 32x32b
  for (int hf = 0; hf < 2; ++hf) {
    uint32_t r[32];
#pragma unroll
    for (int j = 0; j < 16; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) r[2 * j + e] = (warp * 32 + hf * 16 + lane / 4 + (e ? 8 : 0)) * 1000 + 4 * j + (lane % 4);
    tmem_st16x128b_x16(tm + 64 + (static_cast<uint32_t>(warp * 32 + hf * 16) << 16), r);
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e)
        r[4 * j + e] = (warp * 32 + hf * 16 + lane / 4 + (e >= 2 ? 8 : 0)) * 1000 + 8 * j + 2 * (lane % 4) + (e & 1);
    tmem_st16x256b_x8(tm + 128 + (static_cast<uint32_t>(warp * 32 + hf * 16) << 16), r);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  bad = 0;
  int bad2 = 0;
  for (int c0 = 0; c0 < 64; c0 += 32) {
    tmem_ld32(tm + lane_sel + 64 + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) if (v[i] != static_cast<uint32_t>((warp * 32 + lane) * 1000 + c0 + i)) ++bad;
    tmem_ld32(tm + lane_sel + 128 + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) if (v[i] != static_cast<uint32_t>((warp * 32 + lane) * 1000 + c0 + i)) ++bad2;
  }
  atomicAdd(&out[2], bad);
  atomicAdd(&out[3], bad2);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 256); }
}

int main() {
  int* d; cudaMalloc(&d, 16); cudaMemset(d, 0, 16);
  probe<<<1, 128>>>(d);
  int h[4];
  cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  printf("mismatches: ld16x256b %d  ld16x128b %d  st16x128b %d  st16x256b %d\n", h[0], h[1], h[2], h[3]);
  return (h[0] | h[1] | h[2] | h[3]) ? 2 : 0;
}
