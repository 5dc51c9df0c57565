// Microbenchmark: TMEM -> register (tcgen05.ld) and register -> TMEM (tcgen05.st) throughput per SM, alone and with
// a concurrent stream of tcgen05.mma (SS form, 128x64x16) from a fifth warp — the situation of the attention kernels,
// where the softmax / dS warps drain S and dP while the tensor core keeps accumulating.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../include -I../../mca_paper_b200/csrc tmem_rate.cu -o tmem_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace mca;

// warps 0..nw-1 load (mode 0) or store (mode 1) `iters` x 4 x (32 lanes x 32 columns); warp 8 optionally issues MMAs
__global__ void __launch_bounds__(288, 1) tmem_rate_kernel(int nw, int mode, int iters, int mma_iters, int mma_ts, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t holder;
  __shared__ long long tstart[9], tend[9];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&holder, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = holder;
  uint32_t sink = 0;
  if (warp < nw) {
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t col0 = (warp >> 2) * 128;  // warps 4..7 use another column range of the same lane quarter
    uint32_t v[32], v1[32], v2[32], v3[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = i + lane, v1[i] = v2[i] = v3[i] = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (mode == 0) {
        tmem_ld32(tm + lane_sel + col0, v);
        tmem_ld32(tm + lane_sel + col0 + 32, v1);
        tmem_ld32(tm + lane_sel + col0 + 64, v2);
        tmem_ld32(tm + lane_sel + col0 + 96, v3);
        tmem_ld_wait();
        sink += v[0] + v1[31] + v2[7] + v3[9];
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_st32(tm + lane_sel + col0 + c * 32, v);
        tmem_st_wait();
      }
    }
    const long long t1 = clock64();
    if (lane == 0) tstart[warp] = t0, tend[warp] = t1;
  } else if (warp == 8 && mma_iters > 0) {
    const uint32_t idesc = make_idesc_bf16(128, 64, false, false);
    const uint32_t idesc_ts = make_idesc_bf16(128, 64, false, true);
    const uint64_t da = make_smem_desc_sw128(smem_u32(smem), 16, 1024), db = make_smem_desc_sw128(smem_u32(smem + 32768), 16, 1024);
    const uint64_t dbm = make_smem_desc_sw128(smem_u32(smem + 32768), 8192, 1024);
    const long long t0 = clock64();
    for (int i = 0; i < mma_iters; i += 4) {
      if (elect_one()) {
        if (mma_ts) {  // A operand from TMEM (the dV / dK / PV products of the attention kernels), B MN-major
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ts(tm + 448, tm + 384 + k * 8, dbm + k * 128, idesc_ts, 1u);
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tm + 448, da + k * 2, db + k * 2, idesc, 1u);
        }
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (lane == 0) tstart[8] = t0, tend[8] = t1;
  }
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    long long a = tstart[0], b = tend[0];
    for (int w = 1; w < nw; ++w) a = min(a, tstart[w]), b = max(b, tend[w]);
    out[0] = b - a;
    out[1] = mma_iters > 0 ? tend[8] - tstart[8] : 0;
    out[2] = sink;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 64);
  cudaFuncSetAttribute(tmem_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024 + 1024);
  const int iters = 2000;
  printf("%-6s %-5s %-9s %12s %12s %14s\n", "mode", "warps", "mma", "cycles", "mma cycles", "B/cycle/SM");
  for (int mode = 0; mode < 2; ++mode)
    for (int nw : {1, 4, 8})
      for (int mma : {0, 1, 2}) {
        const int mma_iters = mma ? 16000 : 0;
        tmem_rate_kernel<<<1, 288, 64 * 1024 + 1024>>>(nw, mode, iters, mma_iters, mma == 2, out);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        const double bytes = double(nw) * iters * 4 * 32 * 32 * 4;
        printf("%-6s %-5d %-9s %12lld %12lld %14.1f\n", mode == 0 ? "ld" : "st", nw, mma == 0 ? "-" : (mma == 1 ? "SS 128x64" : "TS 128x64"), out[0], out[1],
               bytes / double(out[0]));
      }
  return 0;
}
