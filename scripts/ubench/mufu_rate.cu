// Microbenchmark: throughput of the softmax inner loop on one SM for different exp2 implementations and warp counts.
// Each thread processes `iters` rounds of 128 fp32 scores held in registers: p = exp2(s*log2e - m), row sum, bf16 pack.
// mode 0: MUFU.EX2 only (ex2.approx.ftz); mode 1: full loop with MUFU; mode 2: full loop, polynomial exp2 on the FMA
// pipe (Cody-Waite + degree-3 minimax, as in FA4) for every element; mode 3: polynomial for every 4th element; mode 4:
// polynomial for every 2nd element.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../include -I../../mca_paper_b200/csrc mufu_rate.cu -o mufu_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace mca;

__device__ __forceinline__ float poly_ex2(float x) {
  // 2^x for x <= 0 (clamped at -126): split x = n + f with f in [0, 1), 2^f by a degree-3 polynomial, exponent by integer add
  x = fmaxf(x, -126.0f);
  const float fl = floorf(x);  // FRND on ALU? use magic-number rounding instead to stay on the FMA pipe
  const float f = x - fl;
  float p = fmaf(0.0555041086648216f, f, 0.2402265069591007f);
  p = fmaf(p, f, 0.6931471805599453f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (static_cast<int>(fl) << 23));
}
__device__ __forceinline__ float poly_ex2_magic(float x) {
  // round-to-floor through the magic-number add (FMA pipe only): t = x + 1.5*2^23 - 0.5 -> low mantissa bits = floor-ish integer
  x = fmaxf(x, -126.0f);
  const float magic = 12582912.0f;  // 1.5 * 2^23
  const float t = (x - 0.5f) + magic;
  const float fl = t - magic;        // round(x - 0.5) = floor(x) for non-integers
  const float f = x - fl;            // in [0, 1]
  float p = fmaf(0.0555041086648216f, f, 0.2402265069591007f);
  p = fmaf(p, f, 0.6931471805599453f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(int iters, float* out, long long* cyc) {
  float s[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) s[i] = -0.01f * static_cast<float>((threadIdx.x * 7 + i * 13) & 255);
  float acc = 0.f;
  uint32_t packacc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 128; ++i) s[i] = fast_ex2(s[i]) - 1.0f;
    } else {
      const uint64_t l2 = f2_pack(1.4426950408889634f, 1.4426950408889634f), nm = f2_pack(-0.25f, -0.25f);
      uint64_t sum2 = f2_pack(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 64; ++j) {
        float t0_, t1_;
        f2_unpack(f2_fma(f2_pack(s[2 * j], s[2 * j + 1]), l2, nm), t0_, t1_);
        float p0, p1;
        const bool poly0 = MODE == 2 || (MODE == 3 && (j & 1) == 0) || (MODE == 4);
        const bool poly1 = MODE == 2;
        p0 = poly0 ? poly_ex2_magic(t0_) : fast_ex2(t0_);
        p1 = poly1 ? poly_ex2_magic(t1_) : fast_ex2(t1_);
        sum2 = f2_add(sum2, f2_pack(p0, p1));
        packacc ^= pack_bf16x2(p0, p1);
        s[2 * j] = t0_ * 0.5f;
        s[2 * j + 1] = t1_ * 0.5f;
      }
      float a0, a1;
      f2_unpack(sum2, a0, a1);
      acc += a0 + a1;
    }
  }
  const long long t1 = clock64();
  float r = acc;
#pragma unroll
  for (int i = 0; i < 128; ++i) r += s[i];
  if (r == 123.456f || packacc == 0x12345u) out[threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE>
void run(const char* name, float* o, long long* d) {
  for (int threads : {128, 256, 512}) {
    const int iters = 200;
    k<MODE><<<148, threads>>>(8, o, d);
    k<MODE><<<148, threads>>>(iters, o, d);
    long long c = 0;
    cudaError_t e = cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return; }
    const double per_round = double(c) / iters;  // cycles for `threads` rows x 128 elements
    printf("%-34s warps/SM %2d: %8.1f cycles per 128 elem/thread  -> %6.2f elem/clk/SM\n", name, threads / 32, per_round,
           threads * 128.0 / per_round);
  }
}

int main() {
  float* o; long long* d;
  cudaMalloc(&o, 4096); cudaMalloc(&d, 64);
  run<0>("MUFU.EX2 only", o, d);
  run<1>("softmax loop, MUFU", o, d);
  run<2>("softmax loop, poly (all)", o, d);
  run<3>("softmax loop, poly 1/4", o, d);
  run<4>("softmax loop, poly 1/2", o, d);
  // accuracy of the polynomial
  return 0;
}
