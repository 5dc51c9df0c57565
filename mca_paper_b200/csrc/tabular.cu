// TabularEncoder pieces (TCGA configs): encoders.py:25-37 (nn.Embedding(max_norm=1.0) renormalises looked-up rows
// in place on every forward), encoders.py:55-72 (ContinuousValueEncoder: pad test on the raw value, clamp(max),
// Linear(1,d) -> ReLU; the d x d Linear runs on the tcgen05 GEMM, its LayerNorm + pad-zero + embedding add on
// ln512_fwd).  All bandwidth-bound elementwise work.
#include "mca_b200.h"
#include "ptx.cuh"
#include "runtime.h"

namespace mca {

// rows with ||row|| > max_norm are scaled by max_norm / (||row|| + 1e-7)  (torch embedding_renorm_)
__global__ void __launch_bounds__(256)
embedding_renorm_kernel(float* __restrict__ emb, int rows, int d, float max_norm) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  float* e = emb + static_cast<long long>(r) * d;
  float ss = 0.f;
  for (int c = lane; c < d; c += 32) ss += e[c] * e[c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float norm = sqrtf(ss);
  if (norm > max_norm) {
    const float s = max_norm / (norm + 1e-7f);
    for (int c = lane; c < d; c += 32) e[c] *= s;
  }
}

// h1[r, c] = relu(w1[c] * min(v[r], max_value) + b1[c]) (bf16 GEMM operand); vpad[r] = (v[r] == padding_value)
__global__ void __launch_bounds__(256)
tabular_fwd_kernel(const float* __restrict__ values, const float* __restrict__ w1, const float* __restrict__ b1,
                   __nv_bfloat16* __restrict__ h1, uint8_t* __restrict__ vpad, float max_value, float padding_value,
                   int d, long long rows) {
  const long long n = rows * (d / 2);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / (d / 2);
    const int c = static_cast<int>(i % (d / 2)) * 2;
    const float v = values[r];
    if (c == 0) vpad[r] = v == padding_value ? 1 : 0;
    const float vc = fminf(v, max_value);
    const float a = fmaxf(w1[c] * vc + b1[c], 0.f), b = fmaxf(w1[c + 1] * vc + b1[c + 1], 0.f);
    *reinterpret_cast<uint32_t*>(h1 + r * d + c) = pack_bf16x2(a, b);
  }
}

// dw1[c] += sum_r dh1[r,c] * [pre > 0] * vc ; db1[c] += sum_r dh1[r,c] * [pre > 0]
__global__ void __launch_bounds__(256)
tabular_bwd_kernel(const float* __restrict__ dh1, const float* __restrict__ values, const float* __restrict__ w1,
                   const float* __restrict__ b1, float* __restrict__ dw1, float* __restrict__ db1, float max_value,
                   int d, long long rows, int rows_per_block) {
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    const float w = w1[c], b = b1[c];
    float aw = 0.f, ab = 0.f;
    for (long long r = r0; r < r1; ++r) {
      const float vc = fminf(values[r], max_value);
      if (w * vc + b > 0.f) {
        const float g = dh1[r * d + c];
        aw += g * vc;
        ab += g;
      }
    }
    atomicAdd(dw1 + c, aw);
    atomicAdd(db1 + c, ab);
  }
}

}  // namespace mca

using namespace mca;

extern "C" int mca_embedding_renorm(float* emb, int rows, int d, float max_norm, void* stream) {
  if (rows <= 0 || d <= 0) return MCA_ERR_SHAPE;
  embedding_renorm_kernel<<<(rows + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(emb, rows, d, max_norm);
  return check_launch();
}

extern "C" int mca_tabular_fwd(const float* values, const float* w1, const float* b1, void* h1_bf16, uint8_t* vpad,
                               float max_value, float padding_value, int d, long long rows, void* stream) {
  if (rows <= 0 || (d % 2) != 0) return MCA_ERR_SHAPE;
  tabular_fwd_kernel<<<148 * 4, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      values, w1, b1, reinterpret_cast<__nv_bfloat16*>(h1_bf16), vpad, max_value, padding_value, d, rows);
  return check_launch();
}

extern "C" int mca_tabular_bwd(const float* dh1, const float* values, const float* w1, const float* b1, float* dw1,
                               float* db1, const void* unused0, const void* unused1, float max_value,
                               float padding_value, int d, long long rows, void* stream) {
  (void)unused0, (void)unused1, (void)padding_value;
  if (rows <= 0) return MCA_ERR_SHAPE;
  const int rpb = 64;
  tabular_bwd_kernel<<<static_cast<unsigned>((rows + rpb - 1) / rpb), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      dh1, values, w1, b1, dw1, db1, max_value, d, rows, rpb);
  return check_launch();
}
