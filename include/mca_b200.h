/* mca_b200.h — C ABI of libmca_b200.so, the B200 (sm_100a) implementation of the mca-paper training hot path.
 *
 * The reference (josiahbjorgaard/mca-paper) has no FFI layer: its boundary is the Python nn.Module API of
 * model.py / encoders.py / utils/contrastive_loss_with_temperature.py.  This header is what a
 * Python (ctypes) or C++ host binds instead of the ATen calls those modules make; every entry point cites the
 * reference lines it replaces.  Conventions:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in `_host`
 *   - the library never allocates or frees device memory and never synchronises; work is enqueued on `stream`
 *     (a cudaStream_t passed as void*), so every call is CUDA-graph capturable
 *   - return value: MCA_OK or one of the MCA_ERR_* codes below
 *   - bf16 tensors are raw 16-bit storage (`void*`), fp32 tensors are `float*`
 */
#ifndef MCA_B200_H
#define MCA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  MCA_OK = 0,
  MCA_ERR_SHAPE = 1,     /* bad dimension / divisibility (reference: AssertionError, model.py:186,416,437) */
  MCA_ERR_ALIGN = 2,     /* pointer or stride not 16-byte aligned */
  MCA_ERR_CUDA = 3,      /* launch / driver failure */
  MCA_ERR_NONFINITE = 4, /* reference: Exception on non-finite tokens, encoders.py:197-213 (device flag) */
  MCA_ERR_ARG = 5
};

/* epilogue selectors of mca_gemm_bf16 */
enum {
  MCA_EPI_BF16 = 0,      /* out0(bf16) = alpha*acc + bias */
  MCA_EPI_F32 = 1,       /* out0(f32)[z] = alpha*acc + bias, one slab per k-split z */
  MCA_EPI_RESID = 2,     /* out0(f32) = alpha*acc + aux0(f32); optional out1(bf16) copy */
  MCA_EPI_GEGLU = 3,     /* out1(bf16) = u = acc (interleaved value|gate), out0(bf16) = gelu(gate)*value */
  MCA_EPI_GEGLU_BWD = 4  /* acc = dL/dh, aux0 = u; out0(bf16) = dL/du */
};

int mca_version(void);

/* Dense contraction out[M,N] = A[M,K] * B[N,K]^T on tcgen05 tensor cores (bf16 in, fp32 accumulate in TMEM).
 * Replaces every nn.Linear matmul of the step and its autograd transposes: model.py:83 (to_q,to_kv), :105 (to_out),
 * :49-51 (feed-forward, with model.py:37-38 GEGLU fused as an epilogue), encoders.py:190 (token projection).
 * a_mn_major/b_mn_major = 0: operand stored [rows, K] with K contiguous (ld = row stride in elements);
 *                       = 1: operand stored [K, rows] with rows contiguous (the autograd transposes need no copies).
 * k_splits > 1 (MCA_EPI_F32 only) splits K; slab z of out0 holds the partial sum of split z; the effective split
 * count is mca_gemm_effective_splits(K, k_splits). N must be a multiple of 32. */
int mca_gemm_bf16(const void* A, int a_mn_major, long long lda, const void* B, int b_mn_major, long long ldb, int M,
                  int N, int K, int k_splits, int mode, void* out0, long long ld0, void* out1, long long ld1,
                  const void* aux0, long long ldaux, const float* bias, float alpha, void* stream);
int mca_gemm_effective_splits(int K, int k_splits);

#ifdef __cplusplus
}
#endif
#endif /* MCA_B200_H */
