// Host-side helpers shared by the launchers: device properties, TMA tensor-map encoding (driver entry point is
// resolved through the runtime so the library never links libcuda directly), launch error mapping.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "mca_b200.h"

namespace mca {

int num_sms();

// 2-D bf16 tensor map, 128B swizzle, zero OOB fill. `inner` is the contiguous dimension (elements), `outer`
// the strided one; `row_stride` in elements.
int make_tmap_2d_bf16(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t row_stride,
                      uint32_t box_inner, uint32_t box_outer);

// same for fp32 tensors (box_inner * 4 bytes must be <= 128 for the 128B swizzle)
int make_tmap_2d_f32(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t row_stride,
                     uint32_t box_inner, uint32_t box_outer);

// general form: rank <= 3, elem_bytes 2 (bf16) or 4 (fp32); strides_elems[i] = stride of dim i+1 in elements;
// swizzle_bytes 0 / 64 / 128 (box[0] * elem_bytes must not exceed it when non-zero)
int make_tmap(CUtensorMap* out, int elem_bytes, const void* ptr, int rank, const uint64_t* dims,
              const uint64_t* strides_elems, const uint32_t* box, int swizzle_bytes);

inline int gemm_effective_splits(int K, int k_splits) {
  const int kb_total = (K + 63) / 64;
  if (k_splits < 1) k_splits = 1;
  if (k_splits > kb_total) k_splits = kb_total;
  const int per = (kb_total + k_splits - 1) / k_splits;
  return (kb_total + per - 1) / per;
}

inline int check_launch() { return cudaGetLastError() == cudaSuccess ? MCA_OK : MCA_ERR_CUDA; }

}  // namespace mca
