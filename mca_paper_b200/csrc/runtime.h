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

// MCA_PDL: 0 disables programmatic dependent launch (every kernel starts after its predecessor has completed), 1 (default)
// enables it for the kernels with a real set-up phase (GEMM, attention: 5.01 -> 4.91 ms per step), 2 also for the bandwidth
// kernels (LayerNorm, cast), which only hides their launch latency and measured slightly worse (4.93 - 4.97 ms: their early
// resident CTAs take slots from the draining predecessor)
int pdl_level();

// kern<<<grid, block, smem, stream>>>(args...) with the programmatic-stream-serialization attribute when `pdl` (the kernel
// must call pdl_wait() before its first global-memory access)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int pdl,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = (pdl > 0 && pdl_level() >= pdl) ? 1 : 0;   // pdl = the level from which this launch is programmatic
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace mca
