#include <cstdlib>
#include "runtime.h"

#include <cudaTypedefs.h>
#include <mutex>

namespace mca {

int pdl_level() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MCA_PDL");
    v = (e != nullptr && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1;
  }
  return v;
}


int num_sms() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cached = n > 0 ? n : 148;
  }
  return cached;
}

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t row_stride,
                      uint32_t box_inner, uint32_t box_outer) {
  auto fn = encode_fn();
  if (fn == nullptr) return MCA_ERR_CUDA;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (row_stride * 2) % 16 != 0) return MCA_ERR_ALIGN;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MCA_OK : MCA_ERR_CUDA;
}

int make_tmap_2d_f32(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t row_stride,
                     uint32_t box_inner, uint32_t box_outer) {
  auto fn = encode_fn();
  if (fn == nullptr) return MCA_ERR_CUDA;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (row_stride * 4) % 16 != 0) return MCA_ERR_ALIGN;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride * 4};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MCA_OK : MCA_ERR_CUDA;
}

int make_tmap(CUtensorMap* out, int elem_bytes, const void* ptr, int rank, const uint64_t* dims,
              const uint64_t* strides_elems, const uint32_t* box, int swizzle_bytes) {
  auto fn = encode_fn();
  if (fn == nullptr) return MCA_ERR_CUDA;
  if (rank < 1 || rank > 3 || (elem_bytes != 2 && elem_bytes != 4)) return MCA_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return MCA_ERR_ALIGN;
  cuuint64_t d[3] = {1, 1, 1};
  cuuint64_t st[2] = {0, 0};
  cuuint32_t bx[3] = {1, 1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) d[i] = dims[i], bx[i] = box[i];
  for (int i = 0; i + 1 < rank; ++i) {
    st[i] = strides_elems[i] * static_cast<uint64_t>(elem_bytes);
    if (st[i] % 16 != 0) return MCA_ERR_ALIGN;
  }
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                      : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank,
                  const_cast<void*>(ptr), d, st, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MCA_OK : MCA_ERR_CUDA;
}

}  // namespace mca

extern "C" int mca_version(void) { return 100; }

extern "C" int mca_gemm_effective_splits(int K, int k_splits) { return mca::gemm_effective_splits(K, k_splits); }
