// Host side of the tcgen05 engine: tensor-map (TMA descriptor) encoding through the driver entry point
// (resolved at run time with cudaGetDriverEntryPoint, so the library has no link-time libcuda dependency
// and still loads on a GPU-less build box).
#include <mutex>
#include "umma.cuh"

namespace coper {
namespace umma {

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, bool is_bf16, uint64_t rows, uint64_t cols,
                 uint64_t pitch_elems, uint32_t box_cols, uint32_t box_rows, bool atom32) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return COPER_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((pitch_elems * elem_bytes) & 15) || box_rows > 256 ||
      box_cols * elem_bytes > 128)
    return COPER_ERR_INVALID_ARG;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {pitch_elems * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    g_last_cuda_error = (int)r;
    return COPER_ERR_CUDA;
  }
  return COPER_OK;
}

}  // namespace umma
}  // namespace coper
