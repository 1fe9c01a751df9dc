// a10 scatter / K8(3): gradient of tf.nn.embedding_lookup (models.py:176-180, IndexedSlices made dense at
// models.py:198).  Duplicate head entities / relations in a batch make this a segmented reduction:
//   1. stable radix sort of (row id, position) pairs (cub::DeviceRadixSort — library plumbing),
//   2. one warp per segment head: lanes stride the embedding width (coalesced 128 B row reads), the warp
//      walks its segment in sorted (= original, the sort is stable) order and issues ONE read-modify-write
//      of the destination row.  No atomics, fixed summation order -> bit-reproducible.
// HBM-bound: M*w*4 B read + 2*rows_touched*w*4 B RMW + 8 M B of indices.
#include <cub/device/device_radix_sort.cuh>
#include "common.cuh"

namespace coper {

__global__ void iota_kernel(int32_t* p, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}

__global__ void segscatter_kernel(const int64_t* __restrict__ keys, const int32_t* __restrict__ pos, int M,
                                  const float* __restrict__ src, int width, float* __restrict__ dst, int64_t row_lo,
                                  int64_t row_hi) {
  int seg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // candidate head position
  int lane = threadIdx.x & 31;
  if (seg >= M) return;
  int64_t key = keys[seg];
  if (seg > 0 && keys[seg - 1] == key) return;  // not a segment head
  if (key < row_lo || key >= row_hi) return;    // row owned by another shard
  int end = seg + 1;
  while (end < M && keys[end] == key) ++end;
  float* drow = dst + (key - row_lo) * (int64_t)width;
  for (int c = lane; c < width; c += 32) {
    float acc = 0.f;
    for (int i = seg; i < end; ++i) acc += __ldg(src + (int64_t)pos[i] * width + c);
    drow[c] += acc;
  }
}

struct SegLayout {
  size_t off_keys_out, off_pos_in, off_pos_out, off_cub, cub_bytes, total;
};
static SegLayout seg_layout(int M) {
  SegLayout L;
  size_t o = 0;
  L.off_keys_out = o; o = align_up(o + (size_t)M * sizeof(int64_t), 256);
  L.off_pos_in = o; o = align_up(o + (size_t)M * sizeof(int32_t), 256);
  L.off_pos_out = o; o = align_up(o + (size_t)M * sizeof(int32_t), 256);
  size_t cb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cb, (const int64_t*)nullptr, (int64_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, M);
  L.cub_bytes = cb;
  L.off_cub = o; o = align_up(o + cb, 256);
  L.total = o;
  return L;
}
}  // namespace coper

using namespace coper;

extern "C" {

size_t coper_segscatter_workspace_bytes(int M) { return M <= 0 ? 256 : seg_layout(M).total; }

int coper_segscatter_add(const int64_t* idx, int M, const float* src, int width, float* dst, int64_t row_lo,
                         int64_t row_hi, void* workspace, size_t workspace_bytes, coper_stream_t stream) {
  COPER_CHECK_ARG(idx && src && dst && workspace && M >= 0 && width > 0 && row_hi >= row_lo);
  if (M == 0) return COPER_OK;
  SegLayout L = seg_layout(M);
  if (workspace_bytes < L.total) return COPER_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  char* ws = static_cast<char*>(workspace);
  int64_t* keys_out = reinterpret_cast<int64_t*>(ws + L.off_keys_out);
  int32_t* pos_in = reinterpret_cast<int32_t*>(ws + L.off_pos_in);
  int32_t* pos_out = reinterpret_cast<int32_t*>(ws + L.off_pos_out);
  iota_kernel<<<ceil_div(M, 256), 256, 0, st>>>(pos_in, M);
  int rc = check_launch();
  if (rc) return rc;
  size_t cb = L.cub_bytes;
  rc = check_cuda(cub::DeviceRadixSort::SortPairs(ws + L.off_cub, cb, idx, keys_out, pos_in, pos_out, M, 0, 64, st));
  if (rc) return rc;
  segscatter_kernel<<<ceil_div(M, 8), 256, 0, st>>>(keys_out, pos_out, M, src, width, dst, row_lo, row_hi);
  return check_launch();
}

}  // extern "C"
