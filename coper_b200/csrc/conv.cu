// a5 / K4: the ConvE 3x3 VALID convolution over the 10 x (d/10) "image" of the head-entity embedding
// (models.py:355,373-385) and its backward.  1 input channel, K = KH*KW = 9 taps: far too thin for the
// tensor pipe, so this is a direct CUDA-core kernel: one CTA per query row, image + filters staged in
// shared memory, outputs written coalesced in the (h,w,c) order the fused CPG-FC kernel consumes
// (models.py:404).  < 0.2 % of the step's FLOPs; what matters is that it streams f / dz once.
#include "common.cuh"

namespace coper {

constexpr int kConvThreads = 256;
constexpr int kMaxImg = 2048;    // H*W floats
constexpr int kMaxFilt = 2048;   // KH*KW*C floats

__global__ void __launch_bounds__(kConvThreads) conv_fwd_kernel(const float* __restrict__ x0, int H, int W,
                                                                const float* __restrict__ wc,
                                                                const float* __restrict__ bc, int KH, int KW, int C,
                                                                int per_query, float* __restrict__ z) {
  extern __shared__ float sm[];
  float* img = sm;                 // H*W
  float* wf = sm + H * W;          // KH*KW*C
  float* bs = wf + KH * KW * C;    // C
  int b = blockIdx.x;
  int OH = H - KH + 1, OW = W - KW + 1;
  const float* wsrc = wc + (per_query ? (int64_t)b * KH * KW * C : 0);
  const float* bsrc = bc + (per_query ? (int64_t)b * C : 0);
  for (int i = threadIdx.x; i < H * W; i += blockDim.x) img[i] = x0[(int64_t)b * H * W + i];
  for (int i = threadIdx.x; i < KH * KW * C; i += blockDim.x) wf[i] = wsrc[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) bs[i] = bsrc[i];
  __syncthreads();
  int total = OH * OW * C;
  float* zo = z + (int64_t)b * total;
  for (int o = threadIdx.x; o < total; o += blockDim.x) {
    int c = o % C, hw = o / C;
    int h = hw / OW, w = hw % OW;
    float acc = 0.f;
    for (int i = 0; i < KH; ++i)
      for (int j = 0; j < KW; ++j) acc = fmaf(img[(h + i) * W + (w + j)], wf[(i * KW + j) * C + c], acc);
    zo[o] = acc + bs[c];
  }
}

// backward: one CTA per query row.
//   dx[y][x]     = sum_{i,j,c} dz[y-i][x-j][c] * wc[i][j][c]
//   dwc[i][j][c] = sum_{h,w}   img[h+i][w+j]   * dz[h][w][c]     (per-sample partial)
//   dbc[c]       = sum_{h,w}   dz[h][w][c]                        (per-sample partial)
__global__ void __launch_bounds__(kConvThreads) conv_bwd_kernel(const float* __restrict__ dz,
                                                                const float* __restrict__ x0, int H, int W,
                                                                const float* __restrict__ wc, int KH, int KW, int C,
                                                                int per_query, float* __restrict__ dx0,
                                                                float* __restrict__ dwc_part,
                                                                float* __restrict__ dbc_part) {
  extern __shared__ float sm[];
  int OH = H - KH + 1, OW = W - KW + 1;
  int total = OH * OW * C;
  float* img = sm;                   // H*W
  float* wf = img + H * W;           // KH*KW*C
  float* dzs = wf + KH * KW * C;     // OH*OW*C
  int b = blockIdx.x;
  const float* wsrc = wc + (per_query ? (int64_t)b * KH * KW * C : 0);
  for (int i = threadIdx.x; i < H * W; i += blockDim.x) img[i] = x0[(int64_t)b * H * W + i];
  for (int i = threadIdx.x; i < KH * KW * C; i += blockDim.x) wf[i] = wsrc[i];
  for (int i = threadIdx.x; i < total; i += blockDim.x) dzs[i] = dz[(int64_t)b * total + i];
  __syncthreads();
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  // dx: one warp per pixel, lanes over channels, fixed-order shuffle reduction
  for (int p = warp; p < H * W; p += nwarp) {
    int y = p / W, x = p % W;
    float acc = 0.f;
    for (int i = 0; i < KH; ++i) {
      int h = y - i;
      if (h < 0 || h >= OH) continue;
      for (int j = 0; j < KW; ++j) {
        int w = x - j;
        if (w < 0 || w >= OW) continue;
        for (int c = lane; c < C; c += 32) acc = fmaf(dzs[(h * OW + w) * C + c], wf[(i * KW + j) * C + c], acc);
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) dx0[(int64_t)b * H * W + p] = acc;
  }
  // dwc partial
  for (int t = threadIdx.x; t < KH * KW * C; t += blockDim.x) {
    int c = t % C, ij = t / C;
    int i = ij / KW, j = ij % KW;
    float acc = 0.f;
    for (int h = 0; h < OH; ++h)
      for (int w = 0; w < OW; ++w) acc = fmaf(img[(h + i) * W + (w + j)], dzs[(h * OW + w) * C + c], acc);
    dwc_part[(int64_t)b * KH * KW * C + t] = acc;
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int hw = 0; hw < OH * OW; ++hw) acc += dzs[hw * C + c];
    dbc_part[(int64_t)b * C + c] = acc;
  }
}
}  // namespace coper

using namespace coper;

extern "C" {

int coper_conv_fwd(const float* x0, int B, int H, int W, const float* wc, const float* bc, int KH, int KW, int C,
                   int per_query, float* z, coper_stream_t stream) {
  COPER_CHECK_ARG(x0 && wc && bc && z && B > 0 && H >= KH && W >= KW && KH > 0 && KW > 0 && C > 0);
  if (H * W > kMaxImg || KH * KW * C > kMaxFilt) return COPER_ERR_UNSUPPORTED;
  size_t smem = (size_t)(H * W + KH * KW * C + C) * sizeof(float);
  conv_fwd_kernel<<<B, kConvThreads, smem, as_stream(stream)>>>(x0, H, W, wc, bc, KH, KW, C, per_query, z);
  return check_launch();
}

int coper_conv_bwd(const float* dz, const float* x0, int B, int H, int W, const float* wc, int KH, int KW, int C,
                   int per_query, float* dx0, float* dwc_part, float* dbc_part, coper_stream_t stream) {
  COPER_CHECK_ARG(dz && x0 && wc && dx0 && dwc_part && dbc_part && B > 0 && H >= KH && W >= KW && C > 0);
  int OH = H - KH + 1, OW = W - KW + 1;
  size_t smem = (size_t)(H * W + KH * KW * C + OH * OW * C) * sizeof(float);
  if (H * W > kMaxImg || KH * KW * C > kMaxFilt || smem > 200 * 1024) return COPER_ERR_UNSUPPORTED;
  static bool attr_set = false;
  if (smem > 48 * 1024 && !attr_set) {
    int rc = check_cuda(cudaFuncSetAttribute(conv_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    if (rc) return rc;
    attr_set = true;
  }
  conv_bwd_kernel<<<B, kConvThreads, smem, as_stream(stream)>>>(dz, x0, H, W, wc, KH, KW, C, per_query, dx0, dwc_part,
                                                               dbc_part);
  return check_launch();
}

}  // extern "C"
