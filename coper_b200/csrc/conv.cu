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
  pdl_enter();
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
  pdl_enter();
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

// ------------------------------------------------------------------------------------------------------------------
// Fast path for the configuration every shipped config uses: 3x3 filters, 32 channels, shared filters.
// Thread (c = lane, hs = warp): output row h is produced by sliding a 3x3 register window along w, so an output costs
// 3 shared-memory loads + 9 FMAs and no integer division; for a fixed (h, w) the 32 lanes write 32 consecutive
// channels (128 B).  IMG images per CTA amortise the filter load and cut the number of gradient partial slabs.
constexpr int kFastC = 32;
template <int IMG>
__global__ void __launch_bounds__(256) conv3x3_fwd_kernel(const float* __restrict__ x0, int B, int H, int W,
                                                          const float* __restrict__ wc, const float* __restrict__ bc,
                                                          float* __restrict__ z) {
  pdl_enter();
  extern __shared__ float sm[];                    // IMG * H * W
  const int c = threadIdx.x & 31, hs = threadIdx.x >> 5;
  const int OH = H - 2, OW = W - 2, HW = H * W;
  const int b0 = blockIdx.x * IMG;
  for (int i = threadIdx.x; i < IMG * HW; i += 256) {
    int b = b0 + i / HW;
    sm[i] = b < B ? x0[(int64_t)b0 * HW + i] : 0.f;
  }
  float wr[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) wr[k] = __ldg(wc + k * kFastC + c);
  const float bias = __ldg(bc + c);
  __syncthreads();
  for (int q = 0; q < IMG; ++q) {
    const int b = b0 + q;
    if (b >= B) break;
    const float* img = sm + q * HW;
    float* zo = z + (int64_t)b * OH * OW * kFastC + c;
    for (int h = hs; h < OH; h += 8) {
      const float* r0 = img + h * W;
      float a0 = r0[0], a1 = r0[1], b0_ = r0[W], b1 = r0[W + 1], c0 = r0[2 * W], c1 = r0[2 * W + 1];
      for (int w = 0; w < OW; ++w) {
        const float a2 = r0[w + 2], b2 = r0[W + w + 2], c2 = r0[2 * W + w + 2];
        float acc = bias;
        acc = fmaf(a0, wr[0], acc); acc = fmaf(a1, wr[1], acc); acc = fmaf(a2, wr[2], acc);
        acc = fmaf(b0_, wr[3], acc); acc = fmaf(b1, wr[4], acc); acc = fmaf(b2, wr[5], acc);
        acc = fmaf(c0, wr[6], acc); acc = fmaf(c1, wr[7], acc); acc = fmaf(c2, wr[8], acc);
        zo[(h * OW + w) * kFastC] = acc;
        a0 = a1; a1 = a2; b0_ = b1; b1 = b2; c0 = c1; c1 = c2;
      }
    }
  }
}

// backward fast path.  Per image: dz staged in shared memory; phase 1 (thread = (c, h-slice)) accumulates the 9 filter
// taps + bias gradient in registers with the same sliding window; phase 2 (warp per pixel, lanes = channels) forms dx.
// The per-CTA filter / bias partials (IMG images, 8 h-slices) are combined in shared memory in a fixed order:
// dwc_part / dbc_part get ONE slab per CTA.
// BN = true: `dz` is the gradient w.r.t. the OUTPUT of the batch-norm / relu / dropout block that follows the conv
// (models.py:386-391) and `bn.z` the conv output: the block's backward (coper_bn_act_bwd_apply, keep_pre = 1) is applied
// while the tile is staged in shared memory - its [B, OH*OW*C] result is never written to HBM.
struct ConvBnBwd {
  const float* z;
  const float* a;
  const float* b;
  const float* mean;
  const float* invstd;
  const float* c1;
  const float* c2;
  int relu;
  float keep, inv_keep;
  uint32_t thr;
  const uint64_t* seed_dev;
  uint64_t salt;
};
template <int IMG, bool BN>
__global__ void __launch_bounds__(256) conv3x3_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ x0,
                                                          int B, int H, int W, const float* __restrict__ wc,
                                                          float* __restrict__ dx0, float* __restrict__ dwc_part,
                                                          float* __restrict__ dbc_part, ConvBnBwd bn) {
  pdl_enter();
  extern __shared__ float sm[];
  __shared__ float bnp[BN ? 6 : 1][kFastC];
  const int c = threadIdx.x & 31, hs = threadIdx.x >> 5, lane = c;
  const int OH = H - 2, OW = W - 2, HW = H * W, total = OH * OW * kFastC;
  float* img = sm;                  // HW
  float* dzs = img + HW;            // total
  float* red = dzs + total;         // 8 * 10 * 32
  const int b0 = blockIdx.x * IMG;
  uint64_t seed = 0;
  if (BN) {
    seed = (bn.seed_dev ? *bn.seed_dev : 0ull) + bn.salt;
    if (threadIdx.x < kFastC) {
      const int ch = threadIdx.x;
      bnp[0][ch] = bn.a[ch]; bnp[1][ch] = bn.b[ch]; bnp[2][ch] = bn.mean[ch];
      bnp[3][ch] = bn.invstd[ch]; bnp[4][ch] = bn.c1[ch]; bnp[5][ch] = bn.c2[ch];
    }
  }
  float wr[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) wr[k] = __ldg(wc + k * kFastC + c);
  float g[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) g[k] = 0.f;
  for (int q = 0; q < IMG; ++q) {
    const int b = b0 + q;
    if (b >= B) break;
    __syncthreads();
    for (int i = threadIdx.x; i < HW; i += 256) img[i] = x0[(int64_t)b * HW + i];
    if (!BN) {
      const float4* src = reinterpret_cast<const float4*>(dz + (int64_t)b * total);
      float4* dst = reinterpret_cast<float4*>(dzs);
      for (int i = threadIdx.x; i < total / 4; i += 256) dst[i] = __ldg(src + i);
    } else {
      const float4* src = reinterpret_cast<const float4*>(dz + (int64_t)b * total);
      const float4* zsrc = reinterpret_cast<const float4*>(bn.z + (int64_t)b * total);
      float4* dst = reinterpret_cast<float4*>(dzs);
      for (int i = threadIdx.x; i < total / 4; i += 256) {
        const float4 g4 = __ldg(src + i), z4 = __ldg(zsrc + i);
        const int ch = (i * 4) & (kFastC - 1);
        const uint64_t e = (uint64_t)b * (uint64_t)total + (uint64_t)i * 4u;
        const float gs[4] = {g4.x, g4.y, g4.z, g4.w}, zs[4] = {z4.x, z4.y, z4.z, z4.w};
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          o[k] = bn_bwd_dx(gs[k], drop_factor(bn.keep, bn.inv_keep, bn.thr, seed, e + k), zs[k], bnp[0][ch + k],
                           bnp[1][ch + k], bnp[2][ch + k], bnp[3][ch + k], bnp[4][ch + k], bnp[5][ch + k], bn.relu);
        dst[i] = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
    __syncthreads();
    // phase 1: filter / bias gradient, thread (c, h = hs, hs + 8, ...)
    for (int h = hs; h < OH; h += 8) {
      const float* r0 = img + h * W;
      float a0 = r0[0], a1 = r0[1], b0_ = r0[W], b1 = r0[W + 1], c0 = r0[2 * W], c1 = r0[2 * W + 1];
      for (int w = 0; w < OW; ++w) {
        const float a2 = r0[w + 2], b2 = r0[W + w + 2], c2 = r0[2 * W + w + 2];
        const float d = dzs[(h * OW + w) * kFastC + c];
        g[0] = fmaf(a0, d, g[0]); g[1] = fmaf(a1, d, g[1]); g[2] = fmaf(a2, d, g[2]);
        g[3] = fmaf(b0_, d, g[3]); g[4] = fmaf(b1, d, g[4]); g[5] = fmaf(b2, d, g[5]);
        g[6] = fmaf(c0, d, g[6]); g[7] = fmaf(c1, d, g[7]); g[8] = fmaf(c2, d, g[8]);
        g[9] += d;
        a0 = a1; a1 = a2; b0_ = b1; b1 = b2; c0 = c1; c1 = c2;
      }
    }
    // phase 2: dx, one warp per pixel, lanes over channels, fixed-order shuffle reduction
    for (int y = 0; y < H; ++y) {
      for (int x = hs; x < W; x += 8) {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int h = y - i;
          if (h < 0 || h >= OH) continue;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const int w = x - j;
            if (w < 0 || w >= OW) continue;
            acc = fmaf(dzs[(h * OW + w) * kFastC + lane], wr[i * 3 + j], acc);
          }
        }
        acc = warp_sum(acc);
        if (lane == 0) dx0[(int64_t)b * HW + y * W + x] = acc;
      }
    }
  }
  // combine the 8 h-slices (fixed order) -> one slab per CTA: dwc [9][32], dbc [32]
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 10; ++k) red[(hs * 10 + k) * kFastC + c] = g[k];
  __syncthreads();
  for (int o = threadIdx.x; o < 10 * kFastC; o += 256) {
    float t = 0.f;
#pragma unroll
    for (int s8 = 0; s8 < 8; ++s8) t += red[s8 * 10 * kFastC + o];
    if (o < 9 * kFastC) dwc_part[(int64_t)blockIdx.x * 9 * kFastC + o] = t;
    else dbc_part[(int64_t)blockIdx.x * kFastC + (o - 9 * kFastC)] = t;
  }
}
constexpr int kConvImgPerCta = 2;      // forward: 256 CTAs at B = 512
constexpr int kConvImgPerCtaBwd = 1;   // backward: every image its own CTA (3-4 resident per SM), one partial slab each
}  // namespace coper

using namespace coper;

extern "C" {

int coper_conv_fwd(const float* x0, int B, int H, int W, const float* wc, const float* bc, int KH, int KW, int C,
                   int per_query, float* z, coper_stream_t stream) {
  COPER_CHECK_ARG(x0 && wc && bc && z && B > 0 && H >= KH && W >= KW && KH > 0 && KW > 0 && C > 0);
  if (H * W > kMaxImg || KH * KW * C > kMaxFilt) return COPER_ERR_UNSUPPORTED;
  if (KH == 3 && KW == 3 && C == kFastC && !per_query && H >= 3 && W >= 3) {
    size_t sm_fast = (size_t)kConvImgPerCta * H * W * sizeof(float);
    launch_pdl(conv3x3_fwd_kernel<kConvImgPerCta>, (B + kConvImgPerCta - 1) / kConvImgPerCta, 256, sm_fast,
               as_stream(stream), x0, B, H, W, wc, bc, z);
    return check_launch();
  }
  size_t smem = (size_t)(H * W + KH * KW * C + C) * sizeof(float);
  launch_pdl(conv_fwd_kernel, B, kConvThreads, smem, as_stream(stream), x0, H, W, wc, bc, KH, KW, C, per_query, z);
  return check_launch();
}

int coper_conv_bwd_slabs(int B, int H, int W, int KH, int KW, int C, int per_query) {
  int OH = H - KH + 1, OW = W - KW + 1;
  size_t sm_fast = (size_t)(H * W + OH * OW * C + 8 * 10 * kFastC) * sizeof(float);
  if (KH == 3 && KW == 3 && C == kFastC && !per_query && H >= 3 && W >= 3 && sm_fast <= 47 * 1024 &&
      (OH * OW * C) % 4 == 0)
    return (B + kConvImgPerCtaBwd - 1) / kConvImgPerCtaBwd;
  return B;
}

int coper_conv_bwd(const float* dz, const float* x0, int B, int H, int W, const float* wc, int KH, int KW, int C,
                   int per_query, float* dx0, float* dwc_part, float* dbc_part, coper_stream_t stream) {
  COPER_CHECK_ARG(dz && x0 && wc && dx0 && dwc_part && dbc_part && B > 0 && H >= KH && W >= KW && C > 0);
  int OH = H - KH + 1, OW = W - KW + 1;
  size_t smem = (size_t)(H * W + KH * KW * C + OH * OW * C) * sizeof(float);
  if (H * W > kMaxImg || KH * KW * C > kMaxFilt || smem > 200 * 1024) return COPER_ERR_UNSUPPORTED;
  if (KH == 3 && KW == 3 && C == kFastC && !per_query && H >= 3 && W >= 3) {
    size_t sm_fast = (size_t)(H * W + OH * OW * C + 8 * 10 * kFastC) * sizeof(float);
    if (sm_fast <= 47 * 1024 && (OH * OW * C) % 4 == 0) {
      launch_pdl(conv3x3_bwd_kernel<kConvImgPerCtaBwd, false>, (B + kConvImgPerCtaBwd - 1) / kConvImgPerCtaBwd, 256,
                 sm_fast, as_stream(stream), dz, x0, B, H, W, wc, dx0, dwc_part, dbc_part, ConvBnBwd{});
      return check_launch();
    }
  }
  static bool attr_set_dev[64] = {};            // the attribute is per device, not per process
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return COPER_ERR_CUDA;
  bool& attr_set = attr_set_dev[dev];
  if (smem > 48 * 1024 && !attr_set) {
    int rc = check_cuda(cudaFuncSetAttribute(conv_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    if (rc) return rc;
    attr_set = true;
  }
  launch_pdl(conv_bwd_kernel, B, kConvThreads, smem, as_stream(stream), dz, x0, H, W, wc, KH, KW, C, per_query, dx0,
             dwc_part, dbc_part);
  return check_launch();
}

/* coper_bn_act_bwd_apply (keep_pre = 1) + coper_conv_bwd in one call: dout [B, OH*OW*C] is the gradient w.r.t. the
 * output of the batch-norm / relu / dropout block, z the conv output.  3x3 x 32-channel shared filters: one kernel, the
 * block's input gradient never exists in HBM; any other shape: the two kernels through dz_scratch [B, OH*OW*C]. */
int coper_conv_bwd_bn(const float* dout, const float* z, const float* x0, int B, int H, int W, const float* wc, int KH,
                      int KW, int C, int per_query, const float* a, const float* b, const float* mean,
                      const float* invstd, const float* c1, const float* c2, int relu, float keep_post,
                      const uint64_t* seed_dev, uint64_t salt_post, float* dx0, float* dwc_part, float* dbc_part,
                      float* dz_scratch, coper_stream_t stream) {
  COPER_CHECK_ARG(dout && z && x0 && wc && dx0 && dwc_part && dbc_part && dz_scratch && B > 0 && H >= KH && W >= KW && C > 0);
  COPER_CHECK_ARG(a && b && mean && invstd && c1 && c2 && keep_post > 0.f);
  const int OH = H - KH + 1, OW = W - KW + 1;
  if (KH == 3 && KW == 3 && C == kFastC && !per_query && H >= 3 && W >= 3 && H * W <= kMaxImg) {
    size_t sm_fast = (size_t)(H * W + OH * OW * C + 8 * 10 * kFastC) * sizeof(float);
    if (sm_fast <= 47 * 1024 && (OH * OW * C) % 4 == 0 &&
        ((reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(z)) & 15) == 0) {
      ConvBnBwd bn{z, a, b, mean, invstd, c1, c2, relu, keep_post, 1.0f / keep_post, keep_threshold(keep_post), seed_dev,
                   salt_post};
      launch_pdl(conv3x3_bwd_kernel<kConvImgPerCtaBwd, true>, (B + kConvImgPerCtaBwd - 1) / kConvImgPerCtaBwd, 256,
                 sm_fast, as_stream(stream), dout, x0, B, H, W, wc, dx0, dwc_part, dbc_part, bn);
      return check_launch();
    }
  }
  int rc = coper_bn_act_bwd_apply(dout, z, (int64_t)B * OH * OW, C, a, b, mean, invstd, c1, c2, relu, keep_post, seed_dev,
                                  salt_post, 1.0f, 0, dz_scratch, stream);
  if (rc) return rc;
  return coper_conv_bwd(dz_scratch, x0, B, H, W, wc, KH, KW, C, per_query, dx0, dwc_part, dbc_part, stream);
}

}  // extern "C"
