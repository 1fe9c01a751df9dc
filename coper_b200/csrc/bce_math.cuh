// Label-smoothed sigmoid cross-entropy of one 32-logit chunk held in registers (models.py:448-453; the stable form of
// tf.nn.sigmoid_cross_entropy_with_logits: max(s,0) - s z' + log1p(exp(-|s|))) together with its gradient
// (sigmoid(s) - z') * inv_count.  Shared by the tcgen05 scorer epilogues (umma_entity.cu, umma_fused.cu).
//
// The dense variant is the hot loop of the 1-N scorer: B*N logits per step, each with only d multiply-adds of tensor
// work behind it, so the epilogue is budgeted in issue slots and SFU (MUFU) operations per logit:
//   * 2 MUFU per logit (ex2 for e = exp(-|s|), rcp for 1/(1+e)) instead of 3: the log1p term is not evaluated per
//     element - the 32 factors u = 1 + e (each in [1, 2]) are MULTIPLIED on the FMA pipe and one lg2 per chunk is taken
//     of the product (<= 2^32, no overflow; the 32 roundings add <= 32 * 2^-24 relative error to the product, i.e.
//     ~2e-6 absolute on a chunk sum of ~22),
//   * sum max(s,0) = (sum s + sum |s|) / 2: two adds, the |.| is an operand modifier,
//   * the mean-reduction factor and the negative label are folded into one FFMA: g = sig * ic - neg * ic.
#pragma once
#include <stdint.h>

namespace coper {

__device__ __forceinline__ float mufu_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// every element is a negative (label z' = neg) and all 32 columns exist
__device__ __forceinline__ void bce_chunk_dense(const uint32_t (&r)[32], float bias, float neg, float ic, float (&g)[32],
                                                float& lsum, float& gsum) {
  float ssum = 0.f, asum = 0.f, prod = 1.f, gs = 0.f;
  const float c0 = -neg * ic;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float s = __uint_as_float(r[j]) + bias;
    const float a = fabsf(s);
    const float e = mufu_ex2(a * -1.4426950408889634f);
    const float u = 1.0f + e;
    const float rc = mufu_rcp(u);
    prod *= u;
    ssum += s;
    asum += a;
    const float sig = rc * ((s >= 0.f) ? 1.0f : e);      // relative accuracy is kept for sigmoid -> 0 (trained regime)
    const float gj = fmaf(sig, ic, c0);
    g[j] = gj;
    gs += gj;
  }
  gsum = gs;
  lsum = fmaf(mufu_lg2(prod), 0.6931471805599453f, 0.5f * (ssum + asum)) - neg * ssum;
}

// general chunk: label bits w (bit j set = positive, label pos) and column-valid mask vm (bit j clear = the column
// does not exist: g = 0, no loss)
__device__ __forceinline__ void bce_chunk_general(const uint32_t (&r)[32], float bias, uint32_t w, uint32_t vm, float pos,
                                                  float neg, float ic, float (&g)[32], float& lsum, float& gsum) {
  float ssum = 0.f, psum = 0.f, msum = 0.f, lgsum = 0.f, gs = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float s = __uint_as_float(r[j]) + bias;
    s = ((vm >> j) & 1u) ? s : 0.f;                    // non-existent column: s = 0 (its ln 2 is removed below)
    const float e = mufu_ex2(-fabsf(s) * 1.4426950408889634f);
    const float u = 1.0f + e;
    const float rc = mufu_rcp(u);
    const float lg = mufu_lg2(u);
    const float sig = rc * ((s >= 0.f) ? 1.0f : e);
    ssum += s;
    msum += fmaxf(s, 0.f);
    lgsum += lg;
    const bool on = (w >> j) & 1u;
    psum += on ? s : 0.f;
    const float gj = ((vm >> j) & 1u) ? (sig - (on ? pos : neg)) * ic : 0.f;
    g[j] = gj;
    gs += gj;
  }
  gsum = gs;
  // sum of max(s,0) - s z + ln2 * lg2(1+e)   with z = neg + [positive] (pos - neg)
  lgsum -= (float)(32 - __popc(vm));                   // lg2(1 + e^0) = 1 for every masked column
  lsum = fmaf(lgsum, 0.6931471805599453f, msum) - neg * ssum - (pos - neg) * psum;
}

}  // namespace coper
