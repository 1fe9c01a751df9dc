// a11 / K10: filtered rank (metrics.py:44-51) as a counting kernel.  The reference masks known-true
// tails to -inf, restores the gold logit and takes 1 + position in a full argsort; that equals
//     rank = 1 + #{ n != gold : not filtered, s[n] > s[gold] }
// whenever no unfiltered score ties with the gold one (n_equal reports ties; np.argsort's tie order is
// unspecified).  HBM-bound: each query streams 4 B/entity of scores (float4, coalesced) plus 1 bit/entity
// of filter mask (one 128-bit load covers 128 entities); integer counts -> bit-exact for any sharding.
#include "common.cuh"

namespace coper {

constexpr int kRankThreads = 256;
constexpr int kRankChunk = 256 * 4 * 8;  // entities per CTA: 8 float4 per thread

__global__ void gold_scores_kernel(const float* __restrict__ scores, int64_t ld, int B, int64_t Ns,
                                   const int64_t* __restrict__ e2, int64_t ent_lo, float* __restrict__ gold) {
  pdl_enter();
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int64_t l = e2[b] - ent_lo;
  gold[b] = (l >= 0 && l < Ns) ? scores[(int64_t)b * ld + l] : 0.f;
}

__global__ void __launch_bounds__(kRankThreads) filtered_rank_kernel(
    const float* __restrict__ scores, int64_t ld, int B, int64_t Ns, const int64_t* __restrict__ e2, int64_t ent_lo,
    const float* __restrict__ gold, const uint32_t* __restrict__ filt, int32_t* __restrict__ n_greater,
    int32_t* __restrict__ n_equal) {
  pdl_enter();
  __shared__ int sm_g[32], sm_e[32];
  int b = blockIdx.y;
  int64_t c0 = (int64_t)blockIdx.x * kRankChunk;
  int64_t c1 = c0 + kRankChunk < Ns ? c0 + kRankChunk : Ns;
  const float* row = scores + (int64_t)b * ld;
  int64_t words = (Ns + 31) / 32;
  const uint32_t* frow = filt + (int64_t)b * words;
  float g = gold[b];
  int64_t gl = e2[b] - ent_lo;
  int cg = 0, ce = 0;
  bool vec_ok = ((reinterpret_cast<uintptr_t>(row) & 15) == 0);
  // each thread handles groups of 4 consecutive entities; a warp covers 128 entities = 4 filter words
  for (int64_t n = c0 + (int64_t)threadIdx.x * 4; n < c1; n += (int64_t)kRankThreads * 4) {
    uint32_t w = __ldg(frow + (n >> 5));
    uint32_t fb = (w >> (n & 31)) & 0xFu;  // n % 4 == 0 so the 4 bits never straddle a word
    float v[4];
    if (vec_ok && n + 3 < Ns) {
      float4 t = __ldg(reinterpret_cast<const float4*>(row + n));
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = (n + k < Ns) ? __ldg(row + n + k) : -INFINITY;
      if (n + 3 >= Ns) fb |= (0xFu << (Ns - n)) & 0xFu;  // mask the tail
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      bool valid = !((fb >> k) & 1u) && (n + k != gl);
      cg += (valid && v[k] > g) ? 1 : 0;
      ce += (valid && v[k] == g) ? 1 : 0;
    }
  }
  cg = warp_sum_i(cg);
  ce = warp_sum_i(ce);
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { sm_g[wid] = cg; sm_e[wid] = ce; }
  __syncthreads();
  if (wid == 0) {
    int nw = kRankThreads / 32;
    cg = lane < nw ? sm_g[lane] : 0;
    ce = lane < nw ? sm_e[lane] : 0;
    cg = warp_sum_i(cg);
    ce = warp_sum_i(ce);
    if (lane == 0) {
      if (cg) atomicAdd(n_greater + b, cg);
      if (ce) atomicAdd(n_equal + b, ce);
    }
  }
}
}  // namespace coper

using namespace coper;

extern "C" {

int coper_gold_scores(const float* scores, int64_t ld, int B, int64_t Ns, const int64_t* e2, int64_t ent_lo,
                      float* gold, coper_stream_t stream) {
  COPER_CHECK_ARG(scores && e2 && gold && B > 0 && Ns > 0 && ld >= Ns);
  launch_pdl(gold_scores_kernel, ceil_div(B, 128), 128, 0, as_stream(stream), scores, ld, B, Ns, e2, ent_lo, gold);
  return check_launch();
}

int coper_filtered_rank(const float* scores, int64_t ld, int B, int64_t Ns, const int64_t* e2, int64_t ent_lo,
                        const float* gold, const uint32_t* filter_bits, int32_t* n_greater, int32_t* n_equal,
                        coper_stream_t stream) {
  COPER_CHECK_ARG(scores && e2 && gold && filter_bits && n_greater && n_equal && B > 0 && Ns > 0 && ld >= Ns);
  COPER_CHECK_ARG(B <= 65535);
  dim3 grid(ceil_div(Ns, kRankChunk), B);
  launch_pdl(filtered_rank_kernel, grid, kRankThreads, 0, as_stream(stream), scores, ld, B, Ns, e2, ent_lo, gold,
             filter_bits, n_greater, n_equal);
  return check_launch();
}

}  // extern "C"
