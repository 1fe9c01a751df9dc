// a8 / SURVEY §8f-2: the SAMPLED-label scorer of the reference (models.py:438-443: tf.gather of [B, L] entity rows,
// batched mat-vec, + gathered bias), its label-smoothed sigmoid-BCE (models.py:448-453) and the gradients TF's
// autodiff produces for it — what every shipped big-dataset config trains with (num_labels = 100 / 1000).
//
// The path is gather-bound ([B, L, d] rows of the entity table: 410 MB at B=512, L=1000, d=200), not GEMM-shaped:
//   kernel 1 (one CTA per query): q[b] lives in registers, warps stride over the L sampled entities, each row is read
//     ONCE (coalesced 128-byte segments) and used for the logit (warp reduction), the loss, dL/ds and the running
//     dq[b] += g * E_row; per-warp dq accumulators are combined in shared memory in a fixed order.
//   kernel 2: the gradient of the two gathers is an IndexedSlices in TF (values g[b,l]*q[b] at row lookup[b,l]).
//     The optimizer's sparse rule (utils/amsgrad.py:161-189) and tf.clip_by_global_norm need, per entity row, the SUM
//     of its slices and the sum of their SQUARES: (lookup id, position) pairs are radix-sorted (cub) and one warp per
//     segment walks it in sorted order — no atomics, fixed summation order -> deterministic.
#include <cub/device/device_radix_sort.cuh>
#include "common.cuh"

namespace coper {

constexpr int kSampWarps = 8;
constexpr int kSampMaxK = 8;     // d <= 256: up to 8 strided elements per lane

__global__ void __launch_bounds__(kSampWarps * 32) sampled_score_bce_kernel(
    const float* __restrict__ q, const float* __restrict__ E, const float* __restrict__ bias,
    const int32_t* __restrict__ lookup, const float* __restrict__ labels, int L, int64_t N, int d, float one_minus_eps,
    float inv_num_ent, float inv_count, double* __restrict__ loss_part, float* __restrict__ scores,
    float* __restrict__ g_out, float* __restrict__ dq) {
  __shared__ float red[kSampWarps][256];
  __shared__ double lred[kSampWarps];
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float qr[kSampMaxK], acc[kSampMaxK];
#pragma unroll
  for (int k = 0; k < kSampMaxK; ++k) {
    const int c = lane + 32 * k;
    qr[k] = c < d ? __ldg(q + (int64_t)b * d + c) : 0.f;
    acc[k] = 0.f;
  }
  double lsum = 0.0;
  for (int l = warp; l < L; l += kSampWarps) {
    const int64_t p = (int64_t)b * L + l;
    const int64_t ent = lookup[p];
    const float* row = E + ent * d;
    float er[kSampMaxK];
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < kSampMaxK; ++k) {
      const int c = lane + 32 * k;
      er[k] = c < d ? __ldg(row + c) : 0.f;
      dot = fmaf(qr[k], er[k], dot);
    }
    dot = warp_sum(dot);
    const float s = dot + __ldg(bias + ent);
    const float z = one_minus_eps * __ldg(labels + p) + inv_num_ent;          // models.py:450
    const float e = expf(-fabsf(s));
    const float loss = fmaxf(s, 0.f) - s * z + log1pf(e);
    const float sig = (s >= 0.f) ? 1.0f / (1.0f + e) : e / (1.0f + e);
    const float g = (sig - z) * inv_count;
    if (lane == 0) {
      g_out[p] = g;
      if (scores) scores[p] = s;
      lsum += (double)loss;
    }
#pragma unroll
    for (int k = 0; k < kSampMaxK; ++k) acc[k] = fmaf(g, er[k], acc[k]);
  }
#pragma unroll
  for (int k = 0; k < kSampMaxK; ++k) red[warp][lane + 32 * k] = acc[k];
  if (lane == 0) lred[warp] = lsum;
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kSampWarps; ++w) t += red[w][c];
    dq[(int64_t)b * d + c] = t;
  }
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kSampWarps; ++w) t += lred[w];
    loss_part[b] = t;
  }
}

__global__ void sum_loss_kernel(const double* __restrict__ in, int n, double* out) {
  __shared__ double smd[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += in[i];
  double t = block_sum<double>(acc, smd);
  if (threadIdx.x == 0) *out = t;
}

__global__ void iota32_kernel(int32_t* p, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}

// one warp per sorted position; segment heads walk their segment: slice value = g[pos] * q[pos / L, :]
__global__ void sampled_scatter_kernel(const int32_t* __restrict__ keys, const int32_t* __restrict__ pos, int M, int L,
                                       const float* __restrict__ g, const float* __restrict__ q, int d,
                                       float* __restrict__ dE_sum, float* __restrict__ dE_sq,
                                       float* __restrict__ db_sum, float* __restrict__ db_sq) {
  const int seg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (seg >= M) return;
  const int32_t key = keys[seg];
  if (seg > 0 && keys[seg - 1] == key) return;   // not a segment head
  int end = seg + 1;
  while (end < M && keys[end] == key) ++end;
  float a1[kSampMaxK], a2[kSampMaxK];
#pragma unroll
  for (int k = 0; k < kSampMaxK; ++k) a1[k] = a2[k] = 0.f;
  float b1 = 0.f, b2 = 0.f;
  for (int i = seg; i < end; ++i) {
    const int p = pos[i];
    const float gv = __ldg(g + p);
    const float* qrow = q + (int64_t)(p / L) * d;
    b1 += gv;
    b2 = fmaf(gv, gv, b2);
#pragma unroll
    for (int k = 0; k < kSampMaxK; ++k) {
      const int c = lane + 32 * k;
      if (c < d) {
        const float v = gv * __ldg(qrow + c);
        a1[k] += v;
        a2[k] = fmaf(v, v, a2[k]);
      }
    }
  }
  float* r1 = dE_sum + (int64_t)key * d;
  float* r2 = dE_sq + (int64_t)key * d;
#pragma unroll
  for (int k = 0; k < kSampMaxK; ++k) {
    const int c = lane + 32 * k;
    if (c < d) {
      r1[c] += a1[k];
      r2[c] += a2[k];
    }
  }
  if (lane == 0) {
    db_sum[key] += b1;
    db_sq[key] += b2;
  }
}

struct SampLayout {
  size_t off_loss, off_keys_out, off_pos_in, off_pos_out, off_cub, cub_bytes, total;
};
static SampLayout samp_layout(int B, int L) {
  SampLayout S;
  const int M = B * L;
  size_t o = 0;
  S.off_loss = o; o = align_up(o + (size_t)B * sizeof(double), 256);
  S.off_keys_out = o; o = align_up(o + (size_t)M * sizeof(int32_t), 256);
  S.off_pos_in = o; o = align_up(o + (size_t)M * sizeof(int32_t), 256);
  S.off_pos_out = o; o = align_up(o + (size_t)M * sizeof(int32_t), 256);
  size_t cb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cb, (const int32_t*)nullptr, (int32_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, M);
  S.cub_bytes = cb;
  S.off_cub = o; o = align_up(o + cb, 256);
  S.total = o;
  return S;
}
}  // namespace coper

using namespace coper;

extern "C" {

size_t coper_score_sampled_workspace_bytes(int B, int L) { return (B <= 0 || L <= 0) ? 256 : samp_layout(B, L).total; }

int coper_score_sampled_bce_fwd_bwd(const float* q, const float* E, const float* bias, const int32_t* lookup,
                                    const float* labels, int B, int L, int64_t N, int d, float one_minus_eps,
                                    float inv_num_ent, float inv_count, double* loss_sum, float* scores, float* g,
                                    float* dq, float* dE_sum, float* dE_sq, float* dbias_sum, float* dbias_sq,
                                    void* workspace, size_t workspace_bytes, coper_stream_t stream) {
  COPER_CHECK_ARG(q && E && bias && lookup && labels && loss_sum && g && dq && dE_sum && dE_sq && dbias_sum && dbias_sq);
  COPER_CHECK_ARG(B > 0 && L > 0 && N > 0 && d > 0 && workspace);
  if (d > 32 * kSampMaxK) return COPER_ERR_UNSUPPORTED;
  SampLayout S = samp_layout(B, L);
  if (workspace_bytes < S.total) return COPER_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  char* ws = static_cast<char*>(workspace);
  double* loss_part = reinterpret_cast<double*>(ws + S.off_loss);
  int32_t* keys_out = reinterpret_cast<int32_t*>(ws + S.off_keys_out);
  int32_t* pos_in = reinterpret_cast<int32_t*>(ws + S.off_pos_in);
  int32_t* pos_out = reinterpret_cast<int32_t*>(ws + S.off_pos_out);
  sampled_score_bce_kernel<<<B, kSampWarps * 32, 0, st>>>(q, E, bias, lookup, labels, L, N, d, one_minus_eps,
                                                         inv_num_ent, inv_count, loss_part, scores, g, dq);
  int rc = check_launch();
  if (rc) return rc;
  sum_loss_kernel<<<1, 256, 0, st>>>(loss_part, B, loss_sum);
  if ((rc = check_launch())) return rc;
  const int M = B * L;
  iota32_kernel<<<ceil_div(M, 256), 256, 0, st>>>(pos_in, M);
  if ((rc = check_launch())) return rc;
  int end_bit = 1;
  while (end_bit < 31 && (int64_t(1) << end_bit) < N) ++end_bit;
  size_t cb = S.cub_bytes;
  rc = check_cuda(cub::DeviceRadixSort::SortPairs(ws + S.off_cub, cb, lookup, keys_out, pos_in, pos_out, M, 0, end_bit, st));
  if (rc) return rc;
  sampled_scatter_kernel<<<ceil_div(M, 8), 256, 0, st>>>(keys_out, pos_out, M, L, g, q, d, dE_sum, dE_sq, dbias_sum,
                                                         dbias_sq);
  return check_launch();
}

}  // extern "C"
