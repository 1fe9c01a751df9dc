// CUDA-core fp32 tile engine (COPER_PREC_FP32): 128x128x16 CTA tiles, 256 threads, 8x8 register
// micro-tiles, double-buffered shared memory with register-staged global prefetch.  Operands are
// described by small "source" functors so the same main loop serves the 1-N scorer, its two gradient
// contractions and the Khatri-Rao (context (x) feature) contractions of the fused CPG-FC kernels —
// the per-query weights only ever exist as operand tiles in shared memory.
// This is the exact-fp32 parity path; the tcgen05/TMEM engine (umma_*.cu) is the throughput path.
#pragma once
#include "common.cuh"

namespace coper {
namespace simt {

constexpr int BM = 128, BN = 128, BK = 16, THREADS = 256;
constexpr int LDS = BM + 4;  // padded smem row stride (floats); keeps float4 alignment

struct __align__(16) Smem {
  float a[2][BK][LDS];
  float b[2][BK][LDS];
};

// A source exposes:
//   static constexpr bool kKContig  — consecutive k contiguous in memory (K-major) or consecutive mn (MN-major)
//   __device__ void tile(int t)      — select K-tile t
//   __device__ float at(int mn, int kk) const — element (absolute row/col mn, kk in [0,BK)) of the selected tile
struct SrcK {  // element (mn, k) at p[mn*ld + k]
  static constexpr bool kKContig = true;
  const float* p;
  int64_t ld, mn_ext, k_ext;
  int64_t k0 = 0;
  __device__ __forceinline__ void tile(int t) { k0 = (int64_t)t * BK; }
  __device__ __forceinline__ float at(int mn, int kk) const {
    int64_t k = k0 + kk;
    return (mn < mn_ext && k < k_ext) ? __ldg(p + (int64_t)mn * ld + k) : 0.f;
  }
};
struct SrcMN {  // element (mn, k) at p[k*ld + mn]
  static constexpr bool kKContig = false;
  const float* p;
  int64_t ld, mn_ext, k_ext;
  int64_t k0 = 0;
  __device__ __forceinline__ void tile(int t) { k0 = (int64_t)t * BK; }
  __device__ __forceinline__ float at(int mn, int kk) const {
    int64_t k = k0 + kk;
    return (mn < mn_ext && k < k_ext) ? __ldg(p + k * ld + mn) : 0.f;
  }
};

template <class S>
__device__ __forceinline__ void fetch8(S& src, int t, int mn0, float (&r)[8]) {
  src.tile(t);
  int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int e = tid + i * THREADS;
    int mn, kk;
    if (S::kKContig) { kk = e & (BK - 1); mn = e >> 4; } else { mn = e & (BM - 1); kk = e >> 7; }
    r[i] = src.at(mn0 + mn, kk);
  }
}
template <bool KCONTIG>
__device__ __forceinline__ void stash8(float (*dst)[LDS], const float (&r)[8]) {
  int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int e = tid + i * THREADS;
    int mn, kk;
    if (KCONTIG) { kk = e & (BK - 1); mn = e >> 4; } else { mn = e & (BM - 1); kk = e >> 7; }
    dst[kk][mn] = r[i];
  }
}

// rows / cols owned by this thread's 8x8 micro-tile (two 4-wide strips 64 apart -> conflict-free float4 reads)
__device__ __forceinline__ int mt_row(int i) { int ty = threadIdx.x >> 4; return (i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4); }
__device__ __forceinline__ int mt_col(int j) { int tx = threadIdx.x & 15; return (j < 4) ? tx * 4 + j : 64 + tx * 4 + (j - 4); }

__device__ __forceinline__ void zero_acc(float (&acc)[8][8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
}

// acc += A[m0:m0+128, tiles t0..t1) . B[n0:n0+128, same tiles]^T.  All threads of the CTA must call.
// Leaves the CTA synchronised (safe to reuse sm afterwards).
template <class SA, class SB>
__device__ __forceinline__ void mainloop(SA& A, SB& B, int m0, int n0, int t0, int t1, float (&acc)[8][8], Smem& sm) {
  if (t1 <= t0) return;
  float ra[8], rb[8];
  int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  fetch8(A, t0, m0, ra);
  fetch8(B, t0, n0, rb);
  stash8<SA::kKContig>(sm.a[0], ra);
  stash8<SB::kKContig>(sm.b[0], rb);
  __syncthreads();
  for (int t = t0; t < t1; ++t) {
    int buf = (t - t0) & 1;
    bool more = (t + 1 < t1);
    if (more) {
      fetch8(A, t + 1, m0, ra);
      fetch8(B, t + 1, n0, rb);
    }
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&sm.a[buf][kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&sm.a[buf][kk][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&sm.b[buf][kk][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&sm.b[buf][kk][64 + tx * 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) {
      stash8<SA::kKContig>(sm.a[buf ^ 1], ra);
      stash8<SB::kKContig>(sm.b[buf ^ 1], rb);
    }
    __syncthreads();
  }
}

}  // namespace simt
}  // namespace coper
