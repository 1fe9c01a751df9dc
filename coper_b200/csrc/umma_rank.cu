// a7 + a11 fused on the tensor pipe: filtered-rank counts (metrics.py:44-51) taken straight from the 1-N scorer's
// TMEM accumulators (models.py:433-437).  The [B, N] logits are never written: the epilogue adds the bias, tests the
// 1-bit filter word of its (query, 32-entity chunk) and counts  s > gold  /  s == gold  in integer registers.
//
// The gold logit must be the SAME number the scorer produces for (b, e2[b]), otherwise comparisons against entities
// whose score is within rounding distance of the gold one could flip.  It is therefore produced by the same tcgen05
// instruction sequence: the gold rows E[e2[b]] are gathered (in operand form) into a [B, d] matrix, q . Eg^T runs
// through the identical pipeline configuration, and the diagonal (b, b) is extracted.  Ranks are bit-identical to
// scoring into HBM and running coper_filtered_rank on the stored logits (tests/test_gpu_umma.py).
#include "umma_gemm.cuh"

namespace coper {
using namespace umma;

size_t tc_prepared_bytes(int64_t rows, int cols, int prec);                     // umma_score.cu
int64_t tc_prepared_ld(int cols, int prec);
TcOperand tc_operand(const void* prep, int64_t rows, int cols, int prec);       // umma_gemm.cu

// rows e2[b] - ent_lo of a prepared operand [Ns, ldp] -> [B, ldp] (zero rows when not owned); 16-byte vectors
__global__ void gather_prepared_kernel(const uint4* __restrict__ src, int64_t Ns, int vec_per_row, int planes,
                                       const int64_t* __restrict__ e2, int64_t ent_lo, int B, uint4* __restrict__ dst) {
  int b = blockIdx.x;
  int64_t l = e2[b] - ent_lo;
  bool owned = l >= 0 && l < Ns;
  for (int pl = 0; pl < planes; ++pl) {
    const uint4* s = src + ((int64_t)pl * Ns + (owned ? l : 0)) * vec_per_row;
    uint4* o = dst + ((int64_t)pl * B + b) * vec_per_row;
    for (int v = threadIdx.x; v < vec_per_row; v += blockDim.x) o[v] = owned ? __ldg(s + v) : make_uint4(0, 0, 0, 0);
  }
}

struct DiagEpi : EpiBase {
  const float* bias;        // [Ns]
  const int64_t* e2;
  int64_t ent_lo, Ns;
  float* gold;              // [B]
  __device__ __forceinline__ void chunk(const GemmProblem& p, const TileCoord&, int row, int col,
                                        const uint32_t (&r)[32], int) const {
    if (row >= p.M || row < col || row >= col + 32) return;
    int64_t l = e2[row] - ent_lo;
    float v = 0.f;
    if (l >= 0 && l < Ns) {
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) acc = (col + j == row) ? __uint_as_float(r[j]) : acc;
      v = acc + __ldg(bias + l);
    }
    gold[row] = v;
  }
};

template <int NCH>
struct RankEpi : EpiBase {
  const float* bias;        // [N]
  const uint32_t* filt;     // [M, words]
  int64_t words;
  const int64_t* e2;
  int64_t ent_lo;
  const float* gold;        // [M]
  int32_t* n_greater;       // [M]  (accumulated)
  int32_t* n_equal;
  // thread-local state
  int cur_row, cg, ce, gl;
  float g;
  uint32_t w[NCH];
  __device__ __forceinline__ void flush() {
    if (cur_row >= 0) {
      if (cg) atomicAdd(n_greater + cur_row, cg);
      if (ce) atomicAdd(n_equal + cur_row, ce);
    }
    cg = ce = 0;
  }
  __device__ __forceinline__ void tile_begin(const GemmProblem& p, const TileCoord&, int row, int col0) {
    if (row != cur_row) {
      flush();
      cur_row = row < p.M ? row : -1;
      if (row < p.M) {
        g = __ldg(gold + row);
        int64_t l = e2[row] - ent_lo;
        gl = (l >= 0 && l < p.N) ? (int)l : -1;
      }
    }
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      int col = col0 + i * 32;
      uint32_t x = 0xFFFFFFFFu;                             // everything filtered: padded rows / chunks beyond N
      if (row < p.M && col < p.N) {
        x = __ldg(filt + (int64_t)row * words + (col >> 5));
        if (col + 32 > p.N) x |= 0xFFFFFFFFu << (p.N - col); // tail entities do not exist
      }
      w[i] = x;
    }
  }
  __device__ __forceinline__ void chunk(const GemmProblem& p, const TileCoord&, int, int col, const uint32_t (&r)[32],
                                        int ci) {
    uint32_t x = w[ci];
    if (gl >= col && gl < col + 32) x |= 1u << (gl - col);  // the gold entity itself never counts
    const bool full = col + 32 <= p.N;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float s = __uint_as_float(r[j]) + ((full || col + j < p.N) ? __ldg(bias + col + j) : 0.f);
      const bool valid = !((x >> j) & 1u);
      cg += (valid && s > g) ? 1 : 0;
      ce += (valid && s == g) ? 1 : 0;
    }
  }
  __device__ __forceinline__ void finish(int, int) { flush(); }
};

constexpr int kRankEpiWarps = 8;
template <int PREC>
using RankCfg = GemmCfg<PREC, 128, (PREC == PREC_BF16 ? 6 : 3), kRankEpiWarps, false, false, 0>;
constexpr int kRankNCH = 128 / (kRankEpiWarps / 4) / 32;

size_t umma_rank_workspace_bytes(int B, int d, int prec) { return align_up(tc_prepared_bytes(B, d, prec), 256); }

template <int PREC>
static int gold_impl(const TcOperand& Q, const TcOperand& Eg, const float* bias, int B, int64_t Ns, int d,
                     const int64_t* e2, int64_t ent_lo, float* gold, cudaStream_t st) {
  GemmProblem p{};
  p.M = B; p.N = B; p.K = d; p.groups = 1;
  plan_gemm<RankCfg<PREC>>(p, false);
  DiagEpi epi;
  epi.bias = bias; epi.e2 = e2; epi.ent_lo = ent_lo; epi.Ns = Ns; epi.gold = gold;
  return launch_gemm<RankCfg<PREC>, DiagEpi>(Q, Eg, p, epi, st);
}

int umma_score1n_gold(const void* q_prep, const void* E_prep, const float* bias, int B, int64_t Ns, int d,
                      const int64_t* e2, int64_t ent_lo, float* gold, void* ws, size_t ws_bytes, int prec,
                      cudaStream_t st) {
  if (!ws || ws_bytes < umma_rank_workspace_bytes(B, d, prec) || (reinterpret_cast<uintptr_t>(ws) & 255))
    return COPER_ERR_WORKSPACE;
  int64_t ldp = tc_prepared_ld(d, prec);
  int elem = prec == COPER_PREC_BF16 ? 2 : 4;
  int vec_per_row = (int)(ldp * elem / 16);
  gather_prepared_kernel<<<B, 64, 0, st>>>(static_cast<const uint4*>(E_prep), Ns, vec_per_row,
                                           prec == COPER_PREC_BF16 ? 1 : 2, e2, ent_lo, B, static_cast<uint4*>(ws));
  int rc = check_launch();
  if (rc) return rc;
  TcOperand Q = tc_operand(q_prep, B, d, prec), Eg = tc_operand(ws, B, d, prec);
  if (prec == COPER_PREC_BF16) return gold_impl<PREC_BF16>(Q, Eg, bias, B, Ns, d, e2, ent_lo, gold, st);
  if (prec == COPER_PREC_TF32X3) return gold_impl<PREC_TF32X3>(Q, Eg, bias, B, Ns, d, e2, ent_lo, gold, st);
  return COPER_ERR_UNSUPPORTED;
}

template <int PREC>
static int rank_impl(const TcOperand& Q, const TcOperand& E, const float* bias, int B, int64_t Ns, int d,
                     const int64_t* e2, int64_t ent_lo, const float* gold, const uint32_t* filt, int32_t* n_greater,
                     int32_t* n_equal, cudaStream_t st) {
  GemmProblem p{};
  p.M = B; p.N = (int)Ns; p.K = d; p.groups = 1;
  plan_gemm<RankCfg<PREC>>(p, false);
  RankEpi<kRankNCH> epi;
  epi.bias = bias; epi.filt = filt; epi.words = (Ns + 31) / 32; epi.e2 = e2; epi.ent_lo = ent_lo; epi.gold = gold;
  epi.n_greater = n_greater; epi.n_equal = n_equal;
  epi.cur_row = -2; epi.cg = 0; epi.ce = 0; epi.gl = -1; epi.g = 0.f;
  // a grid that is a multiple of the number of row tiles keeps every CTA on ONE row tile (tiles are enumerated
  // row-tile fastest), so the per-thread counters are flushed once, at the end
  int grid = 148;
  if (p.m_tiles <= 148) grid = 148 / p.m_tiles * p.m_tiles;
  return launch_gemm<RankCfg<PREC>, RankEpi<kRankNCH>>(Q, E, p, epi, st, grid);
}

int umma_score1n_rank(const void* q_prep, const void* E_prep, const float* bias, int B, int64_t Ns, int d,
                      const int64_t* e2, int64_t ent_lo, const float* gold, const uint32_t* filt, int32_t* n_greater,
                      int32_t* n_equal, int prec, cudaStream_t st) {
  if (Ns > 0x7fffffff - 512) return COPER_ERR_UNSUPPORTED;
  TcOperand Q = tc_operand(q_prep, B, d, prec), E = tc_operand(E_prep, Ns, d, prec);
  if (prec == COPER_PREC_BF16)
    return rank_impl<PREC_BF16>(Q, E, bias, B, Ns, d, e2, ent_lo, gold, filt, n_greater, n_equal, st);
  if (prec == COPER_PREC_TF32X3)
    return rank_impl<PREC_TF32X3>(Q, E, bias, B, Ns, d, e2, ent_lo, gold, filt, n_greater, n_equal, st);
  return COPER_ERR_UNSUPPORTED;
}

}  // namespace coper

using namespace coper;

extern "C" {

size_t coper_score1n_rank_workspace_bytes(int B, int d, int prec) {
  if (prec != COPER_PREC_BF16 && prec != COPER_PREC_TF32X3) return 0;
  return umma_rank_workspace_bytes(B, d, prec);
}
int coper_score1n_gold_prepared(const void* q_prep, const void* E_prep, const float* bias, int B, int64_t Ns, int d,
                                const int64_t* e2, int64_t ent_lo, float* gold, void* workspace,
                                size_t workspace_bytes, int prec, coper_stream_t stream) {
  COPER_CHECK_ARG(q_prep && E_prep && bias && e2 && gold && B > 0 && Ns > 0 && d > 0);
  return umma_score1n_gold(q_prep, E_prep, bias, B, Ns, d, e2, ent_lo, gold, workspace, workspace_bytes, prec,
                           as_stream(stream));
}
int coper_score1n_rank_prepared(const void* q_prep, const void* E_prep, const float* bias, int B, int64_t Ns, int d,
                                const int64_t* e2, int64_t ent_lo, const float* gold, const uint32_t* filter_bits,
                                int32_t* n_greater, int32_t* n_equal, int prec, coper_stream_t stream) {
  COPER_CHECK_ARG(q_prep && E_prep && bias && e2 && gold && filter_bits && n_greater && n_equal);
  COPER_CHECK_ARG(B > 0 && Ns > 0 && d > 0);
  return umma_score1n_rank(q_prep, E_prep, bias, B, Ns, d, e2, ent_lo, gold, filter_bits, n_greater, n_equal, prec,
                           as_stream(stream));
}

}  // extern "C"
