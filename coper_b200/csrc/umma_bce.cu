// a7 + a9 + a10 (K6-K8) on the tensor pipe: 1-N scorer + label-smoothed sigmoid-BCE + its gradient
// (models.py:433-437, 448-453, 198) for COPER_PREC_BF16 / COPER_PREC_TF32X3.
//
//   pass 1  S = q.E^T on tcgen05 (TMA-fed, fp32 accumulators in TMEM); the TMEM->register epilogue adds the
//           bias, evaluates the stable BCE and dL/dS = (sigmoid(s) - z') / (B*N) per element and emits
//             * G in tensor-pipe operand form (bf16, or tf32 hi/lo planes) - the logits are never written,
//             * per-warp loss partials (fp64) and per-32-row dbias partials (warp transposing reduction).
//   pass 2  dq = G.E   (K = entities: split-K over the SMs, fixed-order slab reduction -> deterministic)
//   pass 3  dE = G^T.q (M = entities, K = batch), written once.
// Labels are 1-bit rows (coper_csr_to_bits): one 32-bit word per (query, 32-entity chunk).
#include "umma_gemm.cuh"

namespace coper {
using namespace umma;

size_t tc_prepared_bytes(int64_t rows, int cols, int prec);                     // umma_score.cu
int tc_prepare(const float* src, int64_t rows, int cols, int64_t ld_src, int prec, void* dst, cudaStream_t st);
int64_t tc_prepared_ld(int cols, int prec);
TcOperand tc_operand(const void* prep, int64_t rows, int cols, int prec);       // umma_gemm.cu
int tc_gemm_store(int prec, bool a_mn, bool b_mn, const TcOperand& A, const TcOperand& B, const GemmProblem& p,
                  bool split, const StoreEpi& epi, cudaStream_t st);
int tc_plan_splits(int prec, GemmProblem p, bool split);

// lane l ends up with sum over the warp's 32 lanes of v[l] (31 shuffles, fixed order)
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      float send = upper ? v[i] : v[i + off];
      float keep = upper ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

template <int PREC>
struct BceEpi : EpiBase {
  const float* bias;          // [N]
  const uint32_t* bits;       // [M, words]
  int64_t words;
  float pos, neg, inv_count;
  void* G;                    // bf16 [M, ldG]  |  fp32 hi plane [M, ldG] followed by the lo plane
  int64_t ldG;
  float* dbias_part;          // [ceil(M/32), N]
  double* loss_part;          // [grid * epi_warps]
  double loss_acc;            // per-thread running sum (kernel-parameter copy -> thread-local)

  __device__ __forceinline__ void chunk(const GemmProblem& p, const TileCoord&, int row, int col,
                                        const uint32_t (&r)[32], int) {
    const int lane = threadIdx.x & 31;
    const bool rowok = row < p.M;
    const uint32_t w = rowok ? __ldg(bits + (int64_t)row * words + (col >> 5)) : 0u;
    float g[32];
    float lsum = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int n = col + j;
      const bool nok = n < p.N;                           // warp-uniform
      const float s = __uint_as_float(r[j]) + (nok ? __ldg(bias + n) : 0.f);
      const float z = ((w >> j) & 1u) ? pos : neg;
      // tf.nn.sigmoid_cross_entropy_with_logits: max(s,0) - s z + log1p(exp(-|s|))
      const float e = __expf(-fabsf(s));
      const float rcp = __fdividef(1.0f, 1.0f + e);
      const float l = fmaxf(s, 0.f) - s * z + __logf(1.0f + e);
      const float sig = (s >= 0.f) ? rcp : e * rcp;
      const bool ok = rowok && nok;
      g[j] = ok ? (sig - z) * inv_count : 0.f;
      lsum += ok ? l : 0.f;
    }
    loss_acc += (double)lsum;
    if (rowok) {
      if (PREC == PREC_BF16) {
        __nv_bfloat16* o = static_cast<__nv_bfloat16*>(G) + (int64_t)row * ldG + col;   // 64-byte aligned
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 u;
          __nv_bfloat162 t0 = __floats2bfloat162_rn(g[j], g[j + 1]);
          __nv_bfloat162 t1 = __floats2bfloat162_rn(g[j + 2], g[j + 3]);
          __nv_bfloat162 t2 = __floats2bfloat162_rn(g[j + 4], g[j + 5]);
          __nv_bfloat162 t3 = __floats2bfloat162_rn(g[j + 6], g[j + 7]);
          u.x = *reinterpret_cast<uint32_t*>(&t0);
          u.y = *reinterpret_cast<uint32_t*>(&t1);
          u.z = *reinterpret_cast<uint32_t*>(&t2);
          u.w = *reinterpret_cast<uint32_t*>(&t3);
          *reinterpret_cast<uint4*>(o + j) = u;
        }
      } else {
        float* hi = static_cast<float*>(G) + (int64_t)row * ldG + col;
        float* lo = hi + (int64_t)p.M * ldG;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float h[4], l4[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint32_t hb;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(g[j + k]));
            h[k] = __uint_as_float(hb);
            l4[k] = g[j + k] - h[k];
          }
          *reinterpret_cast<float4*>(hi + j) = make_float4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<float4*>(lo + j) = make_float4(l4[0], l4[1], l4[2], l4[3]);
        }
      }
    }
    // dbias partial of this 32-row block: lane l <- sum over rows of column col+l
    float cs = warp_transpose_sum(g, lane);
    if (col + lane < p.N && (row - lane) < p.M) dbias_part[(int64_t)(row >> 5) * p.N + col + lane] = cs;
  }
  __device__ __forceinline__ void finish(int epi_thread, int epi_threads) {
    double t = warp_sum_d(loss_acc);
    if ((epi_thread & 31) == 0) loss_part[(int64_t)blockIdx.x * (epi_threads >> 5) + (epi_thread >> 5)] = t;
  }
};

__global__ void sum_doubles_kernel(const double* __restrict__ in, int n, double* out) {
  __shared__ double smd[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += in[i];
  double t = block_sum<double>(acc, smd);
  if (threadIdx.x == 0) *out = t;
}

constexpr int kBceEpiWarps = 16;
using BceCfgBf16 = GemmCfg<PREC_BF16, 256, 4, kBceEpiWarps, false, false>;
using BceCfgTf32 = GemmCfg<PREC_TF32X3, 128, 3, kBceEpiWarps, false, false>;

struct BceTcLayout {
  size_t off_q, off_E, off_dq, off_dbias, off_loss, total;
  int splits, row_blocks;
};
static BceTcLayout bce_tc_layout(int B, int64_t Ns, int d, int prec) {
  BceTcLayout L;
  GemmProblem p{};
  p.M = B; p.N = d; p.K = (int)Ns; p.groups = 1;
  L.splits = tc_plan_splits(prec, p, true);
  L.row_blocks = (B + 31) / 32;
  size_t o = 0;
  L.off_q = o; o = align_up(o + tc_prepared_bytes(B, d, prec), 256);
  L.off_E = o; o = align_up(o + tc_prepared_bytes(Ns, d, prec), 256);
  L.off_dq = o; o = align_up(o + (size_t)L.splits * B * d * sizeof(float), 256);
  L.off_dbias = o; o = align_up(o + (size_t)L.row_blocks * Ns * sizeof(float), 256);
  L.off_loss = o; o = align_up(o + (size_t)148 * kBceEpiWarps * sizeof(double), 256);
  L.total = o;
  return L;
}

size_t umma_bce_workspace_bytes(int B, int64_t Ns, int d, int prec) { return bce_tc_layout(B, Ns, d, prec).total; }
size_t umma_bce_G_bytes(int B, int64_t Ns, int prec) {
  int64_t ld = (Ns + 31) / 32 * 32;
  return (size_t)B * ld * (prec == COPER_PREC_BF16 ? 2 : 8);
}

template <class Cfg>
static int launch_bce(const TcOperand& Q, const TcOperand& E, int B, int64_t Ns, int d, BceEpi<Cfg::PREC> epi,
                      cudaStream_t st, int* grid_out) {
  GemmProblem p{};
  p.M = B; p.N = (int)Ns; p.K = d; p.groups = 1; p.groups_inner = 0;
  plan_gemm<Cfg>(p, false);
  long long supers = (long long)p.m_tiles * p.n_tiles;
  *grid_out = (int)(supers < 148 ? supers : 148);
  return launch_gemm<Cfg, BceEpi<Cfg::PREC>>(Q, E, p, epi, st);
}

int umma_score1n_bce_fwd_bwd(const float* q, const float* E, const void* E_prepared, const float* bias,
                             const uint32_t* label_bits, int B,
                             int64_t Ns, int d, float pos, float neg, float inv_count, double* loss_sum, void* G,
                             int64_t ldG, float* dq, float* dE, float* dbias, void* ws, size_t ws_bytes, int prec,
                             cudaStream_t st) {
  if (Ns > 0x7fffffff - 512) return COPER_ERR_UNSUPPORTED;
  if (ldG % 32 != 0 || ldG < Ns || (reinterpret_cast<uintptr_t>(G) & 127)) return COPER_ERR_INVALID_ARG;
  BceTcLayout L = bce_tc_layout(B, Ns, d, prec);
  if (!ws || ws_bytes < L.total) return COPER_ERR_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(ws) & 255) return COPER_ERR_INVALID_ARG;
  char* w = static_cast<char*>(ws);
  void* qp = w + L.off_q;
  const void* Ep = E_prepared ? E_prepared : w + L.off_E;
  float* dq_part = reinterpret_cast<float*>(w + L.off_dq);
  float* dbias_part = reinterpret_cast<float*>(w + L.off_dbias);
  double* loss_part = reinterpret_cast<double*>(w + L.off_loss);
  int rc;
  if ((rc = tc_prepare(q, B, d, d, prec, qp, st))) return rc;
  if (!E_prepared && (rc = tc_prepare(E, Ns, d, d, prec, w + L.off_E, st))) return rc;
  TcOperand Qo = tc_operand(qp, B, d, prec), Eo = tc_operand(Ep, Ns, d, prec);
  // ---- pass 1: scores -> loss, G, dbias partials
  int grid = 0;
  if (prec == COPER_PREC_BF16) {
    BceEpi<PREC_BF16> epi;
    epi.bias = bias; epi.bits = label_bits; epi.words = (Ns + 31) / 32; epi.pos = pos; epi.neg = neg;
    epi.inv_count = inv_count; epi.G = G; epi.ldG = ldG; epi.dbias_part = dbias_part; epi.loss_part = loss_part;
    epi.loss_acc = 0.0;
    rc = launch_bce<BceCfgBf16>(Qo, Eo, B, Ns, d, epi, st, &grid);
  } else {
    BceEpi<PREC_TF32X3> epi;
    epi.bias = bias; epi.bits = label_bits; epi.words = (Ns + 31) / 32; epi.pos = pos; epi.neg = neg;
    epi.inv_count = inv_count; epi.G = G; epi.ldG = ldG; epi.dbias_part = dbias_part; epi.loss_part = loss_part;
    epi.loss_acc = 0.0;
    rc = launch_bce<BceCfgTf32>(Qo, Eo, B, Ns, d, epi, st, &grid);
  }
  if (rc) return rc;
  sum_doubles_kernel<<<1, 256, 0, st>>>(loss_part, grid * kBceEpiWarps, loss_sum);
  if ((rc = check_launch())) return rc;
  if ((rc = coper_reduce_partials(dbias_part, L.row_blocks, Ns, 1.0f, 0, dbias, (coper_stream_t)st))) return rc;
  // ---- G as a tensor-pipe operand: stored [B, Ns] with pitch ldG
  TcOperand Go;
  Go.main = G;
  Go.lo = prec == COPER_PREC_TF32X3 ? static_cast<const void*>(static_cast<const float*>(G) + (int64_t)B * ldG) : nullptr;
  Go.rows = (uint64_t)B; Go.cols = (uint64_t)Ns; Go.pitch = (uint64_t)ldG;
  // ---- pass 2: dq = G . E   (A = G K-major, B = E stored [K, N] -> MN-major), split-K slabs
  {
    GemmProblem p{};
    p.M = B; p.N = d; p.K = (int)Ns; p.groups = 1; p.groups_inner = 0;
    StoreEpi epi = make_store_epi(dq_part, d, 0, (long long)B * d);
    if ((rc = tc_gemm_store(prec, false, true, Go, Eo, p, true, epi, st))) return rc;
    if ((rc = coper_reduce_partials(dq_part, L.splits, (int64_t)B * d, 1.0f, 0, dq, (coper_stream_t)st))) return rc;
  }
  // ---- pass 3: dE = G^T . q (A = G stored [K, M] -> MN-major, B = q stored [K, N] -> MN-major)
  {
    GemmProblem p{};
    p.M = (int)Ns; p.N = d; p.K = B; p.groups = 1; p.groups_inner = 0;
    StoreEpi epi = make_store_epi(dE, d, 0, 0);
    if ((rc = tc_gemm_store(prec, true, true, Go, Qo, p, false, epi, st))) return rc;
  }
  return COPER_OK;
}

}  // namespace coper
