// Warp-specialised tcgen05 GEMM building blocks (sm_100a):
//   warp 0  : TMA producer   (one elected lane)  global -> 128B-swizzled smem ring, mbarrier complete_tx
//   warp 1  : MMA issuer     (one elected lane)  tcgen05.mma smem x smem -> fp32 accumulators in TMEM
//   warp 2  : TMEM allocator
//   warps 4+: epilogue       tcgen05.ld TMEM -> registers -> fused epilogue -> HBM
// D[128 x BLOCK_N] = sum_k A[m,k] * B[n,k]; accumulators are double-buffered in TMEM so the epilogue of tile i
// overlaps the MMAs of tile i+1.  Two arithmetic modes:
//   PREC_BF16  : bf16 operands, 1 MMA per k-step (kind::f16)
//   PREC_TF32X3: fp32 operands pre-split into (hi, lo) tf32 planes, 3 MMAs per k-step
//                (lo*hi + hi*lo + hi*hi; kind::tf32) -> fp32-class accuracy on the tensor pipe.
// Operands may be K-major (rows x 128 B tiles) or MN-major (k-rows x 128 B chunks); see umma.cuh.
#pragma once
#include "umma.cuh"

namespace coper {
namespace umma {

constexpr int PREC_BF16 = COPER_PREC_BF16, PREC_TF32X3 = COPER_PREC_TF32X3;
constexpr int BLOCK_M = 128;
constexpr int NUM_NON_EPI_THREADS = 128;

template <int PREC_, int BLOCK_N_, int STAGES_, int EPI_WARPS_, bool A_MN_, bool B_MN_>
struct GemmCfg {
  static constexpr int PREC = PREC_, BLOCK_N = BLOCK_N_, STAGES = STAGES_, EPI_WARPS = EPI_WARPS_;
  static constexpr bool A_MN = A_MN_, B_MN = B_MN_;
  static constexpr int ELEM = (PREC == PREC_BF16) ? 2 : 4;
  static constexpr int CHUNK = 128 / ELEM;          // elements per 128-byte swizzle row: 64 bf16 / 32 tf32
  static constexpr int BLOCK_K = CHUNK;             // k elements per pipeline stage
  static constexpr int UMMA_K = 32 / ELEM;          // 16 bf16 / 8 tf32
  static constexpr int TERMS = (PREC == PREC_TF32X3) ? 2 : 1;   // operand planes (hi, lo)
  static constexpr int A_BYTES = BLOCK_M * 128;     // one plane of the A tile (BLOCK_M x BLOCK_K or BLOCK_K x BLOCK_M)
  static constexpr int B_BYTES = BLOCK_N * 128;
  static constexpr int STAGE_BYTES = TERMS * (A_BYTES + B_BYTES);
  static constexpr int TMEM_COLS = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128
                                   : (2 * BLOCK_N <= 256) ? 256 : 512;
  static constexpr int THREADS = NUM_NON_EPI_THREADS + EPI_WARPS * 32;
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr uint32_t FMT = (PREC == PREC_BF16) ? FMT_BF16 : FMT_TF32;
  static constexpr uint32_t IDESC = make_idesc(FMT, BLOCK_M, BLOCK_N, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
  static_assert(BLOCK_N % 16 == 0 && BLOCK_N >= 16 && BLOCK_N <= 256, "UMMA N");
  static_assert(2 * BLOCK_N <= 512, "TMEM double buffer");
  static_assert(!A_MN || (BLOCK_M % CHUNK == 0), "MN-major A chunking");
  static_assert(!B_MN || (BLOCK_N % CHUNK == 0), "MN-major B chunking");
};

template <class Cfg>
struct SmemLayout {
  uint8_t* base;        // 1024-aligned
  uint64_t* full;       // [STAGES]
  uint64_t* empty;      // [STAGES]
  uint64_t* tfull;      // [2]
  uint64_t* tempty;     // [2]
  uint32_t* tmem_ptr;
  __device__ __forceinline__ explicit SmemLayout(uint8_t* raw) {
    base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)Cfg::STAGES * Cfg::STAGE_BYTES);
    full = bars;
    empty = bars + Cfg::STAGES;
    tfull = bars + 2 * Cfg::STAGES;
    tempty = tfull + 2;
    tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);
  }
  __device__ __forceinline__ uint8_t* a_plane(int stage, int term) const {
    return base + (size_t)stage * Cfg::STAGE_BYTES + term * Cfg::A_BYTES;
  }
  __device__ __forceinline__ uint8_t* b_plane(int stage, int term) const {
    return base + (size_t)stage * Cfg::STAGE_BYTES + Cfg::TERMS * Cfg::A_BYTES + term * Cfg::B_BYTES;
  }
};

// One-time CTA setup: barriers + TMEM allocation.  Returns the TMEM base address.
template <class Cfg>
__device__ __forceinline__ uint32_t cta_setup(const SmemLayout<Cfg>& sm) {
  int warp = threadIdx.x >> 5;
  if (warp == 1 && (threadIdx.x & 31) == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) { mbar_init(&sm.full[i], 1); mbar_init(&sm.empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sm.tfull[i], 1); mbar_init(&sm.tempty[i], Cfg::EPI_WARPS); }
    fence_barrier_init();
  } else if (warp == 2) {
    tmem_alloc(sm.tmem_ptr, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *sm.tmem_ptr;
}
template <class Cfg>
__device__ __forceinline__ void cta_teardown(uint32_t tmem_base) {
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// Producer: fill one stage.  (a_mn0, b_mn0) = first row (M / N index) of the tiles, k0 = first k element.
// K-major operand: tensor map over [rows = MN, cols = K], one box {BLOCK_K, rows}.
// MN-major operand: tensor map over [rows = K, cols = MN], boxes {CHUNK, BLOCK_K} per 128-byte MN chunk.
template <class Cfg>
__device__ __forceinline__ void produce_stage(const SmemLayout<Cfg>& sm, int stage, const CUtensorMap* tA,
                                              const CUtensorMap* tAlo, const CUtensorMap* tB, const CUtensorMap* tBlo,
                                              int a_mn0, int b_mn0, int a_k0, int b_k0) {
  uint64_t* bar = &sm.full[stage];
  mbar_expect_tx(bar, Cfg::STAGE_BYTES);
#pragma unroll
  for (int t = 0; t < Cfg::TERMS; ++t) {
    const CUtensorMap* ma = t ? tAlo : tA;
    const CUtensorMap* mb = t ? tBlo : tB;
    if (!Cfg::A_MN) {
      tma_load_2d(sm.a_plane(stage, t), ma, bar, a_k0, a_mn0);
    } else {
#pragma unroll
      for (int c = 0; c < BLOCK_M / Cfg::CHUNK; ++c)
        tma_load_2d(sm.a_plane(stage, t) + c * (Cfg::BLOCK_K * 128), ma, bar, a_mn0 + c * Cfg::CHUNK, a_k0);
    }
    if (!Cfg::B_MN) {
      tma_load_2d(sm.b_plane(stage, t), mb, bar, b_k0, b_mn0);
    } else {
#pragma unroll
      for (int c = 0; c < Cfg::BLOCK_N / Cfg::CHUNK; ++c)
        tma_load_2d(sm.b_plane(stage, t) + c * (Cfg::BLOCK_K * 128), mb, bar, b_mn0 + c * Cfg::CHUNK, b_k0);
    }
  }
}

// MMA issuer: consume one stage (kvalid = number of valid k elements in it, <= BLOCK_K).
// `first` = this is the first stage of the accumulation (overwrite instead of accumulate).
template <class Cfg>
__device__ __forceinline__ void issue_stage(const SmemLayout<Cfg>& sm, int stage, uint32_t tmem_d, int kvalid,
                                            bool first) {
  int nk = (kvalid + Cfg::UMMA_K - 1) / Cfg::UMMA_K;
  uint32_t a_addr[2], b_addr[2];
#pragma unroll
  for (int t = 0; t < Cfg::TERMS; ++t) {
    a_addr[t] = smem_u32(sm.a_plane(stage, t));
    b_addr[t] = smem_u32(sm.b_plane(stage, t));
  }
  for (int k = 0; k < nk; ++k) {
    // K-major: step 32 bytes inside the 128-byte swizzle row; MN-major: step UMMA_K k-rows of 128 bytes
    uint32_t a_off = Cfg::A_MN ? k * Cfg::UMMA_K * 128 : k * 32;
    uint32_t b_off = Cfg::B_MN ? k * Cfg::UMMA_K * 128 : k * 32;
    uint64_t da[2], db[2];
#pragma unroll
    for (int t = 0; t < Cfg::TERMS; ++t) {
      da[t] = Cfg::A_MN ? make_desc(a_addr[t] + a_off, Cfg::BLOCK_K * 128, 1024) : make_desc_kmajor(a_addr[t] + a_off);
      db[t] = Cfg::B_MN ? make_desc(b_addr[t] + b_off, Cfg::BLOCK_K * 128, 1024) : make_desc_kmajor(b_addr[t] + b_off);
    }
    uint32_t acc = (first && k == 0) ? 0u : 1u;
    if (Cfg::PREC == PREC_BF16) {
      mma_bf16(tmem_d, da[0], db[0], Cfg::IDESC, acc);
    } else {
      mma_tf32(tmem_d, da[1], db[0], Cfg::IDESC, acc);   // lo * hi
      mma_tf32(tmem_d, da[0], db[1], Cfg::IDESC, 1u);    // hi * lo
      mma_tf32(tmem_d, da[0], db[0], Cfg::IDESC, 1u);    // hi * hi
    }
  }
}

struct PipeState {
  int stage = 0;
  uint32_t phase = 0;
  template <int STAGES>
  __device__ __forceinline__ void advance() {
    if (++stage == STAGES) { stage = 0; phase ^= 1; }
  }
};

}  // namespace umma
}  // namespace coper
