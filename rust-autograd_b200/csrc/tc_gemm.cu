// tc_gemm.cu — tcgen05 (5th-gen tensor core) TF32 / 3xTF32 GEMM for sm_100a.
//
//   D[lane, col] = sum_k P[lane, k] * Q[col, k]          (UMMA: A operand = P, B operand = Q, fp32 accumulators in TMEM)
//
// For C[m, n] = op(A)[m, k] . op(B)[k, n] (row-major C) the kernel puts n on the 128 TMEM lanes and m on the TMEM
// columns (P = op(B)^T, Q = op(A)), so that a `tcgen05.ld` of one accumulator column is 32 consecutive floats of a
// C row: the epilogue stores straight from registers with full 128-byte coalescing, no shared-memory staging.
// Operand tiles are fetched by TMA (128-byte swizzle, zero fill out of bounds => no edge special-casing in the
// main loop) into a multi-stage mbarrier ring; one elected thread issues `tcgen05.mma.kind::tf32` (K = 8 per
// instruction, 4 per 32-wide k-block); `tcgen05.commit` releases ring slots and signals the epilogue.
//
// Both operand majors are supported natively (no transposed copies): a K-major tile is one TMA box
// [rows x 32 k]; an MN-major tile is rows/32 boxes of [32 k x 32 rows] (UMMA MN-major SWIZZLE_128B canonical layout).
//
// 3xTF32 (f32-faithful) mode: four extra warps split every landed tile in place into hi = rna_tf32(x) and
// lo = rna_tf32(x - hi) (second smem buffer), then the issuing thread runs lo*hi + hi*lo + hi*hi into the same
// accumulator.  Dropped term lo*lo ~ 2^-22 relative.
//
// Replaces matrixmultiply::sgemm / cblas_sgemm behind MatMul and BatchMatMul
// (reference src/tensor_ops/dot_ops.rs:383-422, 142-380).
#include "tc_common.cuh"

#define TCG_THREADS 192      // warp 0: TMA producer, warp 1: TMEM alloc + MMA issuer, warps 2-5: splitter + epilogue
#define TCG_LANES 128

template <int TN, bool SPLIT> struct TcgCfg {
  static constexpr int P_BYTES = TCG_LANES * TC_BK * 4;     // 16 KB
  static constexpr int Q_BYTES = TN * TC_BK * 4;
  static constexpr int STAGE_BYTES = (P_BYTES + Q_BYTES) * (SPLIT ? 2 : 1);
  static constexpr int STAGES = SPLIT ? (STAGE_BYTES <= 64 * 1024 ? 3 : 2) : (STAGE_BYTES <= 32 * 1024 ? 3 : 4);
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

template <int TN, bool P_MN, bool Q_MN, bool SPLIT>
__global__ void __launch_bounds__(TCG_THREADS) tc_gemm_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmQ,
                                                              float* __restrict__ C, int NL, int NC, int K, int64_t ldc, int64_t bsc, int accumulate) {
  using Cfg = TcgCfg<TN, SPLIT>;
  constexpr int S = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + S * Cfg::STAGE_BYTES);
  uint64_t* full = bars; uint64_t* ready = bars + S; uint64_t* empty = bars + 2 * S; uint64_t* tmem_full = bars + 3 * S;
  uint32_t* tmem_slot = (uint32_t*)(bars + 3 * S + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lane0 = blockIdx.x * TCG_LANES, col0 = blockIdx.y * TN, bz = blockIdx.z;
  const int nk = (K + TC_BK - 1) / TC_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmP); tma_prefetch_desc(&tmQ);
    for (int s = 0; s < S; s++) { mbar_init(&full[s], 1); mbar_init(&ready[s], 128); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, TN); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int kb = 0; kb < nk; kb++) {
        const int s = kb % S; const uint32_t ph = (kb / S) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* st = smem + s * Cfg::STAGE_BYTES;
        uint8_t* pP = st; uint8_t* pQ = st + Cfg::P_BYTES;
        mbar_expect_tx(&full[s], Cfg::P_BYTES + Cfg::Q_BYTES);
        const int k0 = kb * TC_BK;
        if (P_MN) { for (int j = 0; j < TCG_LANES / 32; j++) tma_load_3d(pP + j * 4096, &tmP, &full[s], lane0 + 32 * j, k0, bz); }
        else tma_load_3d(pP, &tmP, &full[s], k0, lane0, bz);
        if (Q_MN) { for (int j = 0; j < TN / 32; j++) tma_load_3d(pQ + j * 4096, &tmQ, &full[s], col0 + 32 * j, k0, bz); }
        else tma_load_3d(pQ, &tmQ, &full[s], k0, col0, bz);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(TCG_LANES, TN, P_MN ? 1 : 0, Q_MN ? 1 : 0);
      for (int kb = 0; kb < nk; kb++) {
        const int s = kb % S; const uint32_t ph = (kb / S) & 1;
        mbar_wait(SPLIT ? &ready[s] : &full[s], ph);
        tc_fence_after();
        const uint32_t st = smem_u32(smem + s * Cfg::STAGE_BYTES);
        const uint32_t aP = st, aQ = st + Cfg::P_BYTES;
        const uint32_t aPl = st + Cfg::P_BYTES + Cfg::Q_BYTES, aQl = aPl + Cfg::P_BYTES;
#pragma unroll
        for (int k = 0; k < TC_BK / 8; k++) {
          const uint64_t dP = P_MN ? umma_desc_mnmajor(aP, k) : umma_desc_kmajor(aP, k);
          const uint64_t dQ = Q_MN ? umma_desc_mnmajor(aQ, k) : umma_desc_kmajor(aQ, k);
          if (SPLIT) {
            const uint64_t dPl = P_MN ? umma_desc_mnmajor(aPl, k) : umma_desc_kmajor(aPl, k);
            const uint64_t dQl = Q_MN ? umma_desc_mnmajor(aQl, k) : umma_desc_kmajor(aQl, k);
            umma_tf32(tmem_base, dPl, dQ, idesc, (kb | k) != 0);
            umma_tf32(tmem_base, dP, dQl, idesc, 1);
            umma_tf32(tmem_base, dP, dQ, idesc, 1);
          } else {
            umma_tf32(tmem_base, dP, dQ, idesc, (kb | k) != 0);
          }
        }
        umma_commit(&empty[s]);            // ring slot reusable once these MMAs have read it
      }
      umma_commit(tmem_full);              // accumulator complete
    }
  } else {
    // ===================== splitter (3xTF32 only), then epilogue =====================
    const int t = threadIdx.x - 64;        // 0..127
    if (SPLIT) {
      for (int kb = 0; kb < nk; kb++) {
        const int s = kb % S; const uint32_t ph = (kb / S) & 1;
        mbar_wait(&full[s], ph);
        float4* hi = (float4*)(smem + s * Cfg::STAGE_BYTES);
        float4* lo = (float4*)(smem + s * Cfg::STAGE_BYTES + Cfg::P_BYTES + Cfg::Q_BYTES);
        constexpr int N4 = (Cfg::P_BYTES + Cfg::Q_BYTES) / 16;
#pragma unroll 4
        for (int i = t; i < N4; i += 128) {
          float4 x = hi[i], h, l;
          h.x = tf32_rna(x.x); h.y = tf32_rna(x.y); h.z = tf32_rna(x.z); h.w = tf32_rna(x.w);
          l.x = tf32_rna(x.x - h.x); l.y = tf32_rna(x.y - h.y); l.z = tf32_rna(x.z - h.z); l.w = tf32_rna(x.w - h.w);
          hi[i] = h; lo[i] = l;
        }
        fence_proxy_async();               // generic-proxy writes -> visible to the tensor core (async proxy)
        mbar_arrive(&ready[s]);
      }
    }
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int q = warp & 3;                // TMEM lane quarter this warp may access
    const int n = lane0 + 32 * q + lane;   // C column handled by this thread
    float* cbase = C + (int64_t)bz * bsc + n;
#pragma unroll 1
    for (int c = 0; c < TN; c += 32) {
      if (col0 + c >= NC) break;
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)c, v);
      tmem_ld_wait();
      if (n < NL) {
#pragma unroll
        for (int j = 0; j < 32; j++) {
          const int m = col0 + c + j;
          if (m < NC) {
            float* p = cbase + (int64_t)m * ldc;
            *p = accumulate ? (*p + v[j]) : v[j];
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TN);
}

// ------------------------------------------------------------------ host side
// operand described as a logical [rows, K] matrix with element strides (rs over rows, ks over k) and batch stride
struct TcOperand { const float* p; int64_t rows, rs, ks, bs; };

static bool tc_operand_ok(const TcOperand& o, int64_t K, int64_t batch, bool& mn_major) {
  if (((uintptr_t)o.p & 15) != 0) return false;
  if (batch > 1 && (o.bs % 4 != 0 || o.bs <= 0)) return false;
  if (o.ks == 1 && o.rs % 4 == 0 && o.rs >= K) { mn_major = false; return true; }        // K contiguous
  if (o.rs == 1 && o.ks % 4 == 0 && o.ks >= o.rows) { mn_major = true; return true; }      // rows contiguous
  return false;
}
static int tc_make_map(CUtensorMap* m, const TcOperand& o, int64_t K, int64_t batch, bool mn_major, int box_rows) {
  uint64_t dims[3], strides[2]; uint32_t box[3];
  if (!mn_major) { dims[0] = K; dims[1] = o.rows; strides[0] = o.rs * 4; box[0] = TC_BK; box[1] = box_rows; }
  else { dims[0] = o.rows; dims[1] = K; strides[0] = o.ks * 4; box[0] = 32; box[1] = TC_BK; }
  dims[2] = batch; strides[1] = (batch > 1 ? o.bs : (mn_major ? o.ks * K : o.rs * o.rows)) * 4; box[2] = 1;
  if (strides[1] == 0) strides[1] = 16;
  return agb_make_tmap(m, o.p, 3, dims, strides, box);
}

template <int TN, bool P_MN, bool Q_MN, bool SPLIT>
static int tc_launch(agb_ctx* ctx, const CUtensorMap& tmP, const CUtensorMap& tmQ, float* C, int NL, int NC, int K, int64_t ldc, int64_t bsc,
                     int64_t batch, int accumulate) {
  using Cfg = TcgCfg<TN, SPLIT>;
  static bool attr = false;
  if (!attr) { AGB_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<TN, P_MN, Q_MN, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM)); attr = true; }
  dim3 grid((NL + TCG_LANES - 1) / TCG_LANES, (NC + TN - 1) / TN, (unsigned)batch);
  tc_gemm_kernel<TN, P_MN, Q_MN, SPLIT><<<grid, TCG_THREADS, Cfg::SMEM, ctx->stream>>>(tmP, tmQ, C, NL, NC, K, ldc, bsc, accumulate);
  AGB_LAUNCHED(ctx);
  return AGB_OK;
}

template <int TN, bool SPLIT>
static int tc_dispatch_major(agb_ctx* ctx, bool pmn, bool qmn, const CUtensorMap& tmP, const CUtensorMap& tmQ, float* C, int NL, int NC, int K,
                             int64_t ldc, int64_t bsc, int64_t batch, int acc) {
  if (!pmn && !qmn) return tc_launch<TN, false, false, SPLIT>(ctx, tmP, tmQ, C, NL, NC, K, ldc, bsc, batch, acc);
  if (!pmn && qmn) return tc_launch<TN, false, true, SPLIT>(ctx, tmP, tmQ, C, NL, NC, K, ldc, bsc, batch, acc);
  if (pmn && !qmn) return tc_launch<TN, true, false, SPLIT>(ctx, tmP, tmQ, C, NL, NC, K, ldc, bsc, batch, acc);
  return tc_launch<TN, true, true, SPLIT>(ctx, tmP, tmQ, C, NL, NC, K, ldc, bsc, batch, acc);
}

// C[m,n] = op(A)[m,k] . op(B)[k,n];  (rsa, csa) strides of op(A) over (m, k), (rsb, csb) strides of op(B) over (k, n)
int agb_tc_gemm(agb_ctx* ctx, int mode, const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, int64_t batch,
                int64_t rsa, int64_t csa, int64_t bsa, int64_t rsb, int64_t csb, int64_t bsb, int64_t bsc, float beta) {
  // worth a 128-lane tile?  (tiny problems stay on the CUDA-core path, which is also exact fp32)
  if (N < 64 || M < 16 || K < 32) return AGB_ERR_UNSUPPORTED;
  if (M > (1ll << 30) || N > (1ll << 30) || K > (1ll << 30) || batch > 65535) return AGB_ERR_UNSUPPORTED;
  if (((uintptr_t)C & 15) != 0) return AGB_ERR_UNSUPPORTED;
  TcOperand P{B, N, csb, rsb, bsb};     // lanes = n : P[n, k] = op(B)[k, n]
  TcOperand Q{A, M, rsa, csa, bsa};     // cols  = m : Q[m, k] = op(A)[m, k]
  bool pmn, qmn;
  if (!tc_operand_ok(P, K, batch, pmn) || !tc_operand_ok(Q, K, batch, qmn)) return AGB_ERR_UNSUPPORTED;
  const bool split = (mode == AGB_MATH_3XTF32);
  const int TN = (M >= 192 && !split) ? 256 : (M > 64 ? 128 : 64);
  CUtensorMap tmP, tmQ;
  int r = tc_make_map(&tmP, P, K, batch, pmn, TCG_LANES); if (r != AGB_OK) return r;
  r = tc_make_map(&tmQ, Q, K, batch, qmn, TN); if (r != AGB_OK) return r;
  const int acc = beta != 0.0f;
  if (split) {
    if (TN == 128) return tc_dispatch_major<128, true>(ctx, pmn, qmn, tmP, tmQ, C, (int)N, (int)M, (int)K, N, bsc, batch, acc);
    return tc_dispatch_major<64, true>(ctx, pmn, qmn, tmP, tmQ, C, (int)N, (int)M, (int)K, N, bsc, batch, acc);
  }
  if (TN == 256) return tc_dispatch_major<256, false>(ctx, pmn, qmn, tmP, tmQ, C, (int)N, (int)M, (int)K, N, bsc, batch, acc);
  if (TN == 128) return tc_dispatch_major<128, false>(ctx, pmn, qmn, tmP, tmQ, C, (int)N, (int)M, (int)K, N, bsc, batch, acc);
  return tc_dispatch_major<64, false>(ctx, pmn, qmn, tmP, tmQ, C, (int)N, (int)M, (int)K, N, bsc, batch, acc);
}
