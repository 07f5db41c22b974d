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
// 3xTF32 (f32-faithful) mode: four extra warps write lo = rna_tf32(x - trunc_tf32(x)) of every landed tile to a second smem
// buffer (the tensor core itself truncates the raw tile to hi), then the issuing thread runs lo*hi + hi*lo + hi*hi; partial sums
// are promoted to fp32 registers every 64 k (tc_tile.cuh).  The kernel bodies live in tc_tile.cuh (shared with tc_conv.cu).
//
// Replaces matrixmultiply::sgemm / cblas_sgemm behind MatMul and BatchMatMul
// (reference src/tensor_ops/dot_ops.rs:383-422, 142-380).
#include <stdlib.h>
#include "tc_tile.cuh"

template <int TN_, bool P_MN_, bool Q_MN_, bool SPLIT_, bool QPRE_ = false, bool PPRE_ = false> struct GemmPol {
  static constexpr int TN = TN_, MT = 1; static constexpr bool SPLIT = SPLIT_, P_MN = P_MN_, Q_MN = Q_MN_, Q_PRESPLIT = QPRE_, P_PRESPLIT = PPRE_;
  static_assert(!PPRE_ || QPRE_, "a pre-split P comes with a pre-split Q (no splitter warps at all)");
  static constexpr int OCC = (SPLIT_ || TN_ > 128) ? 1 : 2;
  static constexpr bool PAIR2 = !SPLIT_ && TN_ == 256;       // CTA pairs: tmQlo = the Q map with TN / 2 rows per box
  static constexpr bool SPLIT_PAIR2 = SPLIT_ && QPRE_ && !PPRE_ && TN_ == 128;      // 3xTF32 CTA pairs: tmQh / tmQlh = half-height boxes of the hi / lo planes
  // splits > 1: split-K for problems with too few output tiles to fill the machine (e.g. the LSTM's 128 x 1024 x 8192 dgrad GEMM is
  // 8 tiles): grid.z = batch * splits, every CTA reduces kb_per_split k-blocks and adds its partial with red.global.add (C pre-zeroed)
  struct Params { CUtensorMap tmP, tmQ, tmQlo /* 3xTF32: lo plane; CTA pairs: half-height Q boxes */, tmQh, tmQlh; float* C; int NL, NC, K; int64_t ldc, bsc; int accumulate; int splits, kb_per_split; MnDescCfg mnc; int64_t part_stride /* > 0: split ks stores its partial at C + ks * part_stride (deterministic mode) */;
                  CUtensorMap tmPw, tmQw; int widep, wideq;
                  CUtensorMap tmPlo, tmPwlo; };     // 3xTF32 with P pre-split in global memory: tmP / tmPw address the hi plane, these the lo plane      // MN-major operand with rows % 32 == 0: the row axis as {32, blocks}, ONE box per tile instead of one per 32 rows
  struct Tile { int lane0, col0, bz, kb0, nkb, ks; };
  __device__ static Tile tile(const Params& p, uint3 blk) {
    const int kb_total = (p.K + TC_BK - 1) / TC_BK;
    const int bz = (int)blk.z / p.splits, ks = (int)blk.z - bz * p.splits;
    const int kb0 = ks * p.kb_per_split;
    return Tile{(int)blk.x * TC_LANES, (int)blk.y * TN, bz, kb0, max(0, min(p.kb_per_split, kb_total - kb0)), ks};
  }
  __device__ static int num_kblocks(const Params&, const Tile& t) { return t.nkb; }
  __device__ static uint32_t p_bytes(const Params&, uint32_t full) { return full; }
  __device__ static void prefetch(const Params& p) { tma_prefetch_desc(&p.tmP); tma_prefetch_desc(&p.tmQ); if (QPRE_) tma_prefetch_desc(&p.tmQlo); if (PPRE_) tma_prefetch_desc(&p.tmPlo); }
  __device__ static void load_q(const CUtensorMap* tm, const Tile& t, int kb, uint8_t* pQ, uint64_t* bar) {
    const int k0 = (t.kb0 + kb) * TC_BK;
    if (Q_MN) { for (int j = 0; j < TN / 32; j++) tma_load_3d(pQ + j * 4096, tm, bar, t.col0 + 32 * j, k0, t.bz); }
    else tma_load_3d(pQ, tm, bar, k0, t.col0, t.bz);
  }
  __device__ static void load_p(const CUtensorMap* tm, const CUtensorMap* tmw, int wide, const Tile& t, int k0, uint8_t* pP, uint64_t* bar) {
    if (P_MN) {
      if (wide) tma_load_4d(pP, tmw, bar, 0, k0, t.lane0 >> 5, t.bz);
      else for (int j = 0; j < TC_LANES / 32; j++) tma_load_3d(pP + j * 4096, tm, bar, t.lane0 + 32 * j, k0, t.bz);
    }
    else tma_load_3d(pP, tm, bar, k0, t.lane0, t.bz);
  }
  __device__ static void load_p_lo(const Params& p, const Tile& t, int kb, uint8_t* pPlo, uint64_t* bar) { load_p(&p.tmPlo, &p.tmPwlo, p.widep, t, (t.kb0 + kb) * TC_BK, pPlo, bar); }
  __device__ static void load(const Params& p, const Tile& t, int kb, uint8_t* pP, uint8_t* pQ, uint64_t* bar) {
    const int k0 = (t.kb0 + kb) * TC_BK;
    load_p(&p.tmP, &p.tmPw, p.widep, t, k0, pP, bar);
    if (Q_MN && p.wideq) tma_load_4d(pQ, &p.tmQw, bar, 0, k0, t.col0 >> 5, t.bz);
    else load_q(&p.tmQ, t, kb, pQ, bar);
  }
  __device__ static void load_q_lo(const Params& p, const Tile& t, int kb, uint8_t* pQlo, uint64_t* bar) { load_q(&p.tmQlo, t, kb, pQlo, bar); }       // 3xTF32 with the Q operand pre-split in global memory
  __device__ static void load_sp2(const Params& p, const Tile& t, int kb, uint8_t* pP, uint8_t* pQh, uint8_t* pQl, uint64_t* bar, int rank) {
    const int k0 = (t.kb0 + kb) * TC_BK, c0 = t.col0 + rank * (TN / 2);
    if (P_MN) { for (int j = 0; j < TC_LANES / 32; j++) tma_load_3d(pP + j * 4096, &p.tmP, bar, t.lane0 + 32 * j, k0, t.bz); }
    else tma_load_3d(pP, &p.tmP, bar, k0, t.lane0, t.bz);
    if (Q_MN) { for (int j = 0; j < TN / 64; j++) { tma_load_3d(pQh + j * 4096, &p.tmQh, bar, c0 + 32 * j, k0, t.bz); tma_load_3d(pQl + j * 4096, &p.tmQlh, bar, c0 + 32 * j, k0, t.bz); } }
    else { tma_load_3d(pQh, &p.tmQh, bar, k0, c0, t.bz); tma_load_3d(pQl, &p.tmQlh, bar, k0, c0, t.bz); }
  }
  // CTA pair: this CTA's 128 P rows + its half of the Q rows, completing on the leader's barrier
  __device__ static void load2(const Params& p, const Tile& t, int kb, uint8_t* pP, uint8_t* pQ, uint32_t bar, int rank) {
    const int k0 = (t.kb0 + kb) * TC_BK, c0 = t.col0 + rank * (TN / 2);
    if (P_MN) {
      if (p.widep) tma_load_4d_2sm(pP, &p.tmPw, bar, 0, k0, t.lane0 >> 5, t.bz);
      else for (int j = 0; j < TC_LANES / 32; j++) tma_load_3d_2sm(pP + j * 4096, &p.tmP, bar, t.lane0 + 32 * j, k0, t.bz);
    }
    else tma_load_3d_2sm(pP, &p.tmP, bar, k0, t.lane0, t.bz);
    if (Q_MN) {
      if (p.wideq) tma_load_4d_2sm(pQ, &p.tmQw, bar, 0, k0, c0 >> 5, t.bz);      // (the pair's wide Q map has TN / 2 rows per box)
      else for (int j = 0; j < TN / 64; j++) tma_load_3d_2sm(pQ + j * 4096, &p.tmQlo, bar, c0 + 32 * j, k0, t.bz);
    }
    else tma_load_3d_2sm(pQ, &p.tmQlo, bar, k0, c0, t.bz);
  }
  // thread `lane` owns C column n = lane0 + lane; v[j] belongs to C row m = col0 + c0 + j: a warp stores 32 consecutive
  // floats of one C row per instruction (128-byte coalesced), no shared-memory staging
  __device__ static void pre_epilogue(const Params&, const Tile&, int, uint32_t*) {}
  __device__ static void store(const Params& p, const Tile& t, int, int lane, int c0, const float* v, uint32_t&) {
    const int n = t.lane0 + lane;
    if (n >= p.NL) return;
    float* cbase = p.C + (int64_t)t.bz * p.bsc + n + (int64_t)t.ks * p.part_stride;      // part_stride > 0 (deterministic split-K): this split's own copy of C
#pragma unroll
    for (int j = 0; j < 32; j++) {
      const int m = t.col0 + c0 + j;
      if (m < p.NC) {
        float* q = cbase + (int64_t)m * p.ldc;
        if (p.splits > 1 && p.part_stride == 0) red_add_f32(q, v[j]); else *q = (p.accumulate && p.splits == 1) ? (*q + v[j]) : v[j];
      }
    }
  }
};

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
  bool a32 = true; agb_mn_cfg(&a32);
  return agb_make_tmap(m, o.p, 3, dims, strides, box, mn_major && a32);
}

struct TcHalfMaps { CUtensorMap hi, lo; };
// wide maps of the MN-major operands of the current agb_tc_gemm call (set before the dispatch, read by tc_launch; single pass only)
struct TcWide { CUtensorMap p, q, plo; int hp = 0, hq = 0; };
static thread_local TcWide g_tc_wide;
static int tc_make_wide_map(CUtensorMap* m, const TcOperand& o, int64_t K, int64_t batch, int box_rows) {
  uint64_t dims[4] = {32, (uint64_t)K, (uint64_t)(o.rows / 32), (uint64_t)batch};
  uint64_t str[3] = {(uint64_t)o.ks * 4, 128, (uint64_t)((batch > 1 ? o.bs : o.ks * K) * 4)};
  if (str[2] == 0) str[2] = 16;
  uint32_t box[4] = {32, TC_BK, (uint32_t)(box_rows / 32), 1};
  return agb_make_tmap(m, o.p, 4, dims, str, box, true);
}       // 3xTF32 CTA pairs: TN / 2-row boxes of the pre-split Q planes
template <int TN, bool P_MN, bool Q_MN, bool SPLIT, bool QPRE = false, bool PPRE = false>
static int tc_launch(agb_ctx* ctx, const CUtensorMap& tmP, const CUtensorMap& tmQ, const CUtensorMap* tmQlo, float* C, int NL, int NC, int K, int64_t ldc, int64_t bsc,
                     int64_t batch, int accumulate, const TcHalfMaps* half = nullptr, const CUtensorMap* tmPlo = nullptr) {
  using Pol = GemmPol<TN, P_MN, Q_MN, SPLIT, QPRE, PPRE>;
  const int gx = (NL + TC_LANES - 1) / TC_LANES, gy = (NC + TN - 1) / TN;
  const int kb_total = (K + TC_BK - 1) / TC_BK;
  // split-K when the output tiles cannot fill the machine and there is K to share (>= 4 k-blocks per split)
  int64_t tiles = (int64_t)gx * gy * batch, cap = (int64_t)ctx->sm_count * Pol::OCC;
  int splits = 1;
  if (tiles * 2 <= cap && kb_total >= 8 && ldc == NL && (batch == 1 || bsc == (int64_t)NC * NL)) {
    static const int min_kb = [] { const char* e = getenv("AGB_GEMM_SPLIT_MIN_KB"); int v = e ? atoi(e) : 4; return v < 1 ? 1 : v; }();      // tuning knob (k-blocks per split)
    static const int occ_cap = [] { const char* e = getenv("AGB_GEMM_SPLIT_WAVES_X2"); int v = e ? atoi(e) : 2; return v < 1 ? 1 : v; }();     // cap = SMs * OCC * v / 2
    int64_t s = cap * occ_cap / 2 / tiles; if (s > kb_total / min_kb) s = kb_total / min_kb; if (s > 1) splits = (int)s;
  }
  int kb_per = (kb_total + splits - 1) / splits; splits = (kb_total + kb_per - 1) / kb_per;
  if ((int64_t)batch * splits > 65535) { splits = 1; kb_per = kb_total; }
  const int64_t cn = (int64_t)batch * NC * NL;
  float* part = nullptr;
  // deterministic split-K through per-split copies of C pays 2 x splits x |C| of traffic plus a launch: right for the small outputs split-K exists for
  // (the classifier's [256, 10], weight gradients), wrong for a 128 x 4096 recurrence GEMM that lives on launch latency (measured: 12.3 -> 24.9 us).
  // Above 1 MB of partials (a 512^3 product already pays +40 % for 4 MB of them) the sums stay on red.global.add (run-to-run differences of fp32 reassociation, like the embedding scatter-add).
  if (splits > 1 && ctx->deterministic && (size_t)splits * cn * sizeof(float) <= (1u << 20)) AGB_TRY(agb_scratch2(ctx, (size_t)splits * cn * sizeof(float), (void**)&part));
  if (splits > 1 && !accumulate && !part) AGB_TRY(agb_memset0(ctx, C, (size_t)cn * sizeof(float)));
  typename Pol::Params prm{tmP, tmQ, tmQlo ? *tmQlo : tmQ, half ? half->hi : tmQ, half ? half->lo : tmQ, part ? part : C, NL, NC, K, ldc, bsc, accumulate, splits, kb_per, agb_mn_cfg(), part ? cn : 0};
  prm.widep = 0; prm.wideq = 0; prm.tmPw = tmP; prm.tmQw = tmQ;
  if (P_MN && g_tc_wide.hp) { prm.tmPw = g_tc_wide.p; prm.widep = 1; }
  prm.tmPlo = tmPlo ? *tmPlo : tmP; prm.tmPwlo = (PPRE && prm.widep) ? g_tc_wide.plo : prm.tmPw;
  if (Q_MN && g_tc_wide.hq && !QPRE) { prm.tmQw = g_tc_wide.q; prm.wideq = 1; }      // (a pre-split Q comes from its hi / lo planes)
  dim3 grid(gx, gy, (unsigned)(batch * splits));
  AGB_TRY(tc_tile_launch<Pol>(ctx, prm, grid));
  if (part) return agb_reduce_partials(ctx, part, C, splits, cn, cn, accumulate);
  return AGB_OK;
}

template <int TN, bool SPLIT, bool QPRE = false, bool PPRE = false>
static int tc_dispatch_major(agb_ctx* ctx, bool pmn, bool qmn, const CUtensorMap& tmP, const CUtensorMap& tmQ, const CUtensorMap* tmQlo, float* C, int NL, int NC, int K,
                             int64_t ldc, int64_t bsc, int64_t batch, int acc, const TcHalfMaps* half = nullptr, const CUtensorMap* tmPlo = nullptr) {
  if (!pmn && !qmn) return tc_launch<TN, false, false, SPLIT, QPRE, PPRE>(ctx, tmP, tmQ, tmQlo, C, NL, NC, K, ldc, bsc, batch, acc, half, tmPlo);
  if (!pmn && qmn) return tc_launch<TN, false, true, SPLIT, QPRE, PPRE>(ctx, tmP, tmQ, tmQlo, C, NL, NC, K, ldc, bsc, batch, acc, half, tmPlo);
  if (pmn && !qmn) return tc_launch<TN, true, false, SPLIT, QPRE, PPRE>(ctx, tmP, tmQ, tmQlo, C, NL, NC, K, ldc, bsc, batch, acc, half, tmPlo);
  return tc_launch<TN, true, true, SPLIT, QPRE, PPRE>(ctx, tmP, tmQ, tmQlo, C, NL, NC, K, ldc, bsc, batch, acc, half, tmPlo);
}

// 3xTF32: hi = rna_tf32(x), lo = rna_tf32(x - hi) planes of a dense operand, written once to scratch when the operand is re-read by
// enough lane tiles to pay for the pass (12 bytes per element against 48 KB of shared-memory traffic per k-block and CTA saved)
__global__ void __launch_bounds__(256) presplit_kernel(const float4* __restrict__ x, float4* __restrict__ hi, float4* __restrict__ lo, int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(x + i); float4 h, l;
    h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
    l.x = tf32_rna(v.x - h.x); l.y = tf32_rna(v.y - h.y); l.z = tf32_rna(v.z - h.z); l.w = tf32_rna(v.w - h.w);
    hi[i] = h; lo[i] = l;
  }
}

// (also the gy planes of the 3xTF32 filter gradient, tc_conv.cu); src / hi / lo 16-byte aligned, n % 4 == 0
int agb_tc_presplit(agb_ctx* ctx, const float* src, float* hi, float* lo, int64_t n) {
  if (n <= 0) return AGB_OK;
  presplit_kernel<<<agb_grid_for(n / 4, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>((const float4*)src, (float4*)hi, (float4*)lo, n / 4);
  AGB_LAUNCHED(ctx);
  return AGB_OK;
}

// A thin operand whose contiguous extent is not a multiple of 4 floats (the [65536, 10] classifier weight of a CNN, its [256, 10]
// logits gradient) breaks TMA's 16-byte pitch rule.  Such an operand is tiny next to the other one, so it is copied once per call into
// scratch with the pitch rounded up to 16 floats (zero padded) and the GEMM runs on the tensor cores instead of the CUDA-core path
// (VGG classifier, batch 256: forward + two gradients 229 us -> the three calls are bound by streaming the 67 MB activation once each).
__global__ void __launch_bounds__(256) pad_pitch_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t outer, int inner, int64_t src_pitch, int dst_pitch) {
  const int64_t n = outer * dst_pitch;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o = i / dst_pitch; const int j = (int)(i - o * dst_pitch);
    dst[i] = j < inner ? __ldg(src + o * src_pitch + j) : 0.0f;
  }
}
// returns true when `o` was re-pointed at a padded copy (scratch cursor advanced)
static bool tc_pad_thin(agb_ctx* ctx, TcOperand& o, int64_t K, float*& cursor, int* status) {
  *status = AGB_OK;
  int64_t outer, inner, pitch; bool k_inner;
  if (o.ks == 1 && o.rs >= K) { outer = o.rows; inner = K; pitch = o.rs; k_inner = true; }           // K contiguous
  else if (o.rs == 1 && o.ks >= o.rows) { outer = K; inner = o.rows; pitch = o.ks; k_inner = false; } // rows contiguous
  else return false;
  if ((pitch % 4 == 0 && ((uintptr_t)o.p & 15) == 0) || inner > 64) return false;
  const int dp = (int)((inner + 15) / 16 * 16);
  pad_pitch_kernel<<<agb_grid_for(outer * dp, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(o.p, cursor, outer, (int)inner, pitch, dp);
  ctx->launches++;
  o.p = cursor; cursor += outer * dp;
  if (k_inner) o.rs = dp; else o.ks = dp;
  return true;
}

// C[m,n] = op(A)[m,k] . op(B)[k,n];  (rsa, csa) strides of op(A) over (m, k), (rsb, csb) strides of op(B) over (k, n)
int agb_tc_gemm(agb_ctx* ctx, int mode, const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, int64_t batch,
                int64_t rsa, int64_t csa, int64_t bsa, int64_t rsb, int64_t csb, int64_t bsb, int64_t bsc, float beta) {
  // worth a 128-lane tile?  (tiny problems stay on the CUDA-core path, which is also exact fp32)
  const bool thin_big = batch == 1 && M >= 16 && (double)M * (double)N * (double)K >= (double)(1 << 26) && (N < 64 || K < 32);
  if ((N < 64 || M < 16 || K < 32) && !thin_big) return AGB_ERR_UNSUPPORTED;
  if (M > (1ll << 30) || N > (1ll << 30) || K > (1ll << 30) || batch > 65535) return AGB_ERR_UNSUPPORTED;
  if (((uintptr_t)C & 15) != 0) return AGB_ERR_UNSUPPORTED;
  TcOperand P{B, N, csb, rsb, bsb};     // lanes = n : P[n, k] = op(B)[k, n]
  TcOperand Q{A, M, rsa, csa, bsa};     // cols  = m : Q[m, k] = op(A)[m, k]
  bool pmn, qmn;
  float* pad_base = nullptr; size_t pad_floats = 0;
  if (thin_big) {        // padded copies of the thin operand(s): at most 64 floats of pitch per outer index
    bool p_ok = tc_operand_ok(P, K, batch, pmn), q_ok = tc_operand_ok(Q, K, batch, qmn);
    if (!p_ok) pad_floats += (size_t)((P.ks == 1 ? P.rows : K)) * 64;
    if (!q_ok) pad_floats += (size_t)((Q.ks == 1 ? Q.rows : K)) * 64;
    if (pad_floats) {
      AGB_TRY(agb_scratch(ctx, pad_floats * sizeof(float), (void**)&pad_base));
      float* cur = pad_base; int st;
      if (!p_ok) { tc_pad_thin(ctx, P, K, cur, &st); AGB_TRY(st); }
      if (!q_ok) { tc_pad_thin(ctx, Q, K, cur, &st); AGB_TRY(st); }
    }
  }
  if (!tc_operand_ok(P, K, batch, pmn) || !tc_operand_ok(Q, K, batch, qmn)) return AGB_ERR_UNSUPPORTED;
  const bool split = (mode == AGB_MATH_3XTF32);
  // 256-wide tiles (one CTA per SM, 128 KB epilogue per tile) only pay off when the k-loop is long enough to hide the epilogue behind it
  const int TN = (M >= 192 && !split && K >= 512) ? 256 : (M > 64 ? 128 : 64);
  CUtensorMap tmP, tmQ;
  int r = tc_make_map(&tmP, P, K, batch, pmn, TC_LANES); if (r != AGB_OK) return r;
  r = tc_make_map(&tmQ, Q, K, batch, qmn, TN); if (r != AGB_OK) return r;
  const int acc = beta != 0.0f;
  g_tc_wide.hp = g_tc_wide.hq = 0;
  static const int wide_env = [] { const char* e = getenv("AGB_GEMM_WIDE"); return (e && e[0] == '0') ? 0 : 1; }();
  if (wide_env) {
    if (pmn && P.rows % 32 == 0 && tc_make_wide_map(&g_tc_wide.p, P, K, batch, TC_LANES) == AGB_OK) g_tc_wide.hp = 1;
    if (qmn && Q.rows % 32 == 0 && tc_make_wide_map(&g_tc_wide.q, Q, K, batch, TN == 256 ? 128 : TN) == AGB_OK) g_tc_wide.hq = 1;      // CTA pairs: half of the Q tile per CTA
  }
  if (split) {
    // Q = op(A) is re-read by every one of the N/128 lane tiles: pre-split it once in global memory when it is dense and N is large
    const int64_t qn = M * K * batch;
    const bool q_dense = qn % 4 == 0 && (batch == 1 || Q.bs == M * K) && (qmn ? (Q.ks == M) : (Q.rs == K));
    if (q_dense && N >= 512 && qn <= (1ll << 31) && pad_base == nullptr) {
      // P = op(B)^T is re-read by every one of the M/TN column tiles: with M large (>= 16 tiles) it is pre-split too (hi = rna_tf32, lo planes), the kernel then has
      // no splitter work at all and its shared memory carries only the TMA writes and the MMA reads (this mode is bound by shared-memory bandwidth)
      const int64_t pn = N * K * batch;
      const bool p_dense = pn % 4 == 0 && (batch == 1 || P.bs == N * K) && (pmn ? (P.ks == N) : (P.rs == K));
      static const int ppre_env = [] { const char* e = getenv("AGB_GEMM_PPRE"); return (e && e[0] == '0') ? 0 : 1; }();
      const bool ppre = ppre_env && p_dense && M >= 2048 && pn <= (1ll << 28);      // measured: batch 64 x 512^3 +11 %, 16 x 1024^3 +3 % (slower), 2048^3 -5 %, 8192^3 -11 %
      float* planes = nullptr;
      AGB_TRY(agb_scratch(ctx, (size_t)(qn * 2 + (ppre ? pn * 2 : 0)) * sizeof(float), (void**)&planes));
      presplit_kernel<<<agb_grid_for(qn / 4, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>((const float4*)A, (float4*)planes, (float4*)(planes + qn), qn / 4);
      AGB_LAUNCHED(ctx);
      CUtensorMap tmQh, tmQl, tmPh, tmPl;
      TcOperand Qh = Q, Ql = Q; Qh.p = planes; Ql.p = planes + qn;
      r = tc_make_map(&tmQh, Qh, K, batch, qmn, TN); if (r != AGB_OK) return r;
      r = tc_make_map(&tmQl, Ql, K, batch, qmn, TN); if (r != AGB_OK) return r;
      if (ppre) {
        float* pp = planes + 2 * qn;
        presplit_kernel<<<agb_grid_for(pn / 4, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>((const float4*)B, (float4*)pp, (float4*)(pp + pn), pn / 4);
        AGB_LAUNCHED(ctx);
        TcOperand Ph = P, Pl = P; Ph.p = pp; Pl.p = pp + pn;
        r = tc_make_map(&tmPh, Ph, K, batch, pmn, TC_LANES); if (r != AGB_OK) return r;
        r = tc_make_map(&tmPl, Pl, K, batch, pmn, TC_LANES); if (r != AGB_OK) return r;
        if (g_tc_wide.hp) {      // the wide (one box per tile) maps of the MN-major planes
          if (tc_make_wide_map(&g_tc_wide.p, Ph, K, batch, TC_LANES) != AGB_OK || tc_make_wide_map(&g_tc_wide.plo, Pl, K, batch, TC_LANES) != AGB_OK) g_tc_wide.hp = 0;
        }
        if (TN == 128) return tc_dispatch_major<128, true, true, true>(ctx, pmn, qmn, tmPh, tmQh, &tmQl, C, (int)N, (int)M, (int)K, N, bsc, batch, acc, nullptr, &tmPl);
        return tc_dispatch_major<64, true, true, true>(ctx, pmn, qmn, tmPh, tmQh, &tmQl, C, (int)N, (int)M, (int)K, N, bsc, batch, acc, nullptr, &tmPl);
      }
      if (TN == 128) {
        TcHalfMaps half;
        r = tc_make_map(&half.hi, Qh, K, batch, qmn, TN / 2); if (r != AGB_OK) return r;
        r = tc_make_map(&half.lo, Ql, K, batch, qmn, TN / 2); if (r != AGB_OK) return r;
        return tc_dispatch_major<128, true, true>(ctx, pmn, qmn, tmP, tmQh, &tmQl, C, (int)N, (int)M, (int)K, N, bsc, batch, acc, &half);
      }
      return tc_dispatch_major<64, true, true>(ctx, pmn, qmn, tmP, tmQh, &tmQl, C, (int)N, (int)M, (int)K, N, bsc, batch, acc);
    }
    if (TN == 128) return tc_dispatch_major<128, true>(ctx, pmn, qmn, tmP, tmQ, nullptr, C, (int)N, (int)M, (int)K, N, bsc, batch, acc);
    return tc_dispatch_major<64, true>(ctx, pmn, qmn, tmP, tmQ, nullptr, C, (int)N, (int)M, (int)K, N, bsc, batch, acc);
  }
  if (TN == 256) {
    CUtensorMap tmQhalf;        // CTA pairs: each CTA of a pair loads TN / 2 rows of the Q tile
    r = tc_make_map(&tmQhalf, Q, K, batch, qmn, TN / 2); if (r != AGB_OK) return r;
    return tc_dispatch_major<256, false>(ctx, pmn, qmn, tmP, tmQ, &tmQhalf, C, (int)N, (int)M, (int)K, N, bsc, batch, acc);
  }
  if (TN == 128) return tc_dispatch_major<128, false>(ctx, pmn, qmn, tmP, tmQ, nullptr, C, (int)N, (int)M, (int)K, N, bsc, batch, acc);
  return tc_dispatch_major<64, false>(ctx, pmn, qmn, tmP, tmQ, nullptr, C, (int)N, (int)M, (int)K, N, bsc, batch, acc);
}
