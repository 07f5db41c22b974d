// tc_tile.cuh — the shared tcgen05 "tile engine" behind the GEMM and implicit-GEMM convolution kernels.
//
//   D[lane, col] = sum_k P[lane, k] * Q[col, k]        128 TMEM lanes x TN TMEM columns per CTA, fp32 accumulators in TMEM
//
// Warp roles (one CTA per output tile):
//   warp 0      TMA producer: a Policy functor issues the `cp.async.bulk.tensor` loads of one 32-wide k-block of P and Q
//               into a multi-stage mbarrier ring (128-byte swizzled tiles, zero fill out of bounds)
//   warp 1      allocates TMEM, then ONE thread issues `tcgen05.mma.kind::tf32` (K = 8 per instruction) and
//               `tcgen05.commit`s ring slots back to the producer and accumulators to the drain warps
//   warps 2-5   SPLIT (3xTF32) only: the tensor core TRUNCATES an fp32 operand to tf32 (the low 13 mantissa bits are ignored), so the
//               landed tile itself serves as hi = trunc_tf32(x); the splitters only write lo = rna_tf32(x - trunc_tf32(x)) (exact
//               difference) to a second buffer — one read and ONE write per element, the tile engine's scarce resource in this mode
//               being shared-memory bandwidth (an N = 128 MMA already reads 128 B/clk).  An operand that is tiny and re-read by many
//               CTAs (a convolution's repacked filter) arrives pre-split instead: Pol::Q_PRESPLIT, its lo plane is a second TMA load.
//               A dense GEMM operand that many tiles re-read is pre-split the same way (tc_gemm.cu); with both sides pre-split the splitters idle.
//               The lo plane of Q lies directly behind Q in the stage, so the issuer runs ONE N = 2 TN instruction P.[Q | Q_lo] (hi*hi in TMEM
//               columns [0, TN), hi*lo in [TN, 2 TN)) and one N = TN instruction P_lo.Q into [TN, 2 TN): P is read from shared memory twice per
//               k-step instead of three times (20 instead of 24 KB for TN = 128) — shared-memory bandwidth is what bounds this mode (dropped lo*lo
//               ~ 2^-22).  The tensor core truncates on every accumulate, so keeping the small cross terms out of the main sum leaves it one
//               truncation per K = 8 step instead of three (measured bias on positive operands, relative to the sum: -2.4e-6 with everything in
//               one accumulator and 256-k chunks, -1.1e-6 with 64-k chunks; scripts/bias_probe.py)
//   last 4      drain/epilogue: `tcgen05.ld` the accumulator and hand 32-column strips to the Policy's store functor
//
// fp32-faithful accumulation (SPLIT): the tensor core adds into the TMEM accumulator with truncation, which biases long
// sums toward zero (measured on B200: relative bias ~5e-9 * K, i.e. 1e-4 at K = 16k; ~half an ulp of the running sum per MMA,
// and the two small cross terms of 3xTF32 pay it too).  The bias is coherent across output elements, so a reduction over the
// GEMM's output (a bias gradient summing 1024 rows) accumulates it: with 256-k chunks the LSTM LM's bias gradient was 1.1e-4 off
// the oracle while every GEMM output was within 1e-5.  The 3xTF32 mode therefore accumulates at most TC_KC k-blocks (64 k = 24
// MMAs) per TMEM buffer set (2 TN columns: main and cross sums), ping-pongs two sets, and the drain warps add each finished chunk into
// fp32 registers with round-to-nearest CUDA-core adds while the tensor core works on the other set.
#pragma once
#include <type_traits>
#include "tc_common.cuh"

#define TC_LANES 128
#define TC_KC 2            // k-blocks per TMEM accumulation chunk in SPLIT mode (2 * 32 = 64 k)

// OCC = CTAs co-resident per SM.  Two co-resident CTAs overlap one CTA's prologue / epilogue (TMEM drain, global stores) with the
// other's MMA main loop without a persistent tile scheduler; the ring depth is what fits in 1/OCC of the 227 KB shared memory.
// MT = 128-lane M-tiles per CTA (1 or 2): MT = 2 shares every Q (filter / B) tile between two accumulators, halving the Q bytes a
// CTA pulls from L2 per output element — the tile engine kernels are bound by L2 -> shared-memory ingest, not by the tensor pipe.
template <int TN, bool SPLIT, int OCC = 1, int MT = 1> struct TcCfg {
  static constexpr int P_BYTES = MT * TC_LANES * TC_BK * 4;            // 16 KB per M-tile
  static constexpr int Q_BYTES = TN * TC_BK * 4;
  static constexpr int STAGE_BYTES = (P_BYTES + Q_BYTES) * (SPLIT ? 2 : 1);
  static constexpr int BUDGET = (224 * 1024) / OCC - 2048;
  static constexpr int STAGES = BUDGET / STAGE_BYTES > 6 ? 6 : BUDGET / STAGE_BYTES;
  static_assert(STAGES >= 2, "tile does not fit the shared-memory budget");
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int THREADS = SPLIT ? 320 : 192;
  static constexpr int TMEM_COLS = SPLIT ? 2 * TN : MT * TN;           // power of two >= 32 for TN in {32,64,128,256}
  static_assert(MT == 1 || (!SPLIT && MT == 2 && MT * TN <= 512), "two M-tiles need 2*TN TMEM columns and the single-pass mode");
  static_assert(!SPLIT || TN <= 128, "3xTF32 keeps TN fp32 partial sums per thread in registers");
};

// Policy interface:
//   static constexpr int TN, OCC; static constexpr bool SPLIT, P_MN, Q_MN, Q_PRESPLIT, PAIR2;
//   (SPLIT_PAIR2) __device__ static void load_sp2(const Params&, const Tile&, int kb, uint8_t* pP, uint8_t* pQh_hi, uint8_t* pQh_lo, uint64_t* bar, int rank);
//   (PAIR2) __device__ static void load2(const Params&, const Tile&, int kb, uint8_t* pP, uint8_t* pQhalf, uint32_t leader_bar, int rank);
//   (Q_PRESPLIT) __device__ static void load_q_lo(const Params&, const Tile&, int kb, uint8_t* pQlo, uint64_t* bar);
//   struct Params { ... CUtensorMap members ...; MnDescCfg mnc; };
//   struct Tile { ... };                                               per-CTA coordinates
//   __device__ static Tile tile(const Params&, uint3 blk);             blk = block coordinates in the logical grid
//   __device__ static int  num_kblocks(const Params&, const Tile&);
//   __device__ static void prefetch(const Params&);
//   __device__ static void load(const Params&, const Tile&, int kb, uint8_t* pP, uint8_t* pQ, uint64_t* bar);   one thread
//   __device__ static void store(const Params&, const Tile&, int lane /*0..127*/, int c0 /*0..TN-32*/, const float* v /*[32]*/);
template <class Pol>
__global__ void __launch_bounds__(TcCfg<Pol::TN, false, Pol::OCC, Pol::MT>::THREADS, Pol::OCC) tc_tile_kernel(const __grid_constant__ typename Pol::Params prm) {
  // one output tile per CTA, single-pass (TF32) mode: the fallback of the persistent kernel below (AGB_TC_PERSIST=0, > 2^31 tiles)
  constexpr int TN = Pol::TN; constexpr bool P_MN = Pol::P_MN, Q_MN = Pol::Q_MN;
  constexpr int MT = Pol::MT;
  static_assert(!Pol::SPLIT, "3xTF32 runs on tc_tile_split_kernel");
  using Cfg = TcCfg<TN, false, Pol::OCC, MT>;
  constexpr int S = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + S * Cfg::STAGE_BYTES);
  uint64_t* full = bars; uint64_t* empty = bars + S;
  uint64_t* acc_full = bars + 2 * S;
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * S + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const typename Pol::Tile tl = Pol::tile(prm, blockIdx);
  const int nk = Pol::num_kblocks(prm, tl);

  if (warp == 0 && lane == 0) {
    Pol::prefetch(prm);
    for (int s = 0; s < S; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(&acc_full[0], 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
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
        mbar_expect_tx(&full[s], Pol::p_bytes(prm, Cfg::P_BYTES) + Cfg::Q_BYTES);
        Pol::load(prm, tl, kb, st, st + Cfg::P_BYTES, &full[s]);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp walks the loop converged and ONE elected lane issues: with warp-uniform control flow the descriptors live in
    // uniform registers and a tcgen05.mma costs ~4 issue slots.  (Issuing from inside an `if (lane == 0)` branch made the compiler
    // wrap every MMA in an ELECT / R2UR.BROADCAST waterfall, ~25 instructions per MMA: measured, the issuer thread — not the
    // tensor pipe or L2 — was the bottleneck of every TN <= 128 kernel.)
    constexpr uint32_t idesc = umma_idesc_tf32(TC_LANES, TN, P_MN ? 1 : 0, Q_MN ? 1 : 0);
    const MnDescCfg mnc = prm.mnc;
    // descriptor = {lo: addr>>4 | LBO>>4 << 16, hi: SBO>>4 | version 1 << 14 | layout << 29}; only the address field changes
    const uint32_t hiK = (1024u >> 4) | (1u << 14) | (2u << 29), loK = (16u >> 4) << 16, stepK = 32u >> 4;
    const uint32_t hiM = (mnc.sbo >> 4) | (1u << 14) | (mnc.layout << 29), loM = (mnc.lbo >> 4) << 16, stepM = mnc.kadv >> 4;
    const uint32_t hiP = P_MN ? hiM : hiK, loP = P_MN ? loM : loK, stepP = P_MN ? stepM : stepK;
    const uint32_t hiQ = Q_MN ? hiM : hiK, loQ = Q_MN ? loM : loK, stepQ = Q_MN ? stepM : stepK;
    const uint32_t smem0 = smem_u32(smem) >> 4;
    int s = 0; uint32_t ph = 0;
    for (int kb = 0; kb < nk; kb++) {
      mbar_wait(&full[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t st = smem0 + (uint32_t)s * (Cfg::STAGE_BYTES >> 4);
        const uint32_t aP = st + loP, aQ = st + (Cfg::P_BYTES >> 4) + loQ;
#pragma unroll
        for (int k = 0; k < TC_BK / 8; k++) {
          const uint64_t dQ = umma_desc_pack(aQ + k * stepQ, hiQ);
#pragma unroll
          for (int mt = 0; mt < MT; mt++)
            umma_tf32(tmem_base + (uint32_t)(mt * TN), umma_desc_pack(aP + (uint32_t)(mt * ((TC_LANES * TC_BK * 4) >> 4)) + k * stepP, hiP), dQ, idesc, !(kb == 0 && k == 0));
        }
        umma_commit(&empty[s]);            // ring slot reusable once these MMAs have read it
        if (kb == nk - 1) umma_commit(&acc_full[0]);
      }
      __syncwarp();
      if (++s == S) { s = 0; ph ^= 1; }
    }
  } else {
    // ===================== drain / epilogue =====================
    const int q = warp & 3;                // TMEM lane quarter this warp may access
    const int row = 32 * q + lane;         // accumulator lane handled by this thread
    const uint32_t tlane = tmem_base + ((uint32_t)(32 * q) << 16);
    // per-thread epilogue side input (e.g. the ReLU mask bits of the dgrad epilogue), fetched while the MMAs are still running
    uint32_t pre[MT * (TN / 32)];
    Pol::pre_epilogue(prm, tl, row, pre);
    if (nk > 0) {
      mbar_wait(&acc_full[0], 0);
      tc_fence_after();
#pragma unroll
      for (int mt = 0; mt < MT; mt++) {
#pragma unroll
        for (int c0 = 0; c0 < TN; c0 += 32) {
          float v[32];
          tmem_ld32(tlane + (uint32_t)(mt * TN + c0), v);
          tmem_ld_wait();
          Pol::store(prm, tl, mt, row, c0, v, pre[mt * (TN / 32) + c0 / 32]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------ 3xTF32 (f32-faithful) persistent kernel
// Warps: 0 TMA producer, 1 MMA issuer, 2-5 splitters, 6-9 drain.  Persistent over the logical grid like tc_tile_persist_kernel; the two
// TMEM accumulator sets ([main | cross], 2 TN columns each) ping-pong per CHUNK (TC_KC k-blocks), the chunk counter runs on across tiles, and the drain warps keep the
// tile's fp32 partial sums in registers (TN <= 128) until its last chunk, then run the Policy's epilogue while the issuer already works
// on the next tile's first two chunks.
template <class Pol, class = void> struct tc_p_presplit { static constexpr bool value = false; };
template <class Pol> struct tc_p_presplit<Pol, std::enable_if_t<Pol::P_PRESPLIT || !Pol::P_PRESPLIT>> { static constexpr bool value = Pol::P_PRESPLIT; };
template <class Pol>
__global__ void __launch_bounds__(320, 1) tc_tile_split_kernel(const __grid_constant__ typename Pol::Params prm, const uint3 lgrid, const int kc /* k-blocks per TMEM chunk */, const int poll /* 1: one lane per warp polls */, const int nstages) {
  constexpr int TN = Pol::TN; constexpr bool P_MN = Pol::P_MN, Q_MN = Pol::Q_MN;
  static_assert(Pol::SPLIT && Pol::MT == 1 && TN <= 128, "3xTF32: one M-tile, TN fp32 partial sums per drain thread");
  constexpr bool QPRE = Pol::Q_PRESPLIT;                 // Q's lo plane comes from global memory (second TMA load), only P is split here
  constexpr bool PPRE = tc_p_presplit<Pol>::value;       // P's planes too (Pol::load_p_lo): the splitter warps idle and the issuer waits for the TMA itself
  static_assert(!PPRE || QPRE, "P pre-split implies Q pre-split");
  using Cfg = TcCfg<TN, true, 1, 1>;
  constexpr int SMAX = Cfg::STAGES;
  const int S = nstages > 0 && nstages < SMAX ? nstages : SMAX;
  // stage = [P | Q | Q_lo | P_lo]: Q_lo directly behind Q, so ONE N = 2 TN instruction computes P.[Q | Q_lo] (the hi*hi sums and one cross term
  // side by side in TMEM) and P is read from shared memory twice per k-step instead of three times — this mode is bound by shared-memory bandwidth
  constexpr int QLO_OFF = Cfg::P_BYTES + Cfg::Q_BYTES, PLO_OFF = Cfg::P_BYTES + 2 * Cfg::Q_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + SMAX * Cfg::STAGE_BYTES);
  uint64_t* full = bars; uint64_t* ready = bars + SMAX; uint64_t* empty = bars + 2 * SMAX;
  uint64_t* acc_full = bars + 3 * SMAX; uint64_t* acc_empty = bars + 3 * SMAX + 2;
  uint32_t* tmem_slot = (uint32_t*)(bars + 3 * SMAX + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t total = lgrid.x * lgrid.y * lgrid.z;

  if (warp == 0 && lane == 0) {
    Pol::prefetch(prm);
    for (int s = 0; s < S; s++) { mbar_init(&full[s], 1); mbar_init(&ready[s], 128); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; b++) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 128); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 4 * TN); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  auto blk_of = [&](uint32_t t) { uint3 b; b.x = t % lgrid.x; const uint32_t r = t / lgrid.x; b.y = r % lgrid.y; b.z = r / lgrid.y; return b; };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (uint32_t t = blockIdx.x; t < total; t += gridDim.x) {
        const typename Pol::Tile tl = Pol::tile(prm, blk_of(t));
        const int nk = Pol::num_kblocks(prm, tl);
        for (int kb = 0; kb < nk; kb++) {
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* st = smem + s * Cfg::STAGE_BYTES;
          mbar_expect_tx(&full[s], Pol::p_bytes(prm, Cfg::P_BYTES) * (PPRE ? 2 : 1) + Cfg::Q_BYTES * (QPRE ? 2 : 1));
          Pol::load(prm, tl, kb, st, st + Cfg::P_BYTES, &full[s]);
          if constexpr (QPRE) Pol::load_q_lo(prm, tl, kb, st + QLO_OFF, &full[s]);
          if constexpr (PPRE) Pol::load_p_lo(prm, tl, kb, st + PLO_OFF, &full[s]);
          if (++s == S) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (converged warp, elected lane; see tc_tile_kernel) =====================
    constexpr uint32_t idesc = umma_idesc_tf32(TC_LANES, TN, P_MN ? 1 : 0, Q_MN ? 1 : 0);            // P_lo . Q
    constexpr uint32_t idesc2 = umma_idesc_tf32(TC_LANES, 2 * TN, P_MN ? 1 : 0, Q_MN ? 1 : 0);       // P . [Q | Q_lo]
    const MnDescCfg mnc = prm.mnc;
    const uint32_t hiK = (1024u >> 4) | (1u << 14) | (2u << 29), loK = (16u >> 4) << 16, stepK = 32u >> 4;
    const uint32_t hiM = (mnc.sbo >> 4) | (1u << 14) | (mnc.layout << 29), loM = (mnc.lbo >> 4) << 16, stepM = mnc.kadv >> 4;
    const uint32_t hiP = P_MN ? hiM : hiK, loP = P_MN ? loM : loK, stepP = P_MN ? stepM : stepK;
    const uint32_t hiQ = Q_MN ? hiM : hiK, loQ = Q_MN ? loM : loK, stepQ = Q_MN ? stepM : stepK;
    const uint32_t smem0 = smem_u32(smem) >> 4;
    int s = 0; uint32_t ph = 0, ch = 0;                  // ch: chunks issued so far by this CTA (all tiles)
    for (uint32_t t = blockIdx.x; t < total; t += gridDim.x) {
      const typename Pol::Tile tl = Pol::tile(prm, blk_of(t));
      const int nk = Pol::num_kblocks(prm, tl);
      if (nk <= 0) continue;
      // TMEM: two accumulator sets of 2 TN columns, ping-ponged per chunk: [hi*hi | hi*lo + lo*hi].  The cross-term sums are 2^-11 of the
      // magnitude; they are promoted to the fp32 registers with their chunk like the main sums.
      for (int kb = 0; kb < nk; kb++) {
        const bool first = (kb % kc) == 0, last = (kb % kc) == kc - 1 || kb == nk - 1;
        const uint32_t buf = ch & 1;
        if (first) { mbar_wait(&acc_empty[buf], ((ch >> 1) & 1) ^ 1); tc_fence_after(); }
        mbar_wait(PPRE ? &full[s] : &ready[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t tacc = tmem_base + buf * (uint32_t)(2 * TN);
          const uint32_t st = smem0 + (uint32_t)s * (Cfg::STAGE_BYTES >> 4);
          const uint32_t aP = st + loP, aQ = st + (Cfg::P_BYTES >> 4) + loQ, aPl = aP + (PLO_OFF >> 4);
#pragma unroll
          for (int k = 0; k < TC_BK / 8; k++) {
            const uint64_t dP = umma_desc_pack(aP + k * stepP, hiP), dQ = umma_desc_pack(aQ + k * stepQ, hiQ), dPl = umma_desc_pack(aPl + k * stepP, hiP);
            umma_tf32(tacc, dP, dQ, idesc2, !(first && k == 0));          // columns [0, TN): P.Q, [TN, 2 TN): P.Q_lo (Q_lo's rows follow Q's at the same pitch)
            umma_tf32(tacc + (uint32_t)TN, dPl, dQ, idesc, 1);            // += P_lo.Q
          }
          umma_commit(&empty[s]);
          if (last) umma_commit(&acc_full[buf]);
        }
        __syncwarp();
        if (last) ch++;
        if (++s == S) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp < 6) {
    // ===================== splitters =====================
    const int tid = threadIdx.x - 64;        // 0..127
    int s = 0; uint32_t ph = 0;
    for (uint32_t t = blockIdx.x; !PPRE && t < total; t += gridDim.x) {
      const typename Pol::Tile tl = Pol::tile(prm, blk_of(t));
      const int nk = Pol::num_kblocks(prm, tl);
      for (int kb = 0; kb < nk; kb++) {
        if (poll) mbar_wait_warp(&full[s], ph); else mbar_wait(&full[s], ph);
        float4* hi = (float4*)(smem + s * Cfg::STAGE_BYTES);
        float4* plo = (float4*)(smem + s * Cfg::STAGE_BYTES + PLO_OFF);
        constexpr int NP4 = Cfg::P_BYTES / 16, NQ4 = Cfg::Q_BYTES / 16;
#pragma unroll 4
        for (int i = tid; i < NP4; i += 128) {             // P: one read, one write (the hardware truncates the raw tile to hi)
          const float4 x = hi[i]; float4 l;
          l.x = tf32_rna(x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u));
          l.y = tf32_rna(x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u));
          l.z = tf32_rna(x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u));
          l.w = tf32_rna(x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u));
          plo[i] = l;
        }
        if constexpr (!QPRE) {
          // Q: rounded hi written back + signed lo.  P's lo always has the sign of its value (truncation), so with a truncated Q the dropped
          // lo*lo term would have the sign of the product — a coherent 2^-22 bias toward zero; a round-to-nearest split on ONE side removes it.
#pragma unroll 4
          for (int i = NP4 + tid; i < NP4 + NQ4; i += 128) {          // Q_lo sits NQ4 float4 behind Q
            const float4 x = hi[i]; float4 h, l;
            h.x = tf32_rna(x.x); h.y = tf32_rna(x.y); h.z = tf32_rna(x.z); h.w = tf32_rna(x.w);
            l.x = tf32_rna(x.x - h.x); l.y = tf32_rna(x.y - h.y); l.z = tf32_rna(x.z - h.z); l.w = tf32_rna(x.w - h.w);
            hi[i] = h; hi[i + NQ4] = l;
          }
        }
        fence_proxy_async();               // generic-proxy writes -> visible to the tensor core (async proxy)
        mbar_arrive(&ready[s]);
        if (++s == S) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ===================== drain / epilogue =====================
    const int q = warp & 3;
    const int row = 32 * q + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(32 * q) << 16);
    uint32_t ch = 0;
    for (uint32_t t = blockIdx.x; t < total; t += gridDim.x) {
      const typename Pol::Tile tl = Pol::tile(prm, blk_of(t));
      const int nk = Pol::num_kblocks(prm, tl);
      if (nk <= 0) continue;
      uint32_t pre[TN / 32];
      Pol::pre_epilogue(prm, tl, row, pre);
      float racc[TN];
#pragma unroll
      for (int j = 0; j < TN; j++) racc[j] = 0.0f;
      const int nchunks = (nk + kc - 1) / kc;
      for (int c = 0; c < nchunks; c++, ch++) {
        const uint32_t buf = ch & 1;
        if (poll) mbar_wait_warp(&acc_full[buf], (ch >> 1) & 1); else mbar_wait(&acc_full[buf], (ch >> 1) & 1);
        tc_fence_after();
        const uint32_t tacc = tlane + buf * (uint32_t)(2 * TN);
#pragma unroll
        for (int c0 = 0; c0 < 2 * TN; c0 += 32) {        // the chunk's hi*hi sums, then its cross-term sums: both added to the fp32 registers
          float v[32];
          tmem_ld32(tacc + (uint32_t)c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j++) racc[(c0 & (TN - 1)) + j] += v[j];
        }
        tc_fence_before();
        mbar_arrive(&acc_empty[buf]);
      }
#pragma unroll
      for (int c0 = 0; c0 < TN; c0 += 32) Pol::store(prm, tl, 0, row, c0, &racc[c0], pre[c0 / 32]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 4 * TN);
}

// ------------------------------------------------------------------------------------------------ persistent variant
// Single-pass (non-SPLIT) tiles only.  CTAs (OCC per SM) walk the logical grid; when two accumulator sets fit in TMEM
// (2 * MT * TN <= 512 columns) the drain / epilogue of tile t overlaps the MMAs of tile t+1, and the per-CTA prologue (barrier
// init, TMEM allocation, descriptor prefetch, pipeline fill) is paid once per SM instead of once per tile.  Measured on the
// one-tile-per-CTA kernel with TN = 256 (one CTA per SM, nothing to overlap with): 17 % of a CTA's life is the epilogue and
// ~10 % the prologue (profiles/ncu_conv_full_r1.csv: tensor pipe 53 % active).
template <class Pol>
__global__ void __launch_bounds__(TcCfg<Pol::TN, false, Pol::OCC, Pol::MT>::THREADS, Pol::OCC)
tc_tile_persist_kernel(const __grid_constant__ typename Pol::Params prm, const uint3 lgrid) {
  constexpr int TN = Pol::TN; constexpr bool P_MN = Pol::P_MN, Q_MN = Pol::Q_MN;
  constexpr int MT = Pol::MT;
  static_assert(!Pol::SPLIT, "the persistent kernel serves the single-pass mode");
  using Cfg = TcCfg<TN, false, Pol::OCC, MT>;
  constexpr int S = Cfg::STAGES;
  constexpr int NACC = (2 * MT * TN * Pol::OCC <= 512) ? 2 : 1;       // accumulator sets per CTA (all co-resident CTAs share 512 columns)
  constexpr int TCOLS = NACC * MT * TN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + S * Cfg::STAGE_BYTES);
  uint64_t* full = bars; uint64_t* empty = bars + S;
  uint64_t* acc_full = bars + 2 * S; uint64_t* acc_empty = bars + 2 * S + 2;
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * S + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t total = lgrid.x * lgrid.y * lgrid.z;

  if (warp == 0 && lane == 0) {
    Pol::prefetch(prm);
    for (int s = 0; s < S; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; b++) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 128); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, TCOLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  auto blk_of = [&](uint32_t t) { uint3 b; b.x = t % lgrid.x; const uint32_t r = t / lgrid.x; b.y = r % lgrid.y; b.z = r / lgrid.y; return b; };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (uint32_t t = blockIdx.x; t < total; t += gridDim.x) {
        const typename Pol::Tile tl = Pol::tile(prm, blk_of(t));
        const int nk = Pol::num_kblocks(prm, tl);
        for (int kb = 0; kb < nk; kb++) {
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* st = smem + s * Cfg::STAGE_BYTES;
          mbar_expect_tx(&full[s], Pol::p_bytes(prm, Cfg::P_BYTES) + Cfg::Q_BYTES);
          Pol::load(prm, tl, kb, st, st + Cfg::P_BYTES, &full[s]);
          if (++s == S) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (converged warp, elected lane; see tc_tile_kernel) =====================
    constexpr uint32_t idesc = umma_idesc_tf32(TC_LANES, TN, P_MN ? 1 : 0, Q_MN ? 1 : 0);
    const MnDescCfg mnc = prm.mnc;
    const uint32_t hiK = (1024u >> 4) | (1u << 14) | (2u << 29), loK = (16u >> 4) << 16, stepK = 32u >> 4;
    const uint32_t hiM = (mnc.sbo >> 4) | (1u << 14) | (mnc.layout << 29), loM = (mnc.lbo >> 4) << 16, stepM = mnc.kadv >> 4;
    const uint32_t hiP = P_MN ? hiM : hiK, loP = P_MN ? loM : loK, stepP = P_MN ? stepM : stepK;
    const uint32_t hiQ = Q_MN ? hiM : hiK, loQ = Q_MN ? loM : loK, stepQ = Q_MN ? stepM : stepK;
    const uint32_t smem0 = smem_u32(smem) >> 4;
    int s = 0; uint32_t ph = 0, ti = 0;
    for (uint32_t t = blockIdx.x; t < total; t += gridDim.x) {
      const typename Pol::Tile tl = Pol::tile(prm, blk_of(t));
      const int nk = Pol::num_kblocks(prm, tl);
      if (nk <= 0) continue;
      const uint32_t ab = ti % NACC;
      mbar_wait(&acc_empty[ab], ((ti / NACC) & 1) ^ 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + ab * (uint32_t)(MT * TN);
      for (int kb = 0; kb < nk; kb++) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t st = smem0 + (uint32_t)s * (Cfg::STAGE_BYTES >> 4);
          const uint32_t aP = st + loP, aQ = st + (Cfg::P_BYTES >> 4) + loQ;
#pragma unroll
          for (int k = 0; k < TC_BK / 8; k++) {
            const uint64_t dQ = umma_desc_pack(aQ + k * stepQ, hiQ);
#pragma unroll
            for (int mt = 0; mt < MT; mt++)
              umma_tf32(tacc + (uint32_t)(mt * TN), umma_desc_pack(aP + (uint32_t)(mt * ((TC_LANES * TC_BK * 4) >> 4)) + k * stepP, hiP), dQ, idesc, !(kb == 0 && k == 0));
          }
          umma_commit(&empty[s]);
          if (kb == nk - 1) umma_commit(&acc_full[ab]);
        }
        __syncwarp();
        if (++s == S) { s = 0; ph ^= 1; }
      }
      ti++;
    }
  } else {
    // ===================== drain / epilogue =====================
    const int q = warp & 3;
    const int row = 32 * q + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(32 * q) << 16);
    uint32_t ti = 0;
    for (uint32_t t = blockIdx.x; t < total; t += gridDim.x) {
      const typename Pol::Tile tl = Pol::tile(prm, blk_of(t));
      const int nk = Pol::num_kblocks(prm, tl);
      if (nk <= 0) continue;
      uint32_t pre[MT * (TN / 32)];
      Pol::pre_epilogue(prm, tl, row, pre);
      const uint32_t ab = ti % NACC;
      mbar_wait(&acc_full[ab], (ti / NACC) & 1);
      tc_fence_after();
      {   // software-pipelined drain (see tc_tile_pair_kernel): chunk i + 1 is loaded from TMEM while chunk i is stored
        constexpr int NCH = MT * (TN / 32);
        const uint32_t tb = tlane + ab * (uint32_t)(MT * TN);
        float va[32], vb[32];
        tmem_ld32(tb, va);
#pragma unroll
        for (int i = 0; i < NCH; i += 2) {
          tmem_ld_wait();
          if (i + 1 < NCH) tmem_ld32(tb + (uint32_t)((i + 1) * 32), vb);
          else { tc_fence_before(); mbar_arrive(&acc_empty[ab]); }
          Pol::store(prm, tl, i / (TN / 32), row, (i % (TN / 32)) * 32, va, pre[i]);
          if (i + 1 < NCH) {
            tmem_ld_wait();
            if (i + 2 < NCH) tmem_ld32(tb + (uint32_t)((i + 2) * 32), va);
            else { tc_fence_before(); mbar_arrive(&acc_empty[ab]); }
            Pol::store(prm, tl, (i + 1) / (TN / 32), row, ((i + 1) % (TN / 32)) * 32, vb, pre[i + 1]);
          }
        }
      }
      ti++;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TCOLS);
}

// ------------------------------------------------------------------------------------------------ CTA-pair variant (cta_group::2)
// Single-pass (TF32) tiles of TN = 256: a cluster of two CTAs (the two SMs of a TPC) owns TWO adjacent M-tiles of the logical grid
// (blk.x = 2 * pair_tile + rank) and runs them as one M = 256 MMA.  Each CTA loads its own 128 P rows and HALF of the Q rows (Pol::load2:
// TN / 2 rows at rank * TN / 2), so the Q bytes a CTA pulls through L2 -> shared memory per output halve and a stage is 32 KB (6-deep ring
// with two 256-column TMEM accumulator sets per CTA) — the one-CTA TN = 256 kernel moves 48 KB per 4 MMAs (94 B/clk against the ~50 B/clk
// an SM can ingest; measured 65-72 % tensor-pipe activity).  The leader issues; both producers' TMA bytes land on the leader's `full`
// barrier; commits are multicast to both CTAs' `empty` / `acc_full`; both CTAs' drain warps release the leader's `acc_empty`.
template <class Pol>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
tc_tile_pair_kernel(const __grid_constant__ typename Pol::Params prm, const uint3 lgrid) {
  constexpr int TN = Pol::TN; constexpr bool P_MN = Pol::P_MN, Q_MN = Pol::Q_MN;
  static_assert(!Pol::SPLIT && Pol::MT == 1 && TN == 256, "CTA pairs serve the single-pass 256-wide tiles");
  constexpr int P_BYTES = TC_LANES * TC_BK * 4, QH_BYTES = (TN / 2) * TC_BK * 4, STAGE = P_BYTES + QH_BYTES;
  constexpr int S = 6;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + S * STAGE);
  uint64_t* full = bars; uint64_t* empty = bars + S;
  uint64_t* acc_full = bars + 2 * S; uint64_t* acc_empty = bars + 2 * S + 2;
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * S + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t xpairs = (lgrid.x + 1) / 2, total = xpairs * lgrid.y * lgrid.z;
  const uint32_t pair_id = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    Pol::prefetch(prm);
    for (int s = 0; s < S; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; b++) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 256); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc2(tmem_slot, 512); tmem_relinquish2(); }
  tc_fence_before();
  cluster_sync_all();                // barriers of BOTH CTAs are initialised before anybody signals across
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  auto blk_of = [&](uint32_t t) { uint3 b; b.x = 2 * (t % xpairs) + rank; const uint32_t r = t / xpairs; b.y = r % lgrid.y; b.z = r / lgrid.y; return b; };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (uint32_t t = pair_id; t < total; t += npairs) {
        const typename Pol::Tile tl = Pol::tile(prm, blk_of(t));
        const int nk = Pol::num_kblocks(prm, tl);
        for (int kb = 0; kb < nk; kb++) {
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* st = smem + s * STAGE;
          if (rank == 0) mbar_expect_tx(&full[s], 2 * (Pol::p_bytes(prm, P_BYTES) + QH_BYTES));
          Pol::load2(prm, tl, kb, st, st + P_BYTES, mapa_shared(smem_u32(&full[s]), 0), (int)rank);
          if (++s == S) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(2 * TC_LANES, TN, P_MN ? 1 : 0, Q_MN ? 1 : 0);
      const MnDescCfg mnc = prm.mnc;
      const uint32_t hiK = (1024u >> 4) | (1u << 14) | (2u << 29), loK = (16u >> 4) << 16, stepK = 32u >> 4;
      const uint32_t hiM = (mnc.sbo >> 4) | (1u << 14) | (mnc.layout << 29), loM = (mnc.lbo >> 4) << 16, stepM = mnc.kadv >> 4;
      const uint32_t hiP = P_MN ? hiM : hiK, loP = P_MN ? loM : loK, stepP = P_MN ? stepM : stepK;
      const uint32_t hiQ = Q_MN ? hiM : hiK, loQ = Q_MN ? loM : loK, stepQ = Q_MN ? stepM : stepK;
      const uint32_t smem0 = smem_u32(smem) >> 4;
      int s = 0; uint32_t ph = 0, ti = 0;
      for (uint32_t t = pair_id; t < total; t += npairs) {
        const typename Pol::Tile tl = Pol::tile(prm, blk_of(t));
        const int nk = Pol::num_kblocks(prm, tl);
        if (nk <= 0) continue;
        const uint32_t ab = ti & 1;
        mbar_wait(&acc_empty[ab], ((ti >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + ab * (uint32_t)TN;
        for (int kb = 0; kb < nk; kb++) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t st = smem0 + (uint32_t)s * (STAGE >> 4);
            const uint32_t aP = st + loP, aQ = st + (P_BYTES >> 4) + loQ;
#pragma unroll
            for (int k = 0; k < TC_BK / 8; k++)
              umma_tf32_2cta(tacc, umma_desc_pack(aP + k * stepP, hiP), umma_desc_pack(aQ + k * stepQ, hiQ), idesc, !(kb == 0 && k == 0));
            umma_commit_2cta(&empty[s], 3);
            if (kb == nk - 1) umma_commit_2cta(&acc_full[ab], 3);
          }
          __syncwarp();
          if (++s == S) { s = 0; ph ^= 1; }
        }
        ti++;
      }
    }
  } else {
    // ===================== drain / epilogue (both CTAs, each its own 128 lanes) =====================
    const int q = warp & 3;
    const int row = 32 * q + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(32 * q) << 16);
    const uint32_t rel0 = mapa_shared(smem_u32(&acc_empty[0]), 0), rel1 = mapa_shared(smem_u32(&acc_empty[1]), 0);
    uint32_t ti = 0;
    for (uint32_t t = pair_id; t < total; t += npairs) {
      const typename Pol::Tile tl = Pol::tile(prm, blk_of(t));
      const int nk = Pol::num_kblocks(prm, tl);
      if (nk <= 0) continue;
      uint32_t pre[TN / 32];
      Pol::pre_epilogue(prm, tl, row, pre);
      const uint32_t ab = ti & 1;
      mbar_wait(&acc_full[ab], (ti >> 1) & 1);
      tc_fence_after();
      // software-pipelined drain: the tcgen05.ld of chunk c+1 is in flight while chunk c is stored (ncu on the 128 -> 256 layer: the drain warps were
      // 94 % busy — the tile's critical path with only 36 k-blocks per tile — and a third of that was LDTM latency), and the accumulator set goes
      // back to the MMA issuer as soon as its last chunk is in registers, before that chunk's stores
      {
        static_assert(TN % 64 == 0, "pair tiles are drained two chunks at a time");
        const uint32_t tb = tlane + ab * (uint32_t)TN;
        float va[32], vb[32];
        tmem_ld32(tb, va);
#pragma unroll
        for (int c0 = 0; c0 < TN; c0 += 64) {
          tmem_ld_wait();
          tmem_ld32(tb + (uint32_t)(c0 + 32), vb);
          Pol::store(prm, tl, 0, row, c0, va, pre[c0 / 32]);
          tmem_ld_wait();
          if (c0 + 64 < TN) tmem_ld32(tb + (uint32_t)(c0 + 64), va);
          else { tc_fence_before(); mbar_arrive_cluster(ab ? rel1 : rel0); }
          Pol::store(prm, tl, 0, row, c0 + 32, vb, pre[c0 / 32 + 1]);
        }
      }
      ti++;
    }
  }
  tc_fence_before();
  cluster_sync_all();                // nobody leaves (or frees TMEM) while the peer may still read its shared memory / signal its barriers
  if (warp == 1) tmem_dealloc2(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------ 3xTF32 on CTA pairs
// The f32-faithful mode is bound by shared-memory bandwidth (an N = 128 MMA reads 128 B/clk, the splitters add to it).  A CTA pair
// running M = 256 x N = 128 MMAs reads, per CTA and MMA, its own 128 P rows (4 KB) but only 64 Q rows (2 KB): 96 instead of 128 B/clk,
// which leaves room for the splitters' P traffic.  Requires Pol::Q_PRESPLIT (both Q planes arrive by TMA, nothing to split on that side).
// Warps per CTA: 0 producer, 1 issuer (leader only), 2-5 splitters, 6-9 drain.  Each CTA's loads complete on its OWN `full` barrier; its
// splitters write P_lo and then one thread arrives on the LEADER's `ready` barrier (count 2), so `ready` = both CTAs' tiles landed and split.
template <class Pol>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
tc_tile_split_pair_kernel(const __grid_constant__ typename Pol::Params prm, const uint3 lgrid, const int kc) {
  constexpr int TN = Pol::TN; constexpr bool P_MN = Pol::P_MN, Q_MN = Pol::Q_MN;
  static_assert(Pol::SPLIT && Pol::Q_PRESPLIT && Pol::MT == 1 && TN == 128, "3xTF32 CTA pairs: pre-split Q, 128-wide tiles");
  constexpr int P_BYTES = TC_LANES * TC_BK * 4, QH_BYTES = (TN / 2) * TC_BK * 4;
  constexpr int STAGE = 2 * P_BYTES + 2 * QH_BYTES;          // [P | Qh | P_lo | Qh_lo]
  constexpr int LO_OFF = P_BYTES + QH_BYTES;
  constexpr int S = 4;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + S * STAGE);
  uint64_t* full = bars; uint64_t* ready = bars + S; uint64_t* empty = bars + 2 * S;
  uint64_t* acc_full = bars + 3 * S; uint64_t* acc_empty = bars + 3 * S + 2;
  uint32_t* tmem_slot = (uint32_t*)(bars + 3 * S + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t xpairs = (lgrid.x + 1) / 2, total = xpairs * lgrid.y * lgrid.z;
  const uint32_t pair_id = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    Pol::prefetch(prm);
    for (int s = 0; s < S; s++) { mbar_init(&full[s], 1); mbar_init(&ready[s], 2); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; b++) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 256); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc2(tmem_slot, 512); tmem_relinquish2(); }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  auto blk_of = [&](uint32_t t) { uint3 b; b.x = 2 * (t % xpairs) + rank; const uint32_t r = t / xpairs; b.y = r % lgrid.y; b.z = r / lgrid.y; return b; };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs, own barrier) =====================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (uint32_t t = pair_id; t < total; t += npairs) {
        const typename Pol::Tile tl = Pol::tile(prm, blk_of(t));
        const int nk = Pol::num_kblocks(prm, tl);
        for (int kb = 0; kb < nk; kb++) {
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* st = smem + s * STAGE;
          mbar_expect_tx(&full[s], Pol::p_bytes(prm, P_BYTES) + 2 * QH_BYTES);
          Pol::load_sp2(prm, tl, kb, st, st + P_BYTES, st + LO_OFF + P_BYTES, &full[s], (int)rank);
          if (++s == S) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader only) =====================
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(2 * TC_LANES, TN, P_MN ? 1 : 0, Q_MN ? 1 : 0);
      const MnDescCfg mnc = prm.mnc;
      const uint32_t hiK = (1024u >> 4) | (1u << 14) | (2u << 29), loK = (16u >> 4) << 16, stepK = 32u >> 4;
      const uint32_t hiM = (mnc.sbo >> 4) | (1u << 14) | (mnc.layout << 29), loM = (mnc.lbo >> 4) << 16, stepM = mnc.kadv >> 4;
      const uint32_t hiP = P_MN ? hiM : hiK, loP = P_MN ? loM : loK, stepP = P_MN ? stepM : stepK;
      const uint32_t hiQ = Q_MN ? hiM : hiK, loQ = Q_MN ? loM : loK, stepQ = Q_MN ? stepM : stepK;
      const uint32_t smem0 = smem_u32(smem) >> 4;
      int s = 0; uint32_t ph = 0, ch = 0, ti = 0;
      for (uint32_t t = pair_id; t < total; t += npairs) {
        const typename Pol::Tile tl = Pol::tile(prm, blk_of(t));
        const int nk = Pol::num_kblocks(prm, tl);
        if (nk <= 0) continue;
        const uint32_t tcross = tmem_base + (uint32_t)((2 + (ti & 1)) * TN);      // TMEM: [main 0 | main 1 | cross 0 | cross 1], see tc_tile_split_kernel
        ti++;
        for (int kb = 0; kb < nk; kb++) {
          const bool first = (kb % kc) == 0, last = (kb % kc) == kc - 1 || kb == nk - 1;
          const uint32_t buf = ch & 1;
          if (first) { mbar_wait(&acc_empty[buf], ((ch >> 1) & 1) ^ 1); tc_fence_after(); }
          mbar_wait(&ready[s], ph);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t tacc = tmem_base + buf * (uint32_t)TN;
            const uint32_t st = smem0 + (uint32_t)s * (STAGE >> 4);
            const uint32_t aP = st + loP, aQ = st + (P_BYTES >> 4) + loQ;
            const uint32_t aPl = aP + (LO_OFF >> 4), aQl = aQ + (LO_OFF >> 4);
#pragma unroll
            for (int k = 0; k < TC_BK / 8; k++) {
              const uint64_t dP = umma_desc_pack(aP + k * stepP, hiP), dQ = umma_desc_pack(aQ + k * stepQ, hiQ);
              const uint64_t dPl = umma_desc_pack(aPl + k * stepP, hiP), dQl = umma_desc_pack(aQl + k * stepQ, hiQ);
              umma_tf32_2cta(tcross, dPl, dQ, idesc, !(kb == 0 && k == 0));
              umma_tf32_2cta(tcross, dP, dQl, idesc, 1);
              umma_tf32_2cta(tacc, dP, dQ, idesc, !(first && k == 0));
            }
            umma_commit_2cta(&empty[s], 3);
            if (last) umma_commit_2cta(&acc_full[buf], 3);
          }
          __syncwarp();
          if (last) ch++;
          if (++s == S) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp < 6) {
    // ===================== splitters (both CTAs): P_lo = rna_tf32(P - trunc_tf32(P)) =====================
    const int tid = threadIdx.x - 64;
    const uint32_t ready0 = mapa_shared(smem_u32(&ready[0]), 0);
    int s = 0; uint32_t ph = 0;
    for (uint32_t t = pair_id; t < total; t += npairs) {
      const typename Pol::Tile tl = Pol::tile(prm, blk_of(t));
      const int nk = Pol::num_kblocks(prm, tl);
      for (int kb = 0; kb < nk; kb++) {
        mbar_wait(&full[s], ph);
        const float4* hi = (const float4*)(smem + s * STAGE);
        float4* lo = (float4*)(smem + s * STAGE + LO_OFF);
#pragma unroll 4
        for (int i = tid; i < P_BYTES / 16; i += 128) {
          const float4 x = hi[i]; float4 l;
          l.x = tf32_rna(x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u));
          l.y = tf32_rna(x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u));
          l.z = tf32_rna(x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u));
          l.w = tf32_rna(x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u));
          lo[i] = l;
        }
        fence_proxy_async();
        asm volatile("bar.sync 1, 128;" ::: "memory");          // the four splitter warps of this CTA
        if (tid == 0) mbar_arrive_cluster(ready0 + (uint32_t)s * 8u);
        if (++s == S) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ===================== drain / epilogue (both CTAs) =====================
    const int q = warp & 3;
    const int row = 32 * q + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(32 * q) << 16);
    const uint32_t rel0 = mapa_shared(smem_u32(&acc_empty[0]), 0), rel1 = mapa_shared(smem_u32(&acc_empty[1]), 0);
    uint32_t ch = 0, ti = 0;
    for (uint32_t t = pair_id; t < total; t += npairs) {
      const typename Pol::Tile tl = Pol::tile(prm, blk_of(t));
      const int nk = Pol::num_kblocks(prm, tl);
      if (nk <= 0) continue;
      uint32_t pre[TN / 32];
      Pol::pre_epilogue(prm, tl, row, pre);
      float racc[TN];
#pragma unroll
      for (int j = 0; j < TN; j++) racc[j] = 0.0f;
      const int nchunks = (nk + kc - 1) / kc;
      const uint32_t tcross = tlane + (uint32_t)((2 + (ti & 1)) * TN);
      ti++;
      for (int c = 0; c < nchunks; c++, ch++) {
        const uint32_t buf = ch & 1;
        mbar_wait(&acc_full[buf], (ch >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < TN; c0 += 32) {
          float v[32];
          tmem_ld32(tlane + buf * (uint32_t)TN + (uint32_t)c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j++) racc[c0 + j] += v[j];
        }
        if (c == nchunks - 1) {
#pragma unroll
          for (int c0 = 0; c0 < TN; c0 += 32) {
            float v[32];
            tmem_ld32(tcross + (uint32_t)c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j++) racc[c0 + j] += v[j];
          }
        }
        tc_fence_before();
        mbar_arrive_cluster(buf ? rel1 : rel0);
      }
#pragma unroll
      for (int c0 = 0; c0 < TN; c0 += 32) Pol::store(prm, tl, 0, row, c0, &racc[c0], pre[c0 / 32]);
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc2(tmem_base, 512);
}

template <class Pol>
static int tc_tile_launch(agb_ctx* ctx, const typename Pol::Params& prm, dim3 grid) {
  const uint64_t total = (uint64_t)grid.x * grid.y * grid.z;
  if constexpr (Pol::SPLIT) {
    using Cfg = TcCfg<Pol::TN, true, 1, 1>;
    static bool attr = false;
    if (!attr) { AGB_CUDA(cudaFuncSetAttribute(tc_tile_split_kernel<Pol>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM)); attr = true; }
    if (total >= (1ull << 31)) return AGB_ERR_UNSUPPORTED;
    const unsigned n = (unsigned)(total < (uint64_t)ctx->sm_count ? total : (uint64_t)ctx->sm_count);
    if (n == 0) return AGB_OK;
    if constexpr (Pol::SPLIT_PAIR2) {     // CTA pairs (pre-split Q, 128-wide tiles): see tc_tile_split_pair_kernel
      // measured (B200): correct, but slower than the one-CTA kernel (GEMM 8192^3: 138 vs 186 useful TFLOP/s; VGG conv layers 125 vs 160-175): the
      // 3xTF32 pipeline is bound by its depth (3-4 stages against a ~4300-clock stage round trip: ncu, AGB_SPLIT_STAGES=2 costs 27 %), and the
      // cross-CTA ready / release handshakes lengthen that round trip.  Kept as an opt-in experiment.
      static const int pair_on = [] { const char* e = getenv("AGB_TC_SPLIT_PAIR"); return (e && e[0] == '1') ? 1 : 0; }();
      const uint64_t ptotal = (uint64_t)((grid.x + 1) / 2) * grid.y * grid.z;
      if (pair_on && grid.x >= 2) {
        constexpr int SMEM2 = 4 * (2 * TC_LANES * TC_BK * 4 + 2 * (Pol::TN / 2) * TC_BK * 4) + 1024 + 256;
        static bool attr3 = false;
        if (!attr3) { AGB_CUDA(cudaFuncSetAttribute(tc_tile_split_pair_kernel<Pol>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2)); attr3 = true; }
        static const int kc2 = [] { const char* e = getenv("AGB_SPLIT_KC"); int v = e ? atoi(e) : TC_KC; return v < 1 ? 1 : v; }();
        const uint64_t cap = (uint64_t)(ctx->sm_count / 2);
        const unsigned np = (unsigned)(ptotal < cap ? ptotal : cap);
        tc_tile_split_pair_kernel<Pol><<<2 * np, 320, SMEM2, ctx->stream>>>(prm, make_uint3(grid.x, grid.y, grid.z), kc2);
        AGB_LAUNCHED(ctx);
        return AGB_OK;
      }
    }
    static const int kc = [] { const char* e = getenv("AGB_SPLIT_KC"); int v = e ? atoi(e) : TC_KC; return v < 1 ? 1 : v; }();      // tuning knob: k-blocks per TMEM accumulation chunk
    static const int poll = [] { const char* e = getenv("AGB_SPLIT_POLL"); return e ? atoi(e) : 1; }();
    static const int nst = [] { const char* e = getenv("AGB_SPLIT_STAGES"); return e ? atoi(e) : 0; }();
    tc_tile_split_kernel<Pol><<<n, 320, Cfg::SMEM, ctx->stream>>>(prm, make_uint3(grid.x, grid.y, grid.z), kc, poll, nst);
    AGB_LAUNCHED(ctx);
    return AGB_OK;
  } else {
    using Cfg = TcCfg<Pol::TN, false, Pol::OCC, Pol::MT>;
    if constexpr (Pol::PAIR2) {           // CTA pairs (cta_group::2): the policy provides load2 and half-Q tensor maps
      static const int pair_on = [] { const char* e = getenv("AGB_TC_PAIR"); return (e && e[0] == '0') ? 0 : 1; }();
      const uint64_t ptotal = (uint64_t)((grid.x + 1) / 2) * grid.y * grid.z;
      if (pair_on && ptotal < (1ull << 31) && ptotal >= 1) {
        constexpr int SMEM2 = 6 * (TC_LANES * TC_BK * 4 + (Pol::TN / 2) * TC_BK * 4) + 1024 + 256;
        static bool attr3 = false;
        if (!attr3) { AGB_CUDA(cudaFuncSetAttribute(tc_tile_pair_kernel<Pol>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2)); attr3 = true; }
        const uint64_t cap = (uint64_t)(ctx->sm_count / 2);
        const unsigned np = (unsigned)(ptotal < cap ? ptotal : cap);
        tc_tile_pair_kernel<Pol><<<2 * np, 192, SMEM2, ctx->stream>>>(prm, make_uint3(grid.x, grid.y, grid.z));
        AGB_LAUNCHED(ctx);
        return AGB_OK;
      }
    }
    static int persist = -1;
    if (persist < 0) { const char* e = getenv("AGB_TC_PERSIST"); persist = (e && e[0] == '0') ? 0 : 1; }
    if (persist && total < (1ull << 31)) {
      static bool attr2 = false;
      if (!attr2) { AGB_CUDA(cudaFuncSetAttribute(tc_tile_persist_kernel<Pol>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM)); attr2 = true; }
      const uint64_t cap = (uint64_t)ctx->sm_count * Pol::OCC;
      const unsigned n = (unsigned)(total < cap ? total : cap);
      tc_tile_persist_kernel<Pol><<<n, Cfg::THREADS, Cfg::SMEM, ctx->stream>>>(prm, make_uint3(grid.x, grid.y, grid.z));
      AGB_LAUNCHED(ctx);
      return AGB_OK;
    }
    static bool attr = false;
    if (!attr) { AGB_CUDA(cudaFuncSetAttribute(tc_tile_kernel<Pol>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM)); attr = true; }
    tc_tile_kernel<Pol><<<grid, Cfg::THREADS, Cfg::SMEM, ctx->stream>>>(prm);
    AGB_LAUNCHED(ctx);
    return AGB_OK;
  }
}
