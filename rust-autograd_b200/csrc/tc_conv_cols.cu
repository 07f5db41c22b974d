// tc_conv_cols.cu — persistent, window-reusing tcgen05 implicit-GEMM convolution for NARROW feature maps (output width 16, 32 or
// 64, stride 1, Cout <= 128, channels-last).  Companion of tc_conv_rows.cu (width >= 128); serves Conv2D fprop and
// Conv2DTranspose / dgrad (flipped filter) with the same epilogues (bias, ReLU, ReLU-mask, per-channel sums).
//
// Why: on these layers the per-tap kernel (ConvFpropPol) is bound by L2 -> shared-memory ingest (it re-reads its 128-pixel input
// window once per tap and its filter slice once per 128 pixels: 288 KB per 32-channel block per 128 output pixels at TN = 128,
// ~13 TB/s chip-wide, tensor pipe 45-55 % busy — profiles/ncu_conv_full_r1.csv).
//
// With an output row narrower than the 128 MMA lanes, one M-tile spans RM = 128 / yw full rows, so the lanes are pixel-linear only
// if the window in shared memory has pitch == yw — no room for a halo column.  Hence kw column-shifted COPIES of the window, each a
// dense [rows][yw pixels][16 channels] TMA box (zero fill = padding) whose x origin is j*d - pad: tap (i, j) of M-tile mt is copy j
// seen through a descriptor shifted by (mt*RM + i*d)*yw pixel rows (any 64-byte row is a legal SWIZZLE_64B start:
// scripts/cuda/umma_shift_probe_sw64.cu).  Two M-tiles per CTA (2*RM output rows) share every filter tap tile.
//   L2 -> smem bytes per 32 channels per 128 output pixels (3x3, yw = 64, TN = 128): 3*(2*RM+2)/(2*RM) * 16 KB + 144/2 KB = 146 KB
//   (per-tap kernel: 288 KB).  16-channel k-blocks (64-byte rows) keep two window slots + the filter ring within 227 KB.
// Persistent CTAs, one MMA-issuer warp per M-tile, two TMEM accumulator sets (2 x 2 x TN columns), transposed coalesced epilogue.
//
// Reference semantics: Conv2D::compute conv2d.rs:115-211, Conv2DTranspose::compute conv2d_transpose.rs:89-247.
#include "tc_common.cuh"
#include <stdlib.h>

struct ColsParams {
  CUtensorMap tmX, tmW;
  float* y; const float* bias; const float* mask; const uint32_t* mask_bits; uint32_t* bits_out; float* csum; float* csum_part; int relu;
  int Cout, yh, yw, kw, pad, dil, tiles_y, cblocks, taps, rm, copy_bytes, copy_stride, a_slot_bytes, nb /* filter ring depth, in stages of G taps */;
  long long num_tiles;
};
template <int TN> struct ColsCfg {
  static constexpr int B_BYTES = TN * 64;                      // [TN o][16 c] per tap
  static constexpr int TMEM_COLS = 4 * TN;                     // 2 sets x 2 M-tiles
  static constexpr int THREADS = 256;                          // warps: 0 TMA-A, 1 MMA tile 0, 2-5 epilogue, 6 TMA-B, 7 MMA tile 1
};

// G = filter taps per ring stage.  G = 1: one tap per stage.  G = 3 (kw == 3): one filter ROW per stage, one TMA box {16 c, TN o, 3 taps} — the issuer
// pays one barrier wait + one commit per 6 MMAs instead of per 2.  With N = 64 an MMA pair occupies the tensor pipe for ~67 clk while one trip
// round the issuer's loop (wait, fence, elect, commit, counters) costs more than that: ncu showed the TN = 64 kernel at 34 % tensor-active with
// the issuers 27 % of their time on b_full and the rest in loop overhead.
#define COLS_NB_MAX 16
// MB: the ReLU sign-bit side channel (mask read as bits, sign bits of the output written) — its own instantiation, so the float-mask kernel keeps the
// instruction schedule it had without it (with both forms in one kernel the ordinary masked dgrad lost 30 %)
template <int TN, int G, bool MB, bool MBO>
__global__ void __launch_bounds__(256, 1) conv_cols_kernel(const __grid_constant__ ColsParams p) {
  using Cfg = ColsCfg<TN>;
  constexpr int NB = COLS_NB_MAX;                              // barrier slots; p.nb stages are in use
  constexpr int ST_BYTES = G * Cfg::B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                                          // [2][a_slot_bytes]: kw copies each
  uint8_t* sB = smem + 2 * p.a_slot_bytes;                     // [p.nb][G][TN * 64]
  uint64_t* bars = (uint64_t*)(sB + p.nb * ST_BYTES);
  uint64_t* a_full = bars; uint64_t* a_empty = bars + 2;
  uint64_t* b_full = bars + 4; uint64_t* b_empty = bars + 4 + NB;
  uint64_t* acc_full = bars + 4 + 2 * NB; uint64_t* acc_empty = bars + 6 + 2 * NB;
  uint32_t* tmem_slot = (uint32_t*)(bars + 8 + 2 * NB);
  float* stage = (float*)((uint8_t*)bars + 512);               // [4 warps][32][36] epilogue transpose tiles
  float* csum_s = stage + 4 * 32 * 36;                         // [4 epilogue warps][TN] per-channel sums of this CTA
  for (int i = threadIdx.x; i < 4 * TN; i += blockDim.x) csum_s[i] = 0.0f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmX); tma_prefetch_desc(&p.tmW);
    for (int s = 0; s < 2; s++) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 2); mbar_init(&acc_full[s], 2); mbar_init(&acc_empty[s], 128); }
    for (int s = 0; s < NB; s++) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 2); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int rt = 2 * p.rm;                                      // output rows per tile

  if (warp == 0) {
    // ===================== TMA producer A: kw column-shifted window copies per 16-channel block =====================
    if (lane == 0) {
      uint32_t ai = 0;
      for (long long t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const int b = (int)(t / p.tiles_y), ty = (int)(t - (long long)b * p.tiles_y);
        const int oy0 = ty * rt;
        for (int cb = 0; cb < p.cblocks; cb++, ai++) {
          const uint32_t as = ai & 1;
          mbar_wait(&a_empty[as], ((ai >> 1) & 1) ^ 1);
          mbar_expect_tx(&a_full[as], (uint32_t)(p.kw * p.copy_bytes));
          for (int j = 0; j < p.kw; j++)
            tma_load_4d(sA + as * p.a_slot_bytes + j * p.copy_stride, &p.tmX, &a_full[as], cb * 16, j * p.dil - p.pad, oy0 - p.pad, b);   // dims {c, w, h, b}
        }
      }
    }
  } else if (warp == 6) {
    // ===================== TMA producer B: filter taps =====================
    if (lane == 0) {
      uint32_t bs = 0, bph = 0;
      for (long long t = blockIdx.x; t < p.num_tiles; t += gridDim.x)
        for (int cb = 0; cb < p.cblocks; cb++)
          for (int tap = 0; tap < p.taps; tap += G) {
            mbar_wait(&b_empty[bs], bph ^ 1);
            mbar_expect_tx(&b_full[bs], ST_BYTES);
            tma_load_3d(sB + bs * ST_BYTES, &p.tmW, &b_full[bs], cb * 16, 0, tap);
            if (++bs == (uint32_t)p.nb) { bs = 0; bph ^= 1; }
          }
    }
  } else if (warp == 1 || warp == 7) {
    // ===================== MMA issuers: one warp per M-tile (converged warp, elected lane; see tc_tile.cuh) =====================
    const int mt = warp == 1 ? 0 : 1;
    constexpr uint32_t idesc = umma_idesc_tf32(128, TN, 0, 0);
    const uint32_t hi = (512u >> 4) | (1u << 14) | (4u << 29), lo0 = (16u >> 4) << 16;       // K-major SWIZZLE_64B: 64-byte rows, SBO 512 B
    const uint32_t row4 = (uint32_t)(p.yw * 64) >> 4;                // one window row (yw pixels x 64 B) in 16-byte units
    const uint32_t sA4 = (smem_u32(sA) >> 4) + lo0 + (uint32_t)(mt * p.rm) * row4, sB4 = (smem_u32(sB) >> 4) + lo0;
    const uint32_t di4 = (uint32_t)p.dil * row4, cp4 = (uint32_t)p.copy_stride >> 4;
    uint32_t ai = 0, bs = 0, bph = 0, it = 0;
    for (long long t = blockIdx.x; t < p.num_tiles; t += gridDim.x, it++) {
      const uint32_t acs = it & 1;
      mbar_wait(&acc_empty[acs], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + (acs * 2u + (uint32_t)mt) * (uint32_t)TN;
      for (int cb = 0; cb < p.cblocks; cb++, ai++) {
        const uint32_t as = ai & 1;
        mbar_wait(&a_full[as], (ai >> 1) & 1);
        tc_fence_after();
        const uint32_t a0 = sA4 + as * ((uint32_t)p.a_slot_bytes >> 4);
        uint32_t arow = a0, atap = a0; int j = 0;                      // tap (i, j): copy j, window row i*d (+ this M-tile's first row)
        for (int tap = 0; tap < p.taps; tap += G) {
          uint32_t at[G];
#pragma unroll
          for (int jj = 0; jj < G; jj++) { at[jj] = atap; if (++j == p.kw) { j = 0; arow += di4; atap = arow; } else atap += cp4; }
          mbar_wait(&b_full[bs], bph);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t q0 = sB4 + bs * (uint32_t)(ST_BYTES >> 4);
#pragma unroll
            for (int jj = 0; jj < G; jj++)
#pragma unroll
              for (int ks = 0; ks < 2; ks++)
                umma_tf32(tacc, umma_desc_pack(at[jj] + ks * 2, hi), umma_desc_pack(q0 + jj * (uint32_t)(Cfg::B_BYTES >> 4) + ks * 2, hi), idesc, !(cb == 0 && tap == 0 && jj == 0 && ks == 0));
            umma_commit(&b_empty[bs]);
            if (tap + G >= p.taps) {
              umma_commit(&a_empty[as]);
              if (cb == p.cblocks - 1) umma_commit(&acc_full[acs]);
            }
          }
          __syncwarp();
          if (++bs == (uint32_t)p.nb) { bs = 0; bph ^= 1; }
        }
      }
    }
  } else if (warp < 6) {
    // ===================== epilogue (transposed through shared memory, see tc_conv_rows.cu) =====================
    const int q = warp & 3;
    const uint32_t tlane = tmem_base + ((uint32_t)(32 * q) << 16);
    float* stg = stage + q * (32 * 36);
    const int pl0 = lane >> 3, ch4 = lane & 7;
    float4 cs[TN / 32];
#pragma unroll
    for (int c = 0; c < TN / 32; c++) cs[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t it = 0;
    for (long long t = blockIdx.x; t < p.num_tiles; t += gridDim.x, it++) {
      const int b = (int)(t / p.tiles_y), ty = (int)(t - (long long)b * p.tiles_y);
      const int oy0 = ty * rt;
      // pixel of (M-tile mt, strip position pl): row = oy0 + mt*rm + (32q + pl) / yw, column = (32q + pl) % yw
      uint32_t pre[2][TN / 32];
      if (MB && p.mask_bits != nullptr) {        // sign bits written by the forward kernel: one word per pixel per 32 channels
        const int cw = p.Cout >> 5;
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
#pragma unroll
          for (int c = 0; c < TN / 32; c++) {
            const int cwi = min(c, cw - 1);
            uint32_t wv[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
              const int r = 32 * q + 4 * k + pl0; const int oy = min(oy0 + mt * p.rm + r / p.yw, p.yh - 1), ox = r % p.yw;
              wv[k] = __ldg(p.mask_bits + (((long long)b * p.yh + oy) * p.yw + ox) * cw + cwi);
            }
            uint32_t bits = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) bits |= ((wv[k] >> (4 * ch4)) & 15u) << (4 * k);
            pre[mt][c] = bits;
          }
        }
      } else if (p.mask != nullptr) {
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
#pragma unroll
          for (int c = 0; c < TN / 32; c++) {
            uint32_t bits = 0; const int o = min(32 * c + 4 * ch4, p.Cout - 4);
            float4 mv[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
              const int r = 32 * q + 4 * k + pl0; const int oy = min(oy0 + mt * p.rm + r / p.yw, p.yh - 1), ox = r % p.yw;
              mv[k] = __ldg((const float4*)(p.mask + (((long long)b * p.yh + oy) * p.yw + ox) * p.Cout + o));
            }
#pragma unroll
            for (int k = 0; k < 8; k++)
              bits |= ((mv[k].x > 0.0f ? 1u : 0u) | (mv[k].y > 0.0f ? 2u : 0u) | (mv[k].z > 0.0f ? 4u : 0u) | (mv[k].w > 0.0f ? 8u : 0u)) << (4 * k);
            pre[mt][c] = bits;
          }
        }
      }
      float4 bv[TN / 32];
#pragma unroll
      for (int c = 0; c < TN / 32; c++) {
        const int o = 32 * c + 4 * ch4;
        bv[c] = (p.bias != nullptr && o + 4 <= p.Cout) ? __ldg((const float4*)(p.bias + o)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      const uint32_t acs = it & 1;
      mbar_wait(&acc_full[acs], (it >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int mt = 0; mt < 2; mt++) {
        uint32_t sgw[8][TN / 32];                                   // sign words of this lane group's 8 pixels (bits_out only)
#pragma unroll
        for (int c = 0; c < TN / 32; c++) {
          float v[32];
          tmem_ld32(tlane + (uint32_t)((acs * 2 + mt) * TN + 32 * c), v);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 8; k++) *(float4*)(stg + lane * 36 + 4 * k) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
          __syncwarp();
          const int o = 32 * c + 4 * ch4;
#pragma unroll
          for (int k = 0; k < 8; k++) {
            const int pl = 4 * k + pl0, r = 32 * q + pl;
            const int oy = oy0 + mt * p.rm + r / p.yw, ox = r % p.yw;
            float4 a = *(const float4*)(stg + pl * 36 + 4 * ch4);
            a.x += bv[c].x; a.y += bv[c].y; a.z += bv[c].z; a.w += bv[c].w;
            if (p.relu) { a.x = fmaxf(a.x, 0.0f); a.y = fmaxf(a.y, 0.0f); a.z = fmaxf(a.z, 0.0f); a.w = fmaxf(a.w, 0.0f); }
            if (p.mask != nullptr) {
              const uint32_t m = pre[mt][c] >> (4 * k);
              a.x = (m & 1u) ? a.x : 0.0f * a.x; a.y = (m & 2u) ? a.y : 0.0f * a.y; a.z = (m & 4u) ? a.z : 0.0f * a.z; a.w = (m & 8u) ? a.w : 0.0f * a.w;
            }
            if (oy < p.yh && o + 4 <= p.Cout) {
              *(float4*)(p.y + (((long long)b * p.yh + oy) * p.yw + ox) * p.Cout + o) = a;
              cs[c].x += a.x; cs[c].y += a.y; cs[c].z += a.z; cs[c].w += a.w;
            }
            if (MBO && p.bits_out != nullptr) {       // sign bits of the stored activation: the 8 lanes of a pixel OR their nibbles into one word (warp-uniform branch)
              uint32_t sg = ((a.x > 0.0f ? 1u : 0u) | (a.y > 0.0f ? 2u : 0u) | (a.z > 0.0f ? 4u : 0u) | (a.w > 0.0f ? 8u : 0u)) << (4 * ch4);
              sg |= __shfl_xor_sync(0xffffffffu, sg, 1); sg |= __shfl_xor_sync(0xffffffffu, sg, 2); sg |= __shfl_xor_sync(0xffffffffu, sg, 4);
              sgw[k][c] = sg;
            }
          }
          __syncwarp();
        }
        if (MBO && p.bits_out != nullptr && ch4 == 0) {                    // one 8- / 16-byte store per pixel (all of its words when Cout == TN)
          const int cw = p.Cout >> 5;
#pragma unroll
          for (int k = 0; k < 8; k++) {
            const int r = 32 * q + 4 * k + pl0; const int oy = oy0 + mt * p.rm + r / p.yw, ox = r % p.yw;
            if (oy < p.yh) {
              uint32_t* d = p.bits_out + (((long long)b * p.yh + oy) * p.yw + ox) * cw;
              if (TN == 128 && cw == 4) *(uint4*)d = make_uint4(sgw[k][0], sgw[k][1], sgw[k][2], sgw[k][3]);
              else if (TN == 64 && cw == 2) *(uint2*)d = make_uint2(sgw[k][0], sgw[k][1]);
              else {
#pragma unroll
                for (int c = 0; c < TN / 32; c++) if (c < cw) d[c] = sgw[k][c];
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[acs]);
    }
    if (p.csum != nullptr) {       // lanes with equal ch4 (4 per warp) -> this warp's slot of the CTA's sums (no atomics: the order of every addition is fixed)
#pragma unroll
      for (int c = 0; c < TN / 32; c++) {
        float4 v = cs[c];
#pragma unroll
        for (int off = 8; off <= 16; off <<= 1) {
          v.x += __shfl_xor_sync(0xffffffffu, v.x, off); v.y += __shfl_xor_sync(0xffffffffu, v.y, off);
          v.z += __shfl_xor_sync(0xffffffffu, v.z, off); v.w += __shfl_xor_sync(0xffffffffu, v.w, off);
        }
        if (lane < 8) *(float4*)(csum_s + q * TN + 32 * c + 4 * ch4) = v;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.csum != nullptr && (int)threadIdx.x < TN && (int)threadIdx.x < p.Cout) {
    const float tot = (csum_s[threadIdx.x] + csum_s[TN + threadIdx.x]) + (csum_s[2 * TN + threadIdx.x] + csum_s[3 * TN + threadIdx.x]);
    if (p.csum_part != nullptr) p.csum_part[(int64_t)blockIdx.x * TN + threadIdx.x] = tot;      // deterministic mode: per-CTA partials, added in CTA order by agb_reduce_partials
    else atomicAdd(p.csum + threadIdx.x, tot);
  }
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int TN, int G, bool MB, bool MBO>
static int cols_launch(agb_ctx* ctx, ColsParams& p, size_t smem) {
  static bool attr = false;
  if (!attr) { AGB_CUDA(cudaFuncSetAttribute(conv_cols_kernel<TN, G, MB, MBO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr = true; }
  long long grid = p.num_tiles < ctx->sm_count ? p.num_tiles : ctx->sm_count;
  conv_cols_kernel<TN, G, MB, MBO><<<(unsigned)grid, ColsCfg<TN>::THREADS, smem, ctx->stream>>>(p);
  AGB_LAUNCHED(ctx);
  return AGB_OK;
}

static int make_sw64_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides, const uint32_t* box) {
  PFN_encodeTiled enc = agb_get_encode();
  AGB_CHECK(enc, AGB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t d[5], s[4]; cuuint32_t b[5], e[5];
  for (int i = 0; i < rank; i++) { d[i] = dims[i]; b[i] = box[i]; e[i] = 1; }
  for (int i = 0; i + 1 < rank; i++) s[i] = strides[i];
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { agb_set_error("cuTensorMapEncodeTiled (SWIZZLE_64B) failed with CUresult %d", (int)r); return AGB_ERR_UNSUPPORTED; }
  return AGB_OK;
}

// x [B,H,W,Cin] channels-last, wr = filter repacked to [tap][Cout][Cin] (see tc_conv.cu), y [B,yh,yw,Cout] channels-last.
// Returns AGB_ERR_UNSUPPORTED when the geometry is outside this kernel's envelope (the caller falls back to the per-tap kernel).
int agb_tc_conv_cols(agb_ctx* ctx, const float* x, const float* wr, float* y, int B, int Cin, int H, int W, int Cout, int yh, int yw, int kh, int kw,
                     int pad, int dil, const float* bias, int relu, const float* mask, float* csum) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("AGB_CONV_COLS"); enabled = (e && e[0] == '0') ? 0 : 1; }
  if (!enabled || !(yw == 64 || yw == 32 || yw == 16) || Cout > 128 || Cin % 4 != 0 || Cout % 4 != 0) return AGB_ERR_UNSUPPORTED;
  ColsParams p;
  p.rm = 128 / yw;
  const int rt = 2 * p.rm, wrows = rt + dil * (kh - 1);
  if (wrows > 256 || (int64_t)B * ((yh + rt - 1) / rt) < ctx->sm_count) return AGB_ERR_UNSUPPORTED;        // too few tiles to fill the machine
  p.copy_bytes = 64 * yw * wrows; p.copy_stride = (p.copy_bytes + 1023) & ~1023; p.a_slot_bytes = kw * p.copy_stride;
  const int TN = Cout > 64 ? 128 : 64;
  static const int g_env = [] { const char* e = getenv("AGB_COLS_G"); return e ? atoi(e) : 0; }();
  const int G = (kw == 3 && (g_env == 3 || (g_env == 0 && TN == 64))) ? 3 : 1;       // one filter row per ring stage where the ring can still be deep enough
  const size_t fixed = 2 * (size_t)p.a_slot_bytes + 1024 + 512 + 4 * 32 * 36 * 4 + 4 * 128 * 4;
  const size_t st_bytes = (size_t)G * TN * 64;
  if (fixed + 2 * st_bytes > 227 * 1024) return AGB_ERR_UNSUPPORTED;
  int nb = (int)((227 * 1024 - fixed) / st_bytes); if (nb > COLS_NB_MAX) nb = COLS_NB_MAX;
  p.nb = nb;
  const size_t smem = fixed + (size_t)nb * st_bytes;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)Cin * 4, (uint64_t)W * Cin * 4, (uint64_t)H * W * Cin * 4};
    uint32_t box[4] = {16, (uint32_t)yw, (uint32_t)wrows, 1};
    AGB_TRY(make_sw64_map(&p.tmX, x, 4, dims, str, box));
  }
  {
    uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout, (uint64_t)(kh * kw)};
    uint64_t str[2] = {(uint64_t)Cin * 4, (uint64_t)Cin * Cout * 4};
    uint32_t box[3] = {16, (uint32_t)TN, (uint32_t)G};
    AGB_TRY(make_sw64_map(&p.tmW, wr, 3, dims, str, box));
  }
  p.y = y; p.bias = bias; p.mask = mask; p.csum = csum; p.relu = relu; p.Cout = Cout; p.yh = yh; p.yw = yw; p.kw = kw; p.pad = pad; p.dil = dil;
  p.tiles_y = (yh + rt - 1) / rt; p.cblocks = (Cin + 15) / 16; p.taps = kh * kw;
  p.num_tiles = (long long)B * p.tiles_y;
  const int64_t ncta = p.num_tiles < ctx->sm_count ? p.num_tiles : ctx->sm_count;
  p.csum_part = nullptr;
  if (csum != nullptr && ctx->deterministic) AGB_TRY(agb_scratch2(ctx, (size_t)ncta * TN * sizeof(float), (void**)&p.csum_part));
  p.mask_bits = nullptr; p.bits_out = nullptr;
  if (Cout % 32 == 0) {
    if (mask != nullptr && ctx->mask_bits != nullptr) { p.mask_bits = ctx->mask_bits; ctx->mask_bits_used = 1; }
    // measured (64 -> 128 @64x64, B = 256): writing the bits from this kernel's transposed epilogue costs 0.12 ms (the epilogue warps become the tile's
    // critical path) and saves 0.14 ms in the next layer's masked dgrad: off by default (AGB_COLS_BITS=1); reading bits written by other kernels stays on
    static const int prod = [] { const char* e = getenv("AGB_COLS_BITS"); return (e && e[0] == '1') ? 1 : 0; }();
    if (prod && ctx->bits_out != nullptr) { p.bits_out = ctx->bits_out; ctx->bits_written = 1; }
  }
  int r;
#define AGB_COLS_DISPATCH(MB_, MBO_) do { if (G == 3) r = TN == 64 ? cols_launch<64, 3, MB_, MBO_>(ctx, p, smem) : cols_launch<128, 3, MB_, MBO_>(ctx, p, smem); \
                                          else r = TN == 64 ? cols_launch<64, 1, MB_, MBO_>(ctx, p, smem) : cols_launch<128, 1, MB_, MBO_>(ctx, p, smem); } while (0)
  if (p.bits_out != nullptr) AGB_COLS_DISPATCH(false, true);          // forward (no mask)
  else if (p.mask_bits != nullptr) AGB_COLS_DISPATCH(true, false);
  else AGB_COLS_DISPATCH(false, false);
#undef AGB_COLS_DISPATCH
  if (r == AGB_OK && p.csum_part != nullptr) r = agb_reduce_partials(ctx, p.csum_part, csum, (int)ncta, Cout, TN, 1);
  return r;
}
