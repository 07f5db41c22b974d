// tc_conv_wgrad_taps.cu — all-taps-per-CTA tcgen05 filter gradient for narrow layers (C <= 64, 64-wide output-channel tiles, 3x3-class filters,
// stride 1, channels-last activations): gw[o,c,i,j] = sum_{b,oy,ox} gy[b,o,oy,ox] x[b,c,oy+i*d-p,ox+j*d-p]
// (Conv2DFilterGrad::compute, conv2d.rs:631-734; the reference loops the batch sequentially with beta = 1).
//
// Why (profiles/ncu_conv_full_r1.csv, ConvWgradPol<64,pair> on the 64->64 @128x128 VGG layer): one CTA per tap pair re-reads x once
// per tap and gy once per tap pair — 16.1 GB through L2 -> smem and 4.3 GB from DRAM for 2.1 GB of operands, 24 KB of TMA traffic per
// 4 MMAs of 32 clocks each, tensor pipe 27 % busy.  Here a CTA owns ALL taps:
//   * k-block = 32 consecutive output pixels of one (b, oy) row.  Per k-block the haloed x window — kh rows x (32 + d(kw-1)) pixels x
//     64 channels — and the gy tile [32 px][64 o] land in shared memory ONCE (MN-major boxes, 128B_BASE32B swizzle);
//   * tap (i, j) is the window seen through a descriptor whose start is shifted by j*d pixel rows (K dimension) inside halo row i.
//     An MN-major operand may start at any 128-byte K-row and use any LBO: the swizzle is a function of absolute shared-memory
//     address bits (verified on B200 by scripts/cuda/umma_shift_probe_mn.cu);
//   * one MMA covers TWO vertically adjacent taps: its 128 lanes are [tap (2p, j): c 0..63 | tap (2p+1, j): c 0..63] because the four
//     32-channel boxes (row 2p g0, row 2p g1, row 2p+1 g0, row 2p+1 g1) sit at one uniform stride (= LBO).  An odd kh pairs the last
//     row with a zeroed row.  3x3: 6 M-tiles x 64 TMEM columns, 24 MMAs per k-block for 34 KB of TMA traffic;
//   * split-K: every CTA walks a contiguous range of k-blocks with the accumulators resident in TMEM, then adds its 9x64x64 partial
//     into gw with red.global.add (gw is zeroed by the launcher).  Two issuer warps (one per row pair), one TMA warp, 4 drain warps.
#include "tc_common.cuh"

#define WT_STAGES 4
#define WT_BOXS 5120                 // stride of one [<=40 px][32 ch] box (1024-aligned)
struct WTapsParams {
  CUtensorMap tmX, tmG;              // x: box {32 c, HP w, 1, 1}; gy: box {32 o, 32 w, 1, 1}  (SWIZZLE_128B_ATOM_32B)
  float* gw; int C, O, T, kh, kw, pad, dil, yh, xblocks, hp, kb_total, kb_per_cta, npair;
  int64_t part_stride;               // > 0 (deterministic mode): CTA column x stores its partial sums at gw + x * part_stride, added in order by agb_reduce_partials
};

__global__ void __launch_bounds__(224, 1) conv_wgrad_taps_kernel(const __grid_constant__ WTapsParams p) {
  constexpr int S = WT_STAGES, TN = 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int rows = 2 * p.npair;                                  // halo rows incl. the zero row that completes the last pair
  const uint32_t x_bytes = (uint32_t)rows * 2 * WT_BOXS, stage_bytes = x_bytes + 2 * 4096;
  uint64_t* bars = (uint64_t*)(smem + S * stage_bytes);
  uint64_t* full = bars; uint64_t* empty = bars + S; uint64_t* acc_full = bars + 2 * S;
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * S + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // rows the TMA never writes (the pad row of an odd kh, and the tails of the boxes) must read as zeros
  for (uint32_t i = threadIdx.x; i < S * stage_bytes / 16; i += blockDim.x) ((float4*)smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  fence_proxy_async();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmX); tma_prefetch_desc(&p.tmG);
    for (int s = 0; s < S; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 2); }
    mbar_init(acc_full, 2);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int q0 = (int)blockIdx.x * p.kb_per_cta, q1 = min(q0 + p.kb_per_cta, p.kb_total);
  const int o0 = (int)blockIdx.y * TN;                           // 64-wide output-channel tile of this CTA
  const int nk = q1 > q0 ? q1 - q0 : 0;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      const uint32_t tx = (uint32_t)p.kh * 2 * (uint32_t)p.hp * 128 + 2 * 4096;
      for (int kb = 0; kb < nk; kb++) {
        const int q = q0 + kb; const int xb = q % p.xblocks; const int r = q / p.xblocks; const int oy = r % p.yh, b = r / p.yh;
        const int ox0 = xb * 32;
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* st = smem + (size_t)s * stage_bytes;
        mbar_expect_tx(&full[s], tx);
        for (int i = 0; i < p.kh; i++)
          for (int g = 0; g < 2; g++)
            tma_load_4d(st + (i * 2 + g) * WT_BOXS, &p.tmX, &full[s], 32 * g, ox0 - p.pad, oy + i * p.dil - p.pad, b);
        for (int g = 0; g < 2; g++) tma_load_4d(st + x_bytes + g * 4096, &p.tmG, &full[s], o0 + 32 * g, ox0, oy, b);
        if (++s == S) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1 || warp == 6) {
    // ===================== MMA issuers: row pair `pp` each (converged warp, elected lane) =====================
    constexpr uint32_t idesc = umma_idesc_tf32(128, TN, 1, 1);
    const uint32_t hiP = (512u >> 4) | (1u << 14) | (1u << 29), loP = ((uint32_t)WT_BOXS >> 4) << 16;      // MN-major, LBO = box stride, SBO 512 B
    const uint32_t hiQ = hiP, loQ = (4096u >> 4) << 16;
    const uint32_t smem0 = smem_u32(smem) >> 4, dj4 = (uint32_t)(p.dil * 128) >> 4;
    int s = 0; uint32_t ph = 0;
    for (int kb = 0; kb < nk; kb++) {
      mbar_wait(&full[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t st = smem0 + (uint32_t)s * (stage_bytes >> 4);
        const uint32_t aQ = st + (x_bytes >> 4) + loQ;
        for (int pp = (warp == 1 ? 0 : 1); pp < p.npair; pp += 2) {
          const uint32_t aRow = st + (uint32_t)(pp * 4) * (WT_BOXS >> 4) + loP;
          for (int j = 0; j < p.kw; j++) {
            const uint32_t tacc = tmem_base + (uint32_t)((pp * p.kw + j) * TN);
            const uint32_t aP = aRow + (uint32_t)j * dj4;
#pragma unroll
            for (int ks = 0; ks < 4; ks++)
              umma_tf32(tacc, umma_desc_pack(aP + ks * 64, hiP), umma_desc_pack(aQ + ks * 64, hiQ), idesc, !(kb == 0 && ks == 0));
          }
        }
        umma_commit(&empty[s]);
        if (kb == nk - 1) umma_commit(acc_full);
      }
      __syncwarp();
      if (++s == S) { s = 0; ph ^= 1; }
    }
  } else if (warp < 6 && nk > 0) {
    // ===================== drain: lanes 0-63 = tap (2pp, j), lanes 64-127 = tap (2pp+1, j); column = o =====================
    const int q = warp & 3;
    const int row = 32 * q + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(32 * q) << 16);
    mbar_wait(acc_full, 0);
    tc_fence_after();
    for (int pp = 0; pp < p.npair; pp++)
      for (int j = 0; j < p.kw; j++) {
        const int i = 2 * pp + (row >> 6), c = row & 63;
        const int tap = i * p.kw + j;
#pragma unroll
        for (int c0 = 0; c0 < TN; c0 += 32) {
          float v[32];
          tmem_ld32(tlane + (uint32_t)((pp * p.kw + j) * TN + c0), v);
          tmem_ld_wait();
          if (i < p.kh && c < p.C) {
#pragma unroll
            for (int e = 0; e < 32; e++) {
              const int o = o0 + c0 + e;
              if (o < p.O) {
                if (p.part_stride > 0) p.gw[(int64_t)blockIdx.x * p.part_stride + ((int64_t)tap * p.O + o) * p.C + c] = v[e];      // partials as [tap][o][c]: a warp stores 128 contiguous bytes
                else red_add_f32(p.gw + ((int64_t)o * p.C + c) * p.T + tap, v[e]);
              }
            }
          }
        }
      }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// img [B,H,W,C] and g [B,yh,yw,O] channels-last, gw [O,C,kh,kw] (plain, zeroed here).  AGB_ERR_UNSUPPORTED outside the envelope.
int agb_tc_conv_wgrad_taps(agb_ctx* ctx, const float* img, const float* g, float* gw, int B, int C, int H, int W, int O, int yh, int yw,
                           int kh, int kw, int pad, int dil) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("AGB_WGRAD_TAPS"); enabled = (e && e[0] == '0') ? 0 : 1; }
  const int npair = (kh + 1) / 2, hp = 32 + dil * (kw - 1);
  if (!enabled || C > 64 || O > 256 || C % 4 != 0 || O % 4 != 0 || npair * kw * 64 > 512 || hp * 128 > WT_BOXS) return AGB_ERR_UNSUPPORTED;
  const size_t stage = (size_t)2 * npair * 2 * WT_BOXS + 2 * 4096;
  const size_t smem = WT_STAGES * stage + 1024 + 256;
  if (smem > 227 * 1024) return AGB_ERR_UNSUPPORTED;
  WTapsParams p;
  {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)C * 4, (uint64_t)W * C * 4, (uint64_t)H * W * C * 4};
    uint32_t box[4] = {32, (uint32_t)hp, 1, 1};
    AGB_TRY(agb_make_tmap(&p.tmX, img, 4, dims, str, box, true));
  }
  {
    uint64_t dims[4] = {(uint64_t)O, (uint64_t)yw, (uint64_t)yh, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)O * 4, (uint64_t)yw * O * 4, (uint64_t)yh * yw * O * 4};
    uint32_t box[4] = {32, 32, 1, 1};
    AGB_TRY(agb_make_tmap(&p.tmG, g, 4, dims, str, box, true));
  }
  p.gw = gw; p.C = C; p.O = O; p.T = kh * kw; p.kh = kh; p.kw = kw; p.pad = pad; p.dil = dil; p.yh = yh; p.xblocks = (yw + 31) / 32; p.hp = hp; p.npair = npair;
  const int64_t kb_total = (int64_t)B * yh * p.xblocks;
  if (kb_total > 2147483647ll) return AGB_ERR_UNSUPPORTED;
  p.kb_total = (int)kb_total;
  const int otiles = (O + 63) / 64;
  int64_t ctas = ctx->sm_count / otiles; if (ctas > kb_total / 8) ctas = kb_total / 8; if (ctas < 1) ctas = 1;      // >= 8 k-blocks per CTA amortise the 36.9 K reds
  p.kb_per_cta = (int)((kb_total + ctas - 1) / ctas);
  ctas = (kb_total + p.kb_per_cta - 1) / p.kb_per_cta;
  const int64_t n = (int64_t)O * C * p.T;
  float* part = nullptr;
  if (ctx->deterministic && ctas > 1) AGB_TRY(agb_scratch2(ctx, (size_t)ctas * n * sizeof(float), (void**)&part));
  if (part) { p.gw = part; p.part_stride = n; } else { p.part_stride = 0; AGB_TRY(agb_memset0(ctx, gw, (size_t)n * sizeof(float))); }
  static bool attr = false;
  if (!attr) { AGB_CUDA(cudaFuncSetAttribute(conv_wgrad_taps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr = true; }
  conv_wgrad_taps_kernel<<<dim3((unsigned)ctas, (unsigned)otiles), 224, smem, ctx->stream>>>(p);
  AGB_LAUNCHED(ctx);
  if (part) return agb_reduce_partials_wgrad(ctx, part, gw, (int)ctas, O, C, p.T);
  return AGB_OK;
}
