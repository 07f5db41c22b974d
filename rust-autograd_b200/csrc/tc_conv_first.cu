// tc_conv_first.cu — first-layer convolution forward (very few input channels: K = C*kh*kw <= 32) on tcgen05, channels-last output.
//
// Why a kernel of its own: with K <= 32 the whole reduction is ONE 32-wide k-block, so the layer is bound by writing y (256 x 64 x 128 x 128
// f32 = 1.07 GB against 50 MB of input): the target is the HBM roofline, not the tensor pipe.  The warp-MMA kernel of conv_small_c.cu gathers
// every A fragment element from global memory with its own bounds check and address arithmetic and ran at 0.37 of the HBM rate (instruction
// issue bound).  Here the im2col tile is BUILT IN SHARED MEMORY: per 128-pixel tile (a segment of one output row) the C*kh input-row segments
// arrive by ONE TMA box {segment, kh rows, C channels} (out-of-bounds zero fill = the padding; a 4-deep mbarrier ring hides the load latency
// — staging them with ordinary loads cost a full memory round trip per tile: 0.96 ms), each builder thread then writes its pixel's K taps as one 128-byte row of
// a K-major SWIZZLE_128B UMMA tile with conflict-free 16-byte stores, and ONE thread issues the four K = 8 MMAs against the filter tile
// that stays resident for the life of the CTA.  Bias + ReLU are applied to the accumulator registers; the epilogue threads write their
// pixel's 32 channels as one swizzled 128-byte row of a shared-memory staging tile and ONE TMA store ({32 channels, 128 pixels} box,
// clipped at the row end / channel count by the hardware) moves it to y — per-thread 16-byte global stores of a pixel-per-thread layout
// touch 32 half-filled sectors per warp instruction and made the epilogue the bottleneck (ncu: builders and issuer waiting on it).
//   warps 0-3  builders: build the A tile (and A_lo in 3xTF32 mode) from the staged rows, double buffered
//   warp  4    MMA issuer (elected lane), TMEM accumulators double buffered
//   warps 5-8  epilogue: tcgen05.ld -> bias / ReLU -> y
//   warp  9    TMA producer (lane 0)
// 3xTF32: the filter is split once per CTA (hi = rna, lo = rna(w - hi)), the builders write lo = rna(v - trunc(v)) next to the raw tile
// (the tensor core truncates the raw values itself) and the issuer runs lo*hi + hi*lo + hi*hi; with K <= 32 there are only 12
// accumulating MMAs per output, no chunked promotion is needed.
// Replaces Conv2D::compute (conv2d.rs:115-211: im2col + sgemm) and the bias add / ReLU behind it for the first layer of a CNN.
#include "tc_common.cuh"
#include <stdlib.h>

#define CF_XS_MAX_FLOATS 2560   // one staged buffer: C * kh row segments of seg floats (rounded up to 32: 128-byte aligned TMA destinations)
#define CF_XS_STAGES 4

struct FirstParams {
  CUtensorMap tmX;               // x as {W, H, C, B}, box {seg, kh (row step dil), C, 1}, no swizzle
  CUtensorMap tmY;               // y (channels-last) as {O, yw, yh, B}, box {32, 128, 1, 1}, 128-byte swizzle (store)
  const float* w; float* y; const float* bias; int relu;
  int B, C, H, W, O, kh, kw, pad, stride, dil, yh, yw, K;
  int tiles_x, seg, xs_floats; int64_t tiles;
  uint32_t* bits;                // optional: sign bits of the stored activation (O % 32 == 0), one word per pixel per 32 channels
};

template <int TN, bool SPLIT, bool BITS>
__global__ void __launch_bounds__(320, 2) conv_first_fprop_kernel(const __grid_constant__ FirstParams p) {
  constexpr int A_BYTES = 128 * 128, B_BYTES = TN * 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = smem;                                        // [TN o][32 k]  K-major, 128B swizzle (hi)
  uint8_t* sBl = sB + B_BYTES;                               // lo plane (SPLIT)
  uint8_t* sA = sBl + (SPLIT ? B_BYTES : 0);                 // 2 x [128 px][32 k]
  uint8_t* sAl = sA + 2 * A_BYTES;                           // 2 x lo (SPLIT)
  uint8_t* sY = sAl + (SPLIT ? 2 * A_BYTES : 0);            // 2 x [128 px][32 ch] store staging (128B swizzle)
  float* xs = (float*)(sY + 2 * A_BYTES);                    // CF_XS_STAGES x [C][kh][seg]
  float* sbias = xs + CF_XS_STAGES * p.xs_floats;            // [TN]
  uint64_t* bars = (uint64_t*)(sbias + TN);
  uint64_t* a_full = bars; uint64_t* a_empty = bars + 2; uint64_t* acc_full = bars + 4; uint64_t* acc_empty = bars + 6;
  uint64_t* x_full = bars + 8; uint64_t* x_empty = bars + 8 + CF_XS_STAGES;
  uint32_t* tmem_slot = (uint32_t*)(bars + 8 + 2 * CF_XS_STAGES);
  __shared__ int s_koff[32];                                 // tap k -> offset into xs of pixel 0's sample, or -1 past K
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int rows = p.C * p.kh;

  if (tid == 0) {
    for (int b = 0; b < 2; b++) { mbar_init(&a_full[b], 128); mbar_init(&a_empty[b], 1); mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 128); }
    for (int b = 0; b < CF_XS_STAGES; b++) { mbar_init(&x_full[b], 1); mbar_init(&x_empty[b], 128); }
    fence_barrier_init();
    tma_prefetch_desc(&p.tmX); tma_prefetch_desc(&p.tmY);
  }
  if (tid < 32) {
    const int k = tid, kk = p.kh * p.kw;
    const int shift = ((-p.pad) % 4 + 4) % 4;          // the TMA box starts at the 16-byte aligned column below ox0 * stride - pad
    if (k < p.K) { const int c = k / kk, r = k - c * kk, i = r / p.kw, j = r - i * p.kw; s_koff[k] = (c * p.kh + i) * p.seg + j * p.dil + shift; }
    else s_koff[k] = -1;
  }
  if (warp == 4) { tmem_alloc(tmem_slot, 2 * TN); tmem_relinquish(); }
  // resident filter tile: element (o, k) at o*128 + ((k/4) ^ (o%8))*16 + (k%4)*4
  for (int i = tid; i < TN * 32; i += blockDim.x) {
    const int o = i >> 5, k = i & 31;
    const float v = (o < p.O && k < p.K) ? __ldg(p.w + (int64_t)o * p.K + k) : 0.0f;
    const int off = o * 128 + (((k >> 2) ^ (o & 7)) << 4) + ((k & 3) << 2);
    if (SPLIT) { const float h = tf32_rna(v); *(float*)(sB + off) = h; *(float*)(sBl + off) = tf32_rna(v - h); }
    else *(float*)(sB + off) = v;
  }
  for (int i = tid; i < TN; i += blockDim.x) sbias[i] = (p.bias != nullptr && i < p.O) ? __ldg(p.bias + i) : 0.0f;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ===================== builders =====================
    int koff[32];                                               // the tap table lives in registers: the gathers do not wait for a table lookup
#pragma unroll
    for (int k = 0; k < 32; k++) koff[k] = s_koff[k];
    const uint32_t sA_u = smem_u32(sA), sAl_u = smem_u32(sAl), row_u = (uint32_t)tid * 128u;
    uint32_t it = 0;
    for (int64_t t = blockIdx.x; t < p.tiles; t += gridDim.x, it++) {
      const int buf = it & 1;
      const int xb_i = it % CF_XS_STAGES; const float* xsb = xs + xb_i * p.xs_floats + tid * p.stride;
      mbar_wait(&x_full[xb_i], (it / CF_XS_STAGES) & 1);       // this tile's input rows have landed
      mbar_wait(&a_empty[buf], ((it >> 1) & 1) ^ 1);           // the MMAs that read this buffer two tiles ago are complete
      const uint32_t A = sA_u + buf * A_BYTES + row_u, Al = sAl_u + buf * A_BYTES + row_u;
#pragma unroll
      for (int ch = 0; ch < 8; ch++) {
        float4 v, l;
        float e[4];
#pragma unroll
        for (int q = 0; q < 4; q++) { const int ko = koff[ch * 4 + q]; e[q] = ko >= 0 ? xsb[ko] : 0.0f; }
        v = make_float4(e[0], e[1], e[2], e[3]);
        const uint32_t off = (uint32_t)((ch ^ (tid & 7)) << 4);
        if (SPLIT) {      // round-to-nearest split on BOTH operands (the builder writes the tile anyway, so hi costs nothing): a truncated hi leaves a lo of
                          // up to 13 bits whose rounding to tf32 is a coherent 2^-21 error per product; with rna hi the lo fits tf32 exactly
          const float4 h = make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
          l.x = tf32_rna(v.x - h.x); l.y = tf32_rna(v.y - h.y); l.z = tf32_rna(v.z - h.z); l.w = tf32_rna(v.w - h.w);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(A + off), "f"(h.x), "f"(h.y), "f"(h.z), "f"(h.w) : "memory");
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(Al + off), "f"(l.x), "f"(l.y), "f"(l.z), "f"(l.w) : "memory");
        } else asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(A + off), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
      }
      fence_proxy_async();
      mbar_arrive(&a_full[buf]);
      mbar_arrive(&x_empty[xb_i]);                             // the staged rows may be overwritten
    }
  } else if (warp == 9) {
    // ===================== TMA producer (its own warp: the tile decode and the ring wait stay off the issuer's instruction stream) =====================
    if (lane == 0) {
      const uint32_t x_bytes = (uint32_t)(p.C * p.kh * p.seg * 4);
      const int shift = ((-p.pad) % 4 + 4) % 4;
      uint32_t pi = 0;
      for (int64_t t = blockIdx.x; t < p.tiles; t += gridDim.x, pi++) {
        const int tx = (int)(t % p.tiles_x); const int64_t r = t / p.tiles_x; const int oy = (int)(r % p.yh), b = (int)(r / p.yh);
        const uint32_t xi = pi % CF_XS_STAGES;
        mbar_wait(&x_empty[xi], ((pi / CF_XS_STAGES) & 1) ^ 1);
        mbar_expect_tx(&x_full[xi], x_bytes);
        tma_load_4d(xs + xi * p.xs_floats, &p.tmX, &x_full[xi], tx * 128 * p.stride - p.pad - shift, oy * p.stride - p.pad, 0, b);
      }
    }
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_tf32(128, TN, 0, 0);
    const uint32_t hiK = (1024u >> 4) | (1u << 14) | (2u << 29), loK = (16u >> 4) << 16;
    const uint32_t aB = (smem_u32(sB) >> 4) + loK, aBl = (smem_u32(sBl) >> 4) + loK;
    uint32_t it = 0;
    for (int64_t t = blockIdx.x; t < p.tiles; t += gridDim.x, it++) {
      const uint32_t buf = it & 1;
      mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
      mbar_wait(&a_full[buf], (it >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t tacc = tmem_base + buf * (uint32_t)TN;
        const uint32_t aA = (smem_u32(sA + buf * A_BYTES) >> 4) + loK, aAl = (smem_u32(sAl + buf * A_BYTES) >> 4) + loK;
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const uint64_t dA = umma_desc_pack(aA + k * 2, hiK), dB = umma_desc_pack(aB + k * 2, hiK);
          if (SPLIT) {
            umma_tf32(tacc, umma_desc_pack(aAl + k * 2, hiK), dB, idesc, k != 0);
            umma_tf32(tacc, dA, umma_desc_pack(aBl + k * 2, hiK), idesc, 1);
            umma_tf32(tacc, dA, dB, idesc, 1);
          } else umma_tf32(tacc, dA, dB, idesc, k != 0);
        }
        umma_commit(&a_empty[buf]);
        umma_commit(&acc_full[buf]);
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3, row = 32 * q + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(32 * q) << 16);
    const bool leader = warp == 5 && lane == 0;
    uint32_t it = 0, chunk = 0;
    for (int64_t t = blockIdx.x; t < p.tiles; t += gridDim.x, it++) {
      const int tx = (int)(t % p.tiles_x); const int64_t r = t / p.tiles_x; const int oy = (int)(r % p.yh), b = (int)(r / p.yh);
      const uint32_t buf = it & 1;
      mbar_wait(&acc_full[buf], (it >> 1) & 1);
      tc_fence_after();
      uint32_t signs[TN / 32];
#pragma unroll
      for (int c0 = 0; c0 < TN; c0 += 32) {
        float v[32];
        signs[c0 >> 5] = 0;
        tmem_ld32(tlane + buf * (uint32_t)TN + (uint32_t)c0, v);
        tmem_ld_wait();
        if (c0 < p.O) {                                         // (warp-uniform)
          uint8_t* stg = sY + (chunk & 1) * A_BYTES;
          if (leader) tma_store_wait_read<1>();                // the store that last read this staging buffer (two chunks ago) is done with it
          asm volatile("bar.sync 2, 128;" ::: "memory");
          uint32_t sign = 0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o4;
            o4.x = v[j] + sbias[c0 + j]; o4.y = v[j + 1] + sbias[c0 + j + 1]; o4.z = v[j + 2] + sbias[c0 + j + 2]; o4.w = v[j + 3] + sbias[c0 + j + 3];
            if (p.relu) { o4.x = fmaxf(o4.x, 0.0f); o4.y = fmaxf(o4.y, 0.0f); o4.z = fmaxf(o4.z, 0.0f); o4.w = fmaxf(o4.w, 0.0f); }
            *(float4*)(stg + row * 128 + (((j >> 2) ^ (row & 7)) << 4)) = o4;
            if (BITS) sign |= ((o4.x > 0.0f ? 1u : 0u) | (o4.y > 0.0f ? 2u : 0u) | (o4.z > 0.0f ? 4u : 0u) | (o4.w > 0.0f ? 8u : 0u)) << j;
          }
          signs[c0 >> 5] = sign;
          fence_proxy_async();
          asm volatile("bar.sync 2, 128;" ::: "memory");
          if (leader) { tma_store_4d(&p.tmY, stg, c0, tx * 128, oy, b); tma_store_commit(); }
          chunk++;
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[buf]);
      if (BITS && tx * 128 + row < p.yw) {          // the pixel's sign words, contiguous: 8 bytes per pixel at O = 64, consecutive pixels -> coalesced
        uint32_t* d = p.bits + (((int64_t)b * p.yh + oy) * p.yw + tx * 128 + row) * (p.O >> 5);
        if (TN == 64 && p.O == 64) *(uint2*)d = make_uint2(signs[0], signs[1]);
        else {
#pragma unroll
          for (int c = 0; c < TN / 32; c++) if (32 * c < p.O) d[c] = signs[c];
        }
      }
    }
    if (leader) tma_store_wait_read<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 2 * TN);
}

template <int TN, bool SPLIT, bool BITS>
static int first_launch_impl(agb_ctx* ctx, const FirstParams& p);
template <int TN, bool SPLIT>
static int first_launch(agb_ctx* ctx, const FirstParams& p) { return p.bits != nullptr ? first_launch_impl<TN, SPLIT, true>(ctx, p) : first_launch_impl<TN, SPLIT, false>(ctx, p); }
template <int TN, bool SPLIT, bool BITS>
static int first_launch_impl(agb_ctx* ctx, const FirstParams& p) {
  constexpr int SMEM_MAX = (SPLIT ? 2 : 1) * (TN * 128 + 2 * 128 * 128) + 2 * 128 * 128 + CF_XS_STAGES * CF_XS_MAX_FLOATS * 4 + TN * 4 + 1024 + 256;
  static_assert(SMEM_MAX <= 227 * 1024, "first-layer tile does not fit shared memory");
  const int SMEM = SMEM_MAX - CF_XS_STAGES * (CF_XS_MAX_FLOATS - p.xs_floats) * 4;
  static bool attr = false;
  if (!attr) { AGB_CUDA(cudaFuncSetAttribute(conv_first_fprop_kernel<TN, SPLIT, BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX)); attr = true; }
  // co-resident CTAs overlap one's build / epilogue phases with the other's (0.328 -> 0.203 ms on the 3 -> 64 @128x128 layer); they share the 512 TMEM
  // columns.  Residency is computed here: cudaOccupancyMaxActiveBlocksPerMultiprocessor answers 1 for this kernel at 95 KB of dynamic shared memory
  // although two CTAs are resident (ncu: launch__occupancy_limit_shared_mem = 2), which had left half of every SM idle.
  static const int occ_env = [] { const char* e = getenv("AGB_CF_OCC"); return e ? atoi(e) : 0; }();
  int occ = (227 * 1024) / (SMEM + 1024); if (occ > 2) occ = 2; if (occ < 1) occ = 1;
  if (occ * 2 * TN > 512) occ = 512 / (2 * TN);
  if (occ_env > 0) occ = occ_env;
  const int64_t cap = (int64_t)ctx->sm_count * occ;
  const unsigned n = (unsigned)(p.tiles < cap ? p.tiles : cap);
  conv_first_fprop_kernel<TN, SPLIT, BITS><<<n, 320, SMEM, ctx->stream>>>(p);
  AGB_LAUNCHED(ctx);
  return AGB_OK;
}

// x NCHW-contiguous [B,C,H,W], w [O, C*kh*kw], y channels-last [B,yh,yw,O] (dense), bias [O] or null.  AGB_ERR_UNSUPPORTED when the geometry
// does not fit (the caller falls back to the warp-MMA kernel).
int agb_tc_conv_first(agb_ctx* ctx, int mode, const float* x, const float* w, float* y, int B, int C, int H, int W, int O, int kh, int kw, int yh, int yw,
                      int pad, int stride, int dil, const float* bias, int relu) {
  static const int enabled = [] { const char* e = getenv("AGB_CONV_FIRST"); return (e && e[0] == '0') ? 0 : 1; }();
  const int K = C * kh * kw;
  const int shift = ((-pad) % 4 + 4) % 4;
  const int seg = (127 * stride + (kw - 1) * dil + 1 + shift + 3) / 4 * 4;       // staged row segment: starts 16-byte aligned, length a multiple of 16 bytes
  if (!enabled || mode == AGB_MATH_FP32 || K > 32 || seg > 256 || C * kh * seg > CF_XS_MAX_FLOATS || (kh - 1) * dil + 1 > 256 || O > 256 || O % 4 != 0 || yw < 32 || W % 4 != 0)
    return AGB_ERR_UNSUPPORTED;
  if ((((uintptr_t)y | (uintptr_t)x) & 15) != 0 || (bias && (((uintptr_t)bias) & 3) != 0)) return AGB_ERR_UNSUPPORTED;
  FirstParams p;
  {
    uint64_t dims[4] = {(uint64_t)W, (uint64_t)H, (uint64_t)C, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)W * 4, (uint64_t)H * W * 4, (uint64_t)C * H * W * 4};
    uint32_t box[4] = {(uint32_t)seg, (uint32_t)((kh - 1) * dil + 1), (uint32_t)C, 1}, es[4] = {1, (uint32_t)dil, 1, 1};
    AGB_TRY(agb_make_tmap(&p.tmX, x, 4, dims, str, box, false, dil > 1 ? es : nullptr, true));
  }
  {
    uint64_t dims[4] = {(uint64_t)O, (uint64_t)yw, (uint64_t)yh, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)O * 4, (uint64_t)yw * O * 4, (uint64_t)yh * yw * O * 4};
    uint32_t box[4] = {32, 128, 1, 1};
    AGB_TRY(agb_make_tmap(&p.tmY, y, 4, dims, str, box, false));
  }
  p.bits = nullptr;
  if (ctx->bits_out != nullptr && O % 32 == 0) { p.bits = ctx->bits_out; ctx->bits_written = 1; }
  p.w = w; p.y = y; p.bias = bias; p.relu = relu; p.B = B; p.C = C; p.H = H; p.W = W; p.O = O; p.kh = kh; p.kw = kw; p.pad = pad; p.stride = stride;
  p.dil = dil; p.yh = yh; p.yw = yw; p.K = K; p.tiles_x = (yw + 127) / 128; p.seg = seg; p.xs_floats = (C * kh * seg + 31) / 32 * 32; p.tiles = (int64_t)B * yh * p.tiles_x;
  const bool split = mode == AGB_MATH_3XTF32;
  if (O <= 64) return split ? first_launch<64, true>(ctx, p) : first_launch<64, false>(ctx, p);
  if (O <= 128) return split ? first_launch<128, true>(ctx, p) : first_launch<128, false>(ctx, p);
  return split ? first_launch<256, true>(ctx, p) : first_launch<256, false>(ctx, p);
}

// =====================================================================================================================
// First-layer FILTER GRADIENT on tcgen05: gw[o, k] = sum over pixels of gy[pixel, o] * patch[pixel, k], K = C*kh*kw <= 32.
// The layer is bound by READING gy (1.07 GB at B = 256, 64 x 128 x 128) — the warp-MMA kernel of conv_small_c.cu gathered its fragments from
// global memory at 0.41 of the HBM rate.  Here, per 128-pixel tile (a segment of one output row):
//   * gy arrives by TMA as the MN-major A operand (M = o, K = pixel): one {32 o, 128 pixels} box per 32 channels, SWIZZLE_128B_ATOM_32B
//     (the only MN-major TF32 layout, tc_common.cuh), a ring of up to 4 stages = up to 148 KB of loads in flight per SM;
//   * the input rows arrive by the forward kernel's TMA box and 128 builder threads write the im2col tile [pixel][32 taps] — as an
//     MN-major B operand (N = tap, K = pixel) it is the same 128-byte-row image the forward kernel builds, with the 32-byte-chunk swizzle;
//   * 16 MMAs (M = 128 lanes of which O are used, N = 32, K = 8 pixels each) accumulate the tile into one of two TMEM buffers; four
//     accumulator warps add every finished tile into fp32 REGISTERS (TMEM accumulation truncates, so nothing is left there for more
//     than one tile), and at the end each CTA writes its [O, 32] partial sums to scratch.
// A second tiny kernel adds the per-CTA partials in CTA order: no atomics, so the result is bit-reproducible run to run
// (the reference's filter gradient is a sequential sum, conv2d.rs:631-734).
// =====================================================================================================================
struct FirstWgradParams {
  CUtensorMap tmX;               // x as {W, H, C, B}, box {seg, kh (row step dil), C, 1}, no swizzle
  CUtensorMap tmG;               // gy (channels-last) as {O, yw, yh, B}, box {32, 128, 1, 1}, SWIZZLE_128B_ATOM_32B
  float* part;                   // [grid][128][32]
  int B, C, H, W, O, kh, kw, pad, stride, dil, yh, yw, K;
  int tiles_x, seg, xs_floats, nbox, stages; int64_t tiles;
};

#define CFW_MAX_STAGES 4
#define CFW_THREADS 320          // warps 0-3 builders, 4 MMA issuer, 5-8 accumulators, 9 TMA producer
__global__ void __launch_bounds__(CFW_THREADS, 1) conv_first_wgrad_kernel(const __grid_constant__ FirstWgradParams p) {
  constexpr int BOX_BYTES = 128 * 128, P_BYTES = 128 * 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int a_stage = p.nbox * BOX_BYTES;
  uint8_t* sA = smem;                                           // stages x nbox x [128 px][32 o]   (MN-major A, 32-byte-chunk swizzle)
  uint8_t* sP = sA + p.stages * a_stage;                        // 2 x [128 px][32 taps]            (MN-major B, same swizzle)
  float* xs = (float*)(sP + 2 * P_BYTES);                       // stages x [C][kh][seg]
  uint64_t* bars = (uint64_t*)(xs + p.stages * p.xs_floats);
  uint64_t* in_full = bars; uint64_t* in_empty = bars + CFW_MAX_STAGES; uint64_t* p_full = bars + 2 * CFW_MAX_STAGES; uint64_t* p_empty = p_full + 2;
  uint64_t* acc_full = p_full + 4; uint64_t* acc_empty = p_full + 6;
  uint32_t* tmem_slot = (uint32_t*)(p_full + 8);
  __shared__ int s_koff[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int S = p.stages;

  if (tid == 0) {
    for (int b = 0; b < S; b++) { mbar_init(&in_full[b], 1); mbar_init(&in_empty[b], 1); }
    for (int b = 0; b < 2; b++) { mbar_init(&p_full[b], 128); mbar_init(&p_empty[b], 1); mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 128); }
    fence_barrier_init();
    tma_prefetch_desc(&p.tmX); tma_prefetch_desc(&p.tmG);
  }
  if (tid < 32) {
    const int k = tid, kk = p.kh * p.kw;
    const int shift = ((-p.pad) % 4 + 4) % 4;
    if (k < p.K) { const int c = k / kk, r = k - c * kk, i = r / p.kw, j = r - i * p.kw; s_koff[k] = (c * p.kh + i) * p.seg + j * p.dil + shift; }
    else s_koff[k] = -1;
  }
  if (warp == 4) { tmem_alloc(tmem_slot, 64); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ===================== builders: patches[pixel][tap] =====================
    int koff[32];                                               // the tap table lives in registers: the gathers do not wait for a table lookup
#pragma unroll
    for (int k = 0; k < 32; k++) koff[k] = s_koff[k];
    const uint32_t sP_u = smem_u32(sP), row_u = (uint32_t)tid * 128u;
    uint32_t it = 0;
    for (int64_t t = blockIdx.x; t < p.tiles; t += gridDim.x, it++) {
      const int buf = it & 1, st = it % S;
      const float* xsb = xs + st * p.xs_floats + tid * p.stride;
      mbar_wait(&in_full[st], (it / S) & 1);
      mbar_wait(&p_empty[buf], ((it >> 1) & 1) ^ 1);
      const uint32_t Pt = sP_u + buf * P_BYTES + row_u;
#pragma unroll
      for (int ch = 0; ch < 8; ch++) {
        const int c16 = ch ^ ((tid >> 2) & 1);                 // threads t and t+4 write different halves of a 32-byte chunk: conflict-free 16-byte stores
        float e[4];
#pragma unroll
        for (int q = 0; q < 4; q++) { const int ko = (tid >> 2) & 1 ? koff[(ch ^ 1) * 4 + q] : koff[ch * 4 + q]; e[q] = ko >= 0 ? xsb[ko] : 0.0f; }
        const uint32_t off = Pt + ((((c16 >> 1) ^ (tid & 3)) << 5) | ((c16 & 1) << 4));
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(off), "f"(e[0]), "f"(e[1]), "f"(e[2]), "f"(e[3]) : "memory");
      }
      fence_proxy_async();
      mbar_arrive(&p_full[buf]);
    }
  } else if (warp == 9) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const uint32_t in_bytes = (uint32_t)(p.C * p.kh * p.seg * 4) + (uint32_t)a_stage;
      const int shift = ((-p.pad) % 4 + 4) % 4;
      uint32_t pi = 0;
      for (int64_t t = blockIdx.x; t < p.tiles; t += gridDim.x, pi++) {
        const int tx = (int)(t % p.tiles_x); const int64_t r = t / p.tiles_x; const int oy = (int)(r % p.yh), b = (int)(r / p.yh);
        const uint32_t st = pi % S;
        mbar_wait(&in_empty[st], ((pi / S) & 1) ^ 1);
        mbar_expect_tx(&in_full[st], in_bytes);
        tma_load_4d(xs + st * p.xs_floats, &p.tmX, &in_full[st], tx * 128 * p.stride - p.pad - shift, oy * p.stride - p.pad, 0, b);
        for (int g = 0; g < p.nbox; g++) tma_load_4d(sA + st * a_stage + g * BOX_BYTES, &p.tmG, &in_full[st], 32 * g, tx * 128, oy, b);
      }
    }
  } else if (warp == 4) {
    // ===================== MMA issuer (nothing else: one warp's instruction stream per tile is the kernel's critical path) =====================
    constexpr uint32_t idesc = umma_idesc_tf32(128, 32, 1, 1);
    const uint32_t hi = (512u >> 4) | (1u << 14) | (1u << 29);           // SBO = 512 (4-pixel atoms), version 1, layout 128B_BASE32B
    const uint32_t loA = ((uint32_t)BOX_BYTES >> 4) << 16, loB = (4096u >> 4) << 16;      // LBO: between 32-channel boxes (B has one box)
    const uint32_t aA0 = (smem_u32(sA) >> 4) + loA, aB0 = (smem_u32(sP) >> 4) + loB;
    uint32_t it = 0, st = 0;
    for (int64_t t = blockIdx.x; t < p.tiles; t += gridDim.x, it++) {
      const uint32_t buf = it & 1, ph = (it >> 1) & 1;
      mbar_wait(&acc_empty[buf], ph ^ 1);
      mbar_wait(&in_full[st], (it / S) & 1);
      mbar_wait(&p_full[buf], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t tacc = tmem_base + buf * 32u;
        const uint32_t aA = aA0 + st * (uint32_t)(a_stage >> 4), aB = aB0 + buf * (uint32_t)(P_BYTES >> 4);
#pragma unroll
        for (int k = 0; k < 16; k++) umma_tf32(tacc, umma_desc_pack(aA + k * 64, hi), umma_desc_pack(aB + k * 64, hi), idesc, k != 0);      // 8 pixels = 1024 B per step
        umma_commit(&in_empty[st]);
        umma_commit(&p_empty[buf]);
        umma_commit(&acc_full[buf]);
      }
      __syncwarp();
      st = st + 1 == (uint32_t)S ? 0 : st + 1;
    }
  } else {
    // ===================== accumulators: TMEM -> fp32 registers, once per tile =====================
    const int q = warp & 3, row = 32 * q + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(32 * q) << 16);
    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; j++) acc[j] = 0.0f;
    uint32_t it = 0;
    for (int64_t t = blockIdx.x; t < p.tiles; t += gridDim.x, it++) {
      const uint32_t buf = it & 1;
      mbar_wait(&acc_full[buf], (it >> 1) & 1);
      tc_fence_after();
      float v[32];
      tmem_ld32(tlane + buf * 32u, v);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&acc_empty[buf]);
#pragma unroll
      for (int j = 0; j < 32; j++) acc[j] += v[j];
    }
    float4* dst = (float4*)(p.part + ((int64_t)blockIdx.x * 128 + row) * 32);
#pragma unroll
    for (int j = 0; j < 32; j += 4) dst[j >> 2] = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 64);
}

// gw[o, k] = sum over CTAs (in CTA order) of part[cta][o][k]
__global__ void __launch_bounds__(256) conv_first_wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ gw, int ncta, int O, int K) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= O * K) return;
  const int o = i / K, k = i - o * K;
  const float* src = part + o * 32 + k;
  float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
  int c = 0;
  for (; c + 3 < ncta; c += 4) {
    s0 += __ldg(src + (int64_t)c * 4096); s1 += __ldg(src + (int64_t)(c + 1) * 4096); s2 += __ldg(src + (int64_t)(c + 2) * 4096); s3 += __ldg(src + (int64_t)(c + 3) * 4096);
  }
  for (; c < ncta; c++) s0 += __ldg(src + (int64_t)c * 4096);
  gw[i] = (s0 + s1) + (s2 + s3);
}

// x NCHW-contiguous [B,C,H,W], gy channels-last dense [B,yh,yw,O], gw [O, C*kh*kw].  TF32 mode only (the 3xTF32 mode keeps the warp-MMA kernel).
int agb_tc_conv_first_wgrad(agb_ctx* ctx, int mode, const float* x, const float* gy, float* gw, int B, int C, int H, int W, int O, int kh, int kw, int yh, int yw,
                            int pad, int stride, int dil) {
  static const int enabled = [] { const char* e = getenv("AGB_CONV_FIRST_WGRAD"); return (e && e[0] == '0') ? 0 : 1; }();
  const int K = C * kh * kw;
  const int shift = ((-pad) % 4 + 4) % 4;
  const int seg = (127 * stride + (kw - 1) * dil + 1 + shift + 3) / 4 * 4;
  if (!enabled || mode != AGB_MATH_TF32 || K > 32 || seg > 256 || C * kh * seg > CF_XS_MAX_FLOATS || (kh - 1) * dil + 1 > 256 || O > 128 || O % 4 != 0 || yw < 32 || W % 4 != 0)
    return AGB_ERR_UNSUPPORTED;
  if ((((uintptr_t)gy | (uintptr_t)x) & 15) != 0) return AGB_ERR_UNSUPPORTED;
  FirstWgradParams p;
  {
    uint64_t dims[4] = {(uint64_t)W, (uint64_t)H, (uint64_t)C, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)W * 4, (uint64_t)H * W * 4, (uint64_t)C * H * W * 4};
    uint32_t box[4] = {(uint32_t)seg, (uint32_t)((kh - 1) * dil + 1), (uint32_t)C, 1}, es[4] = {1, (uint32_t)dil, 1, 1};
    AGB_TRY(agb_make_tmap(&p.tmX, x, 4, dims, str, box, false, dil > 1 ? es : nullptr, true));
  }
  {
    uint64_t dims[4] = {(uint64_t)O, (uint64_t)yw, (uint64_t)yh, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)O * 4, (uint64_t)yw * O * 4, (uint64_t)yh * yw * O * 4};
    uint32_t box[4] = {32, 128, 1, 1};
    AGB_TRY(agb_make_tmap(&p.tmG, gy, 4, dims, str, box, true));
  }
  p.B = B; p.C = C; p.H = H; p.W = W; p.O = O; p.kh = kh; p.kw = kw; p.pad = pad; p.stride = stride; p.dil = dil; p.yh = yh; p.yw = yw; p.K = K;
  p.tiles_x = (yw + 127) / 128; p.seg = seg; p.xs_floats = (C * kh * seg + 31) / 32 * 32; p.tiles = (int64_t)B * yh * p.tiles_x;
  p.nbox = O <= 32 ? 1 : O <= 64 ? 2 : 4;
  const int per_stage = p.nbox * 16384 + p.xs_floats * 4;
  // measured (B200, 3 -> 64 @128x128, B = 256): one CTA per SM with a 4-deep ring 0.200 ms (0.86 of HBM); two co-resident CTAs with 2 stages each 0.251 ms
  static const int st_env = [] { const char* e = getenv("AGB_CFW_STAGES"); return e ? atoi(e) : 0; }();
  int stages = (227 * 1024 - 2 * 16384 - 2048) / per_stage;
  if (st_env > 0) stages = st_env;
  if (stages > CFW_MAX_STAGES) stages = CFW_MAX_STAGES;
  if (stages < 2 || 1024 + stages * per_stage + 2 * 16384 + 512 > 227 * 1024) return AGB_ERR_UNSUPPORTED;
  p.stages = stages;
  int smem = 1024 + stages * per_stage + 2 * 16384 + 512;
  const int reach = 1024 + (stages - 1) * p.nbox * 16384 + 4 * 16384;       // the M = 128 descriptor of the last stage reads four boxes: keep them inside the allocation
  if (smem < reach) smem = reach;
  static int attr_smem = 0;
  if (attr_smem < smem) { AGB_CUDA(cudaFuncSetAttribute(conv_first_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr_smem = smem; }
  int occ = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, conv_first_wgrad_kernel, CFW_THREADS, smem) != cudaSuccess || occ < 1) occ = 1;
  if (occ > 2) occ = 2;
  const int64_t cap = (int64_t)ctx->sm_count * occ;
  const unsigned n = (unsigned)(p.tiles < cap ? p.tiles : cap);
  float* part; AGB_TRY(agb_scratch(ctx, (size_t)n * 128 * 32 * sizeof(float), (void**)&part));
  p.part = part;
  conv_first_wgrad_kernel<<<n, CFW_THREADS, smem, ctx->stream>>>(p);
  AGB_LAUNCHED(ctx);
  conv_first_wgrad_reduce_kernel<<<(O * K + 255) / 256, 256, 0, ctx->stream>>>(part, gw, (int)n, O, K);
  AGB_LAUNCHED(ctx);
  return AGB_OK;
}
