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
//   warp  4    TMA producer (lane 0) + MMA issuer (elected lane), TMEM accumulators double buffered
//   warps 5-8  epilogue: tcgen05.ld -> bias / ReLU -> y
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
};

template <int TN, bool SPLIT>
__global__ void __launch_bounds__(288, 2) conv_first_fprop_kernel(const __grid_constant__ FirstParams p) {
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
    uint32_t it = 0;
    for (int64_t t = blockIdx.x; t < p.tiles; t += gridDim.x, it++) {
      const int buf = it & 1;
      const int xb_i = it % CF_XS_STAGES; const float* xsb = xs + xb_i * p.xs_floats;
      mbar_wait(&x_full[xb_i], (it / CF_XS_STAGES) & 1);       // this tile's input rows have landed
      mbar_wait(&a_empty[buf], ((it >> 1) & 1) ^ 1);           // the MMAs that read this buffer two tiles ago are complete
      uint8_t* A = sA + buf * A_BYTES; uint8_t* Al = sAl + buf * A_BYTES;
      const int px = tid * p.stride;
#pragma unroll
      for (int ch = 0; ch < 8; ch++) {
        float4 v, l;
        float e[4];
#pragma unroll
        for (int q = 0; q < 4; q++) { const int ko = s_koff[ch * 4 + q]; e[q] = ko >= 0 ? xsb[ko + px] : 0.0f; }
        v = make_float4(e[0], e[1], e[2], e[3]);
        const int off = tid * 128 + ((ch ^ (tid & 7)) << 4);
        *(float4*)(A + off) = v;
        if (SPLIT) {
          l.x = tf32_rna(v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u)); l.y = tf32_rna(v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u));
          l.z = tf32_rna(v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u)); l.w = tf32_rna(v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u));
          *(float4*)(Al + off) = l;
        }
      }
      fence_proxy_async();
      mbar_arrive(&a_full[buf]);
      mbar_arrive(&x_empty[xb_i]);                             // the staged rows may be overwritten
    }
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_tf32(128, TN, 0, 0);
    const uint32_t hiK = (1024u >> 4) | (1u << 14) | (2u << 29), loK = (16u >> 4) << 16;
    const uint32_t aB = (smem_u32(sB) >> 4) + loK, aBl = (smem_u32(sBl) >> 4) + loK;
    const uint32_t x_bytes = (uint32_t)(p.C * p.kh * p.seg * 4);
    const int shift = ((-p.pad) % 4 + 4) % 4;
    auto produce = [&](int64_t t, uint32_t pi) {              // stage tile t's input rows (lane 0)
      const int tx = (int)(t % p.tiles_x); const int64_t r = t / p.tiles_x; const int oy = (int)(r % p.yh), b = (int)(r / p.yh);
      const uint32_t xi = pi % CF_XS_STAGES;
      mbar_wait(&x_empty[xi], ((pi / CF_XS_STAGES) & 1) ^ 1);
      mbar_expect_tx(&x_full[xi], x_bytes);
      tma_load_4d(xs + xi * p.xs_floats, &p.tmX, &x_full[xi], tx * 128 * p.stride - p.pad - shift, oy * p.stride - p.pad, 0, b);
    };
    uint32_t it = 0, pi = 0; int64_t tp = blockIdx.x;
    if (lane == 0) for (; pi < CF_XS_STAGES - 1 && tp < p.tiles; pi++, tp += gridDim.x) produce(tp, pi);
    for (int64_t t = blockIdx.x; t < p.tiles; t += gridDim.x, it++) {
      const uint32_t buf = it & 1;
      if (lane == 0 && tp < p.tiles) { produce(tp, pi); pi++; tp += gridDim.x; }
      __syncwarp();
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
#pragma unroll
      for (int c0 = 0; c0 < TN; c0 += 32) {
        float v[32];
        tmem_ld32(tlane + buf * (uint32_t)TN + (uint32_t)c0, v);
        tmem_ld_wait();
        if (c0 < p.O) {                                         // (warp-uniform)
          uint8_t* stg = sY + (chunk & 1) * A_BYTES;
          if (leader) tma_store_wait_read<1>();                // the store that last read this staging buffer (two chunks ago) is done with it
          asm volatile("bar.sync 2, 128;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o4;
            o4.x = v[j] + sbias[c0 + j]; o4.y = v[j + 1] + sbias[c0 + j + 1]; o4.z = v[j + 2] + sbias[c0 + j + 2]; o4.w = v[j + 3] + sbias[c0 + j + 3];
            if (p.relu) { o4.x = fmaxf(o4.x, 0.0f); o4.y = fmaxf(o4.y, 0.0f); o4.z = fmaxf(o4.z, 0.0f); o4.w = fmaxf(o4.w, 0.0f); }
            *(float4*)(stg + row * 128 + (((j >> 2) ^ (row & 7)) << 4)) = o4;
          }
          fence_proxy_async();
          asm volatile("bar.sync 2, 128;" ::: "memory");
          if (leader) { tma_store_4d(&p.tmY, stg, c0, tx * 128, oy, b); tma_store_commit(); }
          chunk++;
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[buf]);
    }
    if (leader) tma_store_wait_read<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 2 * TN);
}

template <int TN, bool SPLIT>
static int first_launch(agb_ctx* ctx, const FirstParams& p) {
  constexpr int SMEM_MAX = (SPLIT ? 2 : 1) * (TN * 128 + 2 * 128 * 128) + 2 * 128 * 128 + CF_XS_STAGES * CF_XS_MAX_FLOATS * 4 + TN * 4 + 1024 + 256;
  static_assert(SMEM_MAX <= 227 * 1024, "first-layer tile does not fit shared memory");
  const int SMEM = SMEM_MAX - CF_XS_STAGES * (CF_XS_MAX_FLOATS - p.xs_floats) * 4;
  static bool attr = false;
  if (!attr) { AGB_CUDA(cudaFuncSetAttribute(conv_first_fprop_kernel<TN, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX)); attr = true; }
  static int occ = 0, occ_smem = 0;
  if (occ_smem != SMEM) {      // co-resident CTAs overlap one's build / epilogue phases with the other's; they share the 512 TMEM columns
    AGB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, conv_first_fprop_kernel<TN, SPLIT>, 288, SMEM)); if (occ < 1) occ = 1; if (occ * 2 * TN > 512) occ = 512 / (2 * TN);
    occ_smem = SMEM;
  }
  const int64_t cap = (int64_t)ctx->sm_count * occ;
  const unsigned n = (unsigned)(p.tiles < cap ? p.tiles : cap);
  conv_first_fprop_kernel<TN, SPLIT><<<n, 288, SMEM, ctx->stream>>>(p);
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
  p.w = w; p.y = y; p.bias = bias; p.relu = relu; p.B = B; p.C = C; p.H = H; p.W = W; p.O = O; p.kh = kh; p.kw = kw; p.pad = pad; p.stride = stride;
  p.dil = dil; p.yh = yh; p.yw = yw; p.K = K; p.tiles_x = (yw + 127) / 128; p.seg = seg; p.xs_floats = (C * kh * seg + 31) / 32 * 32; p.tiles = (int64_t)B * yh * p.tiles_x;
  const bool split = mode == AGB_MATH_3XTF32;
  if (O <= 64) return split ? first_launch<64, true>(ctx, p) : first_launch<64, false>(ctx, p);
  if (O <= 128) return split ? first_launch<128, true>(ctx, p) : first_launch<128, false>(ctx, p);
  return split ? first_launch<256, true>(ctx, p) : first_launch<256, false>(ctx, p);
}
