// tc_conv_rows.cu — persistent, halo-reusing tcgen05 implicit-GEMM convolution for WIDE feature maps (output width >= 128,
// stride 1, channels-last activations).  Serves Conv2D fprop and Conv2DTranspose/dgrad (flipped filter) exactly like
// ConvFpropPol in tc_conv.cu, for the layers where that kernel is bound by L2 -> shared-memory bandwidth.
//
// Why a second kernel (measured, profiles/ncu_conv_full_r1.csv): the per-tap kernel re-fetches its 128-pixel input window from
// L2 once per filter tap — 9x for 3x3 — and its filter slice once per CTA.  On the 64-channel 128x128 VGG layers that is 432 KB
// of TMA traffic per 9.4 MFLOP tile, 14.5 GB per launch at ~10 TB/s: the kernel sits on the chip's L2->SM throughput cap
// (~6300 B/clk) with the tensor pipe 20 % busy.  DRAM traffic is already minimal (x once, y once); the waste is on-chip.
//
// This kernel loads each input element into shared memory ONCE per 32-channel block:
//   * tile = R output rows x 128 consecutive output pixels of one image x TN output channels;
//   * per 32-channel block ONE 4-D TMA box {32 c, 128 + d(kw-1) w, R + d(kh-1) h, 1 b} brings the haloed input window
//     (SWIZZLE_128B, one 128-byte row per pixel, zero fill = padding);
//   * every tap (i, j) of every output row r reads that window through a K-major UMMA descriptor whose start address is shifted
//     by ((r + i d) * pitch + j d) pixel rows.  The 128B swizzle is a function of absolute shared-memory address bits, so a
//     start address that is 128-byte but not 1024-byte aligned addresses the TMA-written data correctly (verified on B200 by
//     scripts/cuda/umma_shift_probe.cu for every shift, base_offset field 0);
//   * the filter tap tiles [TN o][32 c] stream through a small ring and are shared by the R rows (R MMAs per tap and k-step).
//   L2 -> smem traffic per 128-pixel x 64-channel output strip: (R+2)/R * 33 KB + 147/R KB = 140 KB at R = 2 (was 432 KB).
// Persistent CTAs (one per SM) walk the tile list; two TMEM accumulator sets (2 x R x TN columns) let the epilogue of tile t
// (tcgen05.ld, bias / ReLU / ReLU-mask, 128-byte channel runs per pixel) overlap the MMAs of tile t+1.
//
// Reference semantics: Conv2D::compute conv2d.rs:115-211, Conv2DTranspose::compute conv2d_transpose.rs:89-247.
#include "tc_common.cuh"
#include <float.h>

#define ROWS_R 2
struct RowsParams {
  CUtensorMap tmX, tmW;
  float* y; const float* bias; const float* mask; const uint32_t* mask_bits; uint32_t* bits_out; float* csum; float* csum_part; int relu;
  float* pool_y; int* pool_idx; int ph, pw;      // fused max_pool2d(2, 0, 2): pooled output + int32 argmax (logical NCHW offsets into y), y itself not written
  int Cout, yh, yw, kw, pad, dil, tiles_x, tiles_y, otiles, cblocks, taps, pitch, a_box_bytes, a_slot_bytes, nb /* filter ring depth in stages of G taps */;
  long long num_tiles;
};
#define ROWS_NB_MAX 8
template <int TN> struct RowsCfg {
  static constexpr int B_BYTES = TN * 128;
  static constexpr int TMEM_COLS = 2 * ROWS_R * TN;            // 256 (TN = 64) / 512 (TN = 128)
  static constexpr int THREADS = 256;            // warps: 0 TMA-A, 1 MMA row 0, 2-5 epilogue, 6 TMA-B, 7 MMA row 1
};

// G = filter taps per ring stage (3 = one filter row per TMA box / barrier wait / commit when kw == 3: see tc_conv_cols.cu)
// MB: the ReLU sign-bit side channel (mask read as bits, sign bits of the output written) — its own instantiation, so the float-mask kernel keeps the
// instruction schedule it had without it (with both forms in one kernel the ordinary masked dgrad lost 30 %)
template <int TN, int G, bool MB>
__global__ void __launch_bounds__(256, 1) conv_rows_kernel(const __grid_constant__ RowsParams p) {
  using Cfg = RowsCfg<TN>;
  constexpr int NB = ROWS_NB_MAX, R = ROWS_R;                  // barrier slots; p.nb stages in use
  constexpr int ST_BYTES = G * Cfg::B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                                          // [2][a_slot_bytes]
  uint8_t* sB = smem + 2 * p.a_slot_bytes;                     // [p.nb][G][TN * 128]
  uint64_t* bars = (uint64_t*)(sB + p.nb * ST_BYTES);
  uint64_t* a_full = bars; uint64_t* a_empty = bars + 2;
  uint64_t* b_full = bars + 4; uint64_t* b_empty = bars + 4 + NB;
  uint64_t* acc_full = bars + 4 + 2 * NB; uint64_t* acc_empty = bars + 6 + 2 * NB;
  uint32_t* tmem_slot = (uint32_t*)(bars + 8 + 2 * NB);
  float* stage = (float*)((uint8_t*)bars + 256);               // [4 warps][32][36] epilogue transpose tiles
  float* csum_s = stage + 4 * 32 * 36;                         // [4 epilogue warps][TN] per-channel sums of this CTA
  for (int i = threadIdx.x; i < 4 * TN; i += blockDim.x) csum_s[i] = 0.0f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmX); tma_prefetch_desc(&p.tmW);
    for (int s = 0; s < 2; s++) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], R); mbar_init(&acc_full[s], R); mbar_init(&acc_empty[s], 128); }
    for (int s = 0; s < NB; s++) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], R); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long tiles_per_img = (long long)p.tiles_x * p.tiles_y * p.otiles;

  if (warp == 0) {
    // ===================== TMA producer A: haloed input windows =====================
    // (own warp: the next window is requested the moment its slot frees, independent of the filter-tap ring's back-pressure)
    if (lane == 0) {
      uint32_t ai = 0;
      for (long long t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const int b = (int)(t / tiles_per_img); int r = (int)(t - (long long)b * tiles_per_img);
        r /= p.otiles; const int tx = r % p.tiles_x, ty = r / p.tiles_x;
        const int oy0 = ty * R, ox0 = tx * 128;
        for (int cb = 0; cb < p.cblocks; cb++, ai++) {
          const uint32_t as = ai & 1;
          mbar_wait(&a_empty[as], ((ai >> 1) & 1) ^ 1);
          mbar_expect_tx(&a_full[as], (uint32_t)p.a_box_bytes);
          tma_load_4d(sA + as * p.a_slot_bytes, &p.tmX, &a_full[as], cb * 32, ox0 - p.pad, oy0 - p.pad, b);       // dims {c, w, h, b}
        }
      }
    }
  } else if (warp == 6) {
    // ===================== TMA producer B: filter taps =====================
    if (lane == 0) {
      uint32_t bs = 0, bph = 0;
      for (long long t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const int o0 = (int)(t % p.otiles) * TN;
        for (int cb = 0; cb < p.cblocks; cb++)
          for (int tap = 0; tap < p.taps; tap += G) {
            mbar_wait(&b_empty[bs], bph ^ 1);
            mbar_expect_tx(&b_full[bs], ST_BYTES);
            tma_load_3d(sB + bs * ST_BYTES, &p.tmW, &b_full[bs], cb * 32, o0, tap);
            if (++bs == (uint32_t)p.nb) { bs = 0; bph ^= 1; }
          }
      }
    }
  } else if (warp == 1 || warp == 7) {
    // ===================== MMA issuers: one warp per output row =====================
    // A tcgen05.mma of N = 64 occupies the tensor pipe for only ~32 clocks; a single issuing thread (wait + descriptor adds + loop
    // bookkeeping, ~11 instructions per MMA) could not keep up (measured: issuer-bound at 40 % tensor activity).  Row r of the tile
    // has its own accumulator, so each row gets its own issuer warp; ring slots are released when BOTH have committed.
    // Converged warp, one elected lane issues (uniform-register descriptors; see tc_tile.cuh).
    const int r = warp == 1 ? 0 : 1;
    constexpr uint32_t idesc = umma_idesc_tf32(128, TN, 0, 0);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29), lo0 = (16u >> 4) << 16;     // K-major SWIZZLE_128B, LBO 16 B, SBO 1024 B
    const uint32_t row4 = (uint32_t)(p.pitch * 128) >> 4;            // one input row of the halo window, in 16-byte units
    const uint32_t sA4 = (smem_u32(sA) >> 4) + lo0 + (uint32_t)r * row4, sB4 = (smem_u32(sB) >> 4) + lo0;
    const uint32_t dj4 = (uint32_t)(p.dil * 128) >> 4, di4 = (uint32_t)p.dil * row4;
    uint32_t ai = 0, bs = 0, bph = 0, it = 0;
    for (long long t = blockIdx.x; t < p.num_tiles; t += gridDim.x, it++) {
      const uint32_t acs = it & 1;
      mbar_wait(&acc_empty[acs], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + (acs * (uint32_t)R + (uint32_t)r) * (uint32_t)TN;
      for (int cb = 0; cb < p.cblocks; cb++, ai++) {
        const uint32_t as = ai & 1;
        mbar_wait(&a_full[as], (ai >> 1) & 1);
        tc_fence_after();
        const uint32_t a0 = sA4 + as * ((uint32_t)p.a_slot_bytes >> 4);
        uint32_t arow = a0, atap = a0; int j = 0;                      // window origin of tap (i, j) for this row
        for (int tap = 0; tap < p.taps; tap += G) {
          uint32_t at[G];
#pragma unroll
          for (int jj = 0; jj < G; jj++) { at[jj] = atap; if (++j == p.kw) { j = 0; arow += di4; atap = arow; } else atap += dj4; }
          mbar_wait(&b_full[bs], bph);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t q0 = sB4 + bs * (uint32_t)(ST_BYTES >> 4);
#pragma unroll
            for (int jj = 0; jj < G; jj++)
#pragma unroll
              for (int ks = 0; ks < 4; ks++)
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
    // ===================== epilogue =====================
    // tcgen05.ld hands every thread one pixel x 32 channels; written out like that a warp store touches 32 separate sectors.
    // Each warp therefore transposes its 32 x 32 strip through a padded shared-memory tile (conflict-free both ways) and works in
    // the transposed layout: lane -> (pixel = 4 it + lane / 8, channels 4 (lane % 8) .. +3), so every 128-bit global access of a
    // warp covers 4 pixels x 128 contiguous bytes.  Bias / ReLU / ReLU-mask are applied in that layout (4 fixed channels per lane).
    const int q = warp & 3;
    const uint32_t tlane = tmem_base + ((uint32_t)(32 * q) << 16);
    float* stg = stage + q * (32 * 36);
    const int pl0 = lane >> 3, ch4 = lane & 7;
    float4 cs[TN / 32];                                          // per-channel sums of the stored values (csum side output; otiles == 1)
#pragma unroll
    for (int c = 0; c < TN / 32; c++) cs[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t it = 0;
    for (long long t = blockIdx.x; t < p.num_tiles; t += gridDim.x, it++) {
      const int b = (int)(t / tiles_per_img); int rr = (int)(t - (long long)b * tiles_per_img);
      const int ot = rr % p.otiles; rr /= p.otiles; const int tx = rr % p.tiles_x, ty = rr / p.tiles_x;
      const int oy0 = ty * R, ox0 = tx * 128 + 32 * q, o0 = ot * TN;
      // ReLU-mask bits of the dgrad epilogue, fetched (coalesced) before the accumulator is ready:
      // bit 4 it + e of pre[r][c]  <=>  mask_src[b, oy0 + r, ox0 + 4 it + pl0, o0 + 32 c + 4 ch4 + e] > 0
      uint32_t pre[R][TN / 32];
      if (MB && p.mask_bits != nullptr) {
        // the mask as sign bits written by the forward kernel: one word per pixel per 32 channels (8 lanes share it) instead of 128 bytes
        const int cw = p.Cout >> 5;
#pragma unroll
        for (int r = 0; r < R; r++) {
          const long long prow = ((long long)b * p.yh + min(oy0 + r, p.yh - 1)) * p.yw;
#pragma unroll
          for (int c = 0; c < TN / 32; c++) {
            const int cwi = min((o0 >> 5) + c, cw - 1);
            uint32_t wv[8];
#pragma unroll
            for (int k = 0; k < 8; k++) wv[k] = __ldg(p.mask_bits + (prow + min(ox0 + 4 * k + pl0, p.yw - 1)) * cw + cwi);
            uint32_t bits = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) bits |= ((wv[k] >> (4 * ch4)) & 15u) << (4 * k);
            pre[r][c] = bits;
          }
        }
      } else if (p.mask != nullptr) {
        // pull the NEXT tile's mask lines into L2 now (one 128-byte line per 8 lanes): by the time that tile's epilogue starts, its
        // loads below hit L2 instead of waiting ~1.5 us on HBM with the accumulator already finished
        const long long tn = t + gridDim.x;
        if (tn < p.num_tiles && ch4 == 0) {
          const int bn = (int)(tn / tiles_per_img); int rn = (int)(tn - (long long)bn * tiles_per_img);
          const int otn = rn % p.otiles; rn /= p.otiles; const int txn = rn % p.tiles_x, tyn = rn / p.tiles_x;
#pragma unroll
          for (int r = 0; r < R; r++)
#pragma unroll
            for (int c = 0; c < TN / 32; c++)
#pragma unroll
              for (int k = 0; k < 8; k++) {
                const int oyn = min(tyn * R + r, p.yh - 1), oxn = min(txn * 128 + 32 * q + 4 * k + pl0, p.yw - 1);
                const float* pm = p.mask + (((long long)bn * p.yh + oyn) * p.yw + oxn) * p.Cout + min(otn * TN + 32 * c, p.Cout - 4);
                asm volatile("prefetch.global.L2 [%0];" :: "l"(pm));
              }
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
          const int oy = oy0 + r;
#pragma unroll
          for (int c = 0; c < TN / 32; c++) {
            // unconditional loads from clamped addresses (bits of out-of-range pixels / channels are never used): the 8 loads
            // of a strip are independent and issue back to back
            uint32_t bits = 0; const int o = min(o0 + 32 * c + 4 * ch4, p.Cout - 4);
            const float* mrow = p.mask + (((long long)b * p.yh + min(oy, p.yh - 1)) * p.yw) * p.Cout + o;
            float4 mv[8];
#pragma unroll
            for (int k = 0; k < 8; k++) mv[k] = __ldg((const float4*)(mrow + (long long)min(ox0 + 4 * k + pl0, p.yw - 1) * p.Cout));
#pragma unroll
            for (int k = 0; k < 8; k++)
              bits |= ((mv[k].x > 0.0f ? 1u : 0u) | (mv[k].y > 0.0f ? 2u : 0u) | (mv[k].z > 0.0f ? 4u : 0u) | (mv[k].w > 0.0f ? 8u : 0u)) << (4 * k);
            pre[r][c] = bits;
          }
        }
      }
      float4 bv[TN / 32];
#pragma unroll
      for (int c = 0; c < TN / 32; c++) {
        const int o = o0 + 32 * c + 4 * ch4;
        bv[c] = (p.bias != nullptr && o + 4 <= p.Cout) ? __ldg((const float4*)(p.bias + o)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      const uint32_t acs = it & 1;
      mbar_wait(&acc_full[acs], (it >> 1) & 1);
      tc_fence_after();
      if (p.pool_y != nullptr) {
        // conv -> (+bias) -> ReLU -> max_pool2d(2, 0, 2) in one epilogue (examples/cnn_mnist.rs:38-45 layer pattern): the tile's two
        // rows are the two rows of every window, horizontal neighbours sit 8 lanes apart in the transposed layout.  Scan order and
        // tie-breaking follow MaxPool2D::compute (max_pool2d.rs:21-88): strict '>' from -FLT_MAX over (h0,w0), (h0,w0+1), (h1,w0),
        // (h1,w0+1); index = flat offset into the (never materialised) [B,C,yh,yw] conv output.
        const int py = oy0 >> 1;
#pragma unroll
        for (int c = 0; c < TN / 32; c++) {
          float4 a0[8], a1[8];
#pragma unroll
          for (int r = 0; r < 2; r++) {
            float v[32];
            tmem_ld32(tlane + (uint32_t)((acs * R + r) * TN + 32 * c), v);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 8; k++) *(float4*)(stg + lane * 36 + 4 * k) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 8; k++) {
              float4 a = *(const float4*)(stg + (4 * k + pl0) * 36 + 4 * ch4);
              a.x += bv[c].x; a.y += bv[c].y; a.z += bv[c].z; a.w += bv[c].w;
              if (p.relu) { a.x = fmaxf(a.x, 0.0f); a.y = fmaxf(a.y, 0.0f); a.z = fmaxf(a.z, 0.0f); a.w = fmaxf(a.w, 0.0f); }
              if (r == 0) a0[k] = a; else a1[k] = a;
            }
            __syncwarp();
          }
          const int o = o0 + 32 * c + 4 * ch4;
#pragma unroll
          for (int k = 0; k < 8; k++) {
            float4 b0, b1;                 // the odd-column neighbour (lane ^ 8)
            b0.x = __shfl_xor_sync(0xffffffffu, a0[k].x, 8); b0.y = __shfl_xor_sync(0xffffffffu, a0[k].y, 8);
            b0.z = __shfl_xor_sync(0xffffffffu, a0[k].z, 8); b0.w = __shfl_xor_sync(0xffffffffu, a0[k].w, 8);
            b1.x = __shfl_xor_sync(0xffffffffu, a1[k].x, 8); b1.y = __shfl_xor_sync(0xffffffffu, a1[k].y, 8);
            b1.z = __shfl_xor_sync(0xffffffffu, a1[k].z, 8); b1.w = __shfl_xor_sync(0xffffffffu, a1[k].w, 8);
            const int ox = ox0 + 4 * k + pl0, px = ox >> 1;
            if ((pl0 & 1) == 0 && py < p.ph && px < p.pw && o + 4 <= p.Cout) {
              const float w4[4][4] = {{a0[k].x, b0.x, a1[k].x, b1.x}, {a0[k].y, b0.y, a1[k].y, b1.y}, {a0[k].z, b0.z, a1[k].z, b1.z}, {a0[k].w, b0.w, a1[k].w, b1.w}};
              float mx[4]; int mi[4];
#pragma unroll
              for (int e = 0; e < 4; e++) {
                mx[e] = -FLT_MAX; int pos = -1;
#pragma unroll
                for (int u = 0; u < 4; u++) if (w4[e][u] > mx[e]) { mx[e] = w4[e][u]; pos = u; }
                const long long plane = (long long)p.yh * p.yw;
                mi[e] = pos < 0 ? 0 : (int)(((long long)b * p.Cout + (o + e)) * plane + (long long)(oy0 + (pos >> 1)) * p.yw + ox + (pos & 1));
              }
              const long long off = (((long long)b * p.ph + py) * p.pw + px) * p.Cout + o;
              *(float4*)(p.pool_y + off) = make_float4(mx[0], mx[1], mx[2], mx[3]);
              *(int4*)(p.pool_idx + off) = make_int4(mi[0], mi[1], mi[2], mi[3]);
            }
          }
        }
      } else {
#pragma unroll
      for (int r = 0; r < R; r++) {
        const int oy = oy0 + r;
#pragma unroll
        for (int c = 0; c < TN / 32; c++) {
          float v[32];
          tmem_ld32(tlane + (uint32_t)((acs * R + r) * TN + 32 * c), v);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 8; k++) *(float4*)(stg + lane * 36 + 4 * k) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
          __syncwarp();
          const int o = o0 + 32 * c + 4 * ch4;
#pragma unroll
          for (int k = 0; k < 8; k++) {
            const int pl = 4 * k + pl0, ox = ox0 + pl;
            float4 a = *(const float4*)(stg + pl * 36 + 4 * ch4);
            a.x += bv[c].x; a.y += bv[c].y; a.z += bv[c].z; a.w += bv[c].w;
            if (p.relu) { a.x = fmaxf(a.x, 0.0f); a.y = fmaxf(a.y, 0.0f); a.z = fmaxf(a.z, 0.0f); a.w = fmaxf(a.w, 0.0f); }
            if (p.mask != nullptr) {      // 0*a keeps the NaN/Inf semantics of the un-fused multiply
              const uint32_t m = pre[r][c] >> (4 * k);
              a.x = (m & 1u) ? a.x : 0.0f * a.x; a.y = (m & 2u) ? a.y : 0.0f * a.y; a.z = (m & 4u) ? a.z : 0.0f * a.z; a.w = (m & 8u) ? a.w : 0.0f * a.w;
            }
            if (oy < p.yh && ox < p.yw && o + 4 <= p.Cout) {
              *(float4*)(p.y + (((long long)b * p.yh + oy) * p.yw + ox) * p.Cout + o) = a;
              cs[c].x += a.x; cs[c].y += a.y; cs[c].z += a.z; cs[c].w += a.w;
            }
          }
          __syncwarp();
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

template <int TN, int G, bool MB>
static int rows_launch(agb_ctx* ctx, RowsParams& p, size_t smem) {
  static bool attr = false;
  if (!attr) { AGB_CUDA(cudaFuncSetAttribute(conv_rows_kernel<TN, G, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr = true; }
  long long grid = p.num_tiles < ctx->sm_count ? p.num_tiles : ctx->sm_count;
  conv_rows_kernel<TN, G, MB><<<(unsigned)grid, RowsCfg<TN>::THREADS, smem, ctx->stream>>>(p);
  AGB_LAUNCHED(ctx);
  return AGB_OK;
}

// x [B,H,W,Cin] channels-last, wr = filter repacked to [tap][Cout][Cin] (see tc_conv.cu), y [B,yh,yw,Cout] channels-last.
// Returns AGB_ERR_UNSUPPORTED when the geometry is outside this kernel's envelope (the caller falls back to the per-tap kernel).
int agb_tc_conv_rows(agb_ctx* ctx, const float* x, const float* wr, float* y, int B, int Cin, int H, int W, int Cout, int yh, int yw, int kh, int kw,
                     int pad, int dil, const float* bias, int relu, const float* mask, float* csum, float* pool_y, int* pool_idx) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("AGB_CONV_ROWS"); enabled = (e && e[0] == '0') ? 0 : 1; }
  if (!enabled || yw < 128 || Cout > 128 || Cin % 4 != 0 || Cout % 4 != 0) return AGB_ERR_UNSUPPORTED;
  if (pool_y != nullptr && ((int64_t)B * Cout * yh * yw >= (1ll << 31) || (((uintptr_t)pool_y | (uintptr_t)pool_idx) & 15) != 0)) return AGB_ERR_UNSUPPORTED;     // int32 argmax offsets
  const int pitch = 128 + dil * (kw - 1), hrows = ROWS_R + dil * (kh - 1);
  if (pitch > 256 || hrows > 256) return AGB_ERR_UNSUPPORTED;
  RowsParams p;
  p.a_box_bytes = pitch * hrows * 128; p.a_slot_bytes = (p.a_box_bytes + 1023) & ~1023;
  const int TN = Cout > 64 ? 128 : 64;
  static const int g_env = [] { const char* e = getenv("AGB_ROWS_G"); return e ? atoi(e) : 0; }();
  const int G = (kw == 3 && (g_env == 3 || (g_env == 0 && TN == 64))) ? 3 : 1;
  const size_t fixed = 2 * (size_t)p.a_slot_bytes + 1024 + 256 + 4 * 32 * 36 * 4 + 4 * 128 * 4, st_bytes = (size_t)G * TN * 128;
  if (fixed + 2 * st_bytes > 227 * 1024) return AGB_ERR_UNSUPPORTED;
  int nb = (int)((227 * 1024 - fixed) / st_bytes); if (nb > ROWS_NB_MAX) nb = ROWS_NB_MAX;
  p.nb = nb;
  const size_t smem = fixed + (size_t)nb * st_bytes;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)Cin * 4, (uint64_t)W * Cin * 4, (uint64_t)H * W * Cin * 4};
    uint32_t box[4] = {32, (uint32_t)pitch, (uint32_t)hrows, 1};
    AGB_TRY(agb_make_tmap(&p.tmX, x, 4, dims, str, box, false));
  }
  {
    uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout, (uint64_t)(kh * kw)};
    uint64_t str[2] = {(uint64_t)Cin * 4, (uint64_t)Cin * Cout * 4};
    uint32_t box[3] = {32, (uint32_t)TN, (uint32_t)G};
    AGB_TRY(agb_make_tmap(&p.tmW, wr, 3, dims, str, box, false));
  }
  p.y = y; p.bias = bias; p.mask = mask; p.csum = csum; p.relu = relu; p.pool_y = pool_y; p.pool_idx = pool_idx; p.ph = yh / 2; p.pw = yw / 2; p.Cout = Cout; p.yh = yh; p.yw = yw; p.kw = kw; p.pad = pad; p.dil = dil;
  p.tiles_x = (yw + 127) / 128; p.tiles_y = (yh + ROWS_R - 1) / ROWS_R; p.otiles = (Cout + TN - 1) / TN;
  p.cblocks = (Cin + 31) / 32; p.taps = kh * kw; p.pitch = pitch;
  p.num_tiles = (long long)B * p.tiles_x * p.tiles_y * p.otiles;
  const int64_t ncta = p.num_tiles < ctx->sm_count ? p.num_tiles : ctx->sm_count;
  p.csum_part = nullptr;
  if (csum != nullptr && ctx->deterministic) AGB_TRY(agb_scratch2(ctx, (size_t)ncta * TN * sizeof(float), (void**)&p.csum_part));
  p.mask_bits = nullptr; p.bits_out = nullptr;
  if (mask != nullptr && ctx->mask_bits != nullptr && Cout % 32 == 0) { p.mask_bits = ctx->mask_bits; ctx->mask_bits_used = 1; }
  int r;
  const bool mb = p.mask_bits != nullptr || p.bits_out != nullptr;
#define AGB_ROWS_DISPATCH(MB_) do { if (G == 3) r = TN == 64 ? rows_launch<64, 3, MB_>(ctx, p, smem) : rows_launch<128, 3, MB_>(ctx, p, smem); \
                                    else r = TN == 64 ? rows_launch<64, 1, MB_>(ctx, p, smem) : rows_launch<128, 1, MB_>(ctx, p, smem); } while (0)
  if (mb) AGB_ROWS_DISPATCH(true); else AGB_ROWS_DISPATCH(false);
#undef AGB_ROWS_DISPATCH
  if (r == AGB_OK && p.csum_part != nullptr) r = agb_reduce_partials(ctx, p.csum_part, csum, (int)ncta, Cout, TN, 1);
  return r;
}
