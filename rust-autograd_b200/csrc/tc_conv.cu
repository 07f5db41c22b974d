// tc_conv.cu — tcgen05 implicit-GEMM convolutions for sm_100a on CHANNELS-LAST activations (logical NCHW tensors whose
// memory order is N,H,W,C), stride 1, square pad / dilation, built on the tile engine in tc_tile.cuh.
//
// Why channels-last: TMA requires the innermost box coordinate to be 16-byte aligned (measured: a 4-byte-shifted start raises
// an illegal-instruction fault).  In NCHW the 3x3 taps shift the innermost (W) axis by +-1 element, in channels-last they shift
// H and W which are outer axes, and the innermost axis (C) only ever moves in steps of 32 channels.  The im2col matrix of the
// reference (conv_ops/mod.rs:73-124; 9x the input for 3x3) is never materialised: TMA boxes with out-of-bounds zero fill ARE
// the im2col rows (padding = zero fill).
//
// fprop  (Conv2D::compute, conv2d.rs:115-211)   y[b,o,oy,ox] = sum_{c,i,j} w[o,c,i,j] x[b,c,oy+i*d-p,ox+j*d-p]
//   D[lane = pixel, col = o]: a CTA owns a 4x32 pixel patch of one image (128 TMEM lanes) x TN output channels.
//   k-block = (tap (i,j), 32 input channels).  P operand (K-major): ONE 4-D TMA box {32 c, 32 w, 4 h, 1 b} at the tap's offset
//   lands as [pixel][32 c] rows of 128 B (SWIZZLE_128B).  Q operand (K-major): filter repacked per call to wr[tap][o][c].
//   Epilogue: thread = pixel owns 32 consecutive output channels per tcgen05.ld = 128 contiguous bytes of y (float4 stores).
// dgrad  (Conv2DTranspose::compute, conv2d_transpose.rs:89-247), stride 1: the same kernel on gy with the filter repacked
//   flipped and channel-transposed (wr[tap'][c][o]) and pad' = d(k-1) - p.
// wgrad  (Conv2DFilterGrad::compute, conv2d.rs:631-734; the reference loops the batch sequentially with beta = 1):
//   gw[o,c,i,j] = sum_{b,oy,ox} gy[b,o,oy,ox] x[b,c,oy+i*d-p,ox+j*d-p]
//   D[lane = c, col = o] per tap, K = pixels: k-block = 32 consecutive ox of one (b, oy) row.  Both operands are MN-major
//   (channels contiguous): [32 px][32 ch] TMA boxes {32 c, 32 w, 1, 1} (128B_BASE32B swizzle).  With C == 64 two taps share
//   the 128 lanes.  Split-K over (b, oy) across CTAs, fp32 `red.global.add` into the zeroed gw.
#include "tc_tile.cuh"
#include <stdlib.h>
#include <vector>

// ------------------------------------------------------------------------------------------------ filter repack
// mode 0 (fprop): wr[t][o][c] = w[o][c][t];  mode 1 (dgrad): wr[T-1-t][c][o] = w[o][c][t]   (rows = output channel of the GEMM)
// wr_lo != nullptr (3xTF32): the filter arrives pre-split — wr = rna_tf32(w), wr_lo = rna_tf32(w - wr) in the same layout — so the per-tap
// kernel streams both planes by TMA and splits only the activations in shared memory (a round-to-nearest split on this side keeps the
// dropped lo*lo term free of a sign bias, see tc_tile.cuh).
__global__ void __launch_bounds__(256) repack_filter_kernel(const float* __restrict__ w, float* __restrict__ wr, float* __restrict__ wr_lo, int O, int C, int T, int mode) {
  int64_t n = (int64_t)O * C * T;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int t = (int)(i % T); int64_t r = i / T; int c = (int)(r % C); int o = (int)(r / C);
    float v = __ldg(w + i);
    const int64_t d = mode == 0 ? ((int64_t)t * O + o) * C + c : ((int64_t)(T - 1 - t) * C + c) * O + o;
    if (wr_lo != nullptr) { const float h = tf32_rna(v); wr[d] = h; wr_lo[d] = tf32_rna(v - h); }
    else wr[d] = v;
  }
}

// ------------------------------------------------------------------------------------------------ fprop / dgrad policy
#define AGB_CONV_MAX_TAPS 64
// MT_ = 2: the CTA owns an 8x32 pixel patch = two 128-lane M-tiles (rows 0-3 / 4-7, ONE {32 c, 32 w, 8 h} box per k-block) that
// share every filter tile: 1.33x (TN 128) / 1.5x (TN 256) fewer bytes through L2 -> smem per output, the bound of these kernels.
// BITS_: the ReLU sign-bit side channel (mask read as bits / sign bits of the stored activation written).  A separate instantiation: with both mask forms in
// one kernel the compiler scheduled the float-mask loads of the ordinary masked dgrad worse (0.57 -> 0.66 ms on the 256-wide layer).
template <int TN_, bool SPLIT_, int MT_ = 1, bool BITS_ = false> struct ConvFpropPol {
  static constexpr int TN = TN_, MT = MT_; static constexpr bool SPLIT = SPLIT_, P_MN = false, Q_MN = false, Q_PRESPLIT = SPLIT_;
  static constexpr bool SPLIT_PAIR2 = SPLIT_ && TN_ == 128 && MT_ == 1;      // 3xTF32 CTA pairs (tc_tile_split_pair_kernel): tmWh / tmWlh = 64-row boxes of the filter's hi / lo planes
  static constexpr bool PAIR2 = !SPLIT_ && TN_ == 256 && MT_ == 1;      // CTA pairs (tc_tile_pair_kernel): two adjacent pixel tiles, each CTA streams half of the filter rows (tmWlo = half-height boxes)
  static constexpr int OCC = (SPLIT_ || TN_ > 128 || MT_ > 1) ? 1 : 2;
  // The 128 lanes of an M-tile are a {bw w, bh h, bb images} pixel box (TMA writes box elements in exactly that order): 32x4x1 for
  // maps at least 32 wide; narrow maps take whole rows and, when a whole image is smaller than the tile, several images
  // (14x14 -> 14x9x1, 7x7 -> 7x7x2).  Lanes past bw*bh*bb read stale shared memory and are never stored.
  struct Params { CUtensorMap tmX, tmW, tmWlo, tmWh, tmWlh; float* y; const float* bias; const float* mask; float* csum; int relu; int B, Cout, yh, yw, stride, tiles_x, tiles_y, cblocks, taps;
                  int bw, bh, bb; uint32_t p_bytes; MnDescCfg mnc;
                  // tap k of the k-loop: input box origin = tile origin * stride + (dx[k], dy[k]), filter slice wt[k] of the repacked filter
                  short dy[AGB_CONV_MAX_TAPS], dx[AGB_CONV_MAX_TAPS], wt[AGB_CONV_MAX_TAPS];
                  // output pixel (oy, ox) of the tile grid lands at (oy * os + oyo, ox * os + oxo) of a [B, YH, YW, Cout] tensor
                  // (identity for convolutions; the s*s phases of a strided dgrad write interleaved sub-grids)
                  int os, oyo, oxo, YH, YW;
                  const uint32_t* mask_bits; uint32_t* bits_out;      // ReLU sign bits instead of mask_src (1 word per pixel per 32 channels) / of the stored activation
                  float* csum_part; };      // deterministic mode: slot (tile, M-tile, warp) stores its per-channel sums at csum_part[slot * Cout + channel]
  struct Tile { int b, oy0, ox0, o0; };
  __device__ static Tile tile(const Params& p, uint3 blk) {
    int bx = (int)blk.x; int tx = bx % p.tiles_x; int r = bx / p.tiles_x; int ty = r % p.tiles_y;
    return Tile{(r / p.tiles_y) * p.bb, ty * p.bh * MT, tx * p.bw, (int)blk.y * TN};
  }
  __device__ static uint32_t p_bytes(const Params& p, uint32_t) { return p.p_bytes; }
  __device__ static int64_t out_offset(const Params& p, int b, int oy, int ox) {
    return (((int64_t)b * p.YH + (oy * p.os + p.oyo)) * p.YW + (ox * p.os + p.oxo)) * p.Cout;
  }
  // lane -> pixel of M-tile mt
  __device__ static bool pixel(const Params& p, const Tile& t, int mt, int lane, int& b, int& oy, int& ox) {
    int w, h, bi;
    if (p.bw == 32) { w = lane & 31; h = lane >> 5; bi = 0; }          // the common 32x4x1 box: no divisions in the epilogue
    else { w = lane % p.bw; const int q = lane / p.bw; h = q % p.bh; bi = q / p.bh; }
    b = t.b + bi; oy = t.oy0 + p.bh * mt + h; ox = t.ox0 + w;
    return bi < p.bb && b < p.B && oy < p.yh && ox < p.yw;
  }
  __device__ static int num_kblocks(const Params& p, const Tile&) { return p.taps * p.cblocks; }
  __device__ static void prefetch(const Params& p) { tma_prefetch_desc(&p.tmX); tma_prefetch_desc(&p.tmW); if (SPLIT) tma_prefetch_desc(&p.tmWlo); }
  __device__ static void load(const Params& p, const Tile& t, int kb, uint8_t* pP, uint8_t* pQ, uint64_t* bar) {
    const int tap = kb / p.cblocks, cb = kb - tap * p.cblocks;
    tma_load_4d(pP, &p.tmX, bar, cb * 32, t.ox0 * p.stride + p.dx[tap], t.oy0 * p.stride + p.dy[tap], t.b);     // dims {c, w, h, b}; a strided conv walks
    tma_load_3d(pQ, &p.tmW, bar, cb * 32, t.o0, p.wt[tap]);                                                      // the box with element stride s
  }
  __device__ static void load_q_lo(const Params& p, const Tile& t, int kb, uint8_t* pQlo, uint64_t* bar) {          // 3xTF32: the pre-split filter's low plane
    const int tap = kb / p.cblocks, cb = kb - tap * p.cblocks;
    tma_load_3d(pQlo, &p.tmWlo, bar, cb * 32, t.o0, p.wt[tap]);
  }
  __device__ static void load_sp2(const Params& p, const Tile& t, int kb, uint8_t* pP, uint8_t* pQh, uint8_t* pQl, uint64_t* bar, int rank) {     // 3xTF32 CTA pair: own pixels + half of the filter rows, both planes
    const int tap = kb / p.cblocks, cb = kb - tap * p.cblocks;
    tma_load_4d(pP, &p.tmX, bar, cb * 32, t.ox0 * p.stride + p.dx[tap], t.oy0 * p.stride + p.dy[tap], t.b);
    tma_load_3d(pQh, &p.tmWh, bar, cb * 32, t.o0 + rank * (TN / 2), p.wt[tap]);
    tma_load_3d(pQl, &p.tmWlh, bar, cb * 32, t.o0 + rank * (TN / 2), p.wt[tap]);
  }
  __device__ static void load2(const Params& p, const Tile& t, int kb, uint8_t* pP, uint8_t* pQ, uint32_t bar, int rank) {     // CTA pair: own pixels + half of the filter rows
    const int tap = kb / p.cblocks, cb = kb - tap * p.cblocks;
    tma_load_4d_2sm(pP, &p.tmX, bar, cb * 32, t.ox0 * p.stride + p.dx[tap], t.oy0 * p.stride + p.dy[tap], t.b);
    tma_load_3d_2sm(pQ, &p.tmWlo, bar, cb * 32, t.o0 + rank * (TN / 2), p.wt[tap]);
  }
  // dgrad + ReLU backward of the layer below: bit j of pre[c] = (mask_src[pixel, o0 + 32c + j] > 0).  Read while the MMAs run, so the
  // strided (one pixel per thread) loads cost no epilogue latency; default-cached so both halves of a 32-byte sector are used.
  __device__ static void pre_epilogue(const Params& p, const Tile& t, int lane, uint32_t* pre) {
    if (p.mask == nullptr) return;
#pragma unroll
    for (int mt = 0; mt < MT; mt++) {
      int b, oy, ox;
      const bool in = pixel(p, t, mt, lane, b, oy, ox);
      const int64_t moff = out_offset(p, in ? b : 0, in ? oy : 0, in ? ox : 0) + t.o0;
      const float* m = p.mask + moff;
#pragma unroll
      for (int c = 0; c < TN / 32; c++) {
        uint32_t bits = 0;
        const int o = t.o0 + 32 * c;
        if (BITS_ && p.mask_bits != nullptr) { if (in && o < p.Cout) bits = __ldg(p.mask_bits + ((moff + 32 * c) >> 5)); }      // Cout % 32 == 0 (host)
        else if (in && o + 32 <= p.Cout) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 mv = __ldg((const float4*)(m + 32 * c + j));
            bits |= (mv.x > 0.0f ? 1u : 0u) << j; bits |= (mv.y > 0.0f ? 1u : 0u) << (j + 1);
            bits |= (mv.z > 0.0f ? 1u : 0u) << (j + 2); bits |= (mv.w > 0.0f ? 1u : 0u) << (j + 3);
          }
        } else if (in) {
          for (int j = 0; j < 32; j++) if (o + j < p.Cout && __ldg(m + 32 * c + j) > 0.0f) bits |= 1u << j;
        }
        pre[mt * (TN / 32) + c] = bits;
      }
    }
  }
  // `pre_io`: in = the ReLU-mask bits of this chunk (pre_epilogue); out = the sign bits of the stored chunk when bits_out is set.  The caller's words of one
  // M-tile are consecutive, so the LAST chunk writes all of a pixel's words with 128-bit stores (one 4-byte store per chunk at a 32-byte pitch made
  // every warp store touch 32 sectors and cost 0.3 ms per training step).
  __device__ static void store(const Params& p, const Tile& t, int mt, int lane, int c0, const float* v, uint32_t& pre_io) {
    const uint32_t pre = pre_io;
    int b, oy, ox;
    const bool in = pixel(p, t, mt, lane, b, oy, ox);
    if (!in && p.csum == nullptr) return;
    const int o = t.o0 + c0;
    float* dst = p.y + out_offset(p, in ? b : 0, in ? oy : 0, in ? ox : 0) + o;
    // fused epilogue: per-channel bias and ReLU applied to the accumulator registers, so the
    // pre-activation tensors of conv -> add -> relu never travel through HBM
    float r[32];
    if (p.bias != nullptr && o + 32 <= p.Cout && ((((uintptr_t)p.bias) & 15) == 0)) {       // eight 128-bit (warp-uniform) loads instead of 32 scalar ones: the drain is
#pragma unroll                                                                               // the critical path of the short-K layers
      for (int j = 0; j < 32; j += 4) {
        const float4 b4 = __ldg((const float4*)(p.bias + o + j));
        r[j] = v[j] + b4.x; r[j + 1] = v[j + 1] + b4.y; r[j + 2] = v[j + 2] + b4.z; r[j + 3] = v[j + 3] + b4.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; j++) {
        float a = v[j];
        if (p.bias != nullptr && o + j < p.Cout) a += __ldg(p.bias + o + j);
        r[j] = a;
      }
    }
    if (p.relu) {
#pragma unroll
      for (int j = 0; j < 32; j++) r[j] = fmaxf(r[j], 0.0f);
    }
    if (p.mask != nullptr) {           // gx *= (mask_src > 0); 0*r keeps the NaN/Inf semantics of the un-fused multiply
#pragma unroll
      for (int j = 0; j < 32; j++) r[j] = ((pre >> j) & 1u) ? r[j] : 0.0f * r[j];
    }
    if (in) {
      if (o + 32 <= p.Cout) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) *(float4*)(dst + j) = make_float4(r[j], r[j + 1], r[j + 2], r[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; j++) if (o + j < p.Cout) dst[j] = r[j];
      }
    }
    if (BITS_ && p.bits_out != nullptr) {
      uint32_t sign = 0;
#pragma unroll
      for (int j = 0; j < 32; j++) sign |= (r[j] > 0.0f ? 1u : 0u) << j;
      pre_io = sign;
      const int nw = (min(TN, p.Cout - t.o0) + 31) >> 5, ci = c0 >> 5;
      if (ci == nw - 1 && in) {
        const uint32_t* pw = &pre_io - ci;                          // this M-tile's words 0 .. nw-1 (registers: every index is a compile-time constant after unrolling)
        uint32_t* d = p.bits_out + ((out_offset(p, b, oy, ox) + t.o0) >> 5);
        if ((nw & 3) == 0 && ((((uintptr_t)d) & 15) == 0)) {
#pragma unroll
          for (int q = 0; q < TN / 32; q += 4) if (q < nw) *(uint4*)(d + q) = make_uint4(pw[q], pw[q + 1], pw[q + 2], pw[q + 3]);
        } else {
#pragma unroll
          for (int q = 0; q < TN / 32; q++) if (q < nw) d[q] = pw[q];
        }
      }
    }
    if (p.csum != nullptr) {
      // per-channel sums of the stored tile (the bias gradient of the layer below, reduce_sum over b,h,w): butterfly
      // transpose-reduce across the warp — 31 shuffles leave the sum of channel l in lane l — then one red.global.add per lane
      const int wl = lane & 31;
#pragma unroll
      for (int j = 0; j < 32; j++) r[j] = in ? r[j] : 0.0f;
#pragma unroll
      for (int off = 16, n = 32; off >= 1; off >>= 1, n >>= 1) {
        const bool upper = (wl & off) != 0;
#pragma unroll
        for (int i = 0; i < 16; i++) {
          if (i < n / 2) {
            const float send = upper ? r[i] : r[i + n / 2], keep = upper ? r[i + n / 2] : r[i];
            r[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
          }
        }
      }
      if (o + wl < p.Cout) {
        if (p.csum_part != nullptr) {
          const int64_t tlin = ((int64_t)(t.b / p.bb) * p.tiles_y + t.oy0 / (p.bh * MT)) * p.tiles_x + t.ox0 / p.bw;
          p.csum_part[((tlin * MT + mt) * 4 + ((lane >> 5) & 3)) * p.Cout + o + wl] = r[0];
        } else red_add_f32(p.csum + o + wl, r[0]);
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------ wgrad policy
// MT_ = 2 (non-PAIR): the CTA owns two (tap, 128-channel tile) units — two M-tiles that share every gy tile, so gy is pulled from
// L2 half as often (the TN = 256 wgrad moved 48 KB per 4 MMAs: bound by L2 -> smem ingest at 43 % tensor activity).
// QPRE_ (3xTF32): gy arrives as hi = rna_tf32 / lo planes written once by agb_tc_presplit (12 B/element of HBM traffic) instead of being split per landed tile
// (48 of the 192 KB a 128-wide stage moves through shared memory, the bound of this mode): pays when the tile count per gy element (C * taps / 128) is large.
template <int TN_, bool SPLIT_, bool PAIR_, int MT_ = 1, bool QPRE_ = false> struct ConvWgradPol {
  static constexpr int TN = TN_, MT = MT_; static constexpr bool SPLIT = SPLIT_, P_MN = true, Q_MN = true, Q_PRESPLIT = QPRE_;
  static_assert(!QPRE_ || (SPLIT_ && !PAIR_ && MT_ == 1), "pre-split gy: the 3xTF32 one-M-tile kernel");
  static constexpr bool SPLIT_PAIR2 = false;
  static constexpr bool PAIR2 = !SPLIT_ && TN_ == 256 && MT_ == 1 && !PAIR_;      // CTA pairs: two (tap, 128-channel tile) units share every gy tile, each CTA streams half of its columns
  static constexpr int OCC = (SPLIT_ || TN_ > 128 || MT_ > 1) ? 1 : 2;
  static_assert(!(PAIR_ && MT_ > 1), "tap pairing within one M-tile and two M-tiles are alternatives");
  // tmX5 / tmG5 (wide = 1; C % 32 == 0 and O % 32 == 0): the channel axis split into {32, C / 32}, so ONE box {32 c, pixels, 4 blocks} lands as the four (TN / 32)
  // consecutive 4 KB boxes the MMA descriptors expect — 3 TMA instructions per k-block instead of 16: ncu showed the producer warp 90 % busy issuing them
  // (each cp.async.bulk.tensor of a lane-0 branch is an ELECT / uniform-branch waterfall) and the MMA issuer 20 % of its time waiting on `full`
  struct Params { CUtensorMap tmX, tmG, tmX5, tmG5, tmGl, tmG5l /* QPRE_: the lo plane of gy (tmG / tmG5 address the hi plane) */; int wide; float* gw; int C, O, T, kw, pad, dil, stride, yh, xblocks, kb_total, kb_per_split; MnDescCfg mnc; int64_t part_stride /* > 0: split z stores at gw + z * part_stride (deterministic mode) */; };
  struct Tile { int c0, tapA, tapB, o0, q0, q1, c1; };       // MT == 2: unit 0 = (tapA, c0), unit 1 = (tapB, c1)
  __device__ static Tile tile(const Params& p, uint3 blk) {
    Tile t; t.o0 = (int)blk.y * TN; t.c1 = 0;
    if (PAIR_) { t.c0 = 0; t.tapA = 2 * (int)blk.x; t.tapB = t.tapA + 1; }
    else if (MT == 2) {
      const int ctiles = (p.C + 127) / 128, u0 = 2 * (int)blk.x, u1 = u0 + 1;
      t.c0 = (u0 % ctiles) * 128; t.tapA = u0 / ctiles; t.c1 = (u1 % ctiles) * 128; t.tapB = u1 / ctiles;      // tapB may be >= T: absent unit
    }
    else { int ctiles = (p.C + 127) / 128; t.c0 = ((int)blk.x % ctiles) * 128; t.tapA = (int)blk.x / ctiles; t.tapB = -1; }
    t.q0 = (int)blk.z * p.kb_per_split; t.q1 = min(t.q0 + p.kb_per_split, p.kb_total);
    return t;
  }
  __device__ static int num_kblocks(const Params&, const Tile& t) { return t.q1 > t.q0 ? t.q1 - t.q0 : 0; }
  __device__ static uint32_t p_bytes(const Params&, uint32_t full) { return full; }
  __device__ static void prefetch(const Params& p) { tma_prefetch_desc(&p.tmX); tma_prefetch_desc(&p.tmG); if (QPRE_) tma_prefetch_desc(&p.tmGl); }
  __device__ static void load_q_lo(const Params& p, const Tile& t, int kb, uint8_t* pQlo, uint64_t* bar) {
    const int q = t.q0 + kb; const int xb = q % p.xblocks; const int r = q / p.xblocks; const int oy = r % p.yh, b = r / p.yh;
    const int ox0 = xb * 32;
    if (p.wide) { tma_load_5d(pQlo, &p.tmG5l, bar, 0, ox0, t.o0 >> 5, oy, b); return; }
#pragma unroll
    for (int g = 0; g < TN / 32; g++) tma_load_4d(pQlo + g * 4096, &p.tmGl, bar, t.o0 + 32 * g, ox0, oy, b);
  }
  __device__ static void load(const Params& p, const Tile& t, int kb, uint8_t* pP, uint8_t* pQ, uint64_t* bar) {
    const int q = t.q0 + kb; const int xb = q % p.xblocks; const int r = q / p.xblocks; const int oy = r % p.yh, b = r / p.yh;
    const int ox0 = xb * 32;
    const int iA = t.tapA / p.kw, jA = t.tapA - iA * p.kw;
    const int xA = ox0 * p.stride + jA * p.dil - p.pad, yA = oy * p.stride + iA * p.dil - p.pad;       // strided conv: the x box walks w with element stride s
    if (p.wide) {      // one box per operand tile (the channel axis as {32, blocks})
      if (PAIR_) {
        int xB = xA, yB = yA, cB = (p.C >> 5) + 2;                           // tap B absent: fully out of bounds = zero rows
        if (t.tapB < p.T) { const int iB = t.tapB / p.kw, jB = t.tapB - iB * p.kw; xB = ox0 * p.stride + jB * p.dil - p.pad; yB = oy * p.stride + iB * p.dil - p.pad; cB = 0; }
        tma_load_5d(pP, &p.tmX5, bar, 0, xA, 0, yA, b);
        tma_load_5d(pP + 8192, &p.tmX5, bar, 0, xB, cB, yB, b);
      } else {
        tma_load_5d(pP, &p.tmX5, bar, 0, xA, t.c0 >> 5, yA, b);
        if (MT == 2) {
          int xB = xA, yB = yA, cB = (p.C >> 5) + 4;
          if (t.tapB < p.T) { const int iB = t.tapB / p.kw, jB = t.tapB - iB * p.kw; xB = ox0 * p.stride + jB * p.dil - p.pad; yB = oy * p.stride + iB * p.dil - p.pad; cB = t.c1 >> 5; }
          tma_load_5d(pP + 16384, &p.tmX5, bar, 0, xB, cB, yB, b);
        }
      }
      tma_load_5d(pQ, &p.tmG5, bar, 0, ox0, t.o0 >> 5, oy, b);
      return;
    }
    if (PAIR_) {      // lanes 0-63: channels 0..63 at tap A, lanes 64-127: channels 0..63 at tap B
      int xB = xA, yB = yA, cB = p.C + 64;                                   // tap B absent (odd tap count): fully out of bounds = zero rows
      if (t.tapB < p.T) { const int iB = t.tapB / p.kw, jB = t.tapB - iB * p.kw; xB = ox0 * p.stride + jB * p.dil - p.pad; yB = oy * p.stride + iB * p.dil - p.pad; cB = 0; }
      tma_load_4d(pP, &p.tmX, bar, 0, xA, yA, b);
      tma_load_4d(pP + 4096, &p.tmX, bar, 32, xA, yA, b);
      tma_load_4d(pP + 8192, &p.tmX, bar, cB, xB, yB, b);
      tma_load_4d(pP + 12288, &p.tmX, bar, cB + 32, xB, yB, b);
    } else {
#pragma unroll
      for (int g = 0; g < 4; g++) tma_load_4d(pP + g * 4096, &p.tmX, bar, t.c0 + 32 * g, xA, yA, b);
      if (MT == 2) {
        int xB = xA, yB = yA, cB = p.C + 128;                                // absent unit: fully out of bounds = zero rows
        if (t.tapB < p.T) { const int iB = t.tapB / p.kw, jB = t.tapB - iB * p.kw; xB = ox0 * p.stride + jB * p.dil - p.pad; yB = oy * p.stride + iB * p.dil - p.pad; cB = t.c1; }
#pragma unroll
        for (int g = 0; g < 4; g++) tma_load_4d(pP + 16384 + g * 4096, &p.tmX, bar, cB + 32 * g, xB, yB, b);
      }
    }
#pragma unroll
    for (int g = 0; g < TN / 32; g++) tma_load_4d(pQ + g * 4096, &p.tmG, bar, t.o0 + 32 * g, ox0, oy, b);
  }
  __device__ static void load2(const Params& p, const Tile& t, int kb, uint8_t* pP, uint8_t* pQ, uint32_t bar, int rank) {
    const int q = t.q0 + kb; const int xb = q % p.xblocks; const int r = q / p.xblocks; const int oy = r % p.yh, b = r / p.yh;
    const int ox0 = xb * 32;
    const int iA = t.tapA / p.kw, jA = t.tapA - iA * p.kw;
    const int xA = ox0 * p.stride + jA * p.dil - p.pad, yA = oy * p.stride + iA * p.dil - p.pad;
#pragma unroll
    for (int g = 0; g < 4; g++) tma_load_4d_2sm(pP + g * 4096, &p.tmX, bar, t.c0 + 32 * g, xA, yA, b);
#pragma unroll
    for (int g = 0; g < TN / 64; g++) tma_load_4d_2sm(pQ + g * 4096, &p.tmG, bar, t.o0 + rank * (TN / 2) + 32 * g, ox0, oy, b);
  }
  __device__ static void pre_epilogue(const Params&, const Tile&, int, uint32_t*) {}
  __device__ static void store(const Params& p, const Tile& t, int mt, int lane, int c0, const float* v, uint32_t&) {
    int c, tap;
    if (PAIR_) { c = lane & 63; tap = lane < 64 ? t.tapA : t.tapB; }
    else if (MT == 2 && mt == 1) { c = t.c1 + lane; tap = t.tapB; }
    else { c = t.c0 + lane; tap = t.tapA; }
    if (c >= p.C || tap >= p.T) return;
    if (p.part_stride > 0) {       // deterministic mode: this split's partial, laid out [tap][o][c] so that a warp (32 consecutive c) stores 128 contiguous bytes
      float* base = p.gw + (int64_t)(t.q0 / p.kb_per_split) * p.part_stride + (int64_t)tap * p.O * p.C + c;
#pragma unroll
      for (int j = 0; j < 32; j++) { const int o = t.o0 + c0 + j; if (o < p.O) base[(int64_t)o * p.C] = v[j]; }
      return;
    }
#pragma unroll
    for (int j = 0; j < 32; j++) {
      const int o = t.o0 + c0 + j;
      if (o < p.O) red_add_f32(p.gw + ((int64_t)o * p.C + c) * p.T + tap, v[j]);
    }
  }
};

// ------------------------------------------------------------------------------------------------ host side
int agb_tc_conv_rows(agb_ctx* ctx, const float* x, const float* wr, float* y, int B, int Cin, int H, int W, int Cout, int yh, int yw, int kh, int kw,
                     int pad, int dil, const float* bias, int relu, const float* mask, float* csum, float* pool_y, int* pool_idx);
bool agb_tc_conv_eligible(int C, int O, int kh, int kw, int stride, int yw);
int agb_tc_conv_cols(agb_ctx* ctx, const float* x, const float* wr, float* y, int B, int Cin, int H, int W, int Cout, int yh, int yw, int kh, int kw,
                     int pad, int dil, const float* bias, int relu, const float* mask, float* csum);
int agb_tc_conv_wgrad_taps(agb_ctx* ctx, const float* img, const float* g, float* gw, int B, int C, int H, int W, int O, int yh, int yw,
                           int kh, int kw, int pad, int dil);
// channels-last tensor map of a logical [B, C, H, W] activation: dims {c, w, h, b}
static int make_cl_map(CUtensorMap* m, const float* p, int B, int C, int H, int W, uint32_t bc, uint32_t bw, uint32_t bh, bool atom32) {
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  uint64_t str[3] = {(uint64_t)C * 4, (uint64_t)W * C * 4, (uint64_t)H * W * C * 4};
  uint32_t box[4] = {bc, bw, bh, 1};
  return agb_make_tmap(m, p, 4, dims, str, box, atom32);
}

bool agb_tc_conv_fprop_eligible(int C, int O, int kh, int kw, int stride, int yw) {       // forward: strides 1..4
  return stride >= 1 && stride <= 4 && agb_tc_conv_eligible(C, O, kh, kw, 1, yw);
}
bool agb_tc_conv_eligible(int C, int O, int kh, int kw, int stride, int yw) {
  return stride == 1 && kh == kw && C >= 32 && C % 4 == 0 && O >= 32 && O % 4 == 0 && yw >= 4;       // narrow maps: see ConvFpropPol::Params (whole rows / several images per tile)
}

// strided-dgrad phase: explicit tap list and interleaved output sub-grid (nullptr = an ordinary convolution)
struct ConvPhase { int ntaps; short dy[AGB_CONV_MAX_TAPS], dx[AGB_CONV_MAX_TAPS], wt[AGB_CONV_MAX_TAPS]; int os, oyo, oxo, YH, YW; };

template <int TN, bool SPLIT, int MT, bool BITS>
static int fprop_launch_impl(agb_ctx* ctx, const float* x, const float* wr, float* y, int B, int Cin, int H, int W, int Cout, int yh, int yw, int kh, int kw,
                             int pad, int dil, const float* bias, int relu, const float* mask, float* csum, int stride, const ConvPhase* ph);
template <int TN, bool SPLIT, int MT = 1>
static int fprop_launch(agb_ctx* ctx, const float* x, const float* wr, float* y, int B, int Cin, int H, int W, int Cout, int yh, int yw, int kh, int kw,
                        int pad, int dil, const float* bias, int relu, const float* mask, float* csum, int stride = 1, const ConvPhase* ph = nullptr) {
  // the sign-bit instantiation exists for the one-M-tile kernels (single-pass and 3xTF32: the masked dgrads of the f32-faithful step read 1.07 / 0.54 / 0.27 GB of float
  // masks in their pre-epilogues, +0.2 .. 0.9 ms per layer over the unmasked kernel of the same shape)
  const bool bits = MT == 1 && Cout % 32 == 0 && ph == nullptr && ((mask != nullptr && ctx->mask_bits != nullptr) || ctx->bits_out != nullptr);
  if (bits) return fprop_launch_impl<TN, SPLIT, 1, true>(ctx, x, wr, y, B, Cin, H, W, Cout, yh, yw, kh, kw, pad, dil, bias, relu, mask, csum, stride, ph);
  return fprop_launch_impl<TN, SPLIT, MT, false>(ctx, x, wr, y, B, Cin, H, W, Cout, yh, yw, kh, kw, pad, dil, bias, relu, mask, csum, stride, ph);
}
template <int TN, bool SPLIT, int MT, bool BITS>
static int fprop_launch_impl(agb_ctx* ctx, const float* x, const float* wr, float* y, int B, int Cin, int H, int W, int Cout, int yh, int yw, int kh, int kw,
                             int pad, int dil, const float* bias, int relu, const float* mask, float* csum, int stride, const ConvPhase* ph) {
  using Pol = ConvFpropPol<TN, SPLIT, MT, BITS>;
  typename Pol::Params p;
  if (kh * kw > AGB_CONV_MAX_TAPS) return AGB_ERR_UNSUPPORTED;
  int bw = 32, bh = 4, bb = 1;
  if (yw < 32) { bw = yw; bh = yh < 128 / bw ? yh : 128 / bw; bb = bh == yh ? 128 / (bw * bh) : 1; if (bb > B) bb = B; if (bb < 1) bb = 1; }
  if (MT > 1 && (bw != 32 || bb != 1)) return AGB_ERR_UNSUPPORTED;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)Cin * 4, (uint64_t)W * Cin * 4, (uint64_t)H * W * Cin * 4};
    const uint32_t su = (uint32_t)stride;
    uint32_t box[4] = {32, (uint32_t)(bw - 1) * su + 1, (uint32_t)(bh * MT - 1) * su + 1, (uint32_t)bb}, es[4] = {1, su, su, 1};
    if (box[1] > 256 || box[2] > 256) return AGB_ERR_UNSUPPORTED;
    AGB_TRY(agb_make_tmap(&p.tmX, x, 4, dims, str, box, false, stride > 1 ? es : nullptr));
  }
  p.stride = stride; p.bw = bw; p.bh = bh; p.bb = bb; p.B = B; p.p_bytes = (uint32_t)(128 * bw * bh * MT * bb);
  {  // wr[tap][o][c]
    uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout, (uint64_t)(kh * kw)};
    uint64_t str[2] = {(uint64_t)Cin * 4, (uint64_t)Cin * Cout * 4};
    uint32_t box[3] = {32, (uint32_t)TN, 1};
    AGB_TRY(agb_make_tmap(&p.tmW, wr, 3, dims, str, box, false));
    if (SPLIT) AGB_TRY(agb_make_tmap(&p.tmWlo, wr + (size_t)kh * kw * Cout * Cin, 3, dims, str, box, false));      // the low plane follows the repacked filter (repack_filter_kernel)
    else if (Pol::PAIR2) { uint32_t hbox[3] = {32, (uint32_t)TN / 2, 1}; AGB_TRY(agb_make_tmap(&p.tmWlo, wr, 3, dims, str, hbox, false)); }
    else p.tmWlo = p.tmW;
    p.tmWh = p.tmW; p.tmWlh = p.tmWlo;
    if (Pol::SPLIT_PAIR2) {
      uint32_t hbox[3] = {32, (uint32_t)TN / 2, 1};
      AGB_TRY(agb_make_tmap(&p.tmWh, wr, 3, dims, str, hbox, false));
      AGB_TRY(agb_make_tmap(&p.tmWlh, wr + (size_t)kh * kw * Cout * Cin, 3, dims, str, hbox, false));
    }
  }
  p.y = y; p.bias = bias; p.mask = mask; p.csum = csum; p.relu = relu; p.Cout = Cout; p.yh = yh; p.yw = yw;
  p.mask_bits = nullptr; p.bits_out = nullptr;
  if (BITS && Cout % 32 == 0 && ph == nullptr) {
    if (mask != nullptr && ctx->mask_bits != nullptr) { p.mask_bits = ctx->mask_bits; ctx->mask_bits_used = 1; }
    if (ctx->bits_out != nullptr) { p.bits_out = ctx->bits_out; ctx->bits_written = 1; }
  }
  p.tiles_x = (yw + bw - 1) / bw; p.tiles_y = (yh + bh * MT - 1) / (bh * MT); p.cblocks = (Cin + 31) / 32; p.mnc = agb_mn_cfg();
  if (ph == nullptr) {
    p.taps = kh * kw; p.os = 1; p.oyo = 0; p.oxo = 0; p.YH = yh; p.YW = yw;
    for (int t = 0; t < p.taps; t++) { p.dy[t] = (short)((t / kw) * dil - pad); p.dx[t] = (short)((t % kw) * dil - pad); p.wt[t] = (short)t; }
  } else {
    p.taps = ph->ntaps; p.os = ph->os; p.oyo = ph->oyo; p.oxo = ph->oxo; p.YH = ph->YH; p.YW = ph->YW;
    for (int t = 0; t < p.taps; t++) { p.dy[t] = ph->dy[t]; p.dx[t] = ph->dx[t]; p.wt[t] = ph->wt[t]; }
  }
  int64_t nb = (int64_t)p.tiles_x * p.tiles_y * ((B + bb - 1) / bb);
  if (nb > 2147483647ll) return AGB_ERR_UNSUPPORTED;
  dim3 grid((unsigned)nb, (unsigned)((Cout + TN - 1) / TN), 1);
  // deterministic per-channel sums: one slot per (tile, M-tile, epilogue warp), added in a fixed order by two small reductions (64 interleaved groups, then the groups)
  p.csum_part = nullptr;
  const int64_t nparts = nb * MT * 4;
  if (csum != nullptr && ctx->deterministic) AGB_TRY(agb_scratch2(ctx, agb_reduce_partials2_floats(nparts, Cout) * sizeof(float), (void**)&p.csum_part));
  AGB_TRY(tc_tile_launch<Pol>(ctx, p, grid));
  if (p.csum_part != nullptr) AGB_TRY(agb_reduce_partials2(ctx, p.csum_part, nparts, Cout, csum, 1));
  return AGB_OK;
}

// fprop on channels-last buffers: x [B,H,W,C], w [O,C,kh,kw] (plain) -> y [B,yh,yw,O].  flip_transpose != 0: dgrad — `x` is gy
// with C = filter dim 0, w [C, O(=out channels of this GEMM), kh, kw]; the effective padding is dil*(k-1) - pad.
int agb_tc_conv_fprop(agb_ctx* ctx, int mode, const float* x, const float* w, float* y, int B, int C, int H, int W, int O, int kh, int kw,
                      int pad, int stride, int dil, int flip_transpose, const float* bias, int relu, const float* mask, float* csum, float* pool_y, int* pool_idx) {
  const int epad = flip_transpose ? dil * (kh - 1) - pad : pad;
  if (epad < 0) return AGB_ERR_UNSUPPORTED;
  const int yh = (H + 2 * epad - (dil * (kh - 1) + 1)) / stride + 1, yw = (W + 2 * epad - (dil * (kw - 1) + 1)) / stride + 1;
  // strided convolutions: forward only (the tile's TMA box walks the input with element stride s); dgrad of a strided conv is not a conv
  if (yh < 1 || yw < 1 || stride < 1 || stride > 4 || (stride > 1 && flip_transpose) || !agb_tc_conv_eligible(C, O, kh, kw, 1, yw)) return AGB_ERR_UNSUPPORTED;
  if ((((uintptr_t)x | (uintptr_t)y) & 15) != 0) return AGB_ERR_UNSUPPORTED;
  const int T = kh * kw;
  const bool split = mode == AGB_MATH_3XTF32;
  float* wr = nullptr;
  AGB_TRY(agb_scratch(ctx, (size_t)T * O * C * sizeof(float) * (split ? 2 : 1), (void**)&wr));
  {
    int64_t n = (int64_t)O * C * T;
    float* wr_lo = split ? wr + n : nullptr;
    // fprop: w is [O][C][T] -> wr[t][O][C];  dgrad: w is [C(in)][O(out)][T] -> wr[T-1-t][O(out)][C(in)]
    if (!flip_transpose) repack_filter_kernel<<<agb_grid_for(n, 256, ctx->sm_count, 4), 256, 0, ctx->stream>>>(w, wr, wr_lo, O, C, T, 0);
    else repack_filter_kernel<<<agb_grid_for(n, 256, ctx->sm_count, 4), 256, 0, ctx->stream>>>(w, wr, wr_lo, C, O, T, 1);
    AGB_LAUNCHED(ctx);
  }
  if ((split || stride > 1) && pool_y != nullptr) return AGB_ERR_UNSUPPORTED;
  if (!split && stride == 1) {       // wide feature maps: persistent halo-reusing kernel (tc_conv_rows.cu), 3x less L2 -> smem traffic
    int r = agb_tc_conv_rows(ctx, x, wr, y, B, C, H, W, O, yh, yw, kh, kw, epad, dil, bias, relu, mask, csum, pool_y, pool_idx);
    if (r != AGB_ERR_UNSUPPORTED) return r;
    if (pool_y != nullptr) return AGB_ERR_UNSUPPORTED;          // the fused pooling epilogue exists in the wide-map kernel only
    r = agb_tc_conv_cols(ctx, x, wr, y, B, C, H, W, O, yh, yw, kh, kw, epad, dil, bias, relu, mask, csum);     // narrow maps, Cout <= 128
    if (r != AGB_ERR_UNSUPPORTED) return r;
  }
  if (split) {
    if (O > 64) return fprop_launch<128, true>(ctx, x, wr, y, B, C, H, W, O, yh, yw, kh, kw, epad, dil, bias, relu, mask, csum, stride);
    return fprop_launch<64, true>(ctx, x, wr, y, B, C, H, W, O, yh, yw, kh, kw, epad, dil, bias, relu, mask, csum, stride);
  }
  static int m2 = -1;     // two M-tiles per CTA: measured slower than one (3-stage ring, doubled epilogue) — kept as an opt-in experiment
  if (m2 < 0) { const char* e = getenv("AGB_CONV_M2"); m2 = (e && e[0] == '1') ? 1 : 0; }
  const bool tall = m2 && stride == 1 && yh >= 8 && yw >= 32 && (int64_t)B * ((yh + 7) / 8) * ((yw + 31) / 32) >= ctx->sm_count;       // enough 8-row patches to fill the machine
  if (O > 128) return tall ? fprop_launch<256, false, 2>(ctx, x, wr, y, B, C, H, W, O, yh, yw, kh, kw, epad, dil, bias, relu, mask, csum)
                           : fprop_launch<256, false>(ctx, x, wr, y, B, C, H, W, O, yh, yw, kh, kw, epad, dil, bias, relu, mask, csum, stride);
  if (O > 64) return tall ? fprop_launch<128, false, 2>(ctx, x, wr, y, B, C, H, W, O, yh, yw, kh, kw, epad, dil, bias, relu, mask, csum)
                          : fprop_launch<128, false>(ctx, x, wr, y, B, C, H, W, O, yh, yw, kh, kw, epad, dil, bias, relu, mask, csum, stride);
  return fprop_launch<64, false>(ctx, x, wr, y, B, C, H, W, O, yh, yw, kh, kw, epad, dil, bias, relu, mask, csum, stride);
}

// Strided dgrad (Conv2DTranspose::compute, conv2d_transpose.rs:89-247, stride s > 1) as s*s unit-stride convolutions over gy:
//   gx[b, c, s u + py, s v + px] = sum over the taps (i, j) with (py + p - i d) % s == 0 and (px + p - j d) % s == 0 of
//                                  sum_o gy[b, o, u + (py + p - i d)/s, v + (px + p - j d)/s] * w[o, c, i, j]
// Each phase (py, px) is one launch of the per-tap tile kernel with its own tap list, writing an interleaved sub-grid of gx
// (epilogue mask / channel sums address the same sub-grid).  gy [B,yh,yw,O] and gx [B,H,W,C] channels-last, w [O,C,kh,kw] plain.
int agb_tc_conv_dgrad_strided(agb_ctx* ctx, int mode, const float* gy, const float* w, float* gx, int B, int O, int yh, int yw, int C, int H, int W,
                              int kh, int kw, int pad, int stride, int dil, const float* mask, float* csum) {
  const int s = stride, T = kh * kw;
  if (s < 2 || s > 4 || T > AGB_CONV_MAX_TAPS || mode == AGB_MATH_3XTF32) return AGB_ERR_UNSUPPORTED;
  if (!agb_tc_conv_eligible(O, C, kh, kw, 1, (W + s - 1) / s) || (((uintptr_t)gy | (uintptr_t)gx) & 15) != 0) return AGB_ERR_UNSUPPORTED;
  // floor division helper for possibly negative numerators
  auto fdiv = [](int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); };
  std::vector<ConvPhase> phases;
  bool empty_phase = false;
  for (int py = 0; py < s && py < H; py++)
    for (int px = 0; px < s && px < W; px++) {
      ConvPhase ph; ph.ntaps = 0; ph.os = s; ph.oyo = py; ph.oxo = px; ph.YH = H; ph.YW = W;
      for (int i = 0; i < kh; i++) {
        if (((py + pad - i * dil) % s + s) % s != 0) continue;
        for (int j = 0; j < kw; j++) {
          if (((px + pad - j * dil) % s + s) % s != 0) continue;
          ph.dy[ph.ntaps] = (short)fdiv(py + pad - i * dil, s); ph.dx[ph.ntaps] = (short)fdiv(px + pad - j * dil, s);
          ph.wt[ph.ntaps] = (short)(T - 1 - (i * kw + j));            // the dgrad repack stores tap t at slot T-1-t
          ph.ntaps++;
        }
      }
      if (ph.ntaps == 0) empty_phase = true; else phases.push_back(ph);
    }
  float* wr = nullptr;
  AGB_TRY(agb_scratch(ctx, (size_t)T * O * C * sizeof(float), (void**)&wr));
  {
    int64_t n = (int64_t)O * C * T;
    repack_filter_kernel<<<agb_grid_for(n, 256, ctx->sm_count, 4), 256, 0, ctx->stream>>>(w, wr, nullptr, O, C, T, 1);      // wr[T-1-t][c][o]
    AGB_LAUNCHED(ctx);
  }
  if (empty_phase) AGB_TRY(agb_memset0(ctx, gx, (size_t)B * H * W * C * sizeof(float)));                              // filter smaller than the stride: untouched pixels are 0
  for (const ConvPhase& ph : phases) {
    const int uh = (H - ph.oyo + s - 1) / s, uw = (W - ph.oxo + s - 1) / s;      // size of this phase's sub-grid
    int r;
    if (C > 128) r = fprop_launch<256, false>(ctx, gy, wr, gx, B, O, yh, yw, C, uh, uw, kh, kw, 0, dil, nullptr, 0, mask, csum, 1, &ph);
    else if (C > 64) r = fprop_launch<128, false>(ctx, gy, wr, gx, B, O, yh, yw, C, uh, uw, kh, kw, 0, dil, nullptr, 0, mask, csum, 1, &ph);
    else r = fprop_launch<64, false>(ctx, gy, wr, gx, B, O, yh, yw, C, uh, uw, kh, kw, 0, dil, nullptr, 0, mask, csum, 1, &ph);
    if (r != AGB_OK) return r;
  }
  return AGB_OK;
}

int agb_tc_presplit(agb_ctx* ctx, const float* src, float* hi, float* lo, int64_t n);      // tc_gemm.cu
template <int TN, bool SPLIT, bool PAIR, int MT = 1, bool QPRE = false>
static int wgrad_launch(agb_ctx* ctx, const float* img, const float* g, float* gw, int B, int C, int H, int W, int O, int yh, int yw, int kh, int kw, int pad, int dil, int stride = 1) {
  using Pol = ConvWgradPol<TN, SPLIT, PAIR, MT, QPRE>;
  typename Pol::Params p;
  const float* g_lo = g;
  if (QPRE) {      // hi / lo planes of gy in scratch (this launcher uses scratch2 for its partials)
    const int64_t gn = (int64_t)B * yh * yw * O;
    float* planes = nullptr;
    AGB_TRY(agb_scratch(ctx, (size_t)gn * 2 * sizeof(float), (void**)&planes));
    AGB_TRY(agb_tc_presplit(ctx, g, planes, planes + gn, gn));
    g = planes; g_lo = planes + gn;
  }
  {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)C * 4, (uint64_t)W * C * 4, (uint64_t)H * W * C * 4};
    const uint32_t su = (uint32_t)stride;
    uint32_t box[4] = {32, 31 * su + 1, 1, 1}, es[4] = {1, su, 1, 1};
    AGB_TRY(agb_make_tmap(&p.tmX, img, 4, dims, str, box, true, stride > 1 ? es : nullptr));
  }
  p.stride = stride;
  AGB_TRY(make_cl_map(&p.tmG, g, B, O, yh, yw, 32, 32, 1, true));
  p.tmGl = p.tmG;
  if (QPRE) AGB_TRY(make_cl_map(&p.tmGl, g_lo, B, O, yh, yw, 32, 32, 1, true));
  static const int wide_env = [] { const char* e = getenv("AGB_WGRAD_WIDE"); return (e && e[0] == '0') ? 0 : 1; }();
  p.wide = (wide_env && C % 32 == 0 && O % 32 == 0 && !Pol::PAIR2) ? 1 : 0;
  p.tmX5 = p.tmX; p.tmG5 = p.tmG; p.tmG5l = p.tmGl;
  if (p.wide) {
    const uint32_t su = (uint32_t)stride;
    {   // x as {32 c, W, C / 32, H, B}
      uint64_t dims[5] = {32, (uint64_t)W, (uint64_t)(C / 32), (uint64_t)H, (uint64_t)B};
      uint64_t str[4] = {(uint64_t)C * 4, 128, (uint64_t)W * C * 4, (uint64_t)H * W * C * 4};
      uint32_t box[5] = {32, 31 * su + 1, (uint32_t)(PAIR ? 2 : 4), 1, 1}, es[5] = {1, su, 1, 1, 1};
      AGB_TRY(agb_make_tmap(&p.tmX5, img, 5, dims, str, box, true, stride > 1 ? es : nullptr));
    }
    {   // gy as {32 o, yw, O / 32, yh, B}
      uint64_t dims[5] = {32, (uint64_t)yw, (uint64_t)(O / 32), (uint64_t)yh, (uint64_t)B};
      uint64_t str[4] = {(uint64_t)O * 4, 128, (uint64_t)yw * O * 4, (uint64_t)yh * yw * O * 4};
      uint32_t box[5] = {32, 32, (uint32_t)(TN / 32), 1, 1};
      AGB_TRY(agb_make_tmap(&p.tmG5, g, 5, dims, str, box, true));
      if (QPRE) AGB_TRY(agb_make_tmap(&p.tmG5l, g_lo, 5, dims, str, box, true));
    }
  }
  const int T = kh * kw;
  p.gw = gw; p.C = C; p.O = O; p.T = T; p.kw = kw; p.pad = pad; p.dil = dil; p.yh = yh; p.xblocks = (yw + 31) / 32;
  int64_t kb_total = (int64_t)B * yh * p.xblocks;
  if (kb_total > 2147483647ll) return AGB_ERR_UNSUPPORTED;
  p.kb_total = (int)kb_total; p.mnc = agb_mn_cfg();
  const int units = ((C + 127) / 128) * T;
  const int gx = PAIR ? (T + 1) / 2 : (MT == 2 ? (units + 1) / 2 : units), gy_ = (O + TN - 1) / TN;
  static const int waves_x2 = [] { const char* e = getenv("AGB_WGRAD_WAVES_X2"); int v = e ? atoi(e) : 2; return v < 1 ? 1 : v; }();      // tuning knob: split-K tiles per CTA slot, in halves (measured, 256->256 @32x32: one tile per slot 0.40 ms, two 0.47, one half 0.70)
  int64_t want = (int64_t)waves_x2 * ctx->sm_count * Pol::OCC / (2 * (int64_t)gx * gy_); if (want < 1) want = 1;
  int64_t per = (kb_total + want - 1) / want; if (per < 16) per = 16; if (per > kb_total) per = kb_total;
  p.kb_per_split = (int)per;
  int splits = (int)((kb_total + per - 1) / per);
  if (splits > 65535) return AGB_ERR_UNSUPPORTED;
  const int64_t n = (int64_t)O * C * T;
  float* part = nullptr;
  if (ctx->deterministic && splits > 1) AGB_TRY(agb_scratch2(ctx, (size_t)splits * n * sizeof(float), (void**)&part));
  if (part) { p.gw = part; p.part_stride = n; } else { p.part_stride = 0; AGB_TRY(agb_memset0(ctx, gw, (size_t)n * sizeof(float))); }
  dim3 grid((unsigned)gx, (unsigned)gy_, (unsigned)splits);
  AGB_TRY(tc_tile_launch<Pol>(ctx, p, grid));
  if (part) return agb_reduce_partials_wgrad(ctx, part, gw, splits, O, C, T);
  return AGB_OK;
}

// wgrad on channels-last buffers: img [B,H,W,C] (the im2col'd operand), g [B,yh,yw,O] -> gw [O,C,kh,kw] (plain)
int agb_tc_conv_wgrad(agb_ctx* ctx, int mode, const float* img, const float* g, float* gw, int B, int C, int H, int W, int O, int kh, int kw,
                      int pad, int stride, int dil) {
  const int yh = (H + 2 * pad - (dil * (kh - 1) + 1)) / stride + 1, yw = (W + 2 * pad - (dil * (kw - 1) + 1)) / stride + 1;
  if (yh < 1 || yw < 1 || !agb_tc_conv_fprop_eligible(C, O, kh, kw, stride, yw)) return AGB_ERR_UNSUPPORTED;
  if ((((uintptr_t)img | (uintptr_t)g) & 15) != 0) return AGB_ERR_UNSUPPORTED;
  const bool split = mode == AGB_MATH_3XTF32;
  if (!split && stride == 1) {       // narrow layers: all taps per CTA from one haloed window (tc_conv_wgrad_taps.cu)
    int r = agb_tc_conv_wgrad_taps(ctx, img, g, gw, B, C, H, W, O, yh, yw, kh, kw, pad, dil);
    if (r != AGB_ERR_UNSUPPORTED) return r;
  }
  const bool pair = C <= 64;
#define WG(TN_, SP_) (pair ? wgrad_launch<TN_, SP_, true>(ctx, img, g, gw, B, C, H, W, O, yh, yw, kh, kw, pad, dil, stride) \
                           : wgrad_launch<TN_, SP_, false>(ctx, img, g, gw, B, C, H, W, O, yh, yw, kh, kw, pad, dil, stride))
  if (split) {
    // pre-split gy where its planes (12 B/element once) cost well under the shared-memory traffic they save (~17 % of a kernel whose time grows with C * taps)
    static const int qpre_env = [] { const char* e = getenv("AGB_WGRAD_QPRE"); return (e && e[0] == '0') ? 0 : 1; }();
    const int64_t gn = (int64_t)B * yh * yw * O;
    if (qpre_env && !pair && C * kh * kw >= 1152 && gn % 4 == 0 && gn <= (1ll << 30)) {
      if (O > 64) return wgrad_launch<128, true, false, 1, true>(ctx, img, g, gw, B, C, H, W, O, yh, yw, kh, kw, pad, dil, stride);
      return wgrad_launch<64, true, false, 1, true>(ctx, img, g, gw, B, C, H, W, O, yh, yw, kh, kw, pad, dil, stride);
    }
    if (O > 64) return WG(128, true); return WG(64, true);
  }
  static int m2 = -1;
  if (m2 < 0) { const char* e = getenv("AGB_WGRAD_M2"); m2 = (e && e[0] == '0') ? 0 : 1; }
  // measured (B200, B = 256): C256/O256 0.66 -> 0.54 ms, C128/O256 0.37 -> 0.32 ms; with TN = 128 the one-M-tile kernel at two CTAs
  // per SM is faster (0.57 vs 0.78 ms), so only the 256-wide tiles pair up
  // measured (B200, B = 256): CTA pairs lose to two M-tiles per CTA here (C256/O256 0.63 vs 0.55 ms, C128/O256 0.49 vs 0.32 ms): kept as an opt-in experiment
  static const int pair2 = [] { const char* e = getenv("AGB_WGRAD_PAIR"); return (e && e[0] == '1') ? 1 : 0; }();
  if (pair2 && !pair && O > 128) return wgrad_launch<256, false, false, 1>(ctx, img, g, gw, B, C, H, W, O, yh, yw, kh, kw, pad, dil, stride);      // CTA pairs (ConvWgradPol::PAIR2)
  if (m2 && !pair && O > 128) return wgrad_launch<256, false, false, 2>(ctx, img, g, gw, B, C, H, W, O, yh, yw, kh, kw, pad, dil, stride);
  if (O > 128) return WG(256, false);
  if (O > 64) return WG(128, false);
  return WG(64, false);
#undef WG
}
