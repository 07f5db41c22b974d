// conv_small_c.cu — direct CUDA-core kernels for convolutions with very few input channels (the first layer of a CNN:
// C = 1..4, K = C*kh*kw <= 36).  A K that small is bandwidth-bound (SURVEY §7 "first layer ... needs a direct kernel"): the
// output write dominates, so the tensor-core path (32-channel k-blocks) would waste > 90 % of its MMAs.
//   fprop: x NCHW -> y in either layout; thread = output pixel, keeps its K taps in registers, loops over O with the filter in
//          shared memory (warp-wide broadcast reads); channels-last output = 16-byte stores of consecutive channels.
//   wgrad: gw[o,k] = sum_pix gy[pix,o] * patch[pix,k]; a CTA streams 64-pixel slabs of gy (channels-last, fully coalesced) and
//          the matching patches through shared memory, 256 threads own the O x K outputs, one red.global.add per output at the end.
// Reference semantics: conv2d.rs:115-211 (fprop), :631-734 (filter grad), im2col index math conv_ops/mod.rs:73-124.
#include "common.cuh"

#define SC_MAXK 36
struct SmallGeom { int B, C, H, W, O, kh, kw, yh, yw, pad, stride, dil; int64_t ys[4]; /* y strides (b,o,h,w) */ };

// ---- fprop: per CTA a [64 pixels] x [K] patch matrix times the [K] x [O] filter, both staged in shared memory, each thread
//      a 4 (pixel) x 4 (o) register tile: two 128-bit shared loads per 16 FMAs.
#define SCF_PIX 64
__global__ void __launch_bounds__(256) small_c_fprop_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y, SmallGeom g, int K,
                                                            const float* __restrict__ bias, int relu) {
  extern __shared__ __align__(16) float sm[];
  float* ws = sm;                                     // [K][O]   (transposed filter)
  float* vs = sm + K * g.O;                           // [K][SCF_PIX]
  __shared__ int64_t s_xoff[SCF_PIX], s_yoff[SCF_PIX];
  __shared__ int s_oy[SCF_PIX], s_ox[SCF_PIX];
  __shared__ int s_tc[SC_MAXK], s_ti[SC_MAXK], s_tj[SC_MAXK];
  const int kk = g.kh * g.kw;
  for (int i = threadIdx.x; i < g.O * K; i += blockDim.x) { int k = i / g.O, o = i - k * g.O; ws[i] = __ldg(w + o * K + k); }     // o fastest: conflict-free stores
  if (threadIdx.x < K) { int k = threadIdx.x; int c = k / kk, t = k - c * kk; s_tc[k] = c; s_ti[k] = (t / g.kw) * g.dil - g.pad; s_tj[k] = (t % g.kw) * g.dil - g.pad; }
  const int64_t P = (int64_t)g.yh * g.yw, total = (int64_t)g.B * P;
  const int64_t base = (int64_t)blockIdx.x * SCF_PIX;
  if (threadIdx.x < SCF_PIX) {
    int64_t pix = base + threadIdx.x; if (pix >= total) pix = total - 1;
    const int b = (int)(pix / P); const int r = (int)(pix - (int64_t)b * P); const int oy = r / g.yw, ox = r - oy * g.yw;
    s_xoff[threadIdx.x] = (int64_t)b * g.C * g.H * g.W; s_yoff[threadIdx.x] = b * g.ys[0] + oy * g.ys[2] + ox * g.ys[3];
    s_oy[threadIdx.x] = oy * g.stride; s_ox[threadIdx.x] = ox * g.stride;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * SCF_PIX; i += blockDim.x) {
    const int k = i / SCF_PIX, pi = i - k * SCF_PIX;
    const int iy = s_oy[pi] + s_ti[k], ix = s_ox[pi] + s_tj[k];
    vs[i] = ((unsigned)iy < (unsigned)g.H && (unsigned)ix < (unsigned)g.W) ? __ldg(x + s_xoff[pi] + ((int64_t)s_tc[k] * g.H + iy) * g.W + ix) : 0.0f;
  }
  __syncthreads();
  const int to = threadIdx.x & 15, tp = threadIdx.x >> 4;          // 16 o-quads x 16 pixel-quads
  const bool cl = g.ys[1] == 1;
  for (int o0 = 0; o0 < g.O; o0 += 64) {
    const int o = o0 + 4 * to;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int q = 0; q < 4; q++) acc[a][q] = 0.f;
    if (o < g.O) {
      for (int k = 0; k < K; k++) {
        const float4 pv = *(const float4*)(vs + k * SCF_PIX + 4 * tp);
        const float4 wv = *(const float4*)(ws + k * g.O + o);
        const float p_[4] = {pv.x, pv.y, pv.z, pv.w}, w_[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
          for (int q = 0; q < 4; q++) acc[a][q] = fmaf(p_[a], w_[q], acc[a][q]);
      }
      if (bias != nullptr) {
        const float4 bv = __ldg((const float4*)(bias + o));
#pragma unroll
        for (int a = 0; a < 4; a++) { acc[a][0] += bv.x; acc[a][1] += bv.y; acc[a][2] += bv.z; acc[a][3] += bv.w; }
      }
      if (relu) {
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
          for (int q = 0; q < 4; q++) acc[a][q] = fmaxf(acc[a][q], 0.0f);
      }
#pragma unroll
      for (int a = 0; a < 4; a++) {
        const int pi = 4 * tp + a;
        if (base + pi < total) {
          float* yp = y + s_yoff[pi];
          if (cl) *(float4*)(yp + o) = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
          else { for (int q = 0; q < 4; q++) yp[(o + q) * g.ys[1]] = acc[a][q]; }
        }
      }
    }
  }
}

// gy strides gs (b,o,h,w) in either layout; K <= 32, O <= 256 (multiple of 4).
// Register tiling: thread (to, tk) owns a 4 (o) x 4 (k) block of gw for every 64-channel slab of O, so one slab pixel costs
// two 128-bit shared-memory reads per 16 FMAs (a thread-per-output mapping would be shared-memory bound).
#define SCW_PIX 128
__global__ void __launch_bounds__(128) small_c_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ gw, SmallGeom g,
                                                            int K, int K4 /* K rounded up to a multiple of 4 */, int64_t chunk) {
  extern __shared__ __align__(16) float sm[];
  float* gs = sm;                                     // [SCW_PIX][O]
  float* ps = sm + SCW_PIX * g.O;                     // [SCW_PIX][K4]
  __shared__ int64_t s_xoff[SCW_PIX], s_goff[SCW_PIX];
  __shared__ int s_oy[SCW_PIX], s_ox[SCW_PIX];
  __shared__ int s_tc[SC_MAXK], s_ti[SC_MAXK], s_tj[SC_MAXK];
  const int64_t P = (int64_t)g.yh * g.yw, total = (int64_t)g.B * P;
  const int64_t p0 = blockIdx.x * chunk, p1 = min(p0 + chunk, total);
  const int kk = g.kh * g.kw;
  if (threadIdx.x < K4) { int k = threadIdx.x; int c = k / kk, t = k - c * kk; s_tc[k] = c; s_ti[k] = (t / g.kw) * g.dil - g.pad; s_tj[k] = (t % g.kw) * g.dil - g.pad; }
  const int to = threadIdx.x & 15, tk = threadIdx.x >> 4;          // 16 o-quads x 8 k-quads per 64-channel slab
  const int nslab = (g.O + 63) / 64;
  const bool active = 4 * tk < K4;
  float acc[4][4][4];                                  // [slab][o in quad][k in quad]
#pragma unroll
  for (int s = 0; s < 4; s++)
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
      for (int e = 0; e < 4; e++) acc[s][q][e] = 0.f;
  const bool gy_cl = g.ys[1] == 1;
  for (int64_t base = p0; base < p1; base += SCW_PIX) {
    const int np = (int)min((int64_t)SCW_PIX, p1 - base);
    if (threadIdx.x < np) {
      const int64_t pix = base + threadIdx.x;
      const int b = (int)(pix / P); const int r = (int)(pix - (int64_t)b * P); const int oy = r / g.yw, ox = r - oy * g.yw;
      s_xoff[threadIdx.x] = (int64_t)b * g.C * g.H * g.W; s_goff[threadIdx.x] = b * g.ys[0] + oy * g.ys[2] + ox * g.ys[3];
      s_oy[threadIdx.x] = oy * g.stride; s_ox[threadIdx.x] = ox * g.stride;
    }
    __syncthreads();
    if (gy_cl && (g.O & 3) == 0) {                     // channels-last gy: 128-bit coalesced copies
      const int o4n = g.O >> 2;
      for (int i = threadIdx.x; i < np * o4n; i += blockDim.x) {
        const int pi = i / o4n, o4 = i - pi * o4n;
        *(float4*)(gs + pi * g.O + 4 * o4) = __ldg((const float4*)(gy + s_goff[pi] + 4 * o4));
      }
    } else {
      for (int i = threadIdx.x; i < np * g.O; i += blockDim.x) {
        int pi, o;
        if (gy_cl) { pi = i / g.O; o = i - pi * g.O; } else { o = i / np; pi = i - o * np; }
        gs[pi * g.O + o] = __ldg(gy + s_goff[pi] + o * g.ys[1]);
      }
    }
    for (int i = threadIdx.x; i < np * K4; i += blockDim.x) {
      const int k = i / np, pi = i - k * np;           // consecutive threads -> consecutive pixels (coalesced x reads)
      float v = 0.0f;
      if (k < K) {
        const int iy = s_oy[pi] + s_ti[k], ix = s_ox[pi] + s_tj[k];
        if ((unsigned)iy < (unsigned)g.H && (unsigned)ix < (unsigned)g.W) v = __ldg(x + s_xoff[pi] + ((int64_t)s_tc[k] * g.H + iy) * g.W + ix);
      }
      ps[pi * K4 + k] = v;
    }
    __syncthreads();
    if (active) {
      for (int pi = 0; pi < np; pi++) {
        const float4 pv = *(const float4*)(ps + pi * K4 + 4 * tk);
        const float p_[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
        for (int s = 0; s < 4; s++) {
          if (s < nslab) {
            const float4 gv = *(const float4*)(gs + pi * g.O + s * 64 + 4 * to);
            const float g_[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
            for (int q = 0; q < 4; q++)
#pragma unroll
              for (int e = 0; e < 4; e++) acc[s][q][e] = fmaf(g_[q], p_[e], acc[s][q][e]);
          }
        }
      }
    }
    __syncthreads();
  }
  if (active) {
#pragma unroll
    for (int s = 0; s < 4; s++)
#pragma unroll
      for (int q = 0; q < 4; q++)
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const int o = s * 64 + 4 * to + q, k = 4 * tk + e;
          if (s < nslab && o < g.O && k < K) atomicAdd(gw + o * K + k, acc[s][q][e]);
        }
  }
}

// =====================================================================================================================
// Warp-level tensor-core versions (mma.sync m16n8k8 tf32) for the TF32 / 3xTF32 math modes.  K = C*kh*kw <= 32 is far too
// short for a tcgen05 k-pipeline (one 32-wide k-block per tile, the TMEM/TMA set-up would dominate), but the SIMT kernels
// above are issue- and latency-bound (shared-memory staging with block-wide barriers).  Here every warp streams its own
// pixels straight from global memory into MMA fragments — no shared memory, no barriers — so the kernels run at the HBM rate
// of the one big operand (y for fprop, gy for wgrad).
// =====================================================================================================================
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t tf32_hi(uint32_t v) { return v & 0xffffe000u; }
__device__ __forceinline__ uint32_t tf32_lo(uint32_t v) { return __float_as_uint(__uint_as_float(v) - __uint_as_float(v & 0xffffe000u)); }
// d += a*b with the 3-product hi/lo split when SPLIT (same scheme as the tcgen05 tile engine: lo*hi + hi*lo + hi*hi)
template <bool SPLIT>
__device__ __forceinline__ void mma_x(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  if (SPLIT) {
    uint32_t al[4], ah[4], bl[2], bh[2];
#pragma unroll
    for (int i = 0; i < 4; i++) { ah[i] = tf32_hi(a[i]); al[i] = tf32_lo(a[i]); }
#pragma unroll
    for (int i = 0; i < 2; i++) { bh[i] = tf32_hi(b[i]); bl[i] = tf32_lo(b[i]); }
    mma_tf32_16x8x8(d, al, bh); mma_tf32_16x8x8(d, ah, bl); mma_tf32_16x8x8(d, ah, bh);
  } else mma_tf32_16x8x8(d, a, b);
}
__device__ __forceinline__ uint32_t ldg_stream_u32(const float* p) {
  uint32_t v; asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v;
}

// running (b, oy, ox) decode of a flat output-pixel index, advanced by a fixed step without divisions
struct PixCursor {
  int b, oy, ox;
  __device__ __forceinline__ void init(int64_t pix, const SmallGeom& g) {
    const int64_t P = (int64_t)g.yh * g.yw;
    b = (int)(pix / P); const int r = (int)(pix - (int64_t)b * P); oy = r / g.yw; ox = r - oy * g.yw;
  }
  __device__ __forceinline__ void advance(int step, const SmallGeom& g) {
    ox += step;
    while (ox >= g.yw) { ox -= g.yw; if (++oy >= g.yh) { oy = 0; b++; } }
  }
};

// ---- wgrad: gw[o,k] = sum_pix gy[pix,o] * patch[pix,k] as D[16 o x 8 k] += A[16 o x 8 pix] * B[8 pix x 8 k].
//      A warp owns one 64-channel slab of O (4 m-tiles) x all NT k-tiles = 16*NT accumulators/thread and walks a contiguous run
//      of 8-pixel groups.  A-fragments: lanes g=lane/4 -> o, t=lane%4 -> pixel: with channels-last gy one warp load covers
//      4 pixels x 32 B, every sector fully used.  B-fragments gather x through L1 (each x element is reused by kh*kw taps).
template <int NT, bool SPLIT>
__global__ void __launch_bounds__(128, 3) small_c_wgrad_mma_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ gw,
                                                                   SmallGeom g, int K, int nslab, int64_t groups_per_warp, float* __restrict__ part) {
  __shared__ float red[4][64][NT * 8 + 1];
  const int lane = threadIdx.x & 31, gq = lane >> 2, t = lane & 3, wib = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 4 * 64 * (NT * 8 + 1); i += blockDim.x) (&red[0][0][0])[i] = 0.0f;
  __syncthreads();
  const int64_t wg = (int64_t)blockIdx.x * 4 + wib;
  const int slab = (int)(wg % nslab); const int64_t wp = wg / nslab;
  const int64_t total = (int64_t)g.B * g.yh * g.yw, G = (total + 7) >> 3;
  const int64_t g0 = wp * groups_per_warp, g1 = min(g0 + groups_per_warp, G);
  const int kk = g.kh * g.kw; const int64_t HW = (int64_t)g.H * g.W;
  // per-tap (di, dj) packed as two int16 and the flat offset c*H*W + di*W + dj (host checks C*H*W < 2^31); a tap past K gets
  // di = 0x4000, which fails every bounds check
  int tap_d[NT], tap_off[NT];
#pragma unroll
  for (int j = 0; j < NT; j++) {
    const int k = j * 8 + gq;
    const int c = k / kk, r = k - c * kk; int di = (r / g.kw) * g.dil - g.pad; const int dj = (r % g.kw) * g.dil - g.pad;
    tap_off[j] = (int)(c * HW) + di * g.W + dj;
    if (k >= K) di = 0x4000;
    tap_d[j] = (di << 16) | (dj & 0xffff);
  }
  float acc[4][NT][4];
#pragma unroll
  for (int m = 0; m < 4; m++)
#pragma unroll
    for (int j = 0; j < NT; j++)
#pragma unroll
      for (int e = 0; e < 4; e++) acc[m][j][e] = 0.0f;
  const int o_base = slab * 64 + gq;
  if (g0 < g1) {
    PixCursor ca, cb; ca.init(g0 * 8 + t, g); cb.init(min(g0 * 8 + t + 4, total - 1), g);
    // when the first group is also the last (ragged) one, cb may have been clamped: validity is tracked by the flat index below
    uint32_t a[4][4], bf[NT][2];
    auto load = [&](int64_t grp, uint32_t (&A)[4][4], uint32_t (&Bf)[NT][2]) {
      const int64_t pa = grp * 8 + t, pb = pa + 4;
      const bool va = pa < total, vb = pb < total;
      const int64_t ga = ca.b * g.ys[0] + ca.oy * g.ys[2] + ca.ox * g.ys[3], gb = cb.b * g.ys[0] + cb.oy * g.ys[2] + cb.ox * g.ys[3];
#pragma unroll
      for (int m = 0; m < 4; m++) {
        const int o = o_base + m * 16;
        A[m][0] = (va && o < g.O) ? ldg_stream_u32(gy + ga + o * g.ys[1]) : 0u;
        A[m][1] = (va && o + 8 < g.O) ? ldg_stream_u32(gy + ga + (o + 8) * g.ys[1]) : 0u;
        A[m][2] = (vb && o < g.O) ? ldg_stream_u32(gy + gb + o * g.ys[1]) : 0u;
        A[m][3] = (vb && o + 8 < g.O) ? ldg_stream_u32(gy + gb + (o + 8) * g.ys[1]) : 0u;
      }
      const float* xa = x + (int64_t)ca.b * g.C * HW; const float* xb = x + (int64_t)cb.b * g.C * HW;
#pragma unroll
      const int ya0 = ca.oy * g.stride, xa0 = ca.ox * g.stride, yb0 = cb.oy * g.stride, xb0 = cb.ox * g.stride;
      const float* xpa = xa + ya0 * g.W + xa0; const float* xpb = xb + yb0 * g.W + xb0;
#pragma unroll
      for (int j = 0; j < NT; j++) {
        const int di = tap_d[j] >> 16, dj = (int)(short)(tap_d[j] & 0xffff);
        Bf[j][0] = (va && (unsigned)(ya0 + di) < (unsigned)g.H && (unsigned)(xa0 + dj) < (unsigned)g.W) ? __float_as_uint(__ldg(xpa + tap_off[j])) : 0u;
        Bf[j][1] = (vb && (unsigned)(yb0 + di) < (unsigned)g.H && (unsigned)(xb0 + dj) < (unsigned)g.W) ? __float_as_uint(__ldg(xpb + tap_off[j])) : 0u;
      }
    };
    // cb must be the true decode of pa+4 whenever that pixel exists
    if (g0 * 8 + t + 4 < total) cb.init(g0 * 8 + t + 4, g);
    load(g0, a, bf);
    for (int64_t grp = g0; grp < g1; grp++) {
      uint32_t an[4][4], bn[NT][2];
      const bool more = grp + 1 < g1;
      if (more) {
        ca.advance(8, g); cb.advance(8, g);
        if (ca.b >= g.B) { ca.b = g.B - 1; ca.oy = 0; ca.ox = 0; }          // ragged tail: keep addresses in range, loads are predicated off
        if (cb.b >= g.B) { cb.b = g.B - 1; cb.oy = 0; cb.ox = 0; }
        load(grp + 1, an, bn);
      }
#pragma unroll
      for (int m = 0; m < 4; m++)
#pragma unroll
        for (int j = 0; j < NT; j++) mma_x<SPLIT>(acc[m][j], a[m], bf[j]);
      if (more) {
#pragma unroll
        for (int m = 0; m < 4; m++)
#pragma unroll
          for (int e = 0; e < 4; e++) a[m][e] = an[m][e];
#pragma unroll
        for (int j = 0; j < NT; j++) { bf[j][0] = bn[j][0]; bf[j][1] = bn[j][1]; }
      }
    }
  }
  if (part != nullptr) {         // deterministic mode: every warp stores its fragments as partial number wp; agb_reduce_partials2 adds them in order
    float* d = part + wp * ((int64_t)g.O * K);
#pragma unroll
    for (int m = 0; m < 4; m++)
#pragma unroll
      for (int j = 0; j < NT; j++)
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const int o = slab * 64 + m * 16 + gq + (e >> 1) * 8, k = j * 8 + 2 * t + (e & 1);
          if (o < g.O && k < K) d[(int64_t)o * K + k] = acc[m][j][e];
        }
    return;
  }
  // CTA-level reduction in shared memory, then one red.global.add per output element per CTA
  const int rs = slab & 3;       // nslab <= 4
#pragma unroll
  for (int m = 0; m < 4; m++)
#pragma unroll
    for (int j = 0; j < NT; j++) {
      atomicAdd(&red[rs][m * 16 + gq][j * 8 + 2 * t], acc[m][j][0]);
      atomicAdd(&red[rs][m * 16 + gq][j * 8 + 2 * t + 1], acc[m][j][1]);
      atomicAdd(&red[rs][m * 16 + gq + 8][j * 8 + 2 * t], acc[m][j][2]);
      atomicAdd(&red[rs][m * 16 + gq + 8][j * 8 + 2 * t + 1], acc[m][j][3]);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < nslab * 64 * NT * 8; i += blockDim.x) {
    const int k = i % (NT * 8), o64 = (i / (NT * 8)) % 64, s = i / (NT * 8 * 64);
    const int o = s * 64 + o64;
    const float v = red[s][o64][k];
    if (o < g.O && k < K && v != 0.0f) atomicAdd(gw + (int64_t)o * K + k, v);
  }
}

// ---- fprop: y[pix,o] = sum_k patch[pix,k] * w[o,k] as D[16 pix x 8 o] += A[16 pix x 8 k] * B[8 k x 8 o].
//      A warp keeps its 32-channel slab of the filter in registers (KT k-steps x 4 n-tiles x 2) and walks a contiguous run of
//      16-pixel groups; output: lanes t hold channel pairs (2t, 2t+1) -> 8-byte stores, 32 B per pixel per n-tile.
template <int KT, bool SPLIT>
__global__ void __launch_bounds__(128, 3) small_c_fprop_mma_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y,
                                                                   SmallGeom g, int K, int nslab, int64_t groups_per_warp,
                                                                   const float* __restrict__ bias, int relu) {
  const int lane = threadIdx.x & 31, gq = lane >> 2, t = lane & 3, wib = threadIdx.x >> 5;
  const int64_t wg = (int64_t)blockIdx.x * 4 + wib;
  const int slab = (int)(wg % nslab); const int64_t wp = wg / nslab;
  const int64_t total = (int64_t)g.B * g.yh * g.yw, G = (total + 15) >> 4;
  const int64_t g0 = wp * groups_per_warp, g1 = min(g0 + groups_per_warp, G);
  if (g0 >= g1) return;
  const int kk = g.kh * g.kw; const int64_t HW = (int64_t)g.H * g.W;
  // this thread's taps: k = kt*8 + t and kt*8 + t + 4
  int tap_d[KT][2], tap_off[KT][2];
#pragma unroll
  for (int kt = 0; kt < KT; kt++)
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int k = kt * 8 + t + 4 * h;
      const int c = k / kk, r = k - c * kk; int di = (r / g.kw) * g.dil - g.pad; const int dj = (r % g.kw) * g.dil - g.pad;
      tap_off[kt][h] = (int)(c * HW) + di * g.W + dj;
      if (k >= K) di = 0x4000;
      tap_d[kt][h] = (di << 16) | (dj & 0xffff);
    }
  // filter fragments: b0 = w[o = nt*8+gq][k = kt*8+t], b1 = w[o][k+4]
  uint32_t wf[KT][4][2];
  const int o_w = slab * 32 + gq;
#pragma unroll
  for (int kt = 0; kt < KT; kt++)
#pragma unroll
    for (int nt = 0; nt < 4; nt++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int o = o_w + nt * 8, k = kt * 8 + t + 4 * h;
        wf[kt][nt][h] = (o < g.O && k < K) ? __float_as_uint(__ldg(w + (int64_t)o * K + k)) : 0u;
      }
  float bv[4][2];
#pragma unroll
  for (int nt = 0; nt < 4; nt++) {
    const int o = slab * 32 + nt * 8 + 2 * t;
    bv[nt][0] = (bias != nullptr && o < g.O) ? __ldg(bias + o) : 0.0f;
    bv[nt][1] = (bias != nullptr && o + 1 < g.O) ? __ldg(bias + o + 1) : 0.0f;
  }
  const bool pair_store = g.ys[1] == 1 && (g.ys[0] & 1) == 0 && (g.ys[2] & 1) == 0 && (g.ys[3] & 1) == 0 && ((((uintptr_t)y) & 7) == 0);
  PixCursor ca, cb; ca.init(min(g0 * 16 + gq, total - 1), g); cb.init(min(g0 * 16 + gq + 8, total - 1), g);
  uint32_t a[KT][4];
  auto load = [&](int64_t grp, uint32_t (&A)[KT][4]) {
    const int64_t pa = grp * 16 + gq, pb = pa + 8;
    const bool va = pa < total, vb = pb < total;
    const int ya0 = ca.oy * g.stride, xa0 = ca.ox * g.stride, yb0 = cb.oy * g.stride, xb0 = cb.ox * g.stride;
    const float* xpa = x + (int64_t)ca.b * g.C * HW + ya0 * g.W + xa0; const float* xpb = x + (int64_t)cb.b * g.C * HW + yb0 * g.W + xb0;
#pragma unroll
    for (int kt = 0; kt < KT; kt++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int di = tap_d[kt][h] >> 16, dj = (int)(short)(tap_d[kt][h] & 0xffff);
        A[kt][2 * h]     = (va && (unsigned)(ya0 + di) < (unsigned)g.H && (unsigned)(xa0 + dj) < (unsigned)g.W) ? __float_as_uint(__ldg(xpa + tap_off[kt][h])) : 0u;
        A[kt][2 * h + 1] = (vb && (unsigned)(yb0 + di) < (unsigned)g.H && (unsigned)(xb0 + dj) < (unsigned)g.W) ? __float_as_uint(__ldg(xpb + tap_off[kt][h])) : 0u;
      }
  };
  load(g0, a);
  for (int64_t grp = g0; grp < g1; grp++) {
    const int64_t pa = grp * 16 + gq, pb = pa + 8;
    const int64_t ya = ca.b * g.ys[0] + ca.oy * g.ys[2] + ca.ox * g.ys[3], yb = cb.b * g.ys[0] + cb.oy * g.ys[2] + cb.ox * g.ys[3];
    uint32_t an[KT][4];
    const bool more = grp + 1 < g1;
    if (more) {
      ca.advance(16, g); cb.advance(16, g);
      if (ca.b >= g.B) { ca.b = g.B - 1; ca.oy = 0; ca.ox = 0; }
      if (cb.b >= g.B) { cb.b = g.B - 1; cb.oy = 0; cb.ox = 0; }
      load(grp + 1, an);
    }
    float acc[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; nt++) { acc[nt][0] = bv[nt][0]; acc[nt][1] = bv[nt][1]; acc[nt][2] = bv[nt][0]; acc[nt][3] = bv[nt][1]; }
#pragma unroll
    for (int kt = 0; kt < KT; kt++)
#pragma unroll
      for (int nt = 0; nt < 4; nt++) mma_x<SPLIT>(acc[nt], a[kt], wf[kt][nt]);
#pragma unroll
    for (int nt = 0; nt < 4; nt++) {
      const int o = slab * 32 + nt * 8 + 2 * t;
      if (relu) {
#pragma unroll
        for (int e = 0; e < 4; e++) acc[nt][e] = fmaxf(acc[nt][e], 0.0f);
      }
      if (o + 1 < g.O && pair_store) {
        if (pa < total) *(float2*)(y + ya + o) = make_float2(acc[nt][0], acc[nt][1]);
        if (pb < total) *(float2*)(y + yb + o) = make_float2(acc[nt][2], acc[nt][3]);
      } else {
        if (o < g.O) { if (pa < total) y[ya + o * g.ys[1]] = acc[nt][0]; if (pb < total) y[yb + o * g.ys[1]] = acc[nt][2]; }
        if (o + 1 < g.O) { if (pa < total) y[ya + (o + 1) * g.ys[1]] = acc[nt][1]; if (pb < total) y[yb + (o + 1) * g.ys[1]] = acc[nt][3]; }
      }
    }
    if (more) {
#pragma unroll
      for (int kt = 0; kt < KT; kt++)
#pragma unroll
        for (int e = 0; e < 4; e++) a[kt][e] = an[kt][e];
    }
  }
}

static void fill_geom(SmallGeom& g, int B, int C, int H, int W, int O, int kh, int kw, int yh, int yw, int pad, int stride, int dil, const agb_tensor* y) {
  g.B = B; g.C = C; g.H = H; g.W = W; g.O = O; g.kh = kh; g.kw = kw; g.yh = yh; g.yw = yw; g.pad = pad; g.stride = stride; g.dil = dil;
  for (int i = 0; i < 4; i++) g.ys[i] = y->stride[i];
}

bool agb_small_c_eligible(int C, int O, int kh, int kw) { return C <= 4 && C * kh * kw <= SC_MAXK && O <= 256 && O * C * kh * kw <= 2048; }

// x must be NCHW-contiguous (a first-layer input); y may be NCHW or channels-last (strides taken from the descriptor)
int agb_small_c_fprop(agb_ctx* ctx, const float* x, const float* w, agb_tensor* y, int B, int C, int H, int W, int O, int kh, int kw, int yh, int yw,
                      int pad, int stride, int dil, const float* bias, int relu) {
  SmallGeom g; fill_geom(g, B, C, H, W, O, kh, kw, yh, yw, pad, stride, dil, y);
  const int K = C * kh * kw; const int64_t total = (int64_t)B * yh * yw;
  if (ctx->math_mode != AGB_MATH_FP32 && K <= 32 && O <= 256) {       // warp-MMA path
    const int nslab = (O + 31) / 32, KT = (K + 7) / 8; const bool split = ctx->math_mode == AGB_MATH_3XTF32;
    if ((int64_t)C * H * W >= (1ll << 31) - (1 << 20)) return AGB_ERR_UNSUPPORTED;
    const int64_t G = (total + 15) / 16;
    int64_t warps = 4ll * 12 * ctx->sm_count / nslab; if (warps < 1) warps = 1;      // ~4 waves of 3 CTAs x 4 warps per SM
    int64_t gpw = (G + warps - 1) / warps; if (gpw < 4) gpw = 4;
    warps = (G + gpw - 1) / gpw;
    const unsigned blocks = (unsigned)((warps * nslab + 3) / 4);
#define SCF_LAUNCH(KT_, SP_) small_c_fprop_mma_kernel<KT_, SP_><<<blocks, 128, 0, ctx->stream>>>(x, w, y->ptr, g, K, nslab, gpw, bias, relu)
#define SCF_SW(KT_) do { if (split) SCF_LAUNCH(KT_, true); else SCF_LAUNCH(KT_, false); } while (0)
    switch (KT) { case 1: SCF_SW(1); break; case 2: SCF_SW(2); break; case 3: SCF_SW(3); break; default: SCF_SW(4); break; }
#undef SCF_SW
#undef SCF_LAUNCH
    AGB_LAUNCHED(ctx);
    return AGB_OK;
  }
  if (O % 4 != 0 || K > SC_MAXK || (((uintptr_t)y->ptr) & 15) != 0) return AGB_ERR_UNSUPPORTED;
  const size_t smem = (size_t)K * (O + SCF_PIX) * sizeof(float);
  if (smem > 48 * 1024) return AGB_ERR_UNSUPPORTED;
  int64_t blocks = (total + SCF_PIX - 1) / SCF_PIX;
  if (blocks > 2147483647ll) return AGB_ERR_UNSUPPORTED;
  if (bias && (((uintptr_t)bias) & 15)) return AGB_ERR_UNSUPPORTED;
  small_c_fprop_kernel<<<(unsigned)blocks, 256, smem, ctx->stream>>>(x, w, y->ptr, g, K, bias, relu);
  AGB_LAUNCHED(ctx);
  return AGB_OK;
}

int agb_tc_conv_first_wgrad(agb_ctx* ctx, int mode, const float* x, const float* gy, float* gw, int B, int C, int H, int W, int O, int kh, int kw, int yh, int yw,
                            int pad, int stride, int dil);
// gy may be NCHW or channels-last; gw [O, C*kh*kw] contiguous
int agb_small_c_wgrad(agb_ctx* ctx, const float* x, const agb_tensor* gy, float* gw, int B, int C, int H, int W, int O, int kh, int kw, int yh, int yw,
                      int pad, int stride, int dil) {
  SmallGeom g; fill_geom(g, B, C, H, W, O, kh, kw, yh, yw, pad, stride, dil, gy);
  const int K = C * kh * kw; const int64_t total = (int64_t)B * yh * yw;
  if (ctx->math_mode == AGB_MATH_TF32 && gy->stride[1] == 1 && gy->stride[3] == O && gy->stride[2] == (int64_t)yw * O && gy->stride[0] == (int64_t)yh * yw * O) {
    int r = agb_tc_conv_first_wgrad(ctx, ctx->math_mode, x, gy->ptr, gw, B, C, H, W, O, kh, kw, yh, yw, pad, stride, dil);      // tcgen05, gy by TMA (tc_conv_first.cu)
    if (r != AGB_ERR_UNSUPPORTED) return r;
  }
  if (ctx->math_mode != AGB_MATH_FP32 && K <= 32 && O <= 256) {       // warp-MMA path
    const int nslab = (O + 63) / 64, NT = (K + 7) / 8; const bool split = ctx->math_mode == AGB_MATH_3XTF32;
    if ((int64_t)C * H * W >= (1ll << 31) - (1 << 20)) return AGB_ERR_UNSUPPORTED;
    const int64_t G = (total + 7) / 8;
    int64_t warps = 2ll * 12 * ctx->sm_count / nslab; if (warps < 1) warps = 1;      // 2 waves of 3 CTAs x 4 warps per SM
    int64_t gpw = (G + warps - 1) / warps; if (gpw < 8) gpw = 8;
    warps = (G + gpw - 1) / gpw;
    unsigned blocks = (unsigned)((warps * nslab + 3) / 4);
    while ((blocks * 4u) % (unsigned)nslab) blocks++;            // every partial (pixel range) has all its slabs
    const int64_t nwp = (int64_t)blocks * 4 / nslab;
    float* part = nullptr;
    if (ctx->deterministic) AGB_TRY(agb_scratch2(ctx, agb_reduce_partials2_floats(nwp, (int64_t)O * K) * sizeof(float), (void**)&part));
    else AGB_TRY(agb_memset0(ctx, gw, (size_t)O * K * sizeof(float)));
#define SCW_LAUNCH(NT_, SP_) small_c_wgrad_mma_kernel<NT_, SP_><<<blocks, 128, 0, ctx->stream>>>(x, gy->ptr, gw, g, K, nslab, gpw, part)
#define SCW_SW(NT_) do { if (split) SCW_LAUNCH(NT_, true); else SCW_LAUNCH(NT_, false); } while (0)
    switch (NT) { case 1: SCW_SW(1); break; case 2: SCW_SW(2); break; case 3: SCW_SW(3); break; default: SCW_SW(4); break; }
#undef SCW_SW
#undef SCW_LAUNCH
    AGB_LAUNCHED(ctx);
    if (part) return agb_reduce_partials2(ctx, part, nwp, (int64_t)O * K, gw, 0);
    return AGB_OK;
  }
  if (O % 4 != 0 || O > 256 || K > 32) return AGB_ERR_UNSUPPORTED;       // 16 k-pairs per thread row
  const int K2 = (K + 3) & ~3;
  AGB_TRY(agb_memset0(ctx, gw, (size_t)O * K * sizeof(float)));
  int64_t blocks = 8ll * ctx->sm_count; int64_t chunk = (total + blocks - 1) / blocks; chunk = (chunk + SCW_PIX - 1) / SCW_PIX * SCW_PIX; if (chunk < SCW_PIX) chunk = SCW_PIX;
  blocks = (total + chunk - 1) / chunk;
  const size_t smem = (size_t)SCW_PIX * (O + K2) * sizeof(float);
  static bool attr = false;
  if (!attr) { AGB_CUDA(cudaFuncSetAttribute(small_c_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); attr = true; }
  if (smem > 96 * 1024) return AGB_ERR_UNSUPPORTED;
  small_c_wgrad_kernel<<<(unsigned)blocks, 128, smem, ctx->stream>>>(x, gy->ptr, gw, g, K, K2, chunk);
  AGB_LAUNCHED(ctx);
  return AGB_OK;
}
