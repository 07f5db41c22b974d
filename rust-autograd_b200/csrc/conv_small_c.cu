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

// gy may be NCHW or channels-last; gw [O, C*kh*kw] contiguous
int agb_small_c_wgrad(agb_ctx* ctx, const float* x, const agb_tensor* gy, float* gw, int B, int C, int H, int W, int O, int kh, int kw, int yh, int yw,
                      int pad, int stride, int dil) {
  SmallGeom g; fill_geom(g, B, C, H, W, O, kh, kw, yh, yw, pad, stride, dil, gy);
  const int K = C * kh * kw; const int64_t total = (int64_t)B * yh * yw;
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
