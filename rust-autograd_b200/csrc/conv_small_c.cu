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

template <int K>
__global__ void __launch_bounds__(128) small_c_fprop_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y, SmallGeom g) {
  extern __shared__ float ws[];                       // [O][K]
  for (int i = threadIdx.x; i < g.O * K; i += blockDim.x) ws[i] = __ldg(w + i);
  __syncthreads();
  const int64_t P = (int64_t)g.yh * g.yw, total = (int64_t)g.B * P;
  const int64_t pix = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (pix >= total) return;
  const int b = (int)(pix / P); const int r = (int)(pix - (int64_t)b * P); const int oy = r / g.yw, ox = r - oy * g.yw;
  float v[K];
  const int kk = g.kh * g.kw;
#pragma unroll
  for (int k = 0; k < K; k++) {
    const int c = k / kk, t = k - c * kk, i = t / g.kw, j = t - i * g.kw;
    const int iy = oy * g.stride - g.pad + i * g.dil, ix = ox * g.stride - g.pad + j * g.dil;
    v[k] = ((unsigned)iy < (unsigned)g.H && (unsigned)ix < (unsigned)g.W) ? __ldg(x + (((int64_t)b * g.C + c) * g.H + iy) * g.W + ix) : 0.0f;
  }
  float* yp = y + b * g.ys[0] + oy * g.ys[2] + ox * g.ys[3];
  const bool cl = g.ys[1] == 1;
  for (int o0 = 0; o0 < g.O; o0 += 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 4; q++) {
      if (o0 + q < g.O) {
        const float* wr = ws + (o0 + q) * K;
#pragma unroll
        for (int k = 0; k < K; k++) acc[q] = fmaf(wr[k], v[k], acc[q]);
      }
    }
    if (cl && o0 + 4 <= g.O) *(float4*)(yp + o0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    else { for (int q = 0; q < 4; q++) if (o0 + q < g.O) yp[(o0 + q) * g.ys[1]] = acc[q]; }
  }
}

// gy strides gs (b,o,h,w) in either layout; K <= SC_MAXK, O <= 256 (multiple of 4).
// Register tiling: thread (to, tk) owns a 4 (o) x 2 (k) block of gw for every 64-channel slab of O, so one slab pixel costs
// one 128-bit + one 64-bit shared-memory read per 8 FMAs (a thread-per-output mapping would be shared-memory bound).
#define SCW_PIX 64
__global__ void __launch_bounds__(256) small_c_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ gw, SmallGeom g,
                                                            int K, int K2 /* K rounded up to even */, int64_t chunk) {
  extern __shared__ __align__(16) float sm[];
  float* gs = sm;                                     // [SCW_PIX][O]
  float* ps = sm + SCW_PIX * g.O;                     // [SCW_PIX][K2]
  const int64_t P = (int64_t)g.yh * g.yw, total = (int64_t)g.B * P;
  const int64_t p0 = blockIdx.x * chunk, p1 = min(p0 + chunk, total);
  const int kk = g.kh * g.kw;
  const int to = threadIdx.x & 15, tk = threadIdx.x >> 4;          // 16 o-quads x 16 k-pairs per 64-channel slab
  const int nslab = (g.O + 63) / 64;
  const bool active = 2 * tk < K2;
  float acc[4][4][2];                                  // [slab][o in quad][k in pair]
#pragma unroll
  for (int s = 0; s < 4; s++)
#pragma unroll
    for (int q = 0; q < 4; q++) { acc[s][q][0] = 0.f; acc[s][q][1] = 0.f; }
  const bool gy_cl = g.ys[1] == 1;
  for (int64_t base = p0; base < p1; base += SCW_PIX) {
    const int np = (int)min((int64_t)SCW_PIX, p1 - base);
    for (int i = threadIdx.x; i < np * g.O; i += blockDim.x) {
      int pi, o;
      if (gy_cl) { pi = i / g.O; o = i - pi * g.O; } else { o = i / np; pi = i - o * np; }      // coalesced in either layout
      const int64_t pix = base + pi;
      const int b = (int)(pix / P); const int r = (int)(pix - (int64_t)b * P); const int oy = r / g.yw, ox = r - oy * g.yw;
      gs[pi * g.O + o] = __ldg(gy + b * g.ys[0] + o * g.ys[1] + oy * g.ys[2] + ox * g.ys[3]);
    }
    for (int i = threadIdx.x; i < np * K2; i += blockDim.x) {
      const int k = i / np, pi = i - k * np; const int64_t pix = base + pi;        // consecutive threads -> consecutive pixels (coalesced x reads)
      float v = 0.0f;
      if (k < K) {
        const int b = (int)(pix / P); const int r = (int)(pix - (int64_t)b * P); const int oy = r / g.yw, ox = r - oy * g.yw;
        const int c = k / kk, t = k - c * kk, ii = t / g.kw, jj = t - ii * g.kw;
        const int iy = oy * g.stride - g.pad + ii * g.dil, ix = ox * g.stride - g.pad + jj * g.dil;
        if ((unsigned)iy < (unsigned)g.H && (unsigned)ix < (unsigned)g.W) v = __ldg(x + (((int64_t)b * g.C + c) * g.H + iy) * g.W + ix);
      }
      ps[pi * K2 + k] = v;
    }
    __syncthreads();
    if (active) {
      for (int pi = 0; pi < np; pi++) {
        const float2 pv = *(const float2*)(ps + pi * K2 + 2 * tk);
#pragma unroll
        for (int s = 0; s < 4; s++) {
          if (s < nslab) {
            const float4 gv = *(const float4*)(gs + pi * g.O + s * 64 + 4 * to);
            acc[s][0][0] = fmaf(gv.x, pv.x, acc[s][0][0]); acc[s][0][1] = fmaf(gv.x, pv.y, acc[s][0][1]);
            acc[s][1][0] = fmaf(gv.y, pv.x, acc[s][1][0]); acc[s][1][1] = fmaf(gv.y, pv.y, acc[s][1][1]);
            acc[s][2][0] = fmaf(gv.z, pv.x, acc[s][2][0]); acc[s][2][1] = fmaf(gv.z, pv.y, acc[s][2][1]);
            acc[s][3][0] = fmaf(gv.w, pv.x, acc[s][3][0]); acc[s][3][1] = fmaf(gv.w, pv.y, acc[s][3][1]);
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
        for (int e = 0; e < 2; e++) {
          const int o = s * 64 + 4 * to + q, k = 2 * tk + e;
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
                      int pad, int stride, int dil) {
  SmallGeom g; fill_geom(g, B, C, H, W, O, kh, kw, yh, yw, pad, stride, dil, y);
  const int K = C * kh * kw; const int64_t total = (int64_t)B * yh * yw;
  const size_t smem = (size_t)O * K * sizeof(float);
  dim3 grid((unsigned)((total + 127) / 128));
  switch (K) {
#define SC_CASE(KK) case KK: small_c_fprop_kernel<KK><<<grid, 128, smem, ctx->stream>>>(x, w, y->ptr, g); break;
    SC_CASE(1) SC_CASE(2) SC_CASE(3) SC_CASE(4) SC_CASE(8) SC_CASE(9) SC_CASE(12) SC_CASE(16) SC_CASE(18) SC_CASE(25) SC_CASE(27) SC_CASE(32) SC_CASE(36)
#undef SC_CASE
    default: return AGB_ERR_UNSUPPORTED;
  }
  AGB_LAUNCHED(ctx);
  return AGB_OK;
}

// gy may be NCHW or channels-last; gw [O, C*kh*kw] contiguous
int agb_small_c_wgrad(agb_ctx* ctx, const float* x, const agb_tensor* gy, float* gw, int B, int C, int H, int W, int O, int kh, int kw, int yh, int yw,
                      int pad, int stride, int dil) {
  SmallGeom g; fill_geom(g, B, C, H, W, O, kh, kw, yh, yw, pad, stride, dil, gy);
  const int K = C * kh * kw; const int64_t total = (int64_t)B * yh * yw;
  if (O % 4 != 0 || O > 256 || K > 32) return AGB_ERR_UNSUPPORTED;       // 16 k-pairs per thread row
  const int K2 = (K + 1) & ~1;
  AGB_TRY(agb_memset0(ctx, gw, (size_t)O * K * sizeof(float)));
  int64_t blocks = 4ll * ctx->sm_count; int64_t chunk = (total + blocks - 1) / blocks; chunk = (chunk + 63) / 64 * 64; if (chunk < 64) chunk = 64;
  blocks = (total + chunk - 1) / chunk;
  const size_t smem = (size_t)SCW_PIX * (O + K2) * sizeof(float);
  static bool attr = false;
  if (!attr) { AGB_CUDA(cudaFuncSetAttribute(small_c_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); attr = true; }
  if (smem > 96 * 1024) return AGB_ERR_UNSUPPORTED;
  small_c_wgrad_kernel<<<(unsigned)blocks, 256, smem, ctx->stream>>>(x, gy->ptr, gw, g, K, K2, chunk);
  AGB_LAUNCHED(ctx);
  return AGB_OK;
}
