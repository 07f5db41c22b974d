// pool_index.cu — max_pool2d (+argmax) forward/backward/grad-grad and gather / gather-grad.
// All HBM-bound index work: outputs must be bit-exact with the reference.
//
// Reference semantics followed:
//   max-pool scan: strict '>' from T::min_value(), first maximum in row-major window order wins,
//   index = flat offset into the whole input buffer (batch included), stored as float
//                                                        src/tensor_ops/conv_ops/max_pool2d.rs:21-88
//   MaxPool2DGrad: gx = 0; gx[idx[i]] += gy[i]             src/tensor_ops/conv_ops/max_pool2d.rs:111-135
//   MaxPool2DGradGrad: ggy[i] = ggx[idx[i]]                src/tensor_ops/conv_ops/max_pool2d.rs:137-159
//   Gather (select along axis)                             src/tensor_ops/array_ops.rs:353-384
//   GatherGrad (duplicates accumulate)                     src/tensor_ops/array_ops.rs:401-466
// Quirk kept: with pad > 0 the reference's `i*stride - pad` wraps in usize; only pad == 0
// is meaningful and only pad == 0 is accepted here.
#include "common.cuh"
#include <float.h>

// Layouts: the activation tensors may be dense NCHW or channels-last (N,H,W,C memory order); x and y of one call share the
// layout, index buffers are laid out like the POOLED tensor (y / gy / ggy).  Index VALUES are always the reference's logical
// NCHW flat offsets.  Threads walk the pooled tensor in its memory order (coalesced), strides do the rest.
struct PoolDims { int C, xh, xw, yh, yw; int64_t xs[4], ys[4]; };   // strides of x-like and y-like tensors (b, c, h, w)
template <bool CL> __device__ __forceinline__ void pool_decode(int64_t o, const PoolDims& d, int& b, int& c, int& i, int& j) {
  if (CL) { c = (int)(o % d.C); int64_t t = o / d.C; j = (int)(t % d.yw); t /= d.yw; i = (int)(t % d.yh); b = (int)(t / d.yh); }
  else { j = (int)(o % d.yw); int64_t t = o / d.yw; i = (int)(t % d.yh); t /= d.yh; c = (int)(t % d.C); b = (int)(t / d.C); }
}
template <bool CL>
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                          float* __restrict__ idx_f, int32_t* __restrict__ idx_i,
                                                          int64_t n_out, PoolDims d, int size, int stride) {
  int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t gstride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t o = tid; o < n_out; o += gstride) {
    int b, c, i, j; pool_decode<CL>(o, d, b, c, i, j);
    int h0 = i * stride, w0 = j * stride;
    int h1 = h0 + size > d.xh ? d.xh : h0 + size;
    int w1 = w0 + size > d.xw ? d.xw : w0 + size;
    const float* xp = x + b * d.xs[0] + c * d.xs[1];
    const int64_t lbase = ((int64_t)b * d.C + c) * d.xh * d.xw;
    float mx = -FLT_MAX; int64_t mi = 0;
    for (int h = h0; h < h1; h++)
      for (int w = w0; w < w1; w++) {
        float v = __ldg(xp + h * d.xs[2] + w * d.xs[3]);
        if (v > mx) { mx = v; mi = lbase + (int64_t)h * d.xw + w; }
      }
    y[o] = mx;
    if (idx_f) idx_f[o] = (float)mi;
    if (idx_i) idx_i[o] = (int32_t)mi;
  }
}

// channels-last, C % 4 == 0: one thread owns 4 consecutive channels of one output pixel (128-bit loads / stores, 32-bit indexing)
__global__ void __launch_bounds__(256) maxpool_fwd_cl4_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ idx_f, int32_t* __restrict__ idx_i,
                                                              uint32_t n4, int C, int xh, int xw, int yh, int yw, int size, int stride) {
  const uint32_t c4n = (uint32_t)C >> 2;
  for (uint32_t o = blockIdx.x * blockDim.x + threadIdx.x; o < n4; o += gridDim.x * blockDim.x) {
    uint32_t c4 = o % c4n, t = o / c4n; uint32_t j = t % (uint32_t)yw; t /= (uint32_t)yw; uint32_t i = t % (uint32_t)yh, b = t / (uint32_t)yh;
    int h0 = (int)i * stride, w0 = (int)j * stride;
    int h1 = h0 + size > xh ? xh : h0 + size, w1 = w0 + size > xw ? xw : w0 + size;
    const int c = (int)c4 * 4;
    float mx[4] = {-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX}; int mi[4] = {0, 0, 0, 0};
    if (size == 2 && h0 + 2 <= xh && w0 + 2 <= xw) {
      // the 2x2 window of every VGG / cnn_mnist pool: all four 128-bit loads are issued before the first compare (the generic loop below
      // cannot be unrolled — runtime bounds — and paid four dependent memory round trips per thread: 0.52-0.57 of the HBM rate)
      const float* p0 = x + (((size_t)b * xh + h0) * xw + w0) * C + c;
      const float4 v00 = ldg_stream4(p0), v01 = ldg_stream4(p0 + C), v10 = ldg_stream4(p0 + (size_t)xw * C), v11 = ldg_stream4(p0 + (size_t)xw * C + C);
      const int hw = h0 * xw + w0;
#define POOL_STEP(V, HW) do { if (V.x > mx[0]) { mx[0] = V.x; mi[0] = (HW); } if (V.y > mx[1]) { mx[1] = V.y; mi[1] = (HW); } \
                              if (V.z > mx[2]) { mx[2] = V.z; mi[2] = (HW); } if (V.w > mx[3]) { mx[3] = V.w; mi[3] = (HW); } } while (0)
      POOL_STEP(v00, hw); POOL_STEP(v01, hw + 1); POOL_STEP(v10, hw + xw); POOL_STEP(v11, hw + xw + 1);      // scan order of the reference: first maximum wins
#undef POOL_STEP
    } else
    for (int h = h0; h < h1; h++)
      for (int w = w0; w < w1; w++) {
        float4 v = ldg_stream4(x + (((size_t)b * xh + h) * xw + w) * C + c);
        const int hw = h * xw + w;
        if (v.x > mx[0]) { mx[0] = v.x; mi[0] = hw; }
        if (v.y > mx[1]) { mx[1] = v.y; mi[1] = hw; }
        if (v.z > mx[2]) { mx[2] = v.z; mi[2] = hw; }
        if (v.w > mx[3]) { mx[3] = v.w; mi[3] = hw; }
      }
    // logical NCHW flat offset = ((b*C + c)*xh + h)*xw + w ; a window that never fired keeps index 0 (max_pool2d.rs:51-52)
    int64_t lb = ((int64_t)b * C + c) * xh * xw; const int64_t plane = (int64_t)xh * xw;
    int64_t li[4];
#pragma unroll
    for (int q = 0; q < 4; q++) li[q] = mx[q] == -FLT_MAX ? 0 : lb + q * plane + mi[q];
    const size_t off = (size_t)o * 4;
    *(float4*)(y + off) = make_float4(mx[0], mx[1], mx[2], mx[3]);
    if (idx_f) *(float4*)(idx_f + off) = make_float4((float)li[0], (float)li[1], (float)li[2], (float)li[3]);
    if (idx_i) *(int4*)(idx_i + off) = make_int4((int)li[0], (int)li[1], (int)li[2], (int)li[3]);
  }
}

static bool pool_is_cl(const agb_tensor* t) {
  const int64_t C = t->shape[1], H = t->shape[2], W = t->shape[3];
  return !agb_is_contig(t) && (C == 1 || t->stride[1] == 1) && (W == 1 || t->stride[3] == C) && (H == 1 || t->stride[2] == W * C) && (t->shape[0] == 1 || t->stride[0] == H * W * C);
}
static int pool_layout(const char* who, const agb_tensor* t, bool* cl) {
  AGB_CHECK(t->rank == 4, AGB_ERR_INCOMPATIBLE_SHAPE, "%s: tensors must be 4-D", who);
  if (agb_is_contig(t)) { *cl = false; return AGB_OK; }
  AGB_CHECK(pool_is_cl(t), AGB_ERR_UNSUPPORTED, "%s: tensors must be dense NCHW or channels-last", who);
  *cl = true; return AGB_OK;
}

extern "C" int agb_maxpool2d_fwd(agb_ctx* ctx, const agb_tensor* x, agb_tensor* y, float* idx_f32, int32_t* idx_i32,
                                 int size, int pad, int stride) {
  AGB_CHECK(x->rank == 4 && y->rank == 4, AGB_ERR_INCOMPATIBLE_SHAPE, "max_pool2d: input and output must be 4-D");
  AGB_CHECK(pad == 0, AGB_ERR_UNSUPPORTED, "max_pool2d: pad > 0 underflows in the reference (max_pool2d.rs:43,53); only pad == 0 is defined");
  AGB_CHECK(size >= 1 && stride >= 1, AGB_ERR_INVALID_DIMS, "max_pool2d: size and stride must be >= 1");
  bool xcl, ycl; AGB_TRY(pool_layout("max_pool2d", x, &xcl)); AGB_TRY(pool_layout("max_pool2d", y, &ycl));
  int xh = (int)x->shape[2], xw = (int)x->shape[3];
  AGB_CHECK(xh >= size && xw >= size, AGB_ERR_INCOMPATIBLE_SHAPE, "max_pool2d: window larger than input");
  int yh = (xh + 2 * pad - size) / stride + 1, yw = (xw + 2 * pad - size) / stride + 1;
  AGB_CHECK(y->shape[0] == x->shape[0] && y->shape[1] == x->shape[1] && y->shape[2] == yh && y->shape[3] == yw,
            AGB_ERR_INCOMPATIBLE_SHAPE, "max_pool2d: output shape must be [%lld,%lld,%d,%d]", (long long)x->shape[0], (long long)x->shape[1], yh, yw);
  AGB_CHECK(!idx_i32 || agb_numel(x) < (1ll << 31), AGB_ERR_UNSUPPORTED, "max_pool2d: input too large for int32 indices");
  int64_t n = agb_numel(y); if (n == 0) return AGB_OK;
  AgbProfScope prof(ctx, AGB_PROF_POOL, 4.0 * (double)(agb_numel(x) + 2 * n));
  PoolDims d; d.C = (int)x->shape[1]; d.xh = xh; d.xw = xw; d.yh = yh; d.yw = yw;
  for (int k = 0; k < 4; k++) { d.xs[k] = x->stride[k]; d.ys[k] = y->stride[k]; }
  int grid = agb_grid_for(n, 256, ctx->sm_count, 8);
  if (ycl && xcl && d.C % 4 == 0 && n / 4 < (1ll << 31) && agb_numel(x) < (1ll << 32) && ((((uintptr_t)x->ptr | (uintptr_t)y->ptr | (uintptr_t)idx_f32 | (uintptr_t)idx_i32) & 15) == 0)) {
    maxpool_fwd_cl4_kernel<<<agb_grid_occ(ctx, maxpool_fwd_cl4_kernel, n / 4, 256), 256, 0, ctx->stream>>>(x->ptr, y->ptr, idx_f32, idx_i32, (uint32_t)(n / 4), d.C, xh, xw, yh, yw, size, stride);
    AGB_LAUNCHED(ctx);
    return AGB_OK;
  }
  if (ycl) maxpool_fwd_kernel<true><<<grid, 256, 0, ctx->stream>>>(x->ptr, y->ptr, idx_f32, idx_i32, n, d, size, stride);
  else maxpool_fwd_kernel<false><<<grid, 256, 0, ctx->stream>>>(x->ptr, y->ptr, idx_f32, idx_i32, n, d, size, stride);
  AGB_LAUNCHED(ctx);
  return AGB_OK;
}

// logical NCHW flat offset -> memory offset of a tensor with strides xs
__device__ __forceinline__ int64_t pool_phys(int64_t k, const PoolDims& d, bool x_cl) {
  if (!x_cl) return k;
  int w = (int)(k % d.xw); int64_t t = k / d.xw; int h = (int)(t % d.xh); t /= d.xh; int c = (int)(t % d.C); int64_t b = t / d.C;
  return b * d.xs[0] + c * d.xs[1] + h * d.xs[2] + w * d.xs[3];
}
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ idx_f,
                                                          const int32_t* __restrict__ idx_i, const float* __restrict__ gate,
                                                          float* __restrict__ gx, int64_t n, PoolDims d, bool x_cl) {
  int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t gstride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t o = tid; o < n; o += gstride) {
    int64_t k = idx_i ? (int64_t)__ldg(idx_i + o) : (int64_t)__ldg(idx_f + o);
    float g = __ldg(gy + o);
    if (gate != nullptr) g = __ldg(gate + o) > 0.0f ? g : 0.0f * g;      // 0*g keeps NaN/Inf propagation of the un-fused multiply
    atomicAdd(gx + pool_phys(k, d, x_cl), g);
  }
}
// Non-overlapping windows that tile gx exactly (size == stride, xh == yh*size, xw == yw*size — every VGG / cnn_mnist pool): each
// gx element belongs to exactly one window, so the scatter becomes a gather-form full write: no memset, no atomics.
// The reference accumulates gx[idx[i]] += gy[i] (max_pool2d.rs:111-135); with disjoint windows at most one term lands on an
// element, except for windows whose scan never fired (index 0, NaN-only windows) — those are sent down the scatter path.
template <bool CL>
__global__ void __launch_bounds__(256) maxpool_bwd_tiled_kernel(const float* __restrict__ gy, const float* __restrict__ idx_f,
                                                                const int32_t* __restrict__ idx_i, const float* __restrict__ gate,
                                                                float* __restrict__ gx, int64_t n, PoolDims d, int size) {
  int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t gstride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t o = tid; o < n; o += gstride) {
    int b, c, i, j; pool_decode<CL>(o, d, b, c, i, j);
    const int64_t k = idx_i ? (int64_t)__ldg(idx_i + o) : (int64_t)__ldg(idx_f + o);
    float g = __ldg(gy + o);
    if (gate != nullptr) g = __ldg(gate + o) > 0.0f ? g : 0.0f * g;
    const int64_t lbase = ((int64_t)b * d.C + c) * d.xh * d.xw;
    float* xp = gx + b * d.xs[0] + c * d.xs[1];
    for (int dh = 0; dh < size; dh++)
      for (int dw = 0; dw < size; dw++) {
        const int h = i * size + dh, w = j * size + dw;
        xp[h * d.xs[2] + w * d.xs[3]] = (lbase + (int64_t)h * d.xw + w == k) ? g : 0.0f;
      }
  }
}
// channels-last, C % 4 == 0, int32 indices: 4 channels per thread, 128-bit accesses
__global__ void __launch_bounds__(256) maxpool_bwd_tiled_cl4_kernel(const float* __restrict__ gy, const int32_t* __restrict__ idx_i,
                                                                    const float* __restrict__ gate, float* __restrict__ gx, float* __restrict__ csum, float* __restrict__ csum_part,
                                                                    uint32_t n4, int C, int xh, int xw, int yh, int yw, int size) {
  const uint32_t c4n = (uint32_t)C >> 2;
  // csum != NULL (host guarantees 256 % c4n == 0): a thread meets the same 4 channels at every grid-stride step, so the per-channel
  // sums of gx (= sums of the gated gy: every other window element is 0) accumulate in registers
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
  for (uint32_t o = blockIdx.x * blockDim.x + threadIdx.x; o < n4; o += gridDim.x * blockDim.x) {
    uint32_t c4 = o % c4n, t = o / c4n; uint32_t j = t % (uint32_t)yw; t /= (uint32_t)yw; uint32_t i = t % (uint32_t)yh, b = t / (uint32_t)yh;
    const int c = (int)c4 * 4;
    const size_t off = (size_t)o * 4;
    float4 g = ldg_stream4(gy + off);
    const int4 k = __ldg((const int4*)(idx_i + off));
    if (gate != nullptr) {
      const float4 m = ldg_stream4(gate + off);
      g.x = m.x > 0.0f ? g.x : 0.0f * g.x; g.y = m.y > 0.0f ? g.y : 0.0f * g.y; g.z = m.z > 0.0f ? g.z : 0.0f * g.z; g.w = m.w > 0.0f ? g.w : 0.0f * g.w;
    }
    const int plane = xh * xw; const int lb = (int)((b * (uint32_t)C + c) * (uint32_t)plane);     // host checks numel(gx) < 2^31
    for (int dh = 0; dh < size; dh++)
      for (int dw = 0; dw < size; dw++) {
        const int h = (int)i * size + dh, w = (int)j * size + dw;
        const int li = lb + h * xw + w;
        float4 v;
        v.x = (li == k.x) ? g.x : 0.0f; v.y = (li + plane == k.y) ? g.y : 0.0f; v.z = (li + 2 * plane == k.z) ? g.z : 0.0f; v.w = (li + 3 * plane == k.w) ? g.w : 0.0f;
        stg_stream4(gx + (((size_t)b * xh + h) * xw + w) * C + c, v);
        cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w;
      }
  }
  if (csum != nullptr) {
    __shared__ float4 cs_t[256];                       // every thread's four sums; channel c is then added over the threads that met it, in thread order
    cs_t[threadIdx.x] = cs;
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float s = 0.0f;
      for (uint32_t t = (uint32_t)c >> 2; t < 256; t += c4n) { const float4 v = cs_t[t]; s += (c & 3) == 0 ? v.x : (c & 3) == 1 ? v.y : (c & 3) == 2 ? v.z : v.w; }
      if (csum_part != nullptr) csum_part[(size_t)blockIdx.x * C + c] = s;      // deterministic mode: per-CTA partials, added in CTA order by agb_reduce_partials
      else atomicAdd(csum + c, s);
    }
  }
}
// gy and the index buffer share one layout (either); gx may be NCHW or channels-last
extern "C" int agb_maxpool2d_bwd(agb_ctx* ctx, const agb_tensor* gy, const float* idx_f32, const int32_t* idx_i32, agb_tensor* gx) {
  return agb_maxpool2d_bwd_fused(ctx, gy, idx_f32, idx_i32, nullptr, nullptr, gx, 0, 0);
}
// gx = scatter(gy * (gate > 0)) — `gate` (nullable, laid out like gy) is the POOLED forward output when the pooled input was a
// ReLU output: relu'(x[argmax]) == (max > 0), so the ReLU backward that follows a pool backward in conv->relu->pool stacks
// (activation_ops.rs:161-166) costs one extra read of a 1/size^2-sized tensor instead of a 3-array pass.
// size/stride = the forward window (0 = unknown: always the scatter form).
static int pool_channel_sums(agb_ctx* ctx, const agb_tensor* t, bool cl, float* chan_sum) {
  const int64_t B = t->shape[0], C = t->shape[1], HW = t->shape[2] * t->shape[3];
  if (cl) return agb_reduce(ctx, AGB_R_SUM, t->ptr, chan_sum, 1, B * HW, C);
  float* tmp; AGB_TRY(agb_alloc(ctx, (size_t)(B * C) * sizeof(float), (void**)&tmp));
  int r = agb_reduce(ctx, AGB_R_SUM, t->ptr, tmp, B * C, HW, 1);
  if (r == AGB_OK) r = agb_reduce(ctx, AGB_R_SUM, tmp, chan_sum, 1, B, C);
  agb_free(ctx, tmp);
  return r;
}
extern "C" int agb_maxpool2d_bwd_fused(agb_ctx* ctx, const agb_tensor* gy, const float* idx_f32, const int32_t* idx_i32, const float* gate,
                                       float* chan_sum, agb_tensor* gx, int size, int stride) {
  AGB_CHECK((idx_f32 != nullptr) != (idx_i32 != nullptr), AGB_ERR_INVALID_DIMS, "max_pool2d_grad: exactly one index buffer must be given");
  bool gcl, xcl; AGB_TRY(pool_layout("max_pool2d_grad", gy, &gcl)); AGB_TRY(pool_layout("max_pool2d_grad", gx, &xcl));
  int64_t n = agb_numel(gy);
  PoolDims d; d.C = (int)gx->shape[1]; d.xh = (int)gx->shape[2]; d.xw = (int)gx->shape[3]; d.yh = (int)gy->shape[2]; d.yw = (int)gy->shape[3];
  for (int k = 0; k < 4; k++) { d.xs[k] = gx->stride[k]; d.ys[k] = gy->stride[k]; }
  const bool tiled = size >= 1 && size == stride && gcl == xcl && n > 0 && gy->shape[0] == gx->shape[0] && gy->shape[1] == gx->shape[1] &&
                     (int64_t)d.yh * size == d.xh && (int64_t)d.yw * size == d.xw;
  if (tiled) {
    AgbProfScope prof(ctx, AGB_PROF_POOL, 4.0 * (double)(agb_numel(gx) + (gate ? 3 : 2) * n));
    if (gcl && idx_i32 && d.C % 4 == 0 && agb_numel(gx) < (1ll << 31) &&
        ((((uintptr_t)gy->ptr | (uintptr_t)gx->ptr | (uintptr_t)idx_i32 | (uintptr_t)gate) & 15) == 0)) {
      const bool fuse_sum = chan_sum != nullptr && d.C <= 1024 && 256 % (d.C / 4) == 0;
      const int grid = agb_grid_occ(ctx, maxpool_bwd_tiled_cl4_kernel, n / 4, 256);
      float* part = nullptr;
      if (fuse_sum && ctx->deterministic) AGB_TRY(agb_scratch2(ctx, (size_t)grid * d.C * sizeof(float), (void**)&part));
      if (fuse_sum && !part) AGB_TRY(agb_memset0(ctx, chan_sum, (size_t)d.C * sizeof(float)));
      maxpool_bwd_tiled_cl4_kernel<<<grid, 256, 0, ctx->stream>>>(gy->ptr, idx_i32, gate, gx->ptr, fuse_sum ? chan_sum : nullptr, part,
                                                                (uint32_t)(n / 4), d.C, d.xh, d.xw, d.yh, d.yw, size);
      AGB_LAUNCHED(ctx);
      if (part) AGB_TRY(agb_reduce_partials(ctx, part, chan_sum, grid, d.C, d.C, 0));
      return (chan_sum != nullptr && !fuse_sum) ? pool_channel_sums(ctx, gx, true, chan_sum) : AGB_OK;
    } else if (gcl) maxpool_bwd_tiled_kernel<true><<<agb_grid_for(n, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(gy->ptr, idx_f32, idx_i32, gate, gx->ptr, n, d, size);
    else maxpool_bwd_tiled_kernel<false><<<agb_grid_for(n, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(gy->ptr, idx_f32, idx_i32, gate, gx->ptr, n, d, size);
    AGB_LAUNCHED(ctx);
    return chan_sum != nullptr ? pool_channel_sums(ctx, gx, gcl, chan_sum) : AGB_OK;
  }
  AGB_TRY(agb_memset0(ctx, gx->ptr, agb_numel(gx) * sizeof(float)));
  if (n == 0) return chan_sum != nullptr ? agb_memset0(ctx, chan_sum, (size_t)gx->shape[1] * sizeof(float)) : AGB_OK;
  AgbProfScope prof(ctx, AGB_PROF_POOL, 4.0 * (double)(agb_numel(gx) + 3 * n));
  maxpool_bwd_kernel<<<agb_grid_for(n, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(gy->ptr, idx_f32, idx_i32, gate, gx->ptr, n, d, xcl);
  AGB_LAUNCHED(ctx);
  return chan_sum != nullptr ? pool_channel_sums(ctx, gx, xcl, chan_sum) : AGB_OK;
}

__global__ void __launch_bounds__(256) maxpool_gg_kernel(const float* __restrict__ ggx, const float* __restrict__ idx_f,
                                                         const int32_t* __restrict__ idx_i, float* __restrict__ ggy, int64_t n, PoolDims d, bool x_cl) {
  int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t gstride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t o = tid; o < n; o += gstride) {
    int64_t k = idx_i ? (int64_t)__ldg(idx_i + o) : (int64_t)__ldg(idx_f + o);
    ggy[o] = __ldg(ggx + pool_phys(k, d, x_cl));
  }
}
// ggy and the index buffer share one layout; ggx may be NCHW or channels-last
extern "C" int agb_maxpool2d_gradgrad(agb_ctx* ctx, const agb_tensor* ggx, const float* idx_f32, const int32_t* idx_i32, agb_tensor* ggy) {
  AGB_CHECK((idx_f32 != nullptr) != (idx_i32 != nullptr), AGB_ERR_INVALID_DIMS, "max_pool2d_grad_grad: exactly one index buffer must be given");
  bool xcl, ycl; AGB_TRY(pool_layout("max_pool2d_grad_grad", ggx, &xcl)); AGB_TRY(pool_layout("max_pool2d_grad_grad", ggy, &ycl));
  int64_t n = agb_numel(ggy); if (n == 0) return AGB_OK;
  PoolDims d; d.C = (int)ggx->shape[1]; d.xh = (int)ggx->shape[2]; d.xw = (int)ggx->shape[3]; d.yh = (int)ggy->shape[2]; d.yw = (int)ggy->shape[3];
  for (int k = 0; k < 4; k++) { d.xs[k] = ggx->stride[k]; d.ys[k] = ggy->stride[k]; }
  maxpool_gg_kernel<<<agb_grid_for(n, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(ggx->ptr, idx_f32, idx_i32, ggy->ptr, n, d, xcl);
  AGB_LAUNCHED(ctx);
  return AGB_OK;
}

// ---------------------------------------------------------------------------------------------
// gather: out[pre, j, post] = param[pre, idx[j], post]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_kernel(const float* __restrict__ param, const float* __restrict__ indices,
                                                     float* __restrict__ out, int64_t pre, int64_t axis_len, int64_t post,
                                                     int64_t n_idx, int normalize, int* err) {
  int64_t n = pre * n_idx * post;
  int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t gstride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t o = tid; o < n; o += gstride) {
    int64_t q = o % post; int64_t t = o / post; int64_t j = t % n_idx; int64_t p = t / n_idx;
    float f = __ldg(indices + j);
    int64_t k = (int64_t)f;
    if (normalize && k < 0) k += axis_len;
    if (k < 0 || k >= axis_len || f != f) { atomicExch(err, 2); out[o] = nanf(""); continue; }
    out[o] = __ldg(param + (p * axis_len + k) * post + q);
  }
}
extern "C" int agb_gather(agb_ctx* ctx, const float* param, const float* indices, float* out,
                          int64_t pre, int64_t axis_len, int64_t post, int64_t n_idx, int normalize_negative) {
  int64_t n = pre * n_idx * post; if (n == 0) return AGB_OK;
  gather_kernel<<<agb_grid_for(n, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(param, indices, out, pre, axis_len, post, n_idx, normalize_negative, ctx->dev_err);
  AGB_LAUNCHED(ctx);
  return AGB_OK;
}

__global__ void __launch_bounds__(256) gather_grad_kernel(const float* __restrict__ gy, const float* __restrict__ indices,
                                                          float* __restrict__ gx, int64_t pre, int64_t axis_len, int64_t post,
                                                          int64_t n_idx, int* err) {
  int64_t n = pre * n_idx * post;
  int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t gstride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t o = tid; o < n; o += gstride) {
    int64_t q = o % post; int64_t t = o / post; int64_t j = t % n_idx; int64_t p = t / n_idx;
    float f = __ldg(indices + j);
    int64_t k = (int64_t)f;
    if (k < 0) k += axis_len;      // ndarray slices with a negative start count from the end (array_ops.rs:431-436)
    if (k < 0 || k >= axis_len || f != f) { atomicExch(err, 2); continue; }
    atomicAdd(gx + (p * axis_len + k) * post + q, __ldg(gy + o));
  }
}
// rows of >= 4 floats (an embedding table): four values per thread, one 128-bit load and ONE vector reduction (red.global.add.v4.f32, sm_90+) instead of
// four scalar atomics and four rounds of 64-bit index arithmetic
__global__ void __launch_bounds__(256) gather_grad_v4_kernel(const float* __restrict__ gy, const float* __restrict__ indices,
                                                             float* __restrict__ gx, int64_t pre, int64_t axis_len, int64_t post4,
                                                             int64_t n_idx, int* err) {
  const int64_t n4 = pre * n_idx * post4;
  const int64_t gstride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < n4; o += gstride) {
    const int64_t q4 = o % post4; const int64_t t = o / post4; const int64_t j = t % n_idx; const int64_t p = t / n_idx;
    const float f = __ldg(indices + j);
    int64_t k = (int64_t)f;
    if (k < 0) k += axis_len;
    if (k < 0 || k >= axis_len || f != f) { atomicExch(err, 2); continue; }
    const float4 v = ldg_stream4(gy + 4 * o);
    float* d = gx + ((p * axis_len + k) * post4 + q4) * 4;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(d), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  }
}
extern "C" int agb_scatter_add(agb_ctx* ctx, const float* gy, const float* indices, float* gx,
                               int64_t pre, int64_t axis_len, int64_t post, int64_t n_idx) {
  int64_t n = pre * n_idx * post; if (n == 0) return AGB_OK;
  if (post % 4 == 0 && ((((uintptr_t)gy | (uintptr_t)gx) & 15) == 0)) {
    gather_grad_v4_kernel<<<agb_grid_for(n / 4, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(gy, indices, gx, pre, axis_len, post / 4, n_idx, ctx->dev_err);
    AGB_LAUNCHED(ctx);
    return AGB_OK;
  }
  gather_grad_kernel<<<agb_grid_for(n, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(gy, indices, gx, pre, axis_len, post, n_idx, ctx->dev_err);
  AGB_LAUNCHED(ctx);
  return AGB_OK;
}
extern "C" int agb_gather_grad(agb_ctx* ctx, const float* gy, const float* indices, float* gx,
                               int64_t pre, int64_t axis_len, int64_t post, int64_t n_idx) {
  AGB_TRY(agb_memset0(ctx, gx, pre * axis_len * post * sizeof(float)));
  return agb_scatter_add(ctx, gy, indices, gx, pre, axis_len, post, n_idx);
}

__global__ void __launch_bounds__(256) i32_to_f32_kernel(const int32_t* __restrict__ s, float* __restrict__ d, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) d[i] = (float)s[i];
}
extern "C" int agb_convert_i32_f32(agb_ctx* ctx, const int32_t* src, float* dst, int64_t n) {
  if (n == 0) return AGB_OK;
  i32_to_f32_kernel<<<agb_grid_for(n, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(src, dst, n);
  AGB_LAUNCHED(ctx);
  return AGB_OK;
}
