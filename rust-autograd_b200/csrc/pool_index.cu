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

__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                          float* __restrict__ idx_f, int32_t* __restrict__ idx_i,
                                                          int64_t n_out, int xh, int xw, int yh, int yw, int size, int stride) {
  int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t gstride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t o = tid; o < n_out; o += gstride) {
    int j = (int)(o % yw); int64_t t = o / yw; int i = (int)(t % yh); int64_t bc = t / yh;
    int h0 = i * stride, w0 = j * stride;
    int h1 = h0 + size > xh ? xh : h0 + size;
    int w1 = w0 + size > xw ? xw : w0 + size;
    const int64_t base = bc * (int64_t)xh * xw;
    float mx = -FLT_MAX; int64_t mi = 0;
    for (int h = h0; h < h1; h++) {
      const float* row = x + base + (int64_t)h * xw;
      for (int w = w0; w < w1; w++) {
        float v = __ldg(row + w);
        if (v > mx) { mx = v; mi = base + (int64_t)h * xw + w; }
      }
    }
    y[o] = mx;
    if (idx_f) idx_f[o] = (float)mi;
    if (idx_i) idx_i[o] = (int32_t)mi;
  }
}

extern "C" int agb_maxpool2d_fwd(agb_ctx* ctx, const agb_tensor* x, agb_tensor* y, float* idx_f32, int32_t* idx_i32,
                                 int size, int pad, int stride) {
  AGB_CHECK(x->rank == 4 && y->rank == 4, AGB_ERR_INCOMPATIBLE_SHAPE, "max_pool2d: input and output must be 4-D");
  AGB_CHECK(pad == 0, AGB_ERR_UNSUPPORTED, "max_pool2d: pad > 0 underflows in the reference (max_pool2d.rs:43,53); only pad == 0 is defined");
  AGB_CHECK(size >= 1 && stride >= 1, AGB_ERR_INVALID_DIMS, "max_pool2d: size and stride must be >= 1");
  AGB_CHECK(agb_is_contig(x) && agb_is_contig(y), AGB_ERR_UNSUPPORTED, "max_pool2d: tensors must be C-contiguous");
  int xh = (int)x->shape[2], xw = (int)x->shape[3];
  AGB_CHECK(xh >= size && xw >= size, AGB_ERR_INCOMPATIBLE_SHAPE, "max_pool2d: window larger than input");
  int yh = (xh + 2 * pad - size) / stride + 1, yw = (xw + 2 * pad - size) / stride + 1;
  AGB_CHECK(y->shape[0] == x->shape[0] && y->shape[1] == x->shape[1] && y->shape[2] == yh && y->shape[3] == yw,
            AGB_ERR_INCOMPATIBLE_SHAPE, "max_pool2d: output shape must be [%lld,%lld,%d,%d]", (long long)x->shape[0], (long long)x->shape[1], yh, yw);
  AGB_CHECK(!idx_i32 || agb_numel(x) < (1ll << 31), AGB_ERR_UNSUPPORTED, "max_pool2d: input too large for int32 indices");
  int64_t n = agb_numel(y); if (n == 0) return AGB_OK;
  maxpool_fwd_kernel<<<agb_grid_for(n, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(x->ptr, y->ptr, idx_f32, idx_i32, n, xh, xw, yh, yw, size, stride);
  AGB_LAUNCHED(ctx);
  return AGB_OK;
}

__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ idx_f,
                                                          const int32_t* __restrict__ idx_i, float* __restrict__ gx, int64_t n) {
  int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t gstride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t o = tid; o < n; o += gstride) {
    int64_t k = idx_i ? (int64_t)__ldg(idx_i + o) : (int64_t)__ldg(idx_f + o);
    atomicAdd(gx + k, __ldg(gy + o));
  }
}
extern "C" int agb_maxpool2d_bwd(agb_ctx* ctx, const agb_tensor* gy, const float* idx_f32, const int32_t* idx_i32, agb_tensor* gx) {
  AGB_CHECK((idx_f32 != nullptr) != (idx_i32 != nullptr), AGB_ERR_INVALID_DIMS, "max_pool2d_grad: exactly one index buffer must be given");
  AGB_CHECK(agb_is_contig(gy) && agb_is_contig(gx), AGB_ERR_UNSUPPORTED, "max_pool2d_grad: tensors must be C-contiguous");
  AGB_TRY(agb_memset0(ctx, gx->ptr, agb_numel(gx) * sizeof(float)));
  int64_t n = agb_numel(gy); if (n == 0) return AGB_OK;
  maxpool_bwd_kernel<<<agb_grid_for(n, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(gy->ptr, idx_f32, idx_i32, gx->ptr, n);
  AGB_LAUNCHED(ctx);
  return AGB_OK;
}

__global__ void __launch_bounds__(256) maxpool_gg_kernel(const float* __restrict__ ggx, const float* __restrict__ idx_f,
                                                         const int32_t* __restrict__ idx_i, float* __restrict__ ggy, int64_t n) {
  int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t gstride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t o = tid; o < n; o += gstride) {
    int64_t k = idx_i ? (int64_t)__ldg(idx_i + o) : (int64_t)__ldg(idx_f + o);
    ggy[o] = __ldg(ggx + k);
  }
}
extern "C" int agb_maxpool2d_gradgrad(agb_ctx* ctx, const agb_tensor* ggx, const float* idx_f32, const int32_t* idx_i32, agb_tensor* ggy) {
  AGB_CHECK((idx_f32 != nullptr) != (idx_i32 != nullptr), AGB_ERR_INVALID_DIMS, "max_pool2d_grad_grad: exactly one index buffer must be given");
  AGB_CHECK(agb_is_contig(ggx) && agb_is_contig(ggy), AGB_ERR_UNSUPPORTED, "max_pool2d_grad_grad: tensors must be C-contiguous");
  int64_t n = agb_numel(ggy); if (n == 0) return AGB_OK;
  maxpool_gg_kernel<<<agb_grid_for(n, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(ggx->ptr, idx_f32, idx_i32, ggy->ptr, n);
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
extern "C" int agb_gather_grad(agb_ctx* ctx, const float* gy, const float* indices, float* gx,
                               int64_t pre, int64_t axis_len, int64_t post, int64_t n_idx) {
  AGB_TRY(agb_memset0(ctx, gx, pre * axis_len * post * sizeof(float)));
  int64_t n = pre * n_idx * post; if (n == 0) return AGB_OK;
  gather_grad_kernel<<<agb_grid_for(n, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(gy, indices, gx, pre, axis_len, post, n_idx, ctx->dev_err);
  AGB_LAUNCHED(ctx);
  return AGB_OK;
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
