// conv.cu — conv2d fprop / dgrad / wgrad entry points (NCHW, f32, square pad/stride/dilation) and the
// general CUDA-core implicit-GEMM path.  The im2col matrix (reference: conv_ops/mod.rs:73-124) is never
// materialised: loaders compute the (channel, tap, pixel) -> input address mapping on the fly.
// 3x3 / stride 1 / pad 1 layers with channel counts TMA accepts go to the tcgen05 kernels (tc_conv.cu).
//
// Reference semantics followed (src/tensor_ops/conv_ops/):
//   conv2d_extract_params + yh/yw formula          conv2d.rs:346-404
//   Conv2D: y[b] = W[O, C*kh*kw] . im2col(x[b])    conv2d.rs:115-211, mod.rs:73-124
//   Conv2DTranspose: cols = W^T . gy[b]; col2im     conv2d_transpose.rs:89-247, mod.rs:178-223
//     xh = s(yh-1) - 2p + d(kh-1) + 1                conv2d_transpose.rs:55-56
//   Conv2DFilterGrad: gw = sum_b gy[b] . cols[b]^T   conv2d.rs:631-734
//   Conv2DTransposeFilterGrad (roles swapped)        conv2d_transpose.rs:303-431
//   quirk: im2col uses `ph` for the x start (mod.rs:98); pads are always square so it is unobservable.
#include "simt_gemm.cuh"

// tcgen05 kernels (tc_conv.cu): operate on channels-last activation buffers
int agb_tc_conv_fprop(agb_ctx* ctx, int mode, const float* x, const float* w, float* y,
                      int B, int C, int H, int W, int O, int kh, int kw, int pad, int stride, int dil, int flip_transpose, const float* bias, int relu, const float* mask, float* csum, float* pool_y, int* pool_idx);
int agb_tc_conv_dgrad_strided(agb_ctx* ctx, int mode, const float* gy, const float* w, float* gx, int B, int O, int yh, int yw, int C, int H, int W,
                              int kh, int kw, int pad, int stride, int dil, const float* mask, float* csum);
int agb_tc_conv_wgrad(agb_ctx* ctx, int mode, const float* img, const float* g, float* gw,
                      int B, int C, int H, int W, int O, int kh, int kw, int pad, int stride, int dil);
bool agb_tc_conv_eligible(int C, int O, int kh, int kw, int stride, int yw);
bool agb_tc_conv_fprop_eligible(int C, int O, int kh, int kw, int stride, int yw);
// direct kernels for very small input-channel counts (conv_small_c.cu)
bool agb_small_c_eligible(int C, int O, int kh, int kw);
int agb_small_c_fprop(agb_ctx* ctx, const float* x, const float* w, agb_tensor* y, int B, int C, int H, int W, int O, int kh, int kw, int yh, int yw, int pad, int stride, int dil, const float* bias, int relu);
int agb_tc_conv_first(agb_ctx* ctx, int mode, const float* x, const float* w, float* y, int B, int C, int H, int W, int O, int kh, int kw, int yh, int yw,
                      int pad, int stride, int dil, const float* bias, int relu);
int agb_small_c_wgrad(agb_ctx* ctx, const float* x, const agb_tensor* gy, float* gw, int B, int C, int H, int W, int O, int kh, int kw, int yh, int yw, int pad, int stride, int dil);

// ---- activation layouts.  A logical [B,C,H,W] tensor is accepted in two dense memory orders: NCHW (C-contiguous, the
// reference's layout) and channels-last (N,H,W,C).  Each kernel family has a native order (tcgen05: channels-last, CUDA-core
// implicit GEMM: NCHW); operands in the other order go through one tiled transpose (agb_copy_strided) into an arena temporary.
static bool is_nchw(const agb_tensor* t) { return agb_is_contig(t); }
static bool is_channels_last(const agb_tensor* t) {
  if (t->rank != 4) return false;
  const int64_t C = t->shape[1], H = t->shape[2], W = t->shape[3];
  return (C == 1 || t->stride[1] == 1) && (W == 1 || t->stride[3] == C) && (H == 1 || t->stride[2] == W * C) && (t->shape[0] == 1 || t->stride[0] == H * W * C);
}
static void set_layout(agb_tensor* t, bool channels_last) {
  const int64_t C = t->shape[1], H = t->shape[2], W = t->shape[3];
  if (channels_last) { t->stride[0] = H * W * C; t->stride[1] = 1; t->stride[2] = W * C; t->stride[3] = C; }
  else { t->stride[0] = C * H * W; t->stride[1] = H * W; t->stride[2] = W; t->stride[3] = 1; }
}
struct LayoutTmp {     // an input converted on entry, or an output converted back on exit
  agb_ctx* ctx; float* tmp = nullptr; agb_tensor view; const agb_tensor* user = nullptr; bool copy_back = false;
  explicit LayoutTmp(agb_ctx* c) : ctx(c) {}
  int input(const agb_tensor* t, bool want_cl) {
    view = *t;
    if (want_cl ? is_channels_last(t) : is_nchw(t)) return AGB_OK;
    AGB_CHECK(is_nchw(t) || is_channels_last(t), AGB_ERR_UNSUPPORTED, "conv: activations must be dense NCHW or channels-last");
    AGB_TRY(agb_alloc(ctx, (size_t)agb_numel(t) * sizeof(float), (void**)&tmp));
    view.ptr = tmp; set_layout(&view, want_cl);
    return agb_copy_strided(ctx, t, &view);
  }
  int output(agb_tensor* t, bool want_cl) {
    view = *t; user = t;
    if (want_cl ? is_channels_last(t) : is_nchw(t)) return AGB_OK;
    AGB_CHECK(is_nchw(t) || is_channels_last(t), AGB_ERR_UNSUPPORTED, "conv: activations must be dense NCHW or channels-last");
    AGB_TRY(agb_alloc(ctx, (size_t)agb_numel(t) * sizeof(float), (void**)&tmp));
    view.ptr = tmp; set_layout(&view, want_cl); copy_back = true;
    return AGB_OK;
  }
  int finish() {
    int r = AGB_OK;
    if (copy_back) { agb_tensor dst = *user; r = agb_copy_strided(ctx, &view, &dst); copy_back = false; }
    if (tmp) { agb_free(ctx, tmp); tmp = nullptr; }
    return r;
  }
  ~LayoutTmp() { if (tmp) agb_free(ctx, tmp); }
};

struct ConvGeom { int B, C, H, W, O, kh, kw, yh, yw, pad, stride, dil; };

// ---- fprop ----
struct FpropA { const float* w; int64_t K; static const bool K_CONTIG = true;
  __device__ __forceinline__ float load(int, int64_t m, int64_t k) const { return __ldg(w + m * K + k); } };
struct FpropB { const float* x; ConvGeom g; static const bool K_CONTIG = false;
  __device__ __forceinline__ float load(int, int64_t k, int64_t n) const {
    int kk = g.kh * g.kw; int c = (int)(k / kk); int r = (int)(k - (int64_t)c * kk); int i = r / g.kw, j = r - i * g.kw;
    int P = g.yh * g.yw; int b = (int)(n / P); int pix = (int)(n - (int64_t)b * P); int oy = pix / g.yw, ox = pix - oy * g.yw;
    int iy = oy * g.stride - g.pad + i * g.dil, ix = ox * g.stride - g.pad + j * g.dil;
    if ((unsigned)iy >= (unsigned)g.H || (unsigned)ix >= (unsigned)g.W) return 0.0f;
    return __ldg(x + (((int64_t)b * g.C + c) * g.H + iy) * g.W + ix);
  } };
struct FpropC { float* y; ConvGeom g;
  __device__ __forceinline__ void store(int, int64_t m, int64_t n, float v) const {
    int P = g.yh * g.yw; int64_t b = n / P; int64_t pix = n - b * P;
    y[(b * g.O + m) * P + pix] = v;
  } };

// ---- dgrad: gx[b,c,iy,ix] = sum_{o,i,j} W[o,c,i,j] * gy[b,o,oy,ox],  oy*s - p + i*d == iy ----
struct DgradA { const float* w; ConvGeom g; static const bool K_CONTIG = true;
  __device__ __forceinline__ float load(int, int64_t m, int64_t k) const {
    int kk = g.kh * g.kw; int o = (int)(k / kk); int r = (int)(k - (int64_t)o * kk);
    return __ldg(w + ((int64_t)o * g.C + m) * kk + r);
  } };
struct DgradB { const float* gy; ConvGeom g; static const bool K_CONTIG = false;
  __device__ __forceinline__ float load(int, int64_t k, int64_t n) const {
    int kk = g.kh * g.kw; int o = (int)(k / kk); int r = (int)(k - (int64_t)o * kk); int i = r / g.kw, j = r - i * g.kw;
    int P = g.H * g.W; int b = (int)(n / P); int pix = (int)(n - (int64_t)b * P); int iy = pix / g.W, ix = pix - iy * g.W;
    int ty = iy + g.pad - i * g.dil, tx = ix + g.pad - j * g.dil;
    if (ty < 0 || tx < 0) return 0.0f;
    int oy = ty / g.stride, ox = tx / g.stride;
    if (oy * g.stride != ty || ox * g.stride != tx || oy >= g.yh || ox >= g.yw) return 0.0f;
    return __ldg(gy + (((int64_t)b * g.O + o) * g.yh + oy) * g.yw + ox);
  } };
struct DgradC { float* gx; ConvGeom g;
  __device__ __forceinline__ void store(int, int64_t m, int64_t n, float v) const {
    int P = g.H * g.W; int64_t b = n / P; int64_t pix = n - b * P;
    gx[(b * g.C + m) * P + pix] = v;
  } };

// ---- wgrad: gw[o,(c,i,j)] = sum_{b,oy,ox} gr[b,o,oy,ox] * img[b,c,oy*s-p+i*d, ox*s-p+j*d]; split over batch ----
struct WgradA { const float* gr; ConvGeom g; int bchunk; static const bool K_CONTIG = true;
  __device__ __forceinline__ float load(int z, int64_t m, int64_t k) const {
    int P = g.yh * g.yw; int bl = (int)(k / P); int pix = (int)(k - (int64_t)bl * P); int b = z * bchunk + bl;
    if (b >= g.B) return 0.0f;
    return __ldg(gr + ((int64_t)b * g.O + m) * P + pix);
  } };
struct WgradB { const float* img; ConvGeom g; int bchunk; static const bool K_CONTIG = true;
  __device__ __forceinline__ float load(int z, int64_t k, int64_t n) const {
    int P = g.yh * g.yw; int bl = (int)(k / P); int pix = (int)(k - (int64_t)bl * P); int b = z * bchunk + bl;
    if (b >= g.B) return 0.0f;
    int oy = pix / g.yw, ox = pix - oy * g.yw;
    int kk = g.kh * g.kw; int c = (int)(n / kk); int r = (int)(n - (int64_t)c * kk); int i = r / g.kw, j = r - i * g.kw;
    int iy = oy * g.stride - g.pad + i * g.dil, ix = ox * g.stride - g.pad + j * g.dil;
    if ((unsigned)iy >= (unsigned)g.H || (unsigned)ix >= (unsigned)g.W) return 0.0f;
    return __ldg(img + (((int64_t)b * g.C + c) * g.H + iy) * g.W + ix);
  } };
struct WgradC { float* gw; int64_t N; int atomic;
  __device__ __forceinline__ void store(int, int64_t m, int64_t n, float v) const {
    if (atomic) atomicAdd(gw + m * N + n, v); else gw[m * N + n] = v;
  } };

static int check_geom(const char* who, const agb_tensor* x, const agb_tensor* w, int pad, int stride, int dil, ConvGeom& g) {
  AGB_CHECK(x->rank == 4, AGB_ERR_INCOMPATIBLE_SHAPE, "%s: lhs input must be 4D (got rank %d)", who, x->rank);
  AGB_CHECK(w->rank == 4, AGB_ERR_INCOMPATIBLE_SHAPE, "%s: filter must be 4D (got rank %d)", who, w->rank);
  AGB_CHECK(stride >= 1 && dil >= 1 && pad >= 0, AGB_ERR_INVALID_DIMS, "%s: stride/dilation must be >= 1, pad >= 0", who);
  g.B = (int)x->shape[0]; g.C = (int)x->shape[1]; g.H = (int)x->shape[2]; g.W = (int)x->shape[3];
  g.O = (int)w->shape[0]; g.kh = (int)w->shape[2]; g.kw = (int)w->shape[3];
  g.pad = pad; g.stride = stride; g.dil = dil;
  AGB_CHECK(x->shape[1] == w->shape[1], AGB_ERR_INCOMPATIBLE_SHAPE, "%s: input channel dim (%lld) must match filter's second dim (%lld)", who,
            (long long)x->shape[1], (long long)w->shape[1]);
  int eh = dil * (g.kh - 1) + 1, ew = dil * (g.kw - 1) + 1;
  AGB_CHECK(g.H + 2 * pad >= eh && g.W + 2 * pad >= ew, AGB_ERR_INCOMPATIBLE_SHAPE, "%s: kernel larger than padded input", who);
  g.yh = (g.H + 2 * pad - eh) / stride + 1; g.yw = (g.W + 2 * pad - ew) / stride + 1;
  return AGB_OK;
}

extern "C" int agb_conv2d_fprop_f32(agb_ctx* ctx, const agb_tensor* x, const agb_tensor* w, agb_tensor* y, int pad, int stride, int dilation) {
  return agb_conv2d_fprop_fused_f32(ctx, x, w, nullptr, 0, y, pad, stride, dilation);
}

// y = [relu](conv(x, w) [+ bias[o]]): the fused form of Conv2D -> AddOp(bias [1,O,1,1]) -> ReLU (examples/cnn_mnist.rs:38-45)
static int fprop_fused_impl(agb_ctx* ctx, const agb_tensor* x, const agb_tensor* w, const float* bias, int relu, agb_tensor* y, int pad, int stride, int dilation);
extern "C" int agb_conv2d_fprop_fused_f32(agb_ctx* ctx, const agb_tensor* x, const agb_tensor* w, const float* bias, int relu, agb_tensor* y,
                                         int pad, int stride, int dilation) {
  return fprop_fused_impl(ctx, x, w, bias, relu, y, pad, stride, dilation);
}
// the same with the SIGN BITS of the stored activation as a side output (1 bit per element in channels-last order: word (pixel * O + o) / 32,
// bit o % 32): the mask of the ReLU backward, 1/32 of the bytes of the activation it would otherwise be read from.  *bits_written = 1 when the
// kernel that ran wrote them (channels-last y, O % 32 == 0, a tensor-core kernel), else the buffer is untouched.
extern "C" int agb_conv2d_fprop_fused_bits_f32(agb_ctx* ctx, const agb_tensor* x, const agb_tensor* w, const float* bias, int relu, agb_tensor* y,
                                              uint32_t* relu_bits, int* bits_written, int pad, int stride, int dilation) {
  const bool ok = relu_bits != nullptr && y->rank == 4 && y->shape[1] % 32 == 0 && is_channels_last(y) && !is_nchw(y) && ((((uintptr_t)relu_bits) & 3) == 0);
  ctx->bits_out = ok ? relu_bits : nullptr; ctx->bits_written = 0;
  const int r = fprop_fused_impl(ctx, x, w, bias, relu, y, pad, stride, dilation);
  if (bits_written) *bits_written = (r == AGB_OK) ? ctx->bits_written : 0;
  ctx->bits_out = nullptr; ctx->bits_written = 0;
  return r;
}
static int fprop_fused_impl(agb_ctx* ctx, const agb_tensor* x, const agb_tensor* w, const float* bias, int relu, agb_tensor* y, int pad, int stride, int dilation) {
  ConvGeom g; AGB_TRY(check_geom("conv2d", x, w, pad, stride, dilation, g));
  AGB_CHECK(agb_is_contig(w), AGB_ERR_UNSUPPORTED, "conv2d: the filter must be C-contiguous");
  AGB_CHECK(y->rank == 4 && y->shape[0] == g.B && y->shape[1] == g.O && y->shape[2] == g.yh && y->shape[3] == g.yw, AGB_ERR_INCOMPATIBLE_SHAPE,
            "conv2d: output must be [%d,%d,%d,%d]", g.B, g.O, g.yh, g.yw);
  if (agb_numel(y) == 0) return AGB_OK;
  AgbProfScope prof(ctx, AGB_PROF_CONV_FPROP, 2.0 * (double)agb_numel(y) * g.C * g.kh * g.kw);
  if (ctx->math_mode != AGB_MATH_FP32 && agb_tc_conv_fprop_eligible(g.C, g.O, g.kh, g.kw, stride, g.yw)) {
    LayoutTmp lx(ctx), ly(ctx);
    AGB_TRY(lx.input(x, true)); AGB_TRY(ly.output(y, true));
    int r = agb_tc_conv_fprop(ctx, ctx->math_mode, lx.view.ptr, w->ptr, ly.view.ptr, g.B, g.C, g.H, g.W, g.O, g.kh, g.kw, pad, stride, dilation, 0, bias, relu, nullptr, nullptr, nullptr, nullptr);
    if (r == AGB_OK) { AGB_TRY(ly.finish()); return lx.finish(); }
    if (r != AGB_ERR_UNSUPPORTED) return r;
  }
  if (agb_small_c_eligible(g.C, g.O, g.kh, g.kw) && (is_nchw(y) || is_channels_last(y))) {
    LayoutTmp lx(ctx); AGB_TRY(lx.input(x, false));
    if (is_channels_last(y) && !is_nchw(y)) {        // im2col tile built in shared memory + tcgen05 (tc_conv_first.cu): the first layer at the HBM rate
      int r1 = agb_tc_conv_first(ctx, ctx->math_mode, lx.view.ptr, w->ptr, y->ptr, g.B, g.C, g.H, g.W, g.O, g.kh, g.kw, g.yh, g.yw, pad, stride, dilation, bias, relu);
      if (r1 == AGB_OK) { prof.set_cls(AGB_PROF_CONV_SMALLC_FPROP); return lx.finish(); }
      if (r1 != AGB_ERR_UNSUPPORTED) return r1;
    }
    int r = agb_small_c_fprop(ctx, lx.view.ptr, w->ptr, y, g.B, g.C, g.H, g.W, g.O, g.kh, g.kw, g.yh, g.yw, pad, stride, dilation, bias, relu);
    if (r == AGB_OK) { prof.set_cls(AGB_PROF_CONV_SMALLC_FPROP); return lx.finish(); }
    if (r != AGB_ERR_UNSUPPORTED) return r;
  }
  prof.set_cls(AGB_PROF_CONV_SIMT);
  LayoutTmp lx(ctx), ly(ctx);
  AGB_TRY(lx.input(x, false)); AGB_TRY(ly.output(y, false));
  int64_t K = (int64_t)g.C * g.kh * g.kw;
  AGB_TRY(simt_gemm_launch(ctx, FpropA{w->ptr, K}, FpropB{lx.view.ptr, g}, FpropC{ly.view.ptr, g}, g.O, (int64_t)g.B * g.yh * g.yw, K, 1));
  if (bias != nullptr || relu) {       // CUDA-core path: epilogue as in-place elementwise passes over the NCHW result
    agb_tensor yv = ly.view;
    if (bias != nullptr) {
      agb_tensor bt; bt.ptr = const_cast<float*>(bias); bt.rank = 4;
      for (int i = 0; i < 4; i++) { bt.shape[i] = yv.shape[i]; bt.stride[i] = 0; }
      bt.stride[1] = 1;
      AGB_TRY(agb_binary(ctx, AGB_B_ADD, 0.f, 0.f, &yv, &bt, &yv));
    }
    if (relu) AGB_TRY(agb_unary(ctx, AGB_U_RELU, 0.f, 0.f, &yv, &yv));
  }
  AGB_TRY(ly.finish()); return lx.finish();
}

// y_pooled = max_pool2d([relu](conv(x, w) [+ bias]), size 2, pad 0, stride 2) with the pooling done in the conv epilogue: the full-size
// activation is never written.  Only the tensor-core wide-map kernel has this epilogue; everything else answers AGB_ERR_UNSUPPORTED
// BEFORE launching anything, and the caller runs conv and pool separately.
extern "C" int agb_conv2d_fprop_pool_f32(agb_ctx* ctx, const agb_tensor* x, const agb_tensor* w, const float* bias, int relu, agb_tensor* y_pooled,
                                        int32_t* idx_i32, int pad, int stride, int dilation) {
  ConvGeom g; AGB_TRY(check_geom("conv2d", x, w, pad, stride, dilation, g));
  AGB_CHECK(agb_is_contig(w), AGB_ERR_UNSUPPORTED, "conv2d: the filter must be C-contiguous");
  const int ph = g.yh / 2, pw = g.yw / 2;
  AGB_CHECK(y_pooled->rank == 4 && y_pooled->shape[0] == g.B && y_pooled->shape[1] == g.O && y_pooled->shape[2] == ph && y_pooled->shape[3] == pw,
            AGB_ERR_INCOMPATIBLE_SHAPE, "conv2d+max_pool2d: pooled output must be [%d,%d,%d,%d]", g.B, g.O, ph, pw);
  if (ctx->math_mode != AGB_MATH_TF32 || stride != 1 || g.yw < 128 || g.O > 128 || ph < 1 || pw < 1 || idx_i32 == nullptr ||
      !agb_tc_conv_eligible(g.C, g.O, g.kh, g.kw, stride, g.yw) || !is_channels_last(x) || !is_channels_last(y_pooled) || agb_is_contig(y_pooled) ||
      (bias != nullptr && (((uintptr_t)bias) & 15) != 0))
    return AGB_ERR_UNSUPPORTED;
  AgbProfScope prof(ctx, AGB_PROF_CONV_FPROP, 2.0 * (double)g.B * g.O * g.yh * g.yw * g.C * g.kh * g.kw);
  return agb_tc_conv_fprop(ctx, ctx->math_mode, x->ptr, w->ptr, nullptr, g.B, g.C, g.H, g.W, g.O, g.kh, g.kw, pad, stride, dilation, 0, bias, relu, nullptr, nullptr,
                           y_pooled->ptr, idx_i32);
}

extern "C" int agb_conv2d_dgrad_f32(agb_ctx* ctx, const agb_tensor* gy, const agb_tensor* w, agb_tensor* gx, int pad, int stride, int dilation) {
  return agb_conv2d_dgrad_fused_f32(ctx, gy, w, nullptr, nullptr, gx, pad, stride, dilation);
}

// chan_sum[c] = sum over (b, h, w) of t[b, c, h, w] for a dense NCHW or channels-last tensor (un-fused form of the side output)
static int channel_sums(agb_ctx* ctx, const agb_tensor* t, float* chan_sum) {
  const int64_t B = t->shape[0], C = t->shape[1], HW = t->shape[2] * t->shape[3];
  if (!agb_is_contig(t)) return agb_reduce(ctx, AGB_R_SUM, t->ptr, chan_sum, 1, B * HW, C);          // channels-last: [pixels, C]
  float* tmp; AGB_TRY(agb_alloc(ctx, (size_t)(B * C) * sizeof(float), (void**)&tmp));
  int r = agb_reduce(ctx, AGB_R_SUM, t->ptr, tmp, B * C, HW, 1);
  if (r == AGB_OK) r = agb_reduce(ctx, AGB_R_SUM, tmp, chan_sum, 1, B, C);
  agb_free(ctx, tmp);
  return r;
}

// un-fused tail of the fused entry point: gx = (mask_src > 0) * gx in place (AGB_B_RELU_GRAD); gx is dense NCHW or channels-last
static int apply_relu_mask(agb_ctx* ctx, const agb_tensor* mask_src, agb_tensor* gx) {
  if (mask_src == nullptr) return AGB_OK;
  if (agb_is_contig(gx)) return agb_binary(ctx, AGB_B_RELU_GRAD, 0.f, 0.f, mask_src, gx, gx);
  // channels-last gx: elementwise over raw memory; a mask in another memory order is first brought into gx's order
  bool same = true;
  for (int i = 0; i < 4; i++) if (gx->shape[i] != 1 && mask_src->stride[i] != gx->stride[i]) same = false;
  const int64_t n = agb_numel(gx);
  float* tmp = nullptr; const float* mptr = mask_src->ptr;
  if (!same) {
    AGB_TRY(agb_alloc(ctx, (size_t)n * sizeof(float), (void**)&tmp));
    agb_tensor tv = *gx; tv.ptr = tmp;
    int r = agb_copy_strided(ctx, mask_src, &tv);
    if (r != AGB_OK) { agb_free(ctx, tmp); return r; }
    mptr = tmp;
  }
  agb_tensor fm, fg; fm.ptr = const_cast<float*>(mptr); fg.ptr = gx->ptr; fm.rank = fg.rank = 1; fm.shape[0] = fg.shape[0] = n; fm.stride[0] = fg.stride[0] = 1;
  int r = agb_binary(ctx, AGB_B_RELU_GRAD, 0.f, 0.f, &fm, &fg, &fg);
  if (tmp) agb_free(ctx, tmp);
  return r;
}

static int dgrad_fused_impl(agb_ctx* ctx, const agb_tensor* gy, const agb_tensor* w, const agb_tensor* mask_src, float* chan_sum, agb_tensor* gx, int pad, int stride, int dilation);
extern "C" int agb_conv2d_dgrad_fused_f32(agb_ctx* ctx, const agb_tensor* gy, const agb_tensor* w, const agb_tensor* mask_src, float* chan_sum,
                                         agb_tensor* gx, int pad, int stride, int dilation) {
  return dgrad_fused_impl(ctx, gy, w, mask_src, chan_sum, gx, pad, stride, dilation);
}
// the same with the sign bits of mask_src (agb_conv2d_fprop_fused_bits_f32) next to it: kernels that fuse the mask read 4 bytes per pixel per 32
// channels instead of 128; every other path reads mask_src as before.  mask_bits must describe exactly mask_src (same buffer, channels-last).
extern "C" int agb_conv2d_dgrad_fused_bits_f32(agb_ctx* ctx, const agb_tensor* gy, const agb_tensor* w, const agb_tensor* mask_src, const uint32_t* mask_bits,
                                              float* chan_sum, agb_tensor* gx, int pad, int stride, int dilation) {
  const bool ok = mask_bits != nullptr && mask_src != nullptr && mask_src->rank == 4 && mask_src->shape[1] % 32 == 0 && is_channels_last(mask_src) && !is_nchw(mask_src);
  ctx->mask_bits = ok ? mask_bits : nullptr; ctx->mask_bits_used = 0;
  const int r = dgrad_fused_impl(ctx, gy, w, mask_src, chan_sum, gx, pad, stride, dilation);
  ctx->mask_bits = nullptr;
  return r;
}
static int dgrad_fused_impl(agb_ctx* ctx, const agb_tensor* gy, const agb_tensor* w, const agb_tensor* mask_src, float* chan_sum, agb_tensor* gx, int pad, int stride, int dilation) {
  AGB_CHECK(gy->rank == 4, AGB_ERR_INCOMPATIBLE_SHAPE, "conv2d_transpose: Input must be 4D (got rank %d)", gy->rank);
  AGB_CHECK(w->rank == 4, AGB_ERR_INCOMPATIBLE_SHAPE, "conv2d_transpose: Filter must be 4D (got rank %d)", w->rank);
  AGB_CHECK(gy->shape[1] == w->shape[0], AGB_ERR_INCOMPATIBLE_SHAPE,
            "conv2d_transpose: Number of input channels (%lld) must match second filter dim (%lld)", (long long)gy->shape[1], (long long)w->shape[0]);
  AGB_CHECK(stride >= 1 && dilation >= 1 && pad >= 0, AGB_ERR_INVALID_DIMS, "conv2d_transpose: bad pad/stride/dilation");
  ConvGeom g;
  g.B = (int)gy->shape[0]; g.O = (int)gy->shape[1]; g.yh = (int)gy->shape[2]; g.yw = (int)gy->shape[3];
  g.C = (int)w->shape[1]; g.kh = (int)w->shape[2]; g.kw = (int)w->shape[3]; g.pad = pad; g.stride = stride; g.dil = dilation;
  g.H = stride * (g.yh - 1) - 2 * pad + (dilation * (g.kh - 1) + 1);
  g.W = stride * (g.yw - 1) - 2 * pad + (dilation * (g.kw - 1) + 1);
  AGB_CHECK(g.H > 0 && g.W > 0, AGB_ERR_INCOMPATIBLE_SHAPE, "conv2d_transpose: non-positive output size");
  AGB_CHECK(gx->rank == 4 && gx->shape[0] == g.B && gx->shape[1] == g.C && gx->shape[2] == g.H && gx->shape[3] == g.W, AGB_ERR_INCOMPATIBLE_SHAPE,
            "conv2d_transpose: output must be [%d,%d,%d,%d]", g.B, g.C, g.H, g.W);
  AGB_CHECK(agb_is_contig(w), AGB_ERR_UNSUPPORTED, "conv2d_transpose: the filter must be C-contiguous");
  if (mask_src != nullptr) {
    AGB_CHECK(mask_src->rank == 4, AGB_ERR_INCOMPATIBLE_SHAPE, "conv2d_transpose: mask_src must be 4-D");
    for (int i = 0; i < 4; i++)
      AGB_CHECK(mask_src->shape[i] == gx->shape[i], AGB_ERR_INCOMPATIBLE_SHAPE, "conv2d_transpose: mask_src must have the shape of the output");
  }
  bool same_strides = mask_src != nullptr;
  if (mask_src != nullptr) for (int i = 0; i < 4; i++) if (gx->shape[i] != 1 && mask_src->stride[i] != gx->stride[i]) same_strides = false;
  if (agb_numel(gx) == 0) return AGB_OK;
  AgbProfScope prof(ctx, AGB_PROF_CONV_DGRAD, 2.0 * (double)agb_numel(gy) * g.C * g.kh * g.kw);
  if (ctx->math_mode != AGB_MATH_FP32 && stride == 1 && dilation * (g.kh - 1) - pad >= 0 && agb_tc_conv_eligible(g.O, g.C, g.kh, g.kw, stride, g.W)) {
    // stride-1 dgrad == fprop of gy with the spatially flipped, channel-transposed filter and pad' = d(k-1) - p
    LayoutTmp lg(ctx), lx(ctx);
    AGB_TRY(lg.input(gy, true)); AGB_TRY(lx.output(gx, true));
    const bool in_place = lx.view.ptr == gx->ptr;                                                               // output written channels-last, no conversion
    const bool fuse = same_strides && in_place && ((((uintptr_t)mask_src->ptr) & 15) == 0);
    const bool fuse_sum = chan_sum != nullptr && in_place && (mask_src == nullptr || fuse);                     // sums of the FINAL (masked) values
    if (fuse_sum) AGB_TRY(agb_memset0(ctx, chan_sum, (size_t)g.C * sizeof(float)));
    int r = agb_tc_conv_fprop(ctx, ctx->math_mode, lg.view.ptr, w->ptr, lx.view.ptr, g.B, g.O, g.yh, g.yw, g.C, g.kh, g.kw, pad, stride, dilation, 1, nullptr, 0,
                              fuse ? mask_src->ptr : nullptr, fuse_sum ? chan_sum : nullptr, nullptr, nullptr);
    if (r == AGB_OK) {
      AGB_TRY(lx.finish()); AGB_TRY(lg.finish());
      if (!fuse) AGB_TRY(apply_relu_mask(ctx, mask_src, gx));
      return (chan_sum != nullptr && !fuse_sum) ? channel_sums(ctx, gx, chan_sum) : AGB_OK;
    }
    if (r != AGB_ERR_UNSUPPORTED) return r;
  }
  prof.set_cls(AGB_PROF_CONV_SIMT);
  if (ctx->math_mode == AGB_MATH_TF32 && stride > 1) {       // strided dgrad: s*s unit-stride phase convolutions on the tensor cores
    LayoutTmp lg(ctx), lx(ctx);
    AGB_TRY(lg.input(gy, true)); AGB_TRY(lx.output(gx, true));
    const bool in_place = lx.view.ptr == gx->ptr;
    const bool fuse = same_strides && in_place && ((((uintptr_t)mask_src->ptr) & 15) == 0);
    const bool fuse_sum = chan_sum != nullptr && in_place && (mask_src == nullptr || fuse);
    if (fuse_sum) AGB_TRY(agb_memset0(ctx, chan_sum, (size_t)g.C * sizeof(float)));
    int r = agb_tc_conv_dgrad_strided(ctx, ctx->math_mode, lg.view.ptr, w->ptr, lx.view.ptr, g.B, g.O, g.yh, g.yw, g.C, g.H, g.W, g.kh, g.kw, pad, stride, dilation,
                                      fuse ? mask_src->ptr : nullptr, fuse_sum ? chan_sum : nullptr);
    if (r == AGB_OK) {
      AGB_TRY(lx.finish()); AGB_TRY(lg.finish());
      if (!fuse) AGB_TRY(apply_relu_mask(ctx, mask_src, gx));
      return (chan_sum != nullptr && !fuse_sum) ? channel_sums(ctx, gx, chan_sum) : AGB_OK;
    }
    if (r != AGB_ERR_UNSUPPORTED) return r;
  }
  LayoutTmp lg(ctx), lx(ctx);
  AGB_TRY(lg.input(gy, false)); AGB_TRY(lx.output(gx, false));
  int64_t K = (int64_t)g.O * g.kh * g.kw;
  AGB_TRY(simt_gemm_launch(ctx, DgradA{w->ptr, g}, DgradB{lg.view.ptr, g}, DgradC{lx.view.ptr, g}, g.C, (int64_t)g.B * g.H * g.W, K, 1));
  AGB_TRY(lx.finish()); AGB_TRY(lg.finish());
  AGB_TRY(apply_relu_mask(ctx, mask_src, gx));
  return chan_sum != nullptr ? channel_sums(ctx, gx, chan_sum) : AGB_OK;
}

extern "C" int agb_conv2d_wgrad_f32(agb_ctx* ctx, const agb_tensor* img, const agb_tensor* gr, agb_tensor* gw, int pad, int stride, int dilation) {
  ConvGeom g; AGB_TRY(check_geom("conv2d_filter_grad", img, gw, pad, stride, dilation, g));
  AGB_CHECK(gr->rank == 4 && gr->shape[0] == g.B && gr->shape[1] == g.O && gr->shape[2] == g.yh && gr->shape[3] == g.yw, AGB_ERR_INCOMPATIBLE_SHAPE,
            "conv2d_filter_grad: gradient must be [%d,%d,%d,%d]", g.B, g.O, g.yh, g.yw);
  AGB_CHECK(agb_is_contig(gw), AGB_ERR_UNSUPPORTED, "conv2d_filter_grad: the filter gradient must be C-contiguous");
  if (agb_numel(gw) == 0) return AGB_OK;
  AgbProfScope prof(ctx, AGB_PROF_CONV_WGRAD, 2.0 * (double)agb_numel(gr) * g.C * g.kh * g.kw);
  if (ctx->math_mode != AGB_MATH_FP32 && agb_tc_conv_fprop_eligible(g.C, g.O, g.kh, g.kw, stride, g.yw)) {       // strides 1..4
    LayoutTmp li(ctx), lg(ctx);
    AGB_TRY(li.input(img, true)); AGB_TRY(lg.input(gr, true));
    int r = agb_tc_conv_wgrad(ctx, ctx->math_mode, li.view.ptr, lg.view.ptr, gw->ptr, g.B, g.C, g.H, g.W, g.O, g.kh, g.kw, pad, stride, dilation);
    if (r == AGB_OK) { AGB_TRY(li.finish()); return lg.finish(); }
    if (r != AGB_ERR_UNSUPPORTED) return r;
  }
  if (agb_small_c_eligible(g.C, g.O, g.kh, g.kw) && (is_nchw(gr) || is_channels_last(gr))) {
    LayoutTmp li(ctx); AGB_TRY(li.input(img, false));
    int r = agb_small_c_wgrad(ctx, li.view.ptr, gr, gw->ptr, g.B, g.C, g.H, g.W, g.O, g.kh, g.kw, g.yh, g.yw, pad, stride, dilation);
    if (r == AGB_OK) { prof.set_cls(AGB_PROF_CONV_SMALLC_WGRAD); return li.finish(); }
    if (r != AGB_ERR_UNSUPPORTED) return r;
  }
  prof.set_cls(AGB_PROF_CONV_SIMT);
  LayoutTmp li(ctx), lg(ctx);
  AGB_TRY(li.input(img, false)); AGB_TRY(lg.input(gr, false));
  int64_t N = (int64_t)g.C * g.kh * g.kw; int64_t P = (int64_t)g.yh * g.yw;
  bool big = (g.O >= 96 && N >= 96); int tile = big ? 128 : 64;
  int64_t tiles = ((g.O + tile - 1) / tile) * ((N + tile - 1) / tile);
  int64_t Z = (2 * (int64_t)ctx->sm_count + tiles - 1) / tiles; if (Z > g.B) Z = g.B; if (Z < 1) Z = 1;
  int bchunk = (int)((g.B + Z - 1) / Z); Z = (g.B + bchunk - 1) / bchunk;
  if (Z > 1) AGB_TRY(agb_memset0(ctx, gw->ptr, agb_numel(gw) * sizeof(float)));
  AGB_TRY(simt_gemm_launch(ctx, WgradA{lg.view.ptr, g, bchunk}, WgradB{li.view.ptr, g, bchunk}, WgradC{gw->ptr, N, Z > 1}, g.O, N, (int64_t)bchunk * P, Z));
  AGB_TRY(li.finish()); return lg.finish();
}

extern "C" int agb_conv_prefers_channels_last(int in_channels, int out_channels, int kh, int kw, int stride, int out_w) {
  if (agb_tc_conv_fprop_eligible(in_channels, out_channels, kh, kw, stride, out_w)) return 1;
  return (agb_small_c_eligible(in_channels, out_channels, kh, kw) && out_channels >= 32 && out_channels % 4 == 0 && out_w >= 16) ? 1 : 0;
}

// ---- im2col materialisation (only for user-visible evaluation of Conv2D's 2nd output) ----
__global__ void __launch_bounds__(256) im2col_kernel(const float* __restrict__ x, float* __restrict__ cols, ConvGeom g, int64_t n) {
  int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t o = tid; o < n; o += gs) {
    // cols [B, C, kh, kw, yh, yw]  (conv_ops/mod.rs:93-121: for c, kh, kw, yh, yw)
    int ox = (int)(o % g.yw); int64_t t = o / g.yw; int oy = (int)(t % g.yh); t /= g.yh;
    int j = (int)(t % g.kw); t /= g.kw; int i = (int)(t % g.kh); t /= g.kh; int c = (int)(t % g.C); int64_t b = t / g.C;
    int iy = oy * g.stride - g.pad + i * g.dil, ix = ox * g.stride - g.pad + j * g.dil;
    float v = 0.0f;
    if ((unsigned)iy < (unsigned)g.H && (unsigned)ix < (unsigned)g.W) v = __ldg(x + ((b * g.C + c) * g.H + iy) * g.W + ix);
    cols[o] = v;
  }
}
extern "C" int agb_im2col_f32(agb_ctx* ctx, const agb_tensor* x, agb_tensor* cols, int kh, int kw, int pad, int stride, int dilation) {
  AGB_CHECK(x->rank == 4 && cols->rank == 6, AGB_ERR_INCOMPATIBLE_SHAPE, "im2col: x must be 4D and cols 6D");
  AGB_CHECK(agb_is_contig(x) && agb_is_contig(cols), AGB_ERR_UNSUPPORTED, "im2col: tensors must be contiguous");
  ConvGeom g; g.B = (int)x->shape[0]; g.C = (int)x->shape[1]; g.H = (int)x->shape[2]; g.W = (int)x->shape[3];
  g.O = 0; g.kh = kh; g.kw = kw; g.pad = pad; g.stride = stride; g.dil = dilation;
  g.yh = (g.H + 2 * pad - (dilation * (kh - 1) + 1)) / stride + 1; g.yw = (g.W + 2 * pad - (dilation * (kw - 1) + 1)) / stride + 1;
  int64_t n = (int64_t)g.B * g.C * kh * kw * g.yh * g.yw;
  AGB_CHECK(agb_numel(cols) == n, AGB_ERR_INCOMPATIBLE_SHAPE, "im2col: cols must have %lld elements", (long long)n);
  if (n == 0) return AGB_OK;
  im2col_kernel<<<agb_grid_for(n, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(x->ptr, cols->ptr, g, n);
  AGB_LAUNCHED(ctx);
  return AGB_OK;
}
