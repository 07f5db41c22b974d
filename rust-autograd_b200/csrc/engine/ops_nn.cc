// ops_nn.cc — dense contractions, convolution family, pooling, dropout, optimizer update ops and the Optimizer trait.
// Reference files mirrored: tensor_ops/dot_ops.rs, tensor_ops/conv_ops/{conv2d,conv2d_transpose,max_pool2d}.rs,
// tensor_ops/random_ops.rs:218-245, tensor_ops/gradient_descent_ops/*.rs, optimizers/*.rs.
#include "agx.h"
#include <cstdlib>
#include <algorithm>

namespace agx {
#define REFNAME(mod, nm) "autograd::tensor_ops::" mod "::" nm
static NdArray on_dev(Device* d, NdArray a) { d->ensure_device(a); return a; }

// ================================================================================================ MatMul / BatchMatMul
struct MatMul : Op {                   // dot_ops.rs:554-629.  Transposes = stride swaps; the kernel consumes strided views directly.
  bool ta, tb, batched;
  const char* name() const override { return batched ? REFNAME("dot_ops", "BatchMatMul") : REFNAME("dot_ops", "MatMul"); }
  void compute(ComputeContext& c) override {
    NdArray a = on_dev(c.dev, c.input(0)), b = on_dev(c.dev, c.input(1));
    if (!batched) {
      if (a.ndim() != 2) throw OpError(AGB_ERR_INCOMPATIBLE_SHAPE, "matmul: lhs input's ndim must be 2");     // :568-573
      if (b.ndim() != 2) throw OpError(AGB_ERR_INCOMPATIBLE_SHAPE, "matmul: rhs input's ndim must be 2");
    } else {
      if (a.ndim() < 2) throw OpError(AGB_ERR_INCOMPATIBLE_SHAPE, "BatchMatMul: Left-hand-side input's ndim must be >= 2");
      if (b.ndim() < 2) throw OpError(AGB_ERR_INCOMPATIBLE_SHAPE, "BatchMatMul: Right-hand-side input's ndim must be >= 2");
      if (a.ndim() != b.ndim()) throw OpError(AGB_ERR_INCOMPATIBLE_SHAPE, "Input shapes mismatch: ranks differ");
      // non-collapsible batch dims are deep-copied, like batch_mat_mul_requires_copy (dot_ops.rs:444-453,524-531)
      auto collapsible = [](const NdArray& t) { for (int i = t.ndim() - 4; i >= 0; i--) if (t.shape[i] != 1 && t.stride[i] != t.stride[i + 1] * t.shape[i + 1]) return false; return true; };
      if (!collapsible(a)) a = c.dev->copy(a);
      if (!collapsible(b)) b = c.dev->copy(b);
    }
    // one term of a gradient accumulation sum_t A_t^T * G_t (MatMul::grad under AddN, gradient.rs:168-173): deferred, so that the sum
    // runs as ONE long-K GEMM over the stacked operands instead of T short ones plus an add (fuse.cc)
    if (!batched && ta && !tb && c.run->fuse && c.run->sole_consumer_sums(c.node)) { NdArray r = expr_gemm_ta(c, a, b); if (r.expr) { c.append_output(r); return; } }
    int R = a.ndim();
    int64_t m = ta ? a.shape[R - 1] : a.shape[R - 2], n = tb ? b.shape[R - 2] : b.shape[R - 1];
    Shape out(a.shape.begin(), a.shape.end() - 2); out.push_back(m); out.push_back(n);
    NdArray y = c.dev->empty(out);
    agb_tensor da = a.desc(), db = b.desc(), dy = y.desc();
    check_status(agb_gemm_f32(c.dev->ctx, ta ? 1 : 0, tb ? 1 : 0, &da, &db, &dy, 0.0f));
    c.append_output(y);
  }
  bool plain_matmul(bool* tb_) const override { if (batched || ta) return false; *tb_ = tb; return true; }
  void grad(GradientContext& c) override {       // dot_ops.rs:608-628,697-717: the forward op's own flags are ignored (sic)
    Tensor gy = c.output_grad();
    auto mk = [&](Tensor l, Tensor r, bool tl, bool tr) { auto* op = new MatMul(); op->ta = tl; op->tb = tr; op->batched = batched; return TensorBuilder(c.graph()).append_input(l, false).append_input(r, false).build(op); };
    c.append_input_grad(mk(gy, c.input(1), false, true));
    c.append_input_grad(mk(c.input(0), gy, true, false));
  }
};
Tensor T::matmul(Tensor a, Tensor b) { auto* op = new MatMul(); op->ta = op->tb = false; op->batched = false; return TensorBuilder(a.graph).append_input(a, false).append_input(b, false).build(op); }
Tensor T::batch_matmul_t(Tensor a, Tensor b, bool ta, bool tb) { auto* op = new MatMul(); op->ta = ta; op->tb = tb; op->batched = true; return TensorBuilder(a.graph).append_input(a, false).append_input(b, false).build(op); }

struct TensordotPreprocess : Op {      // dot_ops.rs:720-799: host shape math, 5 tiny outputs
  const char* name() const override { return REFNAME("dot_ops", "TensordotPreprocess"); }
  void compute(ComputeContext& c) override {
    NdArray x0 = c.input(0), x1 = c.input(1), a0 = c.input(2), a1 = c.input(3);
    auto norm = [&](NdArray& ax, int nd) { std::vector<int64_t> v = as_ints(c.dev, ax); for (auto& e : v) e = e < 0 ? e + nd : e; return v; };
    std::vector<int64_t> axes0 = norm(a0, x0.ndim()), axes1 = norm(a1, x1.ndim());
    auto pre = [&](const Shape& shp, const std::vector<int64_t>& axes, bool flip, std::vector<float>& perm, std::vector<float>& new_shape, std::vector<float>& free_dims) {
      std::vector<int64_t> free;
      for (int64_t i = 0; i < (int64_t)shp.size(); i++) if (std::find(axes.begin(), axes.end(), i) == axes.end()) free.push_back(i);
      int64_t pf = 1, pa = 1;
      for (auto i : free) { pf *= shp[i]; free_dims.push_back((float)shp[i]); }
      for (auto i : axes) pa *= shp[i];
      std::vector<int64_t> first = flip ? axes : free, second = flip ? free : axes;
      for (auto i : first) perm.push_back((float)i);
      for (auto i : second) perm.push_back((float)i);
      new_shape = flip ? std::vector<float>{(float)pa, (float)pf} : std::vector<float>{(float)pf, (float)pa};
    };
    std::vector<float> perm0, ns0, fd0, perm1, ns1, fd1;
    pre(x0.shape, axes0, false, perm0, ns0, fd0); pre(x1.shape, axes1, true, perm1, ns1, fd1);
    fd0.insert(fd0.end(), fd1.begin(), fd1.end());
    auto H = [](std::vector<float> v) { int64_t n = (int64_t)v.size(); return NdArray::from_host({n}, std::move(v), true); };
    c.append_output(H(fd0)); c.append_output(H(perm0)); c.append_output(H(perm1)); c.append_output(H(ns0)); c.append_output(H(ns1));
  }
  void grad(GradientContext& c) override { for (int i = 0; i < 4; i++) c.append_none(); }
};
Tensor T::tensordot(Tensor a, Tensor b, Tensor a_axes, Tensor b_axes) {      // mod.rs:1914-1946
  Graph* g = a.graph;
  Tensor pre = TensorBuilder(g).append_input(a, false).append_input(b, false).append_input(a_axes, false).append_input(b_axes, false).build(new TensordotPreprocess());
  Tensor final_shape = nth_tensor(pre, 0), perm_a = nth_tensor(pre, 1), perm_b = nth_tensor(pre, 2), nsa = nth_tensor(pre, 3), nsb = nth_tensor(pre, 4);
  Tensor ar = reshape(transpose(a, perm_a), nsa), br = reshape(transpose(b, perm_b), nsb);
  return reshape(matmul(ar, br), final_shape);
}

// ================================================================================================ convolution family
// Conv2D's second output in the reference is the materialised im2col buffer (conv2d.rs:484, consumed by Conv2DFilterGrad via
// nth_tensor(y, 1), :571-582).  north_star forbids materialising it, so output #1 is a VIRTUAL tensor: a descriptor holding a
// reference to x plus the window geometry.  Conv2DFilterGrad / Conv2DWithCols recognise it and run implicit GEMM on x.
struct Im2colRef { NdArray x; int kh, kw, pad, stride, dil; };

NdArray materialize_im2col(Device* dev, const NdArray& cols) {      // only when user code evaluates nth_tensor(conv, 1)
  const Im2colRef& r = *cols.virt;
  NdArray x = dev->contiguous(r.x), out = dev->empty(cols.shape);
  agb_tensor tx = x.desc(), tc = out.desc();
  check_status(agb_im2col_f32(dev->ctx, &tx, &tc, r.kh, r.kw, r.pad, r.stride, r.dil));
  return out;
}
static NdArray real_cols(Device* dev, const NdArray& cols) { return cols.virt && !cols.on_device() ? materialize_im2col(dev, cols) : cols; }

struct ConvParams { int pad, stride, dilation; };

// ---- deferred epilogues ("elementwise fusion ... must keep unfused intermediates available").
// The reference graph of a conv layer is Conv2D -> AddOp(bias [1,O,1,1]) -> ReLU, and ReLU's gradient is
// mul(greater(pre_activation, 0), gy) (activation_ops.rs:161-166).  Conv2D / AddOp / `greater(., 0)` return a LAZY array (shape
// only); the op that completes the pattern runs ONE fused kernel:
//     ReLU(lazy conv[+bias])      -> conv kernel with bias + ReLU in its epilogue (agb_conv2d_fprop_fused_f32)
//     MulOp(lazy mask, gy)        -> AGB_B_RELU_GRAD
//     greater(lazy conv+bias, 0)  -> mask of the ReLU OUTPUT (x > 0  <=>  relu(x) > 0), no pre-activation needed
//     MulOp(lazy mask, lazy Conv2DTranspose)  -> dgrad kernel with the compare-and-zero in its epilogue (agb_conv2d_dgrad_fused_f32)
//     MulOp(lazy mask, lazy MaxPool2DGrad)    -> gather-form pool backward gated by (pooled output > 0) (agb_maxpool2d_bwd_fused)
//     MaxPool2D(lazy ReLU(conv [+ bias]))     -> pooling in the conv epilogue (agb_conv2d_fprop_pool_f32): the full-size activation
//                                                is written only if some other consumer asks for it
// ReLU itself stays deferred (`relu_deferred`) until its first consumer shows what it is: MaxPool2D fuses, anything else runs the
// conv + bias + ReLU kernel once and caches the value.  The ReLU-gradient mask refers to that deferred node (`src_lazy`), so the
// backward pass of a conv -> relu -> pool block needs neither the pre-activation nor the activation.
// Any other consumer gets the exact un-fused value through materialize_lazy (ComputeContext::input), so every intermediate of
// the reference graph remains observable; fusion only removes HBM round trips.
struct Lazy {
  int kind = 0;                        // 1 = conv [+ bias], 2 = (src > 0) mask, 3 = Conv2DTranspose(x = gy, w), 4 = MaxPool2DGrad(x = gy, idx)
  NdArray x, w, bias; bool has_bias = false; ConvParams p{0, 1, 1};
  NdArray relu_out; bool has_relu = false;       // filled when the fused conv + bias + ReLU kernel ran
  bool relu_deferred = false;                     // kind 1: this node is ReLU(conv [+ bias]); `value` (when computed) is the activation
  std::shared_ptr<Lazy> relu_child;               // kind 1, pre-activation node: the deferred ReLU that consumed it
  NdArray src; std::shared_ptr<Lazy> src_lazy;    // kind 2: mask source as an array, or as a deferred ReLU node
  NdArray idx; int pool_size = 0, pool_stride = 0; // kind 4
  NdArray value; bool has_value = false;          // cache of the un-fused value
  bool no_bits = false;                           // kind 1: the activation is being materialised for a max-pool (its ReLU backward is gated by the pooled output): no sign bits wanted
};
// activations enter the conv / pool entry points either NCHW-contiguous or channels-last; anything else is deep-copied
static bool is_cl4(const NdArray& a) { std::vector<int> o; return a.ndim() == 4 && a.on_device() && a.dense_order(o) && o == std::vector<int>({0, 2, 3, 1}) && !a.is_contiguous(); }
static NdArray act_layout(Device* dev, NdArray a) { if (a.ndim() == 4 && (a.is_contiguous() || is_cl4(a))) return a; return dev->contiguous(a); }
static NdArray act_empty(Device* dev, const Shape& s, bool channels_last) { return channels_last ? dev->empty_ordered(s, {0, 2, 3, 1}) : dev->empty(s); }
static Tensor mk_conv_transpose(Graph* g, Tensor gy, Tensor w, ConvParams p);
static Tensor mk_conv(Graph* g, Tensor x, Tensor w, ConvParams p);
static Tensor mk_filter_grad(Graph* g, Tensor cols, Tensor gy, Tensor w, Tensor bp_x, Tensor bp_gy, ConvParams p);
static Tensor mk_conv_with_cols(Graph* g, Tensor cols, Tensor w, Tensor bp_x, Tensor bp_w, ConvParams p);

struct PoolRef { NdArray y; const float* x_dptr; Shape x_shape, x_stride; std::shared_ptr<Lazy> x_lazy; };      // x_lazy: the pooled input was a deferred ReLU (fused conv + pool)
// gx = conv2d_transpose(gy, w) [* (mask_src > 0)]
static NdArray run_dgrad(Device* dev, const Lazy& L, const NdArray* mask_src) {
  const NdArray &gy = L.x, &w = L.w; const ConvParams& p = L.p;
  int64_t xh = p.stride * (gy.shape[2] - 1) - 2 * p.pad + (p.dilation * (w.shape[2] - 1) + 1);     // follows the code (conv2d_transpose.rs:55-56)
  int64_t xw = p.stride * (gy.shape[3] - 1) - 2 * p.pad + (p.dilation * (w.shape[3] - 1) + 1);
  bool cl = p.stride == 1 && agb_conv_prefers_channels_last((int)w.shape[0], (int)w.shape[1], (int)w.shape[2], (int)w.shape[3], 1, (int)xw);
  if (p.stride > 1 && p.stride <= 4) {       // strided dgrad runs as phase convolutions on the tensor cores in TF32 mode: keep gx channels-last there
    int mode = 0; agb_get_math_mode(dev->ctx, &mode);
    cl = mode == AGB_MATH_TF32 && agb_conv_prefers_channels_last((int)w.shape[0], (int)w.shape[1], (int)w.shape[2], (int)w.shape[3], 1, (int)((xw + p.stride - 1) / p.stride));
  }
  if (mask_src) cl = !mask_src->is_contiguous();          // write gx in the mask's memory order so the epilogue can read it in place
  NdArray gx = act_empty(dev, {gy.shape[0], w.shape[1], xh, xw}, cl);
  agb_tensor tg = gy.desc(), tw = w.desc(), tx = gx.desc(), tm;
  if (mask_src) tm = mask_src->desc();
  // with the ReLU mask fused this is the gradient of a conv -> add(bias) -> relu layer's pre-activation: its per-channel sums
  // (the bias gradient) come out of the same epilogue
  NdArray cs; if (mask_src) cs = dev->empty({gx.shape[1]});
  // sign bits of the mask source, when the forward kernel wrote them next to this very buffer
  const uint32_t* bits = (mask_src && mask_src->relu_bits && mask_src->relu_bits_of == mask_src->dptr) ? (const uint32_t*)mask_src->relu_bits->dptr : nullptr;
  if (bits) check_status(agb_conv2d_dgrad_fused_bits_f32(dev->ctx, &tg, &tw, &tm, bits, cs.dptr, &tx, p.pad, p.stride, p.dilation));
  else
  check_status(agb_conv2d_dgrad_fused_f32(dev->ctx, &tg, &tw, mask_src ? &tm : nullptr, mask_src ? cs.dptr : nullptr, &tx, p.pad, p.stride, p.dilation));
  if (mask_src) gx.chan_sum = std::make_shared<NdArray>(cs);
  return gx;
}
// gx = max_pool2d_grad(gy, idx) [gated by pooled output > 0]
static NdArray run_pool_grad(Device* dev, const Lazy& L, bool gated) {
  NdArray gy = L.x; const NdArray& idx = L.idx;
  const bool icl = is_cl4(idx);
  if (icl != is_cl4(gy) || !(gy.is_contiguous() || is_cl4(gy))) {     // the kernel walks gy and the index buffer together
    NdArray t = act_empty(dev, gy.shape, icl); agb_tensor ts = gy.desc(), td = t.desc();
    check_status(agb_copy_strided(dev->ctx, &ts, &td)); gy = t;
  }
  int64_t xh = L.pool_stride * (gy.shape[2] - 1) - 2 * L.p.pad + L.pool_size, xw = L.pool_stride * (gy.shape[3] - 1) - 2 * L.p.pad + L.pool_size;     // (max_pool2d.rs:263-264)
  NdArray gx = act_empty(dev, {gy.shape[0], gy.shape[1], xh, xw}, icl);
  agb_tensor tg = gy.desc(), tx = gx.desc();
  NdArray cs; if (gated) cs = dev->empty({gx.shape[1]});
  check_status(agb_maxpool2d_bwd_fused(dev->ctx, &tg, idx.i32 ? nullptr : idx.dptr, idx.i32 ? (const int32_t*)idx.dptr : nullptr,
                                       gated ? idx.pool->y.dptr : nullptr, gated ? cs.dptr : nullptr, &tx, L.pool_size, L.pool_stride));
  if (gated) gx.chan_sum = std::make_shared<NdArray>(cs);
  return gx;
}
NdArray lazy_mask_src(Device* dev, const NdArray& m);
static NdArray run_conv_fused(Device* dev, const Lazy& L, bool relu) {
  const NdArray &x = L.x, &w = L.w;
  int64_t yh = (x.shape[2] + 2 * L.p.pad - (L.p.dilation * (w.shape[2] - 1) + 1)) / L.p.stride + 1;
  int64_t yw = (x.shape[3] + 2 * L.p.pad - (L.p.dilation * (w.shape[3] - 1) + 1)) / L.p.stride + 1;
  NdArray y = act_empty(dev, {x.shape[0], w.shape[0], yh, yw}, agb_conv_prefers_channels_last((int)x.shape[1], (int)w.shape[0], (int)w.shape[2], (int)w.shape[3], L.p.stride, (int)yw));
  agb_tensor tx = x.desc(), tw = w.desc(), ty = y.desc();
  static const bool want_bits = [] { const char* e = getenv("AGX_RELU_BITS"); return !(e && e[0] == '0'); }();
  if (want_bits && relu && !L.no_bits && y.shape[1] % 32 == 0 && is_cl4(y)) {       // a ReLU activation: let the kernel leave its sign bits next to it (1/32 of the bytes) for the masked dgrad
    NdArray bits = dev->empty({(y.size() + 31) / 32});
    int written = 0;
    check_status(agb_conv2d_fprop_fused_bits_f32(dev->ctx, &tx, &tw, L.has_bias ? L.bias.dptr : nullptr, 1, &ty, (uint32_t*)bits.dptr, &written, L.p.pad, L.p.stride, L.p.dilation));
    if (written) { y.relu_bits = std::make_shared<NdArray>(bits); y.relu_bits_of = y.dptr; }
    return y;
  }
  check_status(agb_conv2d_fprop_fused_f32(dev->ctx, &tx, &tw, L.has_bias ? L.bias.dptr : nullptr, relu ? 1 : 0, &ty, L.p.pad, L.p.stride, L.p.dilation));
  return y;
}
NdArray materialize_lazy(Device* dev, const NdArray& a) {
  Lazy& L = *a.lazy;
  if (!L.has_value) {
    if (L.kind == 1) {
      L.value = run_conv_fused(dev, L, L.relu_deferred);
      if (L.relu_deferred) { L.relu_out = L.value; L.has_relu = true; }
    }
    else if (L.kind == 3) L.value = run_dgrad(dev, L, nullptr);
    else if (L.kind == 4) L.value = run_pool_grad(dev, L, false);
    else {
      NdArray zero = dev->full({}, 0.0f), src = lazy_mask_src(dev, a);
      std::vector<int> order;
      if (!src.dense_order(order)) { src = dev->contiguous(src); src.dense_order(order); }
      NdArray y = dev->empty_ordered(src.shape, order);          // same memory order as the source
      agb_tensor ta, tb, ty;
      ta.ptr = src.dptr; ta.rank = 1; ta.shape[0] = src.size(); ta.stride[0] = 1;
      tb.ptr = zero.dptr; tb.rank = 1; tb.shape[0] = src.size(); tb.stride[0] = 0;
      ty.ptr = y.dptr; ty.rank = 1; ty.shape[0] = src.size(); ty.stride[0] = 1;
      check_status(agb_binary(dev->ctx, AGB_B_GT, 0.f, 0.f, &ta, &tb, &ty));
      L.value = y;
    }
    L.has_value = true;
  }
  return L.value;
}
// hooks used by ops_basic.cc (AddOp / ReLU / greater / MulOp)
NdArray lazy_conv_add_bias(const NdArray& conv, const NdArray& bias) {      // returns an invalid (no lazy) array when the pattern does not apply
  NdArray r;
  if (!conv.lazy || conv.lazy->kind != 1 || conv.lazy->has_bias || conv.lazy->has_relu || conv.lazy->relu_deferred || conv.lazy->has_value) return r;
  if (conv.ndim() != 4 || bias.ndim() != 4 || !bias.on_device() || !bias.is_contiguous()) return r;
  if (bias.shape[0] != 1 || bias.shape[1] != conv.shape[1] || bias.shape[2] != 1 || bias.shape[3] != 1) return r;
  if (((uintptr_t)bias.dptr) & 15) return r;
  auto L = std::make_shared<Lazy>(*conv.lazy); L->bias = bias; L->has_bias = true;
  r.shape = conv.shape; r.stride = NdArray::contiguous_strides(r.shape); r.lazy = L;
  return r;
}
NdArray lazy_conv_relu(Device* dev, const NdArray& a) {                      // ReLU(conv [+ bias]): stays deferred, see above
  (void)dev;
  Lazy& L = *a.lazy;
  if (L.has_value || L.relu_deferred) return NdArray();          // pre-activation already materialised / ReLU of a ReLU: plain kernel
  auto L2 = std::make_shared<Lazy>(L); L2->relu_deferred = true; L2->relu_child.reset();
  a.lazy->relu_child = L2;
  NdArray r; r.shape = a.shape; r.stride = NdArray::contiguous_strides(r.shape); r.lazy = L2;
  return r;
}
NdArray lazy_gt0_mask(const NdArray& src_or_lazy) {                          // greater(x, 0) whose multiply has not arrived yet
  NdArray r; auto M = std::make_shared<Lazy>(); M->kind = 2;
  if (src_or_lazy.lazy) {
    Lazy& S = *src_or_lazy.lazy;
    if (S.kind != 1) return NdArray();
    if (S.relu_deferred) M->src_lazy = src_or_lazy.lazy;                     // greater(relu(x), 0)
    else if (S.relu_child) M->src_lazy = S.relu_child;                       // greater(x, 0) with relu(x) deferred: x > 0 <=> relu(x) > 0
    else if (S.has_relu) M->src = S.relu_out;
    else return NdArray();
  } else M->src = src_or_lazy;
  r.shape = src_or_lazy.shape; r.stride = NdArray::contiguous_strides(r.shape); r.lazy = M;
  return r;
}
// the array a mask compares with 0 (materialises a deferred ReLU on first use; cached there)
NdArray lazy_mask_src(Device* dev, const NdArray& m) {
  Lazy& M = *m.lazy;
  if (!M.src_lazy) return M.src;
  NdArray t; t.shape = m.shape; t.stride = NdArray::contiguous_strides(t.shape); t.lazy = M.src_lazy;
  return materialize_lazy(dev, t);
}
// MulOp(mask, lazy producer of gy): returns the fused result, or an invalid array when the pattern does not apply
NdArray lazy_fuse_mask(Device* dev, const NdArray& mask, const NdArray& prod) {
  Lazy& L = *prod.lazy;
  if (L.has_value || mask.shape != prod.shape || mask.ndim() != 4) return NdArray();
  Lazy& M = *mask.lazy;
  if (L.kind == 4) {
    // the gate is the pooled OUTPUT: valid only when the mask source is the very tensor that was pooled (then x[argmax] == y)
    const NdArray& idx = L.idx;
    if (!idx.pool || L.pool_size != L.pool_stride || L.p.pad != 0) return NdArray();
    if (is_cl4(idx) != is_cl4(idx.pool->y) || idx.pool->y.shape != idx.shape) return NdArray();
    bool same = false;
    if (M.src_lazy) {
      if (idx.pool->x_lazy == M.src_lazy) same = true;                       // fused conv + pool: identity of the deferred ReLU node
      else if (M.src_lazy->has_value) { const NdArray& v = M.src_lazy->value; same = idx.pool->x_dptr == v.dptr && idx.pool->x_shape == v.shape && idx.pool->x_stride == v.stride; }
    } else same = idx.pool->x_dptr != nullptr && idx.pool->x_dptr == M.src.dptr && idx.pool->x_shape == M.src.shape && idx.pool->x_stride == M.src.stride;
    if (!same) return NdArray();
    return run_pool_grad(dev, L, true);
  }
  if (L.kind == 3) {
    const NdArray src = lazy_mask_src(dev, mask);
    if (src.shape != prod.shape || !src.on_device() || !(src.is_contiguous() || is_cl4(src))) return NdArray();
    return run_dgrad(dev, L, &src);
  }
  return NdArray();
}
bool lazy_is_mask(const NdArray& a) { return a.lazy && a.lazy->kind == 2 && !a.lazy->has_value; }
bool lazy_is_conv(const NdArray& a) { return a.lazy && a.lazy->kind == 1; }

static int64_t conv_out(int64_t x, int64_t k, ConvParams p) { return (x + 2 * p.pad - (p.dilation * (k - 1) + 1)) / p.stride + 1; }

struct Conv2D : Op {                   // conv2d.rs:531-586
  ConvParams p;
  const char* name() const override { return REFNAME("conv_ops::conv2d", "Conv2D"); }
  void compute(ComputeContext& c) override {
    NdArray x = act_layout(c.dev, on_dev(c.dev, c.input(0))), w = c.dev->contiguous(on_dev(c.dev, c.input(1)));   // deep_copy if not a dense layout (:436-452)
    if (x.ndim() != 4) throw OpError(AGB_ERR_INCOMPATIBLE_SHAPE, "conv2d: lhs input must be 4D");
    if (w.ndim() != 4) throw OpError(AGB_ERR_INCOMPATIBLE_SHAPE, "conv2d: filter must be 4D");
    if (x.shape[1] != w.shape[1]) throw OpError(AGB_ERR_INCOMPATIBLE_SHAPE, "conv2d: input channel dim must match filter's second dim");
    int64_t yh = conv_out(x.shape[2], w.shape[2], p), yw = conv_out(x.shape[3], w.shape[3], p);
    if (yh < 1 || yw < 1) throw OpError(AGB_ERR_INCOMPATIBLE_SHAPE, "conv2d: kernel larger than padded input");
    NdArray y;      // deferred: the bias add / ReLU that usually follow are folded into the conv kernel's epilogue
    y.shape = {x.shape[0], w.shape[0], yh, yw}; y.stride = NdArray::contiguous_strides(y.shape);
    y.lazy = std::make_shared<Lazy>(); y.lazy->kind = 1; y.lazy->x = x; y.lazy->w = w; y.lazy->p = p;
    NdArray cols; cols.shape = {x.shape[0], x.shape[1], w.shape[2], w.shape[3], yh, yw}; cols.stride = NdArray::contiguous_strides(cols.shape);
    cols.virt = std::make_shared<Im2colRef>(Im2colRef{x, (int)w.shape[2], (int)w.shape[3], p.pad, p.stride, p.dilation});
    c.append_output(y); c.append_output(cols);
  }
  void grad(GradientContext& c) override {
    Graph* g = c.graph(); Tensor gy = c.output_grad(), y = c.output(), x = c.input(0), w = c.input(1);
    c.append_input_grad(mk_conv_transpose(g, gy, w, p));
    c.append_input_grad(mk_filter_grad(g, T::nth_tensor(y, 1), gy, w, x, gy, p));
  }
};
struct Conv2DWithCols : Op {           // conv2d.rs:589-628
  ConvParams p;
  const char* name() const override { return REFNAME("conv_ops::conv2d", "Conv2DWithCols"); }
  void compute(ComputeContext& c) override {
    NdArray cols = c.input(0), w = c.dev->contiguous(on_dev(c.dev, c.input(1)));
    if (cols.ndim() != 6 || w.ndim() != 4) throw OpError(AGB_ERR_INCOMPATIBLE_SHAPE, "Conv2DWithCols: cols must be 6-D and the filter 4-D");
    if (cols.virt) {
      NdArray x = act_layout(c.dev, cols.virt->x);
      NdArray y = act_empty(c.dev, {x.shape[0], w.shape[0], cols.shape[4], cols.shape[5]}, is_cl4(x));
      agb_tensor tx = x.desc(), tw = w.desc(), ty = y.desc();
      check_status(agb_conv2d_fprop_f32(c.dev->ctx, &tx, &tw, &ty, cols.virt->pad, cols.virt->stride, cols.virt->dil));
      c.append_output(y); return;
    }
    cols = c.dev->contiguous(on_dev(c.dev, cols));       // materialised cols supplied by the caller: y[b] = W . cols[b]
    int64_t B = cols.shape[0], K = cols.shape[1] * cols.shape[2] * cols.shape[3], P = cols.shape[4] * cols.shape[5], O = w.shape[0];
    NdArray y = c.dev->empty({B, O, cols.shape[4], cols.shape[5]});
    NdArray w2 = w.reshaped({1, O, K}); w2.stride[0] = 0;  NdArray wb = w2; wb.shape[0] = B;
    NdArray c3 = cols.reshaped({B, K, P}), y3 = y.reshaped({B, O, P});
    agb_tensor ta = wb.desc(), tb = c3.desc(), ty = y3.desc();
    check_status(agb_gemm_f32(c.dev->ctx, 0, 0, &ta, &tb, &ty, 0.0f));
    c.append_output(y);
  }
  void grad(GradientContext& c) override {
    Graph* g = c.graph(); Tensor cols = c.input(0), w = c.input(1), y = c.output(), gy = c.output_grad();
    c.append_input_grad(mk_conv_transpose(g, gy, w, p));
    Tensor bp0 = g->tensor(g->inner(y.id).get_backprop_inputs()[0].id);
    c.append_input_grad(mk_filter_grad(g, cols, gy, w, bp0, gy, p));
  }
};
struct Conv2DFilterGrad : Op {         // conv2d.rs:736-776
  ConvParams p;
  const char* name() const override { return REFNAME("conv_ops::conv2d", "Conv2DFilterGrad"); }
  void compute(ComputeContext& c) override {
    NdArray cols = c.input(0), gy = act_layout(c.dev, on_dev(c.dev, c.input(1))), w = c.input(2);
    NdArray gw = c.dev->empty(w.shape);
    if (cols.virt) {
      NdArray x = act_layout(c.dev, cols.virt->x);
      agb_tensor tx = x.desc(), tg = gy.desc(), tw = gw.desc();
      check_status(agb_conv2d_wgrad_f32(c.dev->ctx, &tx, &tg, &tw, cols.virt->pad, cols.virt->stride, cols.virt->dil));
      c.append_output(gw); return;
    }
    cols = c.dev->contiguous(on_dev(c.dev, cols)); gy = c.dev->contiguous(gy);       // gw = sum_b gy[b] . cols[b]^T, beta = 1 over the batch like conv2d.rs:703-722
    int64_t B = cols.shape[0], K = cols.shape[1] * cols.shape[2] * cols.shape[3], P = cols.shape[4] * cols.shape[5], O = gy.shape[1];
    NdArray g3 = gy.reshaped({B, O, P}), c3 = cols.reshaped({B, K, P}), gw2 = gw.reshaped({O, K});
    for (int64_t b = 0; b < B; b++) {
      NdArray gb = g3.sliced(0, b, 1).reshaped({O, P}), cb = c3.sliced(0, b, 1).reshaped({K, P});
      gb.dptr = g3.dptr + b * O * P; cb.dptr = c3.dptr + b * K * P;
      agb_tensor ta = gb.desc(), tb = cb.desc(), ty = gw2.desc();
      check_status(agb_gemm_f32(c.dev->ctx, 0, 1, &ta, &tb, &ty, b == 0 ? 0.0f : 1.0f));
    }
    c.append_output(gw);
  }
  void grad(GradientContext& c) override {
    Graph* g = c.graph(); Tensor cols = c.input(0), gy = c.input(1), ggw = c.output_grad(), y = c.output();
    c.append_input_grad(mk_conv_transpose(g, gy, ggw, p));
    Tensor bp0 = g->tensor(g->inner(y.id).get_backprop_inputs()[0].id);
    c.append_input_grad(mk_conv_with_cols(g, cols, ggw, bp0, ggw, p));
  }
};
struct Conv2DTransposeFilterGrad;
struct Conv2DTranspose : Op {          // conv2d_transpose.rs:249-300
  ConvParams p;
  const char* name() const override { return REFNAME("conv_ops::conv2d_transpose", "Conv2DTranspose"); }
  void compute(ComputeContext& c) override {
    NdArray gy = act_layout(c.dev, on_dev(c.dev, c.input(0))), w = c.dev->contiguous(on_dev(c.dev, c.input(1)));
    if (gy.ndim() != 4) throw OpError(AGB_ERR_INCOMPATIBLE_SHAPE, "conv2d_transpose: Input must be 4D");
    if (w.ndim() != 4) throw OpError(AGB_ERR_INCOMPATIBLE_SHAPE, "conv2d_transpose: Filter must be 4D");
    if (gy.shape[1] != w.shape[0]) throw OpError(AGB_ERR_INCOMPATIBLE_SHAPE, "conv2d_transpose: Number of input channels must match second filter dim");
    int64_t xh = p.stride * (gy.shape[2] - 1) - 2 * p.pad + (p.dilation * (w.shape[2] - 1) + 1);     // follows the code (:55-56)
    int64_t xw = p.stride * (gy.shape[3] - 1) - 2 * p.pad + (p.dilation * (w.shape[3] - 1) + 1);
    if (xh < 1 || xw < 1) throw OpError(AGB_ERR_INCOMPATIBLE_SHAPE, "conv2d_transpose: non-positive output size");
    NdArray gx;     // deferred: when this is a backward conv, the ReLU-grad multiply that usually follows joins the kernel's epilogue
    gx.shape = {gy.shape[0], w.shape[1], xh, xw}; gx.stride = NdArray::contiguous_strides(gx.shape);
    gx.lazy = std::make_shared<Lazy>(); gx.lazy->kind = 3; gx.lazy->x = gy; gx.lazy->w = w; gx.lazy->p = p;
    c.append_output(gx);
  }
  void grad(GradientContext& c) override;
};
struct Conv2DTransposeFilterGrad : Op {   // conv2d_transpose.rs:433-480: inputs (gy, x, w)
  ConvParams p;
  const char* name() const override { return REFNAME("conv_ops::conv2d_transpose", "Conv2DTransposeFilterGrad"); }
  void compute(ComputeContext& c) override {
    NdArray gy = act_layout(c.dev, on_dev(c.dev, c.input(0))), x = act_layout(c.dev, on_dev(c.dev, c.input(1))), w = c.input(2);
    NdArray gw = c.dev->empty(w.shape);
    agb_tensor ti = gy.desc(), tg = x.desc(), tw = gw.desc();     // roles swapped: gy is im2col'd, x multiplies it
    check_status(agb_conv2d_wgrad_f32(c.dev->ctx, &ti, &tg, &tw, p.pad, p.stride, p.dilation));
    c.append_output(gw);
  }
  void grad(GradientContext& c) override {
    Graph* g = c.graph(); Tensor gy = c.input(0), gw = c.output_grad(), x = c.input(1);
    c.append_input_grad(mk_conv_transpose(g, x, gw, p));
    c.append_input_grad(mk_conv(g, gy, gw, p));
    c.append_none();
  }
};
void Conv2DTranspose::grad(GradientContext& c) {
  Graph* g = c.graph(); Tensor x = c.input(0), w = c.input(1), gy = c.output_grad();
  c.append_input_grad(mk_conv(g, gy, w, p));
  auto* op = new Conv2DTransposeFilterGrad(); op->p = p;
  c.append_input_grad(TensorBuilder(g).append_input(gy, false).append_input(x, false).append_input(T::stop_gradient(w), false).build(op));
}
static Tensor mk_conv(Graph* g, Tensor x, Tensor w, ConvParams p) { auto* op = new Conv2D(); op->p = p; return TensorBuilder(g).append_input(x, false).append_input(w, false).build(op); }
static Tensor mk_conv_transpose(Graph* g, Tensor gy, Tensor w, ConvParams p) { auto* op = new Conv2DTranspose(); op->p = p; return TensorBuilder(g).append_input(gy, false).append_input(w, false).build(op); }
static Tensor mk_filter_grad(Graph* g, Tensor cols, Tensor gy, Tensor w, Tensor bp_x, Tensor bp_gy, ConvParams p) {
  auto* op = new Conv2DFilterGrad(); op->p = p;
  return TensorBuilder(g).append_input(cols, false).append_input(gy, false).append_input(w, false).append_backprop_input(bp_x).append_backprop_input(bp_gy).build(op);
}
static Tensor mk_conv_with_cols(Graph* g, Tensor cols, Tensor w, Tensor bp_x, Tensor bp_w, ConvParams p) {
  auto* op = new Conv2DWithCols(); op->p = p;
  return TensorBuilder(g).append_input(cols, false).append_input(w, false).append_backprop_input(bp_x).append_backprop_input(bp_w).build(op);
}
Tensor T::conv2d(Tensor x, Tensor w, int pad, int stride, int dilation) { return mk_conv(x.graph, x, w, ConvParams{pad, stride, dilation}); }
Tensor T::conv2d_transpose(Tensor x, Tensor w, int pad, int stride, int dilation) { return mk_conv_transpose(x.graph, x, w, ConvParams{pad, stride, dilation}); }

// ================================================================================================ max pooling
struct MaxPool2DGradGrad : Op {        // max_pool2d.rs:297-337
  int size, pad, stride;
  const char* name() const override { return REFNAME("conv_ops::max_pool2d", "MaxPool2DGradGrad"); }
  void compute(ComputeContext& c) override {
    c.accept_i32 = true;
    NdArray ggx = act_layout(c.dev, on_dev(c.dev, c.input(0))), idx = act_layout(c.dev, on_dev(c.dev, c.input(1)));
    int64_t yh = (ggx.shape[2] + 2 * pad - size) / stride + 1, yw = (ggx.shape[3] + 2 * pad - size) / stride + 1;
    NdArray ggy = act_empty(c.dev, {ggx.shape[0], ggx.shape[1], yh, yw}, is_cl4(idx));      // laid out like the index buffer
    agb_tensor tx = ggx.desc(), ty = ggy.desc();
    check_status(agb_maxpool2d_gradgrad(c.dev->ctx, &tx, idx.i32 ? nullptr : idx.dptr, idx.i32 ? (const int32_t*)idx.dptr : nullptr, &ty));
    c.append_output(ggy);
  }
  void grad(GradientContext& c) override { c.append_none(); c.append_none(); }
};
struct MaxPool2DGrad : Op {            // max_pool2d.rs:245-295
  int size, pad, stride;
  const char* name() const override { return REFNAME("conv_ops::max_pool2d", "MaxPool2DGrad"); }
  void compute(ComputeContext& c) override {
    c.accept_i32 = true;
    NdArray gy = on_dev(c.dev, c.input(0)), idx = act_layout(c.dev, on_dev(c.dev, c.input(1)));
    if (gy.ndim() != 4 || idx.ndim() != 4 || gy.shape != idx.shape) throw OpError(AGB_ERR_INCOMPATIBLE_SHAPE, "max_pool2d_grad: gy and the index buffer must be 4-D and of one shape");
    int64_t xh = stride * (gy.shape[2] - 1) - 2 * pad + size, xw = stride * (gy.shape[3] - 1) - 2 * pad + size;     // (:263-264)
    NdArray gx;     // deferred: a following ReLU-grad multiply is folded into the scatter (see Lazy)
    gx.shape = {gy.shape[0], gy.shape[1], xh, xw}; gx.stride = NdArray::contiguous_strides(gx.shape);
    gx.lazy = std::make_shared<Lazy>(); gx.lazy->kind = 4; gx.lazy->x = gy; gx.lazy->idx = idx; gx.lazy->pool_size = size; gx.lazy->pool_stride = stride; gx.lazy->p.pad = pad;
    c.append_output(gx);
  }
  void grad(GradientContext& c) override {
    auto* op = new MaxPool2DGradGrad(); op->size = size; op->pad = pad; op->stride = stride;
    c.append_input_grad(TensorBuilder(c.graph()).append_input(c.output_grad(), false).append_input(c.input(1), false).build(op));
    c.append_none();
  }
};
struct MaxPool2D : Op {                // max_pool2d.rs:166-243
  int size, pad, stride;
  const char* name() const override { return REFNAME("conv_ops::max_pool2d", "MaxPool2D"); }
  void compute(ComputeContext& c) override {
    c.accept_lazy = true;
    NdArray x = c.input(0);
    if (x.ndim() != 4) throw OpError(AGB_ERR_INCOMPATIBLE_SHAPE, "max_pool2d: input must be 4-D");
    if (x.lazy) {
      // ReLU(conv [+ bias]) feeding a 2x2 / stride-2 pool: pooling runs in the conv epilogue, the activation is not written
      Lazy& L = *x.lazy;
      if (L.kind == 1 && L.relu_deferred && !L.has_value && size == 2 && pad == 0 && stride == 2 && x.shape[2] >= 2 && x.shape[3] >= 2 && is_cl4(L.x)) {
        const Shape ps = {x.shape[0], x.shape[1], x.shape[2] / 2, x.shape[3] / 2};
        NdArray y = act_empty(c.dev, ps, true), idx = act_empty(c.dev, ps, true);
        agb_tensor tx = L.x.desc(), tw = L.w.desc(), ty = y.desc();
        int st = agb_conv2d_fprop_pool_f32(c.dev->ctx, &tx, &tw, L.has_bias ? L.bias.dptr : nullptr, 1, &ty, (int32_t*)idx.dptr, L.p.pad, L.p.stride, L.p.dilation);
        if (st == AGB_OK) {
          idx.i32 = true;
          idx.pool = std::make_shared<PoolRef>(PoolRef{y, nullptr, x.shape, NdArray::contiguous_strides(x.shape), x.lazy});
          c.append_output(y); c.append_output(idx); return;
        }
        if (st != AGB_ERR_UNSUPPORTED) check_status(st);
      }
      if (L.kind == 1 && L.relu_deferred && !L.has_value) L.no_bits = true;
      x = materialize_lazy(c.dev, x);
    }
    x = act_layout(c.dev, on_dev(c.dev, x));
    int64_t yh = (x.shape[2] + 2 * pad - size) / stride + 1, yw = (x.shape[3] + 2 * pad - size) / stride + 1;
    const bool cl = is_cl4(x);
    NdArray y = act_empty(c.dev, {x.shape[0], x.shape[1], yh, yw}, cl), idx = act_empty(c.dev, {x.shape[0], x.shape[1], yh, yw}, cl);
    agb_tensor tx = x.desc(), ty = y.desc();
    // indices stay int32 on the device: the reference's float-encoded flat offsets lose bits above 2^24 elements
    // (max_pool2d.rs:74-75; a 256x64x128x128 VGG activation has 2.7e8), API-visible values are converted on fetch
    if (x.size() < (1ll << 31)) { check_status(agb_maxpool2d_fwd(c.dev->ctx, &tx, &ty, nullptr, (int32_t*)idx.dptr, size, pad, stride)); idx.i32 = true; }
    else check_status(agb_maxpool2d_fwd(c.dev->ctx, &tx, &ty, idx.dptr, nullptr, size, pad, stride));
    idx.pool = std::make_shared<PoolRef>(PoolRef{y, x.dptr, x.shape, x.stride, nullptr});
    c.append_output(y); c.append_output(idx);
  }
  void grad(GradientContext& c) override {
    auto* op = new MaxPool2DGrad(); op->size = size; op->pad = pad; op->stride = stride;
    c.append_input_grad(TensorBuilder(c.graph()).append_input(c.output_grad(), false).append_input(T::nth_tensor(c.output(), 1), false).build(op));
  }
};
Tensor T::max_pool2d(Tensor x, int size, int pad, int stride) { auto* op = new MaxPool2D(); op->size = size; op->pad = pad; op->stride = stride; return TensorBuilder(x.graph).append_input(x, false).build(op); }

// ================================================================================================ dropout
struct Dropout : Op {                  // random_ops.rs:218-245: NOT inverted; outputs (y, mask); eval mode scales by (1 - ratio)
  float ratio; bool train; uint64_t seed; std::shared_ptr<StreamCell> cell;
  std::shared_ptr<StreamCell> stream_cell() const override { return cell; }
  const char* name() const override { return REFNAME("random_ops", "Dropout"); }
  void compute(ComputeContext& c) override {
    NdArray x = c.dev->contiguous(on_dev(c.dev, c.input(0)));
    if (!train) { NdArray y = c.dev->empty(x.shape); agb_tensor tx = x.desc(), ty = y.desc(); check_status(agb_unary(c.dev->ctx, AGB_U_SCALE, 1.0f - ratio, 0.f, &tx, &ty)); c.append_output(y); return; }
    NdArray y = c.dev->empty(x.shape), mask = c.dev->empty(x.shape);
    agb_tensor tx = x.desc(), ty = y.desc(), tm = mask.desc();
    // The reference seeds a XorShift stream with a constant at every op CONSTRUCTION (mod.rs:2895-2905) and the op's rng then advances with
    // every evaluation: a graph rebuilt per step draws the same mask every step, a persistent graph (or a replayed step graph) a fresh one.
    // Same here: device Philox keyed by (seed, node id, evaluations of this op instance so far); the count lives in device memory so that
    // CUDA-graph replays advance it too.  The stream's values are parity-unpinned.
    if (!cell) cell = c.dev->new_stream_cell();
    check_status(agb_dropout_stream(c.dev->ctx, &tx, &ty, &tm, ratio, seed ? seed : 0x5EEDull, (uint64_t)c.node << 32, cell->ptr));
    c.append_output(y); c.append_output(mask);
  }
  void grad(GradientContext& c) override { c.append_input_grad(T::mul(c.output_grad(), T::nth_tensor(c.output(), 1))); }
};
Tensor T::dropout(Tensor x, float ratio, bool train, uint64_t seed) { auto* op = new Dropout(); op->ratio = ratio; op->train = train; op->seed = seed; return TensorBuilder(x.graph).append_input(x, false).build(op); }

// random_ops.rs:6-214: RandomNormal / RandomUniform / StandardNormal / StandardUniform / Bernoulli / Exponential / LogNormal / Gamma.
// The op owns its stream position like the reference's ArrayRng (RefCell<R>): every evaluation of the SAME node continues the stream,
// a node built with the default rng starts from the crate's fixed default seed (ndarray_ext.rs:250-264), so two default-constructed
// nodes draw the same values (as there).  Stream values are parity-unpinned: device Philox instead of XorShift.
struct RandomOp : Op {
  int kind; float p0, p1; uint64_t seed; std::shared_ptr<StreamCell> cell;
  std::shared_ptr<StreamCell> stream_cell() const override { return cell; }
  const char* name() const override {
    static const char* N[] = {REFNAME("random_ops", "RandomUniform"), REFNAME("random_ops", "RandomNormal"), REFNAME("random_ops", "Bernoulli"),
                              REFNAME("random_ops", "Exponential"), REFNAME("random_ops", "LogNormal"), REFNAME("random_ops", "Gamma")};
    return N[kind];
  }
  void compute(ComputeContext& c) override {
    NdArray sh = c.input(0);
    NdArray y = c.dev->empty(as_shape(c.dev, sh));
    agb_tensor ty = y.desc();
    if (!cell) cell = c.dev->new_stream_cell();       // evaluations so far, kept on the device (advances under CUDA-graph replay too)
    check_status(agb_random_stream(c.dev->ctx, kind, p0, p1, seed ? seed : 0x5EEDull, 0, cell->ptr, &ty));
    c.append_output(y);
  }
  void grad(GradientContext& c) override { c.append_none(); }
};
Tensor T::random(Graph* g, int kind, Tensor shape, float p0, float p1, uint64_t seed) {
  if (kind < 0 || kind >= AGB_RAND_COUNT) throw Panic("random: unknown distribution");
  auto* op = new RandomOp(); op->kind = kind; op->p0 = p0; op->p1 = p1; op->seed = seed;
  return TensorBuilder(g).append_input(shape, false).set_shape(shape).build(op);
}

// ================================================================================================ optimizer update ops
// One op node per variable as in the reference (AdamOp inputs: param(mut), grad, m(mut), v(mut), t(mut); adam.rs:11-58), but
// compute() only REGISTERS the update; Graph::eval flushes all registered updates of the run as a single fused multi-tensor
// kernel once every gradient exists (preceded by the NCCL gradient all-reduce in data-parallel runs).  This is the "clean"
// ordering of SURVEY §3.5: every gradient is taken at the pre-update weights.
enum { OPT_ADAM = 0, OPT_SGD = 1, OPT_MOMENTUM = 2, OPT_ADAGRAD = 3 };
// Data parallel: the gradients registered so far and not yet reduced are packed into one arena and summed by ONE NCCL all-reduce on
// the communication stream (agb_allreduce_sum_async), ordered after the kernels that produced them.  The evaluator reaches the update ops in
// variable order, i.e. the early layers' gradients complete first while the late layers' filter gradients are still to be computed, so every
// bucket but the last overlaps with compute (VGG stack: 9.6 MB of gradients = two buckets; only the second one is exposed).
static const int64_t AR_BUCKET_BYTES = 4 << 20;
static void reduce_bucket(Evaluation& run, Device* dev) {
  if (run.ar_next >= run.pending.size()) return;
  int64_t total = 0; for (size_t i = run.ar_next; i < run.pending.size(); i++) total += (run.pending[i].g.size() + 3) / 4 * 4;
  NdArray arena = dev->empty({total}); int64_t off = 0;
  for (size_t i = run.ar_next; i < run.pending.size(); i++) {
    PendingUpdate& u = run.pending[i];
    check_status(agb_d2d(dev->ctx, arena.dptr + off, u.g.dptr, (size_t)u.g.size() * sizeof(float)));
    NdArray v = arena.sliced(0, off, u.g.size()); v.shape = u.g.shape; v.stride = NdArray::contiguous_strides(v.shape);
    off += (u.g.size() + 3) / 4 * 4; u.g = v;
  }
  check_status(agb_allreduce_sum_async(dev->ctx, arena.dptr, total));
  run.ar_buckets.push_back(arena); run.ar_next = run.pending.size();
}
struct UpdateOp : Op {
  int kind; float h[4];
  const char* name() const override {
    return kind == OPT_ADAM ? REFNAME("gradient_descent_ops::adam", "AdamOp") : kind == OPT_SGD ? REFNAME("gradient_descent_ops::sgd", "SGDOp")
         : kind == OPT_MOMENTUM ? REFNAME("gradient_descent_ops::sgd", "MomentumSGDOp") : REFNAME("gradient_descent_ops::adagrad", "AdaGradOp");
  }
  void compute(ComputeContext& c) override {
    PendingUpdate u; u.kind = kind; for (int i = 0; i < 4; i++) u.h[i] = h[i];
    u.p = c.input_mut(0); u.g = c.dev->contiguous(on_dev(c.dev, c.input(1)));
    if (u.g.shape != u.p.shape) { if (u.g.size() != u.p.size()) throw OpError(AGB_ERR_INCOMPATIBLE_SHAPE, "optimizer: gradient shape does not match the variable"); }
    if (kind == OPT_ADAM) { u.s0 = c.input_mut(2); u.s1 = c.input_mut(3); u.t = c.input_mut(4); }
    else if (kind != OPT_SGD) u.s0 = c.input_mut(2);
    c.run->pending.push_back(u);
    if (c.run->graph->env->world > 1) {       // enough gradients for a bucket: start summing them while the remaining ones are still being computed
      int64_t bytes = 0; for (size_t i = c.run->ar_next; i < c.run->pending.size(); i++) bytes += c.run->pending[i].g.size() * (int64_t)sizeof(float);
      if (bytes >= AR_BUCKET_BYTES) reduce_bucket(*c.run, c.dev);
    }
    c.append_empty_output();
  }
  void grad(GradientContext& c) override { for (int i = 0; i < c.num_inputs(); i++) c.append_none(); }
};

void flush_pending_updates(Evaluation& run, VariableEnvironment* env) {
  if (run.pending.empty()) return;
  Device* dev = run.dev;
  dev->small_copies.clear();          // the variables are about to change: no cached copy of a view may outlive this point
  float gscale = 1.0f;
  if (env->world > 1) {
    // data parallel: the last bucket, then the compute stream waits for every bucket's sum; the optimizer kernel reads them
    // scaled by 1/world
    reduce_bucket(run, dev);
    check_status(agb_allreduce_wait(dev->ctx));
    gscale = 1.0f / (float)env->world;
  }
  for (int kind = 0; kind < 4; kind++) {
    std::vector<PendingUpdate*> us; for (auto& u : run.pending) if (u.kind == kind) us.push_back(&u);
    size_t i = 0;
    while (i < us.size()) {       // one launch per group of updates sharing hyper-parameters (normally: all of them)
      size_t j = i; std::vector<float*> p, s0, s1, t; std::vector<const float*> g; std::vector<int64_t> n;
      while (j < us.size() && std::equal(us[j]->h, us[j]->h + 4, us[i]->h)) {
        p.push_back(us[j]->p.dptr); g.push_back(us[j]->g.dptr); n.push_back(us[j]->p.size());
        s0.push_back(us[j]->s0.dptr); s1.push_back(us[j]->s1.dptr); t.push_back(us[j]->t.dptr); j++;
      }
      const float* h = us[i]->h; int cnt = (int)p.size();
      if (kind == OPT_ADAM) check_status(agb_multi_tensor_adam(dev->ctx, cnt, p.data(), g.data(), s0.data(), s1.data(), t.data(), n.data(), h[0], h[1], h[2], h[3], gscale));
      else if (kind == OPT_SGD) check_status(agb_multi_tensor_sgd(dev->ctx, cnt, p.data(), g.data(), n.data(), h[0], gscale));
      else if (kind == OPT_MOMENTUM) check_status(agb_multi_tensor_momentum(dev->ctx, cnt, p.data(), g.data(), s0.data(), n.data(), h[0], h[1], gscale));
      else check_status(agb_multi_tensor_adagrad(dev->ctx, cnt, p.data(), g.data(), s0.data(), n.data(), h[0], gscale));
      i = j;
    }
  }
  run.pending.clear(); run.ar_buckets.clear(); run.ar_next = 0;
}

// ---- Optimizer trait (optimizers/mod.rs:49-99) ----
void Optimizer::update(const std::vector<Tensor>& params, const std::vector<Tensor>& grads, Graph* g, const std::vector<Feed>& feeds) {
  std::vector<Tensor> ups = compute_updates(params, grads, g);
  std::vector<EvalResult> rs = eval(g, ups, feeds, false);
  for (auto& r : rs) if (!r.ok) throw OpError(r.err_code, r.err_msg);      // `r.unwrap()`
}
Tensor Optimizer::get_update_op(const std::vector<Tensor>& params, const std::vector<Tensor>& grads, Graph* g) { return T::add_n(compute_updates(params, grads, g)); }

namespace {
std::string vid_name(VariableID v, const char* suffix) { return std::to_string(v.v) + suffix; }
VariableID var_of(Tensor t) { VariableID v = t.graph->inner(t.id).variable_id; if (!v.valid()) throw Panic("Got non-variable tensor"); return v; }
Tensor update_node(Graph* g, int kind, const float h[4], Tensor param, Tensor grad, const std::vector<Tensor>& state) {
  auto* op = new UpdateOp(); op->kind = kind; for (int i = 0; i < 4; i++) op->h[i] = h[i];
  TensorBuilder b(g); b.append_input(param, true).append_input(grad, false);
  for (auto& s : state) b.append_input(s, true);
  return b.build(op);
}
void make_state(VariableEnvironment* env, const std::vector<VariableID>& vars, const std::string& ns, const std::vector<const char*>& suffixes, bool with_t) {
  for (auto vid : vars) {          // optimizers/adam.rs:84-103: "{vid}m", "{vid}v" zeros like the variable, "{vid}t" = 1.0 (0-d)
    if (vid.v < 0 || vid.v >= (int)env->array_list.size()) throw Panic("variable array not found");
    Shape shp = env->array_list[vid.v].shape; int64_t n = env->array_list[vid.v].size();
    std::vector<float> z((size_t)n, 0.f);
    for (auto sfx : suffixes) env->set(ns, vid_name(vid, sfx), shp, z.data());
    if (with_t) { float one = 1.0f; env->set(ns, vid_name(vid, "t"), {}, &one); }
  }
}
struct Adam : Optimizer {
  float alpha, eps, b1, b2; std::string ns;
  std::vector<Tensor> compute_updates(const std::vector<Tensor>& params, const std::vector<Tensor>& grads, Graph* g) override {   // optimizers/adam.rs:116-153
    if (params.size() != grads.size()) throw Panic("assertion failed: num_params == grads.len()");
    std::vector<Tensor> ret; float h[4] = {alpha, eps, b1, b2};
    for (size_t i = 0; i < params.size(); i++) {
      VariableID v = var_of(params[i]);
      ret.push_back(update_node(g, OPT_ADAM, h, params[i], grads[i], {g->variable_by_name(vid_name(v, "m"), ns), g->variable_by_name(vid_name(v, "v"), ns), g->variable_by_name(vid_name(v, "t"), ns)}));
    }
    return ret;
  }
};
struct SGD : Optimizer {               // optimizers/sgd.rs
  float lr;
  std::vector<Tensor> compute_updates(const std::vector<Tensor>& params, const std::vector<Tensor>& grads, Graph* g) override {
    if (params.size() != grads.size()) throw Panic("assertion failed: num_params == grads.len()");
    std::vector<Tensor> ret; float h[4] = {lr, 0, 0, 0};
    for (size_t i = 0; i < params.size(); i++) { var_of(params[i]); ret.push_back(update_node(g, OPT_SGD, h, params[i], grads[i], {})); }
    return ret;
  }
};
struct StatefulSGD : Optimizer {       // MomentumSGD ("{vid}v", optimizers/momentum_sgd.rs) and AdaGrad ("{vid}h", optimizers/adagrad.rs)
  int kind; float lr, momentum; std::string ns; const char* sfx;
  std::vector<Tensor> compute_updates(const std::vector<Tensor>& params, const std::vector<Tensor>& grads, Graph* g) override {
    if (params.size() != grads.size()) throw Panic("assertion failed: num_params == grads.len()");
    std::vector<Tensor> ret; float h[4] = {lr, momentum, 0, 0};
    for (size_t i = 0; i < params.size(); i++) { VariableID v = var_of(params[i]); ret.push_back(update_node(g, kind, h, params[i], grads[i], {g->variable_by_name(vid_name(v, sfx), ns)})); }
    return ret;
  }
};
}  // namespace
Optimizer* make_adam(VariableEnvironment* env, const std::vector<VariableID>& vars, const std::string& ns, float alpha, float eps, float b1, float b2) {
  make_state(env, vars, ns, {"m", "v"}, true);
  auto* a = new Adam(); a->alpha = alpha; a->eps = eps; a->b1 = b1; a->b2 = b2; a->ns = ns; return a;
}
Optimizer* make_sgd(float lr) { auto* s = new SGD(); s->lr = lr; return s; }
Optimizer* make_momentum_sgd(VariableEnvironment* env, const std::vector<VariableID>& vars, const std::string& ns, float lr, float momentum) {
  make_state(env, vars, ns, {""}, false);      // state is named "{vid}" (optimizers/momentum_sgd.rs:65)
  auto* s = new StatefulSGD(); s->kind = OPT_MOMENTUM; s->lr = lr; s->momentum = momentum; s->ns = ns; s->sfx = ""; return s;
}
Optimizer* make_adagrad(VariableEnvironment* env, const std::vector<VariableID>& vars, const std::string& ns, float lr) {
  make_state(env, vars, ns, {""}, false);      // state is named "{vid}" (optimizers/adagrad.rs:48)
  auto* s = new StatefulSGD(); s->kind = OPT_ADAGRAD; s->lr = lr; s->momentum = 0; s->ns = ns; s->sfx = ""; return s;
}

void grad_helper(const std::vector<Tensor>& losses, const std::string& ns, Graph* g, std::vector<Tensor>& vars, std::vector<Tensor>& grads) {   // optimizers/mod.rs:21-46
  std::vector<Tensor> ys; for (auto& l : losses) ys.push_back(T::sum_all(l));
  std::vector<Tensor> xs; for (auto vid : g->env->current_var_ids(ns)) xs.push_back(g->variable_by_id(vid));     // var_tensors_by_name (variable.rs:768-780)
  std::vector<Tensor> gs = compute_gradients(ys, xs, nullptr, g);
  for (size_t i = 0; i < xs.size(); i++) if (gs[i].valid()) { vars.push_back(xs[i]); grads.push_back(gs[i]); }
}

}  // namespace agx
