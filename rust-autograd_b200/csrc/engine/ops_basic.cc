// ops_basic.cc — Op::compute / Op::grad for the bandwidth-bound op families and host-metadata ops, plus their
// `tensor_ops` constructors.  compute() launches libagb200 kernels (include/agb200.h); grad() builds graph nodes exactly
// like the reference's Op::grad compositions.  Reference files mirrored:
//   binary_ops.rs, math_ops.rs, activation_ops.rs, reduction_ops.rs, xent_ops.rs, array_ops.rs, const_gen_ops.rs,
//   gradient_ops.rs, graph_ops.rs, hook_ops.rs  (all under /root/reference/src/tensor_ops/).
#include "agx.h"
#include <algorithm>
#include <math.h>
#include <set>

namespace agx {

// deferred-epilogue hooks (ops_nn.cc)
NdArray lazy_conv_add_bias(const NdArray& conv, const NdArray& bias);
NdArray lazy_conv_relu(Device* dev, const NdArray& a);
NdArray lazy_gt0_mask(const NdArray& src_or_lazy);
bool lazy_is_mask(const NdArray& a);
NdArray lazy_fuse_mask(Device* dev, const NdArray& mask, const NdArray& prod);
bool lazy_is_conv(const NdArray& a);
NdArray lazy_mask_src(Device* dev, const NdArray& a);

// ================================================================================================ device helpers
static NdArray on_dev(Device* d, NdArray a) { d->ensure_device(a); return a; }

// Elementwise kernels preserve the memory order of their (dominant) input: a channels-last activation stays channels-last
// through bias add / ReLU / gradient masks, so no layout change is ever paid between two tensor-core convolutions.
static bool is_identity(const std::vector<int>& o) { for (size_t i = 0; i < o.size(); i++) if (o[i] != (int)i) return false; return true; }
static agb_tensor flat_desc(const NdArray& a) { agb_tensor t; t.ptr = a.dptr; t.rank = 1; t.shape[0] = a.size(); t.stride[0] = 1; return t; }
static agb_tensor permuted_desc(const agb_tensor& t, const std::vector<int>& order) {
  agb_tensor r = t; for (int i = 0; i < t.rank; i++) { r.shape[i] = t.shape[order[i]]; r.stride[i] = t.stride[order[i]]; } return r;
}
static NdArray dev_unary(Device* d, int op, NdArray x, float p0 = 0.f, float p1 = 0.f) {
  if (x.expr) x = expr_materialize(d, x);
  d->ensure_device(x);
  std::vector<int> order;
  if (x.dense_order(order) && !is_identity(order)) {          // dense but permuted: run on the flat memory, keep the strides
    NdArray y = d->empty_ordered(x.shape, order);
    agb_tensor tx = flat_desc(x), ty = flat_desc(y);
    check_status(agb_unary(d->ctx, op, p0, p1, &tx, &ty));
    return y;
  }
  NdArray y = d->empty(x.shape);
  agb_tensor tx = x.desc(), ty = y.desc();
  check_status(agb_unary(d->ctx, op, p0, p1, &tx, &ty));
  return y;
}
static Shape broadcast_shape(const Shape& a, const Shape& b, const char* who) {
  size_t n = std::max(a.size(), b.size()); Shape r(n);
  for (size_t i = 0; i < n; i++) {
    int64_t da = i + a.size() >= n ? a[i + a.size() - n] : 1, db = i + b.size() >= n ? b[i + b.size() - n] : 1;
    if (da != db && da != 1 && db != 1) throw OpError(AGB_ERR_INCOMPATIBLE_SHAPE, std::string(who) + ": shapes are not broadcast-compatible");
    r[i] = da == 1 ? db : da;
  }
  return r;
}
static agb_tensor broadcast_desc(const NdArray& a, const Shape& out) {
  agb_tensor t; t.ptr = a.dptr; t.rank = (int)out.size();
  if (t.rank > AGB_MAX_RANK) throw OpError(AGB_ERR_INVALID_DIMS, "rank exceeds AGB_MAX_RANK");
  int pad = (int)out.size() - a.ndim();
  for (int i = 0; i < t.rank; i++) {
    t.shape[i] = out[i];
    if (i < pad) t.stride[i] = 0;
    else t.stride[i] = (a.shape[i - pad] == 1 && out[i] != 1) ? 0 : a.stride[i - pad];
  }
  return t;
}
static NdArray dev_binary(Device* d, int op, NdArray a, NdArray b, float p0 = 0.f, float p1 = 0.f, const char* who = "binary op") {
  if (a.expr) a = expr_materialize(d, a);
  if (b.expr) b = expr_materialize(d, b);
  d->ensure_device(a); d->ensure_device(b);
  Shape out = broadcast_shape(a.shape, b.shape, who);
  agb_tensor ta = broadcast_desc(a, out), tb = broadcast_desc(b, out);
  // output memory order = that of the first full-size operand with a permuted dense layout
  std::vector<int> order;
  const NdArray* dom = nullptr;
  if (a.shape == out && a.dense_order(order) && !is_identity(order)) dom = &a;
  else if (b.shape == out && b.dense_order(order) && !is_identity(order)) dom = &b;
  if (dom) {
    // the other operand is a smaller dense array in a DIFFERENT memory order (cnn_mnist.rs:36-45 adds full-map biases [1, C, H, W] to
    // channels-last activations): re-lay it once in the dominant order so that the add collapses to a vectorised row-broadcast pass
    NdArray* oth = dom == &a ? &b : &a; std::vector<int> oo; int nontrivial = 0;
    for (auto dd : oth->shape) if (dd != 1) nontrivial++;
    if (oth->ndim() == (int)out.size() && nontrivial > 1 && oth->size() < dom->size() && oth->size() <= (1 << 22) && oth->dense_order(oo) && oo != order) {
      NdArray t = d->empty_ordered(oth->shape, order);
      agb_tensor ts = oth->desc(), td = t.desc();
      check_status(agb_copy_strided(d->ctx, &ts, &td));
      *oth = t; ta = broadcast_desc(a, out); tb = broadcast_desc(b, out);
    }
    NdArray y = d->empty_ordered(out, order);
    agb_tensor pa = permuted_desc(ta, order), pb = permuted_desc(tb, order), py = permuted_desc(y.desc(), order);
    check_status(agb_binary(d->ctx, op, p0, p1, &pa, &pb, &py));
    return y;
  }
  NdArray y = d->empty(out);
  agb_tensor ty = y.desc();
  check_status(agb_binary(d->ctx, op, p0, p1, &ta, &tb, &ty));
  return y;
}
static NdArray dev_broadcast_to(Device* d, NdArray a, const Shape& out) {
  d->ensure_device(a);
  NdArray y = d->empty(out);
  agb_tensor ta = broadcast_desc(a, out), ty = y.desc();
  check_status(agb_copy_strided(d->ctx, &ta, &ty));
  return y;
}

// ---- host arithmetic on META arrays only (shape / axes vectors of <= a few elements; SURVEY §8 a27: these must never
//      become device kernels).  Data tensors never take this path.
static bool all_meta(const NdArray& a) { return a.meta && a.has_host(); }
static NdArray host_binary(int op, const NdArray& a, const NdArray& b) {
  Shape out = broadcast_shape(a.shape, b.shape, "binary op");
  int64_t n = 1; for (auto s : out) n *= s;
  std::vector<float> r((size_t)n);
  Shape sa = NdArray::contiguous_strides(a.shape), sb = NdArray::contiguous_strides(b.shape);
  for (int64_t i = 0; i < n; i++) {
    int64_t rem = i, ia = 0, ib = 0;
    for (int k = (int)out.size() - 1; k >= 0; k--) {
      int64_t c = rem % out[k]; rem /= out[k];
      int ka = k - ((int)out.size() - a.ndim()), kb = k - ((int)out.size() - b.ndim());
      if (ka >= 0 && a.shape[ka] != 1) ia += c * sa[ka];
      if (kb >= 0 && b.shape[kb] != 1) ib += c * sb[kb];
    }
    float x = (*a.host)[ia], y = (*b.host)[ib];
    r[i] = op == AGB_B_ADD ? x + y : op == AGB_B_SUB ? x - y : op == AGB_B_MUL ? x * y : x / y;
  }
  return NdArray::from_host(out, r, true);
}

static std::vector<int> norm_axes(Device* d, NdArray& axes, int ndim) {     // ndarray_ext::normalize_negative_axes
  std::vector<int64_t> v = as_ints(d, axes); std::vector<int> r;
  for (auto a : v) {
    int ax = normalize_negative_axis(a, ndim);
    if (ax < 0 || ax >= ndim) throw Panic("Invalid index value");
    r.push_back(ax);
  }
  return r;
}

// reduce `x` (contiguous) over the sorted axis set, highest group first (impl_reduce_forward!, reduction_ops.rs:54-108)
static NdArray dev_reduce_axes(Device* d, int op, NdArray x, std::vector<int> axes, bool keep_dims) {
  std::sort(axes.begin(), axes.end()); axes.erase(std::unique(axes.begin(), axes.end()), axes.end());
  x = on_dev(d, x);
  {   // dense permuted input (channels-last): reduce in memory order, hand the result back as a strided logical view
    std::vector<int> order;
    if (x.dense_order(order) && !is_identity(order)) {
      NdArray xp = x; std::vector<int> pos(order.size());
      for (size_t i = 0; i < order.size(); i++) { xp.shape[i] = x.shape[order[i]]; xp.stride[i] = x.stride[order[i]]; pos[order[i]] = (int)i; }
      std::vector<int> paxes; for (int a : axes) paxes.push_back(pos[a]);
      NdArray rp = dev_reduce_axes(d, op, xp, paxes, true);           // physical order, reduced axes kept as 1
      NdArray r = rp;                                                   // un-permute: logical axis a lives at physical slot pos[a]
      for (size_t a = 0; a < order.size(); a++) { r.shape[a] = rp.shape[pos[a]]; r.stride[a] = rp.stride[pos[a]]; }
      if (!keep_dims) {
        Shape s2, st2;
        for (size_t a = 0; a < order.size(); a++) if (std::find(axes.begin(), axes.end(), (int)a) == axes.end()) { s2.push_back(r.shape[a]); st2.push_back(r.stride[a]); }
        r.shape = s2; r.stride = st2;
      }
      return r;
    }
  }
  x = d->contiguous(x);
  Shape cur = x.shape; NdArray curr = x;
  int i = (int)axes.size() - 1;
  while (i >= 0) {
    int hi = axes[i], lo = hi;
    while (i > 0 && axes[i - 1] == lo - 1) { lo--; i--; }      // merge adjacent axes into one [outer, r, inner] pass
    i--;
    int64_t outer = 1, r = 1, inner = 1;
    for (int k = 0; k < lo; k++) outer *= cur[k];
    for (int k = lo; k <= hi; k++) r *= cur[k];
    for (int k = hi + 1; k < (int)cur.size(); k++) inner *= cur[k];
    Shape ns;
    for (int k = 0; k < (int)cur.size(); k++) { if (k < lo || k > hi) ns.push_back(cur[k]); else if (keep_dims) ns.push_back(1); }
    NdArray y = d->empty(ns);
    check_status(agb_reduce(d->ctx, op, curr.dptr, y.dptr, outer, r, inner));
    curr = y; cur = ns;
    if (keep_dims) { /* axis indices unchanged */ } else { /* lower axes keep their indices: we go from the highest down */ }
  }
  return curr;
}
static NdArray host_reduce_axes(int op, const NdArray& x, std::vector<int> axes, bool keep_dims) {
  std::sort(axes.begin(), axes.end()); axes.erase(std::unique(axes.begin(), axes.end()), axes.end());
  Shape out_keep = x.shape; for (int a : axes) out_keep[a] = 1;
  int64_t n = 1; for (auto s : out_keep) n *= s;
  float init = op == AGB_R_PROD ? 1.f : op == AGB_R_MIN ? 3.402823466e+38f : op == AGB_R_MAX ? -3.402823466e+38f : 0.f;
  std::vector<float> r((size_t)n, init);
  Shape so = NdArray::contiguous_strides(out_keep);
  int64_t total = x.size(), len = 1; for (int a : axes) len *= x.shape[a];
  for (int64_t i = 0; i < total; i++) {
    int64_t rem = i, o = 0;
    for (int k = x.ndim() - 1; k >= 0; k--) { int64_t c = rem % x.shape[k]; rem /= x.shape[k]; if (out_keep[k] != 1) o += c * so[k]; }
    float v = (*x.host)[i];
    r[o] = op == AGB_R_PROD ? r[o] * v : op == AGB_R_MIN ? fminf(r[o], v) : op == AGB_R_MAX ? fmaxf(r[o], v) : r[o] + v;
  }
  if (op == AGB_R_MEAN) for (auto& v : r) v *= 1.0f / (float)len;
  Shape out; for (int k = 0; k < x.ndim(); k++) { bool red = std::find(axes.begin(), axes.end(), k) != axes.end(); if (!red) out.push_back(x.shape[k]); else if (keep_dims) out.push_back(1); }
  return NdArray::from_host(out, r, true);
}

// ================================================================================================ const / source-like ops
#define REFNAME(mod, nm) "autograd::tensor_ops::" mod "::" nm

struct ConvertToTensor : Op {          // const_gen_ops.rs:84-90
  NdArray arr;
  const char* name() const override { return REFNAME("const_gen_ops", "ConvertToTensor"); }
  void compute(ComputeContext& c) override { c.append_output(arr); }     // device copy is made lazily and cached in `arr`'s consumers
  void grad(GradientContext&) override {}
};
struct ScalarOp : Op {                 // const_gen_ops.rs:16-25
  float val;
  const char* name() const override { return REFNAME("const_gen_ops", "Scalar"); }
  void compute(ComputeContext& c) override { c.append_output(NdArray::scalar_host(val)); }
  void grad(GradientContext& c) override { c.append_none(); }
};
struct FillOp : Op {                   // Zeros / Ones, const_gen_ops.rs:27-82
  float v;
  const char* name() const override { return v == 0.f ? REFNAME("const_gen_ops", "Zeros") : REFNAME("const_gen_ops", "Ones"); }
  void compute(ComputeContext& c) override { NdArray s = c.input(0); Shape shp = as_shape(c.dev, s); c.append_output(v == 0.f ? c.dev->zeros(shp) : c.dev->full(shp, v)); }
  void grad(GradientContext& c) override { c.append_none(); }
};

Tensor T::convert_to_tensor(Graph* g, const Shape& shape, const std::vector<float>& data, bool meta) {
  int64_t n = 1; for (auto s : shape) n *= s;
  if ((int64_t)data.size() != n) throw Panic("convert_to_tensor: data length does not match the shape");
  auto* sop = new ConvertToTensor(); std::vector<float> sh; for (auto s : shape) sh.push_back((float)s);
  sop->arr = NdArray::from_host({(int64_t)shape.size()}, sh, true);               // shape_of(&arr), :2395-2397
  Tensor st = TensorBuilder(g).build(sop);
  auto* op = new ConvertToTensor(); op->arr = NdArray::from_host(shape, data, meta);
  return TensorBuilder(g).set_shape(st).build(op);
}
Tensor T::as_tensor(Graph* g, const std::vector<int64_t>& ints) {
  std::vector<float> v; for (auto i : ints) v.push_back((float)i);
  return convert_to_tensor(g, {(int64_t)ints.size()}, v, true);
}
Tensor T::scalar(Graph* g, float v) {
  auto* op = new ScalarOp(); op->val = v;
  return TensorBuilder(g).set_shape(convert_to_tensor(g, {0}, {}, true)).build(op);    // scalar_shape() = ndarray of shape [0], :2415-2423
}
Tensor T::zeros(Graph* g, Tensor shape) { auto* op = new FillOp(); op->v = 0.f; return TensorBuilder(g).append_input(shape, false).build(op); }
Tensor T::ones(Graph* g, Tensor shape) { auto* op = new FillOp(); op->v = 1.f; return TensorBuilder(g).append_input(shape, false).build(op); }

// ================================================================================================ metadata ops (host)
struct ShapeOp : Op {                  // array_ops.rs:148-159
  const char* name() const override { return REFNAME("array_ops", "Shape"); }
  void compute(ComputeContext& c) override { c.accept_lazy = true; c.accept_expr = true; NdArray x = c.input(0); std::vector<float> v; for (auto s : x.shape) v.push_back((float)s); c.append_output(NdArray::from_host({(int64_t)v.size()}, v, true)); }
  void grad(GradientContext& c) override { c.append_none(); }
  bool metadata_only() const override { return true; }
};
struct RankOp : Op {
  const char* name() const override { return REFNAME("array_ops", "Rank"); }
  void compute(ComputeContext& c) override { c.accept_lazy = true; c.accept_expr = true; c.append_output(NdArray::scalar_host((float)c.input(0).ndim(), true)); }
  void grad(GradientContext& c) override { c.append_none(); }
  bool metadata_only() const override { return true; }
};
struct SizeOp : Op {
  const char* name() const override { return REFNAME("array_ops", "Size"); }
  void compute(ComputeContext& c) override { c.accept_lazy = true; c.accept_expr = true; c.append_output(NdArray::scalar_host((float)c.input(0).size(), true)); }
  void grad(GradientContext& c) override { c.append_none(); }
  bool metadata_only() const override { return true; }
};
struct InferBinOpShape : Op {          // array_ops.rs:107-146
  const char* name() const override { return REFNAME("array_ops", "InferBinOpShape"); }
  void compute(ComputeContext& c) override {
    NdArray af = c.input(0), bf = c.input(1);
    Shape a = as_shape(c.dev, af), b = as_shape(c.dev, bf);
    bool as = is_scalar_shape(a), bs = is_scalar_shape(b);
    if (!as && !bs) {
      if (a.size() != b.size()) throw OpError(AGB_ERR_INCOMPATIBLE_SHAPE, "InferBinOpShape: rank of lhs and rhs must match.");
      std::vector<float> m; for (size_t i = 0; i < a.size(); i++) m.push_back((float)std::max(a[i], b[i]));
      c.append_output(NdArray::from_host({(int64_t)a.size()}, m, true));
    } else if (!as) c.append_output_view(af);
    else c.append_output_view(bf);
  }
  void grad(GradientContext& c) override { c.append_none(); c.append_none(); }
};
struct SetDiff1D : Op {                // array_ops.rs:241-279
  const char* name() const override { return REFNAME("array_ops", "SetDiff1D"); }
  void compute(ComputeContext& c) override {
    NdArray a = c.input(0), b = c.input(1);
    std::set<int64_t> sa, sb; for (auto v : as_ints(c.dev, a)) sa.insert(v); for (auto v : as_ints(c.dev, b)) sb.insert(v);
    std::vector<float> r; for (auto v : sa) if (!sb.count(v)) r.push_back((float)v);
    c.append_output(NdArray::from_host({(int64_t)r.size()}, r, a.meta && b.meta));
  }
  void grad(GradientContext& c) override { c.append_none(); c.append_none(); }
};

Tensor T::shape(Tensor x) {
  TensorInternal& n = x.graph->inner(x.id);
  if (n.shape >= 0) return x.graph->tensor(n.shape);                // :270-272
  return TensorBuilder(x.graph).append_input(x, false).set_differentiable(false).build(new ShapeOp());
}
Tensor T::rank(Tensor x) { return TensorBuilder(x.graph).append_input(x, false).set_differentiable(false).build(new RankOp()); }
Tensor T::size(Tensor x) { return TensorBuilder(x.graph).append_input(x, false).set_differentiable(false).build(new SizeOp()); }
Tensor T::setdiff1d(Tensor a, Tensor b) { return TensorBuilder(a.graph).append_input(a, false).append_input(b, false).build(new SetDiff1D()); }
static Tensor infer_bin_op_shape(Graph* g, Tensor sa, Tensor sb) { return TensorBuilder(g).append_input(sa, false).append_input(sb, false).build(new InferBinOpShape()); }

// ================================================================================================ pass-through ops
struct IdentityOp : Op {               // activation_ops.rs:169-181 (also nth_tensor)
  const char* name() const override { return REFNAME("activation_ops", "Identity"); }
  void compute(ComputeContext& c) override {
    c.accept_i32 = true; c.accept_expr = true; NdArray x = c.input(0);
    if (x.expr) { c.append_output(expr_passthrough(c, x)); return; }       // the same pending expression, now with this node's consumers too
    c.append_output_view(x);
  }
  void grad(GradientContext& c) override { c.append_input_grad(c.output_grad()); }
};
struct StopGradient : Op {             // gradient_ops.rs:4-16
  const char* name() const override { return REFNAME("gradient_ops", "StopGradient"); }
  void compute(ComputeContext& c) override { c.append_output_view(c.input(0)); }
  void grad(GradientContext& c) override { c.append_none(); }
};
struct ControlDependency : Op {        // graph_ops.rs:5-17
  const char* name() const override { return REFNAME("graph_ops", "ControlDependency"); }
  void compute(ComputeContext& c) override { c.append_output_view(c.input(0)); }
  void grad(GradientContext& c) override { c.append_input_grad(c.output_grad()); }
};
struct HookOp : Op {                   // hook_ops.rs:5-31: the callback sees host values => explicit D2H sync point
  std::function<void(const NdArray&, const std::vector<float>&)> f;
  const char* name() const override { return REFNAME("hook_ops", "HookOp"); }
  void compute(ComputeContext& c) override { NdArray x = c.input(0); const std::vector<float>& h = c.dev->ensure_host(x); f(x, h); c.append_output_view(x); }
  void grad(GradientContext& c) override { c.append_input_grad(c.output_grad()); }
};
Tensor T::nth_tensor(Tensor x, int n) { return TensorBuilder(x.graph).append_input_with_selector(x, false, n).build(new IdentityOp()); }
Tensor T::identity(Tensor x) { return TensorBuilder(x.graph).append_input(x, false).set_shape(shape(x)).build(new IdentityOp()); }
Tensor T::stop_gradient(Tensor x) { return TensorBuilder(x.graph).append_input(x, false).set_differentiable(false).build(new StopGradient()); }
Tensor T::hook(Tensor x, std::function<void(const NdArray&, const std::vector<float>&)> f) { auto* op = new HookOp(); op->f = std::move(f); return TensorBuilder(x.graph).append_input(x, false).build(op); }
Tensor T::control_dependencies(Tensor x, const std::vector<Tensor>& deps) {      // mod.rs:2951-2971: rewires input 0 in place
  Graph* g = x.graph;
  if (g->inner(x.id).incoming_nodes.empty()) throw Panic("Source tensor cannot depend on any other tensors.");
  TensorBuilder b(g); b.append_input(g->tensor(g->inner(x.id).incoming_nodes[0].id), false);
  for (auto& d : deps) b.append_input(d, false);
  Tensor n = b.build(new ControlDependency());
  g->inner(x.id).incoming_nodes[0].id = n.id;
  return x;
}

// ================================================================================================ binary arithmetic
static Tensor maybe_reduce(Tensor target_shape, Tensor x, Graph* g);
static bool scalar_value(const NdArray& a, float* v) {          // rank-0 (or [1]) array whose value is known on the host
  if (a.size() == 1 && a.has_host() && a.ndim() <= 1) { *v = (*a.host)[0]; return true; }
  return false;
}
struct BinArith : Op {                 // AddOp/SubOp/MulOp/DivOp, binary_ops.rs:147-290,304-347
  int kind;
  explicit BinArith(int k) : kind(k) {}
  const char* name() const override {
    return kind == AGB_B_ADD ? REFNAME("binary_ops", "AddOp") : kind == AGB_B_SUB ? REFNAME("binary_ops", "SubOp") : kind == AGB_B_MUL ? REFNAME("binary_ops", "MulOp") : REFNAME("binary_ops", "DivOp");
  }
  void compute(ComputeContext& c) override {
    c.accept_lazy = (kind == AGB_B_ADD || kind == AGB_B_MUL); c.accept_expr = true;
    NdArray a = c.input(0), b = c.input(1);
    if (a.lazy || b.lazy) {
      if (a.expr) a = expr_materialize(c.dev, a);
      if (b.expr) b = expr_materialize(c.dev, b);
      if (kind == AGB_B_ADD) {            // Conv2D + bias[1,O,1,1]: stays deferred, the bias joins the conv epilogue
        NdArray bb = lazy_is_conv(a) ? b : a; if (!bb.lazy) c.dev->ensure_device(bb);
        NdArray r = lazy_is_conv(a) && !b.lazy ? lazy_conv_add_bias(a, bb) : (lazy_is_conv(b) && !a.lazy ? lazy_conv_add_bias(b, bb) : NdArray());
        if (r.lazy) { c.append_output(r); return; }
      } else {                            // (x > 0) * gy: one ReLU-grad kernel instead of compare + multiply
        const NdArray& m = lazy_is_mask(a) ? a : b; NdArray o = lazy_is_mask(a) ? b : a;
        if (lazy_is_mask(m) && o.shape == m.shape && (o.lazy || o.on_device())) {
          if (o.lazy) {                   // the producer of gy is itself deferred (dgrad / pool backward): fuse the mask into it
            NdArray r = lazy_fuse_mask(c.dev, m, o);
            if (r.on_device()) { c.append_output(r); return; }
            o = materialize_lazy(c.dev, o);
          }
          c.append_output(dev_binary(c.dev, AGB_B_RELU_GRAD, lazy_mask_src(c.dev, m), o)); return;
        }
      }
      if (a.lazy) a = materialize_lazy(c.dev, a);
      if (b.lazy) b = materialize_lazy(c.dev, b);
    }
    if (all_meta(a) && all_meta(b)) { c.append_output(host_binary(kind, a, b)); return; }
    float s;
    // every device path first tries to join a pending elementwise expression (fuse.cc); the single-op kernel is the fallback
    auto unary = [&](int op, const NdArray& x, float p) { NdArray r = expr_unary(c, op, p, x); return r.expr ? r : dev_unary(c.dev, op, x, p); };
    const bool a_sc = a.ndim() == 0 && scalar_value(a, &s);
    if (a_sc && b.size() != 1) {       // scalar (op) tensor fast paths: the scalar travels as a kernel parameter
      NdArray y = kind == AGB_B_ADD ? unary(AGB_U_ADD_SCALAR, b, s) : kind == AGB_B_SUB ? unary(AGB_U_RSUB_SCALAR, b, s)
                : kind == AGB_B_MUL ? unary(AGB_U_SCALE, b, s) : unary(AGB_U_RDIV_SCALAR, b, s);
      c.append_output(y); return;
    }
    const bool b_sc = (b.ndim() == 0 || (kind == AGB_B_DIV && b.ndim() == 1 && b.shape[0] == 1)) && scalar_value(b, &s);
    if (b_sc && !(all_meta(a))) {
      NdArray y = kind == AGB_B_ADD ? unary(AGB_U_ADD_SCALAR, a, s) : kind == AGB_B_SUB ? unary(AGB_U_ADD_SCALAR, a, -s)
                : kind == AGB_B_MUL ? unary(AGB_U_SCALE, a, s) : unary(AGB_U_SCALE, a, 1.0f / s);   // Div by scalar = multiply by reciprocal (:251-255)
      c.append_output(y); return;
    }
    if (a.ndim() == b.ndim() && a.ndim() > 0) { NdArray r = expr_binary(c, kind, a, b); if (r.expr) { c.append_output(r); return; } }
    c.append_output(dev_binary(c.dev, kind, a, b, 0.f, 0.f, name()));
  }
  void grad(GradientContext& c) override {
    Graph* g = c.graph(); Tensor x0 = c.input(0), x1 = c.input(1), gy = c.output_grad();
    Tensor s0 = T::shape(x0), s1 = T::shape(x1);
    if (kind == AGB_B_ADD) { c.append_input_grad(maybe_reduce(s0, gy, g)); c.append_input_grad(maybe_reduce(s1, gy, g)); }
    else if (kind == AGB_B_SUB) { c.append_input_grad(maybe_reduce(s0, gy, g)); c.append_input_grad(T::unary("neg", maybe_reduce(s1, gy, g))); }
    else if (kind == AGB_B_MUL) { Tensor g0 = T::mul(gy, x1), g1 = T::mul(gy, x0); c.append_input_grad(maybe_reduce(s0, g0, g)); c.append_input_grad(maybe_reduce(s1, g1, g)); }
    else {
      Tensor g0 = T::div(gy, x1), g1 = T::mul(T::mul(T::unary("neg", x0), T::unary("pow", x1, -2.f)), gy);
      c.append_input_grad(maybe_reduce(s0, g0, g)); c.append_input_grad(maybe_reduce(s1, g1, g));
    }
  }
};
static Tensor bin(int kind, Tensor a, Tensor b) {
  Graph* g = a.graph;
  return TensorBuilder(g).set_shape(infer_bin_op_shape(g, T::shape(a), T::shape(b))).append_input(a, false).append_input(b, false).build(new BinArith(kind));
}
Tensor T::add(Tensor a, Tensor b) { return bin(AGB_B_ADD, a, b); }
Tensor T::sub(Tensor a, Tensor b) { return bin(AGB_B_SUB, a, b); }
Tensor T::mul(Tensor a, Tensor b) { return bin(AGB_B_MUL, a, b); }
Tensor T::div(Tensor a, Tensor b) { return bin(AGB_B_DIV, a, b); }

struct MaybeBroadcast;
struct MaybeReduceSum : Op {           // binary_ops.rs:39-105
  const char* name() const override { return REFNAME("binary_ops", "MaybeReduceSum"); }
  void compute(ComputeContext& c) override {
    c.accept_expr = true;
    NdArray gy = c.input(0), sh = c.input(1);
    Shape orig_ = as_shape(c.dev, sh);
    if (orig_ == gy.shape) { if (gy.expr) c.append_output(expr_passthrough(c, gy)); else c.append_output_view(gy); return; }
    if (gy.expr) gy = expr_materialize(c.dev, gy);
    bool target_scalar = is_scalar_shape(orig_);
    Shape orig = target_scalar ? Shape(gy.shape.size(), 1) : orig_;
    if (orig == gy.shape) { c.append_output_view(d_reshape(c, gy, orig_)); return; }
    if (orig.size() != gy.shape.size()) throw Panic("bug of MaybeReduceSum probably");
    std::vector<int> axes;
    for (size_t i = 0; i < orig.size(); i++) {
      if (orig[i] == 1 && gy.shape[i] > 1) axes.push_back((int)i);
      else if (orig[i] != gy.shape[i]) throw Panic("bug of MaybeReduceSum probably");
    }
    Shape fin = orig_.size() == 1 && orig_[0] == 0 ? Shape{} : orig_;     // shape [0] (scalar_shape) denotes a 0-d target
    if (c.run->fuse && c.run->sole_consumer_sums(c.node) && axes == std::vector<int>({0})) { NdArray r = expr_colsum(c, on_dev(c.dev, gy), fin); if (r.expr) { c.append_output(r); return; } }
    if (gy.chan_sum && gy.ndim() == 4 && axes == std::vector<int>({0, 2, 3}) && gy.chan_sum->size() == gy.shape[1]) {
      c.append_output(gy.chan_sum->reshaped(fin)); return;               // bias gradient already produced by the fused epilogue that wrote gy
    }
    NdArray r = dev_reduce_axes(c.dev, AGB_R_SUM, gy, axes, true);
    if (!r.is_contiguous()) r = c.dev->contiguous(r);
    c.append_output(r.reshaped(fin));
  }
  static NdArray d_reshape(ComputeContext& c, NdArray a, const Shape& s) {
    Shape fin = s.size() == 1 && s[0] == 0 ? Shape{} : s;
    if (!a.is_contiguous()) a = c.dev->contiguous(a);
    return a.reshaped(fin);
  }
  void grad(GradientContext& c) override;
};
struct MaybeBroadcast : Op {           // binary_ops.rs:108-145
  const char* name() const override { return REFNAME("binary_ops", "MaybeBroadcast"); }
  void compute(ComputeContext& c) override {
    NdArray sh = c.input(1); Shape target = as_shape(c.dev, sh);
    NdArray x = c.input(0);
    if (x.shape == target) { c.append_output_view(x); return; }
    if (is_scalar_shape(x.shape)) x = x.reshaped(Shape(target.size(), 1));
    if (x.shape.size() != target.size()) throw OpError(AGB_ERR_INCOMPATIBLE_SHAPE, "PreprocessBinOpGradGrad: Can't broadcast.");
    for (size_t i = 0; i < target.size(); i++) if (x.shape[i] != target[i] && x.shape[i] != 1) throw OpError(AGB_ERR_INCOMPATIBLE_SHAPE, "PreprocessBinOpGradGrad: Can't broadcast.");
    c.append_output(dev_broadcast_to(c.dev, x, target));
  }
  void grad(GradientContext& c) override { c.append_input_grad(maybe_reduce(T::shape(c.input(0)), c.output_grad(), c.graph())); c.append_none(); }
};
void MaybeReduceSum::grad(GradientContext& c) {
  Tensor gx = TensorBuilder(c.graph()).append_input(c.output_grad(), false).append_input(T::shape(c.input(0)), false).build(new MaybeBroadcast());
  c.append_input_grad(gx); c.append_none();
}
static Tensor maybe_reduce(Tensor target_shape, Tensor x, Graph* g) {      // binary_ops.rs:292-302
  return TensorBuilder(g).append_input(x, false).append_input(target_shape, false).set_shape(target_shape).build(new MaybeReduceSum());
}

// ================================================================================================ unary math / activations
struct UnaryInfo { const char* fn; int op; const char* ref; };
static const UnaryInfo UNARY[] = {
  {"sin", AGB_U_SIN, REFNAME("math_ops", "Sin")}, {"cos", AGB_U_COS, REFNAME("math_ops", "Cos")}, {"tan", AGB_U_TAN, REFNAME("math_ops", "Tan")},
  {"asin", AGB_U_ASIN, REFNAME("math_ops", "Asin")}, {"acos", AGB_U_ACOS, REFNAME("math_ops", "Acos")}, {"atan", AGB_U_ATAN, REFNAME("math_ops", "Atan")},
  {"sinh", AGB_U_SINH, REFNAME("math_ops", "Sinh")}, {"cosh", AGB_U_COSH, REFNAME("math_ops", "Cosh")}, {"tanh", AGB_U_TANH, REFNAME("math_ops", "Tanh")},
  {"asinh", AGB_U_ASINH, REFNAME("math_ops", "Asinh")}, {"acosh", AGB_U_ACOSH, REFNAME("math_ops", "Acosh")}, {"atanh", AGB_U_ATANH, REFNAME("math_ops", "Atanh")},
  {"exp", AGB_U_EXP, REFNAME("math_ops", "Exp")}, {"exp2", AGB_U_EXP2, REFNAME("math_ops", "Exp2")}, {"exp10", AGB_U_EXP10, REFNAME("math_ops", "Exp10")},
  {"ln", AGB_U_LN, REFNAME("math_ops", "Ln")}, {"log2", AGB_U_LOG2, REFNAME("math_ops", "Log2")}, {"log10", AGB_U_LOG10, REFNAME("math_ops", "Log10")},
  {"sqrt", AGB_U_SQRT, REFNAME("math_ops", "Sqrt")}, {"pow", AGB_U_POW, REFNAME("math_ops", "Pow")}, {"neg", AGB_U_NEG, REFNAME("math_ops", "NegOp")},
  {"abs", AGB_U_ABS, REFNAME("math_ops", "Abs")}, {"sign", AGB_U_SIGN, REFNAME("math_ops", "Sign")}, {"floor", AGB_U_FLOOR, REFNAME("math_ops", "Floor")},
  {"ceil", AGB_U_CEIL, REFNAME("math_ops", "Ceil")}, {"inv", AGB_U_INV, REFNAME("math_ops", "Inv")}, {"inv_sqrt", AGB_U_INVSQRT, REFNAME("math_ops", "InvSqrt")},
  {"square", AGB_U_SQUARE, REFNAME("math_ops", "Square")}, {"sigmoid", AGB_U_SIGMOID, REFNAME("activation_ops", "Sigmoid")},
  {"relu", AGB_U_RELU, REFNAME("activation_ops", "ReLU")}, {"softplus", AGB_U_SOFTPLUS, REFNAME("activation_ops", "Softplus")},
  {"elu", AGB_U_ELU, REFNAME("activation_ops", "ELU")},
  {"lgamma", AGB_U_LGAMMA, REFNAME("math_ops", "Lgamma")}, {"digamma", AGB_U_DIGAMMA, REFNAME("math_ops", "Digamma")},      // math_ops.rs:1021-1060
};
struct ELUGrad : Op {                  // activation_ops.rs:204-226
  float alpha;
  const char* name() const override { return REFNAME("activation_ops", "ELUGrad"); }
  void compute(ComputeContext& c) override { NdArray x = c.input(0), gy = c.input(1); c.append_output(dev_binary(c.dev, AGB_B_ELU_GRAD, x, gy, alpha)); }
  void grad(GradientContext& c) override { c.append_none(); c.append_none(); }
};
struct UnaryOp : Op {
  const UnaryInfo* info; float p0;
  const char* name() const override { return info->ref; }
  void compute(ComputeContext& c) override {
    c.accept_lazy = info->op == AGB_U_RELU; c.accept_expr = true;
    NdArray x = c.input(0);
    if (x.lazy) {
      if (lazy_is_conv(x)) { NdArray y = lazy_conv_relu(c.dev, x); if (y.lazy) { c.append_output(y); return; } }      // stays deferred (ops_nn.cc)
      x = materialize_lazy(c.dev, x);
    }
    if (info->op == AGB_U_NEG && all_meta(x)) { std::vector<float> v = *x.host; for (auto& e : v) e = -e; c.append_output(NdArray::from_host(x.shape, v, true)); return; }
    if (!all_meta(x)) { NdArray r = expr_unary(c, info->op, p0, x); if (r.expr) { c.append_output(r); return; } }
    c.append_output(dev_unary(c.dev, info->op, x, p0));
  }
  void grad(GradientContext& c) override {
    using namespace T;
    Graph* g = c.graph(); Tensor x = c.input(0), y = c.output(), gy = c.output_grad();
    auto S = [&](float v) { return scalar(g, v); };
    auto U = [&](const char* n, Tensor t, float p = 0.f) { return unary(n, t, p); };
    Tensor gx;
    switch (info->op) {                 // math_ops.rs / activation_ops.rs Op::grad bodies
      case AGB_U_SIN: gx = mul(U("cos", x), gy); break;
      case AGB_U_COS: gx = U("neg", mul(U("sin", x), gy)); break;
      case AGB_U_TAN: gx = div(gy, U("square", U("cos", x))); break;
      case AGB_U_ASIN: gx = mul(U("inv_sqrt", sub(S(1.f), U("square", x))), gy); break;
      case AGB_U_ACOS: gx = mul(U("neg", U("inv_sqrt", sub(S(1.f), U("square", x)))), gy); break;
      case AGB_U_ATAN: gx = mul(U("inv", add(U("square", x), S(1.f))), gy); break;
      case AGB_U_SINH: gx = mul(U("cosh", x), gy); break;
      case AGB_U_COSH: gx = mul(U("sinh", x), gy); break;
      case AGB_U_TANH: gx = mul(gy, sub(S(1.f), U("square", y))); break;
      case AGB_U_ASINH: gx = mul(U("inv", U("sqrt", add(U("square", x), S(1.f)))), gy); break;
      case AGB_U_ACOSH: gx = mul(U("inv", U("sqrt", sub(U("square", x), S(1.f)))), gy); break;
      case AGB_U_ATANH: gx = mul(U("inv", sub(S(1.f), U("square", x))), gy); break;
      case AGB_U_EXP: gx = mul(y, gy); break;
      case AGB_U_EXP2: gx = mul(mul(S(logf(2.f)), y), gy); break;
      case AGB_U_EXP10: gx = mul(mul(S(logf(10.f)), y), gy); break;
      case AGB_U_LN: gx = div(gy, x); break;
      case AGB_U_LOG2: gx = div(gy, mul(S(logf(2.f)), x)); break;
      case AGB_U_LOG10: gx = div(gy, mul(S(logf(10.f)), x)); break;
      case AGB_U_SQRT: gx = mul(gy, mul(S(0.5f), U("pow", x, -0.5f))); break;
      case AGB_U_POW: gx = mul(mul(gy, S(p0)), U("pow", x, p0 - 1.f)); break;
      case AGB_U_NEG: gx = U("neg", gy); break;
      case AGB_U_ABS: gx = mul(gy, U("sign", x)); break;
      case AGB_U_INV: gx = mul(U("neg", U("square", y)), gy); break;
      case AGB_U_INVSQRT: gx = mul(mul(S(-0.5f), U("pow", x, -1.5f)), gy); break;
      case AGB_U_SQUARE: gx = mul(mul(S(2.f), x), gy); break;
      case AGB_U_SIGMOID: gx = mul(gy, sub(y, U("square", y))); break;
      case AGB_U_RELU: gx = mul(cmp("greater", x, S(0.f)), gy); break;
      case AGB_U_SOFTPLUS: { Tensor a = U("exp", x); gx = mul(gy, div(a, add(a, S(1.f)))); break; }
      case AGB_U_LGAMMA: gx = mul(gy, U("digamma", x)); break;            // math_ops.rs:1047-1052 (Digamma itself has no gradient)
      case AGB_U_ELU: { auto* op = new ELUGrad(); op->alpha = p0; gx = TensorBuilder(g).append_input(x, false).append_input(gy, false).set_shape(shape(gy)).build(op); break; }
      default: c.append_none(); return;    // Sign / Floor / Ceil: None
    }
    c.append_input_grad(gx);
  }
};
Tensor T::unary(const std::string& nm, Tensor x, float p0) {
  for (auto& u : UNARY) if (nm == u.fn) {
    auto* op = new UnaryOp(); op->info = &u; op->p0 = p0;
    TensorBuilder b(x.graph); b.append_input(x, false);
    if (nm != "neg") b.set_shape(shape(x));             // mod.rs: every unary constructor but `neg` sets the shape
    return b.build(op);
  }
  throw Panic("unknown unary op: " + nm);
}

struct ClipGrad : Op {                 // array_ops.rs:556-574
  float lo, hi;
  const char* name() const override { return REFNAME("array_ops", "ClipGrad"); }
  void compute(ComputeContext& c) override { NdArray x = c.input(0), gy = c.input(1); c.append_output(dev_binary(c.dev, AGB_B_CLIP_GRAD, x, gy, lo, hi)); }
  void grad(GradientContext& c) override { c.append_none(); c.append_none(); }
};
struct Clip : Op {                     // array_ops.rs:537-554
  float lo, hi;
  const char* name() const override { return REFNAME("array_ops", "Clip"); }
  void compute(ComputeContext& c) override { c.append_output(dev_unary(c.dev, AGB_U_CLIP, c.input(0), lo, hi)); }
  void grad(GradientContext& c) override {
    auto* op = new ClipGrad(); op->lo = lo; op->hi = hi; Tensor gy = c.output_grad();
    c.append_input_grad(TensorBuilder(c.graph()).set_shape(T::shape(gy)).append_input(c.input(0), false).append_input(gy, false).build(op));
  }
};
Tensor T::clip(Tensor x, float lo, float hi) { auto* op = new Clip(); op->lo = lo; op->hi = hi; return TensorBuilder(x.graph).append_input(x, false).build(op); }

// ================================================================================================ compare / select
struct CmpInfo { const char* fn; int op; const char* ref; };
static const CmpInfo CMP[] = {
  {"equal", AGB_B_EQ, REFNAME("math_ops", "Equal")}, {"not_equal", AGB_B_NE, REFNAME("math_ops", "NotEqual")}, {"greater", AGB_B_GT, REFNAME("math_ops", "Greater")},
  {"lesser", AGB_B_LT, REFNAME("math_ops", "Lesser")}, {"greater_equal", AGB_B_GE, REFNAME("math_ops", "GreaterEqual")}, {"lesser_equal", AGB_B_LE, REFNAME("math_ops", "LesserEqual")},
  {"maximum", AGB_B_MAX, REFNAME("math_ops", "Maximum")}, {"minimum", AGB_B_MIN, REFNAME("math_ops", "Minimum")},
};
struct CmpOp : Op {                    // impl_cmp_op!, math_ops.rs:86-184
  const CmpInfo* info;
  const char* name() const override { return info->ref; }
  void compute(ComputeContext& c) override {
    c.accept_lazy = info->op == AGB_B_GT; c.accept_expr = true;
    NdArray a = c.input(0), b = c.input(1);
    if (info->op == AGB_B_GT && !a.expr && !b.expr) {          // greater(x, scalar 0): the ReLU-gradient mask; deferred until its multiply arrives
      float z;
      if (!b.lazy && b.ndim() == 0 && scalar_value(b, &z) && z == 0.0f && a.ndim() > 0 && (a.lazy || a.on_device())) {
        NdArray r = lazy_gt0_mask(a);
        if (r.lazy) { c.append_output(r); return; }
      }
      if (a.lazy) a = materialize_lazy(c.dev, a);
      if (b.lazy) b = materialize_lazy(c.dev, b);
    }
    bool as = is_scalar_shape(a.shape), bs = is_scalar_shape(b.shape);
    if (!as && !bs) {
      if (a.ndim() != b.ndim()) throw Panic(std::string("Tensor ranks mismatch: ") + info->ref);
      if (a.size() > b.size()) throw Panic(std::string("Tensor ranks mismatch: ") + info->ref);     // only lhs -> rhs broadcasting (:134-148)
      if (a.size() == b.size() && a.shape != b.shape) throw Panic(std::string("shape mismatch: ") + info->ref);
    }
    if (a.lazy) a = materialize_lazy(c.dev, a);
    if (b.lazy) b = materialize_lazy(c.dev, b);
    if (!as && !bs && !(all_meta(a) && all_meta(b))) {
      NdArray r = expr_binary(c, info->op, a, b);
      if (r.expr) { c.append_output(r); return; }
    } else if (as != bs && a.ndim() + b.ndim() > 0) {            // compare against a host-known scalar: it travels as an immediate
      float sv; const NdArray& sc = as ? a : b; const NdArray& full = as ? b : a;
      if (sc.ndim() == 0 && scalar_value(sc, &sv) && full.ndim() > 0 && !all_meta(full)) {
        NdArray r = expr_binary_imm(c, info->op, full, sv, as);
        if (r.expr) { c.append_output(r); return; }
      }
    }
    NdArray y = dev_binary(c.dev, info->op, a, b, 0.f, 0.f, info->ref);
    if (as && bs) y = y.reshaped(a.shape == Shape{0} ? Shape{} : a.shape);
    c.append_output(y);
  }
  void grad(GradientContext& c) override {
    if (info->op == AGB_B_MAX || info->op == AGB_B_MIN) {        // min_max_grad, math_ops.rs:198-209
      Tensor gy = c.output_grad(), y = c.output();
      c.append_input_grad(T::mul(T::cmp("equal", c.input(0), y), gy));
      c.append_input_grad(T::mul(T::cmp("equal", c.input(1), y), gy));
    } else c.append_none();                                      // none_grad appends a single None (:186-195)
  }
};
Tensor T::cmp(const std::string& nm, Tensor a, Tensor b) {
  for (auto& u : CMP) if (nm == u.fn) { auto* op = new CmpOp(); op->info = &u; return TensorBuilder(a.graph).append_input(a, false).append_input(b, false).build(op); }
  throw Panic("unknown compare op: " + nm);
}

// ================================================================================================ AddN
struct AddN : Op {                     // array_ops.rs:503-535
  const char* name() const override { return REFNAME("array_ops", "AddN"); }
  void compute(ComputeContext& c) override {
    int n = c.num_inputs();
    if (n == 1) { c.append_output_view(c.input(0)); return; }
    c.accept_expr = true;
    std::vector<NdArray> xs; bool same = true, all_empty_scalars = true;
    for (int i = 0; i < n; i++) { xs.push_back(c.input(i)); if (xs[i].shape != xs[0].shape) same = false; if (!(xs[i].ndim() == 0 && xs[i].has_host() && !xs[i].on_device())) all_empty_scalars = false; }
    if (same && !all_empty_scalars) { NdArray y; if (expr_sum_pads(c, xs, &y) || expr_sum_gemms(c, xs, &y) || expr_sum_scatters(c, xs, &y) || expr_sum_colsums(c, xs, &y)) { c.append_output(y); return; } }
    if (same && !all_empty_scalars && n <= 6 && xs[0].ndim() > 0) {      // a short sum joins the pending expression as the same left fold
      NdArray acc = expr_binary(c, AGB_B_ADD, xs[0], xs[1]);
      for (int i = 2; i < n && acc.expr; i++) acc = expr_binary(c, AGB_B_ADD, acc, xs[i]);
      if (acc.expr) { c.append_output(acc); return; }
    }
    for (auto& x : xs) if (x.expr) x = expr_materialize(c.dev, x);
    if (all_empty_scalars) {           // sum of optimizer-op placeholders (get_update_op = add_n(update_ops), optimizers/mod.rs:87-98)
      float s = 0.f; for (auto& x : xs) s += (*x.host)[0];
      c.append_output(NdArray::scalar_host(s)); return;
    }
    if (same) {
      std::vector<agb_tensor> ds; std::vector<const agb_tensor*> ps;
      for (auto& x : xs) c.dev->ensure_device(x);
      std::vector<int> order; bool same_layout = xs[0].dense_order(order) && !is_identity(order);
      for (auto& x : xs) if (x.stride != xs[0].stride) same_layout = false;
      if (same_layout) {          // all inputs share one permuted dense layout: add the flat buffers
        for (auto& x : xs) ds.push_back(flat_desc(x));
        for (auto& dd : ds) ps.push_back(&dd);
        NdArray y = c.dev->empty_ordered(xs[0].shape, order); agb_tensor ty = flat_desc(y);
        check_status(agb_add_n(c.dev->ctx, n, ps.data(), &ty));
        c.append_output(y); return;
      }
      for (auto& x : xs) { x = c.dev->contiguous(x); ds.push_back(x.desc()); }
      for (auto& d : ds) ps.push_back(&d);
      NdArray y = c.dev->empty(xs[0].shape); agb_tensor ty = y.desc();
      check_status(agb_add_n(c.dev->ctx, n, ps.data(), &ty));
      c.append_output(y); return;
    }
    NdArray acc = dev_binary(c.dev, AGB_B_ADD, xs[0], xs[1]);    // broadcasting left fold, like `&a + &b; base += ..`
    for (int i = 2; i < n; i++) acc = dev_binary(c.dev, AGB_B_ADD, acc, xs[i]);
    c.append_output(acc);
  }
  void grad(GradientContext& c) override { for (int i = 0; i < c.num_inputs(); i++) c.append_input_grad(c.output_grad()); }
  bool sums_inputs() const override { return true; }
};
Tensor T::add_n(const std::vector<Tensor>& xs) {
  if (xs.empty()) throw Panic("add_n: empty input");
  if (xs.size() == 1) return xs[0];
  TensorBuilder b(xs[0].graph);
  for (auto& x : xs) b.append_input(x, false);
  return b.set_shape(shape(xs[0])).build(new AddN());
}

// ================================================================================================ reductions
static const char* REDUCE_REF[] = {REFNAME("reduction_ops", "ReduceSum"), REFNAME("reduction_ops", "ReduceMean"), REFNAME("reduction_ops", "ReduceProd"),
                                   REFNAME("reduction_ops", "ReduceMin"), REFNAME("reduction_ops", "ReduceMax")};
struct ReduceGradCommon : Op {         // reduction_ops.rs:459-512
  bool should_make_broadcast_dims;
  const char* name() const override { return REFNAME("reduction_ops", "ReduceGradCommon"); }
  void compute(ComputeContext& c) override {
    NdArray gy = c.input(0), sh = c.input(1);
    Shape target = as_shape(c.dev, sh);
    if (gy.shape == target) { c.append_output_view(gy); return; }
    if (should_make_broadcast_dims || is_scalar_shape(gy.shape)) {
      NdArray ax = c.input(2);
      std::vector<int> axes = norm_axes(c.dev, ax, (int)target.size());
      std::sort(axes.begin(), axes.end());
      Shape gs = gy.shape; if (gs.size() == 1 && gs[0] == 0) gs.clear();
      for (int a : axes) { if (a > (int)gs.size()) throw Panic("ReduceGradCommon: bad axes"); gs.insert(gs.begin() + a, 1); }
      if (!gy.is_contiguous()) gy = c.dev->contiguous(gy);
      gy = gy.reshaped(gs);
    }
    if (gy.shape.size() != target.size()) throw Panic("ReduceGradCommon: cannot broadcast");
    c.append_output(dev_broadcast_to(c.dev, gy, target));
  }
  void grad(GradientContext& c) override;
};
struct ReduceOp : Op {                 // ReduceSum/Mean/Prod/Min/Max, reduction_ops.rs:161-330
  int kind; bool keep_dims;
  const char* name() const override { return REDUCE_REF[kind]; }
  void compute(ComputeContext& c) override {
    NdArray x = c.input(0), ax = c.input(1);
    if (is_scalar_shape(x.shape)) { c.append_output_view(x); return; }                           // view of the input (:63-64)
    std::vector<int> axes = norm_axes(c.dev, ax, x.ndim());
    if (axes.empty()) { c.append_output_view(x); return; }                                       // (:66-68, 193-196)
    if (all_meta(x)) { c.append_output(host_reduce_axes(kind, x, axes, keep_dims)); return; }
    if (kind == AGB_R_MEAN) {          // sum, then multiply by 1/len with len accumulated in f32 (:198-209)
      float len = 1.f; for (int a : axes) len *= (float)x.shape[a];
      NdArray s = dev_reduce_axes(c.dev, AGB_R_SUM, x, axes, keep_dims);
      c.append_output(dev_unary(c.dev, AGB_U_SCALE, s, 1.0f / len));
    } else c.append_output(dev_reduce_axes(c.dev, kind, x, axes, keep_dims));
  }
  void grad(GradientContext& c) override {
    Graph* g = c.graph(); Tensor x = c.input(0), axes = c.input(1), gy = c.output_grad();
    auto rgc = [&](Tensor t) { auto* op = new ReduceGradCommon(); op->should_make_broadcast_dims = !keep_dims;
                               return TensorBuilder(g).append_input(t, false).append_input(T::shape(x), false).append_input(axes, false).build(op); };
    if (kind == AGB_R_SUM) c.append_input_grad(rgc(gy));
    else if (kind == AGB_R_MEAN) {     // :217-238
      Tensor reduction_len = T::reduce("prod", T::gather_common(T::shape(x), axes, 0), T::as_tensor(g, {0}), false);
      c.append_input_grad(T::div(rgc(gy), reduction_len));
    } else if (kind == AGB_R_PROD) c.append_input_grad(T::div(rgc(T::mul(gy, c.output())), x));     // :256-273
    else c.append_input_grad(T::mul(T::cmp("equal", x, rgc(c.output())), rgc(gy)));                  // min_max_grad :332-363 (ties: every position)
    c.append_none();
  }
};
void ReduceGradCommon::grad(GradientContext& c) {
  auto* op = new ReduceOp(); op->kind = AGB_R_SUM; op->keep_dims = should_make_broadcast_dims;      // (sic) :498-511
  c.append_input_grad(TensorBuilder(c.graph()).append_input(c.output_grad(), false).append_input(c.input(2), false).build(op));
  c.append_none(); c.append_none();
}
Tensor T::reduce(const std::string& nm, Tensor x, Tensor axes, bool keep_dims) {
  static const char* names[] = {"sum", "mean", "prod", "min", "max"};
  for (int k = 0; k < 5; k++) if (nm == names[k]) { auto* op = new ReduceOp(); op->kind = k; op->keep_dims = keep_dims; return TensorBuilder(x.graph).append_input(x, false).append_input(axes, false).build(op); }
  throw Panic("unknown reduction: " + nm);
}

struct ReduceSumToScalarGrad;
struct ReduceSumToScalar : Op {        // reduction_ops.rs:123-137
  const char* name() const override { return REFNAME("reduction_ops", "ReduceSumToScalar"); }
  void compute(ComputeContext& c) override {
    NdArray x = c.input(0);
    if (all_meta(x)) { float s = 0; for (auto v : *x.host) s += v; c.append_output(NdArray::scalar_host(s, true)); return; }
    x = c.dev->contiguous(on_dev(c.dev, x));
    NdArray y = c.dev->empty({});
    if (x.size() == 0) { c.append_output(c.dev->zeros({})); return; }
    check_status(agb_reduce(c.dev->ctx, AGB_R_SUM, x.dptr, y.dptr, 1, x.size(), 1));
    c.append_output(y);
  }
  void grad(GradientContext& c) override;
};
struct ReduceSumToScalarGrad : Op {    // reduction_ops.rs:139-159
  const char* name() const override { return REFNAME("reduction_ops", "ReduceSumToScalarGrad"); }
  void compute(ComputeContext& c) override {
    NdArray sh = c.input(1); Shape shp = as_shape(c.dev, sh);
    NdArray gy = c.input(0); float v;
    if (scalar_value(gy, &v)) { c.append_output(c.dev->full(shp, v)); return; }
    c.append_output(dev_broadcast_to(c.dev, gy.reshaped(Shape(shp.size(), 1)), shp));
  }
  void grad(GradientContext& c) override { c.append_input_grad(TensorBuilder(c.graph()).append_input(c.output_grad(), false).build(new ReduceSumToScalar())); c.append_none(); }
};
void ReduceSumToScalar::grad(GradientContext& c) {
  c.append_input_grad(TensorBuilder(c.graph()).append_input(c.output_grad(), false).append_input(T::shape(c.input(0)), false).build(new ReduceSumToScalarGrad()));
}
Tensor T::sum_all(Tensor x) { return TensorBuilder(x.graph).append_input(x, false).build(new ReduceSumToScalar()); }
Tensor T::mean_all(Tensor x) { return div(sum_all(x), size(x)); }

struct ArgOp : Op {                    // ArgMax / ArgMin, reduction_ops.rs:365-457
  bool is_max; int axis; bool keep_dim;
  const char* name() const override { return is_max ? REFNAME("reduction_ops", "ArgMax") : REFNAME("reduction_ops", "ArgMin"); }
  void compute(ComputeContext& c) override {
    NdArray x = c.dev->contiguous(on_dev(c.dev, c.input(0)));
    int ax = normalize_negative_axis(axis, x.ndim());
    if (ax < 0 || ax >= x.ndim()) throw Panic("argmax/argmin: axis out of range");
    int64_t outer = 1, inner = 1; for (int k = 0; k < ax; k++) outer *= x.shape[k]; for (int k = ax + 1; k < x.ndim(); k++) inner *= x.shape[k];
    Shape ns; for (int k = 0; k < x.ndim(); k++) { if (k != ax) ns.push_back(x.shape[k]); else if (keep_dim) ns.push_back(1); }
    NdArray y = c.dev->empty(ns);
    check_status(agb_argreduce(c.dev->ctx, is_max ? 1 : 0, x.dptr, y.dptr, outer, x.shape[ax], inner));
    c.append_output(y);
  }
  void grad(GradientContext& c) override { c.append_none(); }
};
Tensor T::argmax(Tensor x, int axis, bool keep) { auto* op = new ArgOp(); op->is_max = true; op->axis = axis; op->keep_dim = keep; return TensorBuilder(x.graph).append_input(x, false).build(op); }
Tensor T::argmin(Tensor x, int axis, bool keep) { auto* op = new ArgOp(); op->is_max = false; op->axis = axis; op->keep_dim = keep; return TensorBuilder(x.graph).append_input(x, false).build(op); }

// ================================================================================================ softmax family / xent
static void axis_view(const NdArray& x, int axis, int64_t& outer, int64_t& r, int64_t& inner) {
  outer = inner = 1; r = x.shape[axis];
  for (int k = 0; k < axis; k++) outer *= x.shape[k];
  for (int k = axis + 1; k < x.ndim(); k++) inner *= x.shape[k];
}
struct SoftmaxLike : Op {              // Softmax (activation_ops.rs:61-111), LogSoftmax (xent_ops.rs:17-31), LogSumExp (math_ops.rs:540-609)
  int kind; int axis; bool keep_dims;  // 0 softmax, 1 log_softmax, 2 logsumexp
  const char* name() const override { return kind == 0 ? REFNAME("activation_ops", "Softmax") : kind == 1 ? REFNAME("xent_ops", "LogSoftmax") : REFNAME("math_ops", "LogSumExp"); }
  void compute(ComputeContext& c) override {
    NdArray x = c.dev->contiguous(on_dev(c.dev, c.input(0)));
    int ax = normalize_negative_axis(axis, x.ndim());
    if (ax < 0 || ax >= x.ndim()) throw Panic("softmax: axis out of range");
    int64_t o, r, in; axis_view(x, ax, o, r, in);
    if (kind == 2) {
      Shape ns; for (int k = 0; k < x.ndim(); k++) { if (k != ax) ns.push_back(x.shape[k]); else if (keep_dims) ns.push_back(1); }
      NdArray y = c.dev->empty(ns);
      check_status(agb_logsumexp(c.dev->ctx, x.dptr, y.dptr, o, r, in));
      c.append_output(y);
    } else {
      NdArray y = c.dev->empty(x.shape);
      check_status((kind == 0 ? agb_softmax : agb_log_softmax)(c.dev->ctx, x.dptr, y.dptr, o, r, in));
      c.append_output(y);
    }
  }
  void grad(GradientContext& c) override {
    using namespace T; Graph* g = c.graph(); Tensor gy = c.output_grad(), y = c.output();
    if (kind == 0) { Tensor s = reduce("sum", mul(y, gy), as_tensor(g, {axis}), true); c.append_input_grad(mul(sub(gy, s), y)); }           // activation_ops.rs:105-110
    else if (kind == 1) c.append_input_grad(sub(gy, mul(unary("exp", y), reduce("sum", gy, as_tensor(g, {1}), true))));                     // xent_ops.rs:24-30 (axis 1 hard-coded)
    else c.append_input_grad(mul(softmax(c.input(0), axis), gy));                                                                           // math_ops.rs:602-608
  }
};
Tensor T::softmax(Tensor x, int axis) { auto* op = new SoftmaxLike(); op->kind = 0; op->axis = axis; op->keep_dims = true; return TensorBuilder(x.graph).append_input(x, false).build(op); }
Tensor T::log_softmax(Tensor x, int axis) { auto* op = new SoftmaxLike(); op->kind = 1; op->axis = axis; op->keep_dims = true; return TensorBuilder(x.graph).set_shape(shape(x)).append_input(x, false).build(op); }
Tensor T::reduce_logsumexp(Tensor x, int axis, bool keep) { auto* op = new SoftmaxLike(); op->kind = 2; op->axis = axis; op->keep_dims = keep; return TensorBuilder(x.graph).append_input(x, false).build(op); }

struct SparseSoftmaxCrossEntropyGrad : Op {   // xent_ops.rs:139-158
  const char* name() const override { return REFNAME("xent_ops", "SparseSoftmaxCrossEntropyGrad"); }
  void compute(ComputeContext& c) override {
    NdArray log_x = c.dev->contiguous(on_dev(c.dev, c.input(0))), t = c.dev->contiguous(on_dev(c.dev, c.input(1))), gy = c.input(2);
    if (log_x.ndim() != 2) throw OpError(AGB_ERR_INCOMPATIBLE_SHAPE, "SparseSoftmaxCrossEntropyGrad: log_x must be 2-D");
    int64_t B = log_x.shape[0], C = log_x.shape[1];
    float v; NdArray gyd;
    if (scalar_value(gy, &v)) gyd = c.dev->full({1}, v); else gyd = c.dev->contiguous(on_dev(c.dev, gy));
    if (gyd.size() != 1 && gyd.size() != B) gyd = dev_broadcast_to(c.dev, gyd, {B, 1});
    NdArray gx = c.dev->empty({B, C});
    check_status(agb_sparse_xent_bwd(c.dev->ctx, log_x.dptr, t.dptr, gyd.dptr, gyd.size(), gx.dptr, B, C));
    c.append_output(gx);
  }
  void grad(GradientContext& c) override { c.append_none(); c.append_none(); }
  // T time steps of an unrolled RNN: one launch over the stacked rows (the log_x blocks are slices of the stacked forward output)
  const char* stack_key() const override { return "sparse_xent_grad"; }
  bool compute_stacked(Device* dev, Evaluation& run, const std::vector<std::vector<NdArray>>& ins, std::vector<std::vector<NdArray>>* outs) override {
    std::vector<NdArray> xs, ts; const NdArray& gy0 = ins[0][2];
    for (auto& in : ins) {
      if (in.size() != 3 || !in[0].on_device() || in[0].ndim() != 2 || in[0].shape != ins[0][0].shape || !stackable(in[0]) || in[1].size() != in[0].shape[0] || !in[1].on_device()) return false;
      if (in[2].dptr != gy0.dptr || in[2].shape != gy0.shape || in[2].stride != gy0.stride || in[2].host != gy0.host) return false;      // one upstream gradient (the same array) for all
      xs.push_back(in[0]); ts.push_back(in[1]);
    }
    const int64_t n = (int64_t)ins.size(), B = xs[0].shape[0], C = xs[0].shape[1];
    float v; NdArray gyd;
    if (scalar_value(gy0, &v)) gyd = dev->full({1}, v);
    else if (gy0.size() == 1) gyd = dev->contiguous(on_dev(dev, gy0));
    else if (gy0.size() == B) {      // the same [B, 1] gradient for every step: tiled once
      NdArray g1 = dev->contiguous(on_dev(dev, gy0)); gyd = dev->empty({n * B});
      agb_tensor ts_, td; ts_.ptr = g1.dptr; ts_.rank = 2; ts_.shape[0] = n; ts_.shape[1] = B; ts_.stride[0] = 0; ts_.stride[1] = 1;
      td.ptr = gyd.dptr; td.rank = 2; td.shape[0] = n; td.shape[1] = B; td.stride[0] = B; td.stride[1] = 1;
      check_status(agb_copy_strided(dev->ctx, &ts_, &td));
    } else return false;
    NdArray X = stack_rows(run, dev, xs), L = stack_vectors(run, dev, ts), gx = dev->empty({n * B, C});
    check_status(agb_sparse_xent_bwd(dev->ctx, X.dptr, L.dptr, gyd.dptr, gyd.size(), gx.dptr, n * B, C));
    for (int64_t i = 0; i < n; i++) outs->push_back({gx.sliced(0, i * B, B)});
    return true;
  }
};
struct SparseSoftmaxCrossEntropy : Op {       // xent_ops.rs:63-137
  const char* name() const override { return REFNAME("xent_ops", "SparseSoftmaxCrossEntropy"); }
  void compute(ComputeContext& c) override {
    NdArray x = c.input(0), t = c.input(1);
    if (x.ndim() != 2) throw OpError(AGB_ERR_INCOMPATIBLE_SHAPE, "SparseSoftmaxCrossEntropy: given first argument's ndim is not 2");
    if (!(t.ndim() == 1 || (t.ndim() == 2 && t.shape[1] == 1)))
      throw OpError(AGB_ERR_INCOMPATIBLE_SHAPE, "SparseSoftmaxCrossEntropy: second argument's shape must be (batch_size, 1) or (batch_size,).");
    x = c.dev->contiguous(on_dev(c.dev, x)); t = c.dev->contiguous(on_dev(c.dev, t));
    int64_t B = x.shape[0], C = x.shape[1];
    if (t.size() != B) throw Panic("Batch size mismatch: inputs vs labels");
    NdArray loss = c.dev->empty({B, 1}), log_x = c.dev->empty({B, C});
    check_status(agb_sparse_xent_fwd(c.dev->ctx, x.dptr, t.dptr, loss.dptr, log_x.dptr, B, C));
    c.append_output(loss); c.append_output(log_x);
  }
  const char* stack_key() const override { return "sparse_xent"; }
  bool compute_stacked(Device* dev, Evaluation& run, const std::vector<std::vector<NdArray>>& ins, std::vector<std::vector<NdArray>>* outs) override {
    std::vector<NdArray> xs, ts;
    for (auto& in : ins) {
      if (in.size() != 2 || !in[0].on_device() || in[0].ndim() != 2 || in[0].shape != ins[0][0].shape || !stackable(in[0]) || !in[1].on_device()) return false;
      if (!(in[1].ndim() == 1 || (in[1].ndim() == 2 && in[1].shape[1] == 1)) || in[1].size() != in[0].shape[0]) return false;      // the members raise their own errors
      xs.push_back(in[0]); ts.push_back(in[1]);
    }
    const int64_t n = (int64_t)ins.size(), B = xs[0].shape[0], C = xs[0].shape[1];
    NdArray X = stack_rows(run, dev, xs), L = stack_vectors(run, dev, ts);
    NdArray loss = dev->empty({n * B, 1}), log_x = dev->empty({n * B, C});
    check_status(agb_sparse_xent_fwd(dev->ctx, X.dptr, L.dptr, loss.dptr, log_x.dptr, n * B, C));
    for (int64_t i = 0; i < n; i++) outs->push_back({loss.sliced(0, i * B, B), log_x.sliced(0, i * B, B)});
    return true;
  }
  void grad(GradientContext& c) override {
    using namespace T; Graph* g = c.graph();
    Tensor t = c.input(1), gy = c.output_grad(), log_x = nth_tensor(c.output(), 1);
    Tensor gx1 = TensorBuilder(g).append_input(log_x, false).append_input(t, false).append_input(gy, false).build(new SparseSoftmaxCrossEntropyGrad());
    Tensor x = unary("exp", log_x);
    Tensor sum = reduce("sum", mul(x, log_x), as_tensor(g, {1}), true);
    Tensor gx2 = mul(mul(x, gy), sub(sum, log_x));        // built but normally never evaluated
    c.append_input_grad(gx1); c.append_input_grad(gx2);
  }
};
struct SoftmaxCrossEntropy : Op {             // xent_ops.rs:160-202
  const char* name() const override { return REFNAME("xent_ops", "SoftmaxCrossEntropy"); }
  void compute(ComputeContext& c) override {
    NdArray x = c.dev->contiguous(on_dev(c.dev, c.input(0))), t = c.dev->contiguous(on_dev(c.dev, c.input(1)));
    if (x.ndim() != 2) throw Panic("x must be 2-ranked tensor");
    if (t.ndim() != 2) throw Panic("t must be 2-ranked tensor");
    int64_t B = x.shape[0], C = x.shape[1];
    NdArray loss = c.dev->empty({B}), log_x = c.dev->empty({B, C});
    check_status(agb_softmax_xent_fwd(c.dev->ctx, x.dptr, t.dptr, loss.dptr, log_x.dptr, B, C));
    c.append_output(loss); c.append_output(log_x);
  }
  void grad(GradientContext& c) override {
    using namespace T; Graph* g = c.graph();
    Tensor output = c.output(), log_x = nth_tensor(output, 1), gy = c.output_grad(), x = unary("exp", log_x), t = c.input(1);
    Tensor gx1 = mul(sub(x, t), gy);
    Tensor sum = reduce("sum", mul(x, log_x), as_tensor(g, {-1}), true);
    Tensor gx2 = mul(mul(gy, sub(sum, log_x)), output);
    c.append_input_grad(gx1); c.append_input_grad(gx2);
  }
};
struct SigmoidCrossEntropy : Op {             // xent_ops.rs:33-61
  const char* name() const override { return REFNAME("xent_ops", "SigmoidCrossEntropy"); }
  void compute(ComputeContext& c) override {
    NdArray x = c.input(0), t = c.input(1);
    if (x.shape != t.shape) throw Panic("x.shape must match t.shape");
    c.append_output(dev_binary(c.dev, AGB_B_SIGMOID_XENT, x, t));
  }
  void grad(GradientContext& c) override {
    using namespace T; Graph* g = c.graph(); Tensor x = c.input(0), t = c.input(1), gy = c.output_grad();
    Tensor e = unary("exp", x);
    c.append_input_grad(mul(sub(div(e, add(scalar(g, 1.f), e)), t), gy));
    c.append_input_grad(unary("neg", mul(gy, t)));
  }
};
Tensor T::sparse_softmax_cross_entropy(Tensor y, Tensor t) { return TensorBuilder(y.graph).append_input(y, false).append_input(t, false).build(new SparseSoftmaxCrossEntropy()); }
Tensor T::softmax_cross_entropy(Tensor y, Tensor t) { return TensorBuilder(y.graph).append_input(y, false).append_input(t, false).build(new SoftmaxCrossEntropy()); }
Tensor T::sigmoid_cross_entropy(Tensor y, Tensor t) { return TensorBuilder(y.graph).set_shape(shape(y)).append_input(y, false).append_input(t, false).build(new SigmoidCrossEntropy()); }

// ================================================================================================ views and copies (array_ops.rs)
struct Reshape : Op {                  // array_ops.rs:183-239
  const char* name() const override { return REFNAME("array_ops", "Reshape"); }
  void compute(ComputeContext& c) override {
    NdArray x = c.input(0), sa = c.input(1);
    const std::vector<float>& sv = c.dev->ensure_host(sa);
    float prod = 1.f; for (auto v : sv) prod *= v;
    Shape target;
    for (auto v : sv) target.push_back(v != -1.f ? (int64_t)v : (prod == 0.f ? 0 : x.size() / (int64_t)(-prod)));   // -1 inference (:187-197)
    int64_t n = 1; for (auto d : target) n *= d;
    if (n != x.size()) { throw OpError(AGB_ERR_INCOMPATIBLE_SHAPE, "reshape failed: element counts differ"); }
    if (x.virt) throw Panic("reshape of the virtual im2col tensor");
    if (!x.on_device() || x.is_contiguous()) c.append_output_view(x.reshaped(target));
    else c.append_output(c.dev->contiguous(x).reshaped(target));        // deep_copy for non-standard layouts (:212-225)
  }
  void grad(GradientContext& c) override {
    c.append_input_grad(TensorBuilder(c.graph()).append_input(c.output_grad(), false).append_input(T::shape(c.input(0)), false).build(new Reshape()));
    c.append_none();
  }
};
Tensor T::reshape(Tensor x, Tensor shp) { return TensorBuilder(x.graph).append_input(x, false).append_input(shp, false).build(new Reshape()); }
Tensor T::flatten(Tensor x) { return TensorBuilder(x.graph).append_input(x, false).append_input(scalar(x.graph, -1.f), false).set_shape(shape(x)).build(new Reshape()); }   // (sic) mod.rs:1344-1352

struct Transpose : Op {                // math_ops.rs:426-466: a stride permutation, never a copy
  bool invert_axes;
  const char* name() const override { return REFNAME("math_ops", "Transpose"); }
  void compute(ComputeContext& c) override {
    NdArray x = c.input(0), pa = c.input(1);
    std::vector<int64_t> perm = as_ints(c.dev, pa);
    if ((int)perm.size() != x.ndim()) throw OpError(AGB_ERR_INCOMPATIBLE_SHAPE, "transpose: inputs's ndim and axes's length must match");
    std::vector<int> dims(perm.size(), 0);
    for (size_t i = 0; i < perm.size(); i++) {
      if (perm[i] < 0 || perm[i] >= (int64_t)perm.size()) throw Panic("transpose: bad axis");
      if (invert_axes) dims[perm[i]] = (int)i; else dims[i] = (int)perm[i];
    }
    c.dev->ensure_device(x);
    c.append_output_view(x.permuted(dims));
  }
  void grad(GradientContext& c) override {
    auto* op = new Transpose(); op->invert_axes = !invert_axes;
    c.append_input_grad(TensorBuilder(c.graph()).append_input(c.output_grad(), false).append_input(c.input(1), false).set_shape(T::shape(c.input(0))).build(op));
    c.append_none();
  }
};
Tensor T::transpose(Tensor x, Tensor perm) { auto* op = new Transpose(); op->invert_axes = false; return TensorBuilder(x.graph).append_input(x, false).append_input(perm, false).build(op); }

struct SqueezeExpand : Op {            // Squeeze / ExpandDims, array_ops.rs:826-882
  bool expand;
  const char* name() const override { return expand ? REFNAME("array_ops", "ExpandDims") : REFNAME("array_ops", "Squeeze"); }
  void compute(ComputeContext& c) override {
    NdArray x = c.input(0), aa = c.input(1);
    std::vector<int64_t> axes = as_ints(c.dev, aa);
    std::sort(axes.begin(), axes.end());
    NdArray r = x;
    if (x.on_device()) r.host.reset();
    if (expand) {
      for (auto i : axes) {
        int ax = (int)(i < 0 ? x.ndim() + i : i);
        if (ax < 0 || ax > (int)r.shape.size()) throw Panic("expand_dims: axis out of range");
        r.shape.insert(r.shape.begin() + ax, 1); r.stride.insert(r.stride.begin() + ax, 0);
      }
    } else {
      int adjust = 0;
      for (auto i : axes) {
        int ax = (int)(i < 0 ? (int64_t)r.shape.size() + i : i) - adjust;      // (sic) array_ops.rs:838-844
        if (ax < 0 || ax >= (int)r.shape.size() || r.shape[ax] != 1) throw Panic("Can't squeeze a dim whose size != 1");
        r.shape.erase(r.shape.begin() + ax); r.stride.erase(r.stride.begin() + ax); adjust++;
      }
    }
    if (!r.on_device()) { r.stride = NdArray::contiguous_strides(r.shape); r.host = x.host; }
    c.append_output_view(r);
  }
  void grad(GradientContext& c) override {
    c.append_input_grad(expand ? T::squeeze(c.output_grad(), c.input(1)) : T::expand_dims(c.output_grad(), c.input(1)));
    c.append_none();
  }
};
Tensor T::squeeze(Tensor x, Tensor axes) { auto* op = new SqueezeExpand(); op->expand = false; return TensorBuilder(x.graph).append_input(x, false).append_input(axes, false).build(op); }
Tensor T::expand_dims(Tensor x, Tensor axes) { auto* op = new SqueezeExpand(); op->expand = true; return TensorBuilder(x.graph).append_input(x, false).append_input(axes, false).build(op); }

// ndarray Slice{start, end: Option, step 1} with python-style negative indices
struct SliceElem { int64_t start; bool has_end; int64_t end; };
static void resolve_slice(const SliceElem& e, int64_t len, int64_t& s, int64_t& n) {
  int64_t a = e.start < 0 ? e.start + len : e.start, b = !e.has_end ? len : (e.end < 0 ? e.end + len : e.end);
  if (a < 0 || a > len || b < 0 || b > len) throw Panic("slice: index out of bounds");
  s = a; n = b > a ? b - a : 0;
}
static NdArray apply_slices(Device* d, NdArray x, const std::vector<SliceElem>& idx) {
  if ((int)idx.size() != x.ndim()) throw Panic("slice: number of indices must match the tensor's rank");
  d->ensure_device(x);
  for (int k = 0; k < x.ndim(); k++) { int64_t s, n; resolve_slice(idx[k], x.shape[k], s, n); x = x.sliced(k, s, n); }
  return x;
}
struct SliceGrad : Op {                // SliceGrad / SplitGrad, array_ops.rs:726-749,803-825
  std::vector<SliceElem> indices; int split_axis = -1000; int64_t s0 = 0, s1 = 0; bool is_split = false;
  const char* name() const override { return is_split ? REFNAME("array_ops", "SplitGrad") : REFNAME("array_ops", "SliceGrad"); }
  void compute(ComputeContext& c) override {
    c.accept_expr = true; c.accept_lazy = true;
    NdArray x = c.input(0), gy = c.input(1);          // x: only its shape is used
    if (gy.lazy) gy = materialize_lazy(c.dev, gy);
    std::vector<SliceElem> idx = indices;
    if (is_split) { int ax = normalize_negative_axis(split_axis, x.ndim()); idx.assign(x.ndim(), SliceElem{0, false, 0}); idx[ax] = SliceElem{s0, true, s1}; }
    if ((int)idx.size() == x.ndim() && x.ndim() == gy.ndim()) {      // deferred: AddN may sum the pieces of one gradient in place (fuse.cc)
      std::vector<int64_t> start(x.ndim()); bool match = true;
      for (int k = 0; k < x.ndim(); k++) { int64_t s, n; resolve_slice(idx[k], x.shape[k], s, n); start[k] = s; if (n != gy.shape[k]) match = false; }
      if (match) { NdArray r = expr_pad(c, x.shape, start, gy); if (r.expr) { c.append_output(r); return; } }
    }
    NdArray gx = c.dev->zeros(x.shape);
    NdArray region = apply_slices(c.dev, gx, idx);
    if (region.shape != gy.shape) throw Panic("SliceGrad: gradient shape does not match the sliced region");
    if (gy.expr && expr_materialize_into(c.dev, gy, region)) { c.append_output(gx); return; }       // the fused program writes the region directly
    if (gy.expr) gy = expr_materialize(c.dev, gy);
    c.dev->ensure_device(gy);
    agb_tensor ts = gy.desc(), td = region.desc();
    check_status(agb_copy_strided(c.dev->ctx, &ts, &td));
    c.append_output(gx);
  }
  void grad(GradientContext& c) override { c.append_none(); if (!is_split) c.append_none(); }
};
struct SliceOp : Op {                  // Slice / Split, array_ops.rs:698-724,780-801
  std::vector<SliceElem> indices; int split_axis = -1000; int64_t s0 = 0, s1 = 0; bool is_split = false;
  const char* name() const override { return is_split ? REFNAME("array_ops", "Split") : REFNAME("array_ops", "Slice"); }
  void compute(ComputeContext& c) override {
    c.accept_expr = true;
    NdArray x = c.input(0);
    std::vector<SliceElem> idx = indices;
    if (is_split) { int ax = normalize_negative_axis(split_axis, x.ndim()); if (ax < 0 || ax >= x.ndim()) throw Panic("Wrong split axis"); idx.assign(x.ndim(), SliceElem{0, false, 0}); idx[ax] = SliceElem{s0, true, s1}; }
    if (x.expr) {                        // a slice of a pending expression stays pending (fuse.cc); otherwise the value is needed now
      if ((int)idx.size() == x.ndim()) {
        std::vector<int64_t> st(x.ndim()), ln(x.ndim());
        for (int k = 0; k < x.ndim(); k++) resolve_slice(idx[k], x.shape[k], st[k], ln[k]);
        NdArray r = expr_slice(c, x, st, ln);
        if (r.expr) { c.append_output(r); return; }
      }
      x = expr_materialize(c.dev, x);
    }
    c.append_output_view(apply_slices(c.dev, x, idx));
  }
  void grad(GradientContext& c) override {
    auto* op = new SliceGrad(); op->indices = indices; op->split_axis = split_axis; op->s0 = s0; op->s1 = s1; op->is_split = is_split;
    Tensor x = c.input(0);
    c.append_input_grad(TensorBuilder(c.graph()).append_input(x, false).append_input(c.output_grad(), false).set_shape(T::shape(x)).build(op));
  }
};
Tensor T::slice(Tensor x, const std::vector<int64_t>& starts, const std::vector<int64_t>& ends) {
  if (starts.size() != ends.size()) throw Panic("slice: starts.len() must match ends.len()");
  auto* op = new SliceOp();
  for (size_t i = 0; i < starts.size(); i++) {       // end-index rule: -1 -> to the end, e < -1 -> e + 1 (mod.rs:2181-2190)
    int64_t e = ends[i];
    if (e == -1) op->indices.push_back(SliceElem{starts[i], false, 0});
    else op->indices.push_back(SliceElem{starts[i], true, e < -1 ? e + 1 : e});
  }
  return TensorBuilder(x.graph).append_input(x, false).build(op);
}
std::vector<Tensor> T::split(Tensor x, const std::vector<int64_t>& sizes, int axis) {
  std::vector<Tensor> r; int64_t start = 0;
  for (auto sz : sizes) { auto* op = new SliceOp(); op->is_split = true; op->split_axis = axis; op->s0 = start; op->s1 = start + sz; start += sz; r.push_back(TensorBuilder(x.graph).append_input(x, false).build(op)); }
  return r;
}

struct ConcatGrad : Op {               // array_ops.rs:622-677.  The reference slices [start, region_len) — correct only for index 0;
  int index, axis;                     // this implementation takes the intended region [start, start + len).
  const char* name() const override { return REFNAME("array_ops", "ConcatGrad"); }
  void compute(ComputeContext& c) override {
    NdArray gy = on_dev(c.dev, c.input(0));
    int ax = normalize_negative_axis(axis, gy.ndim());
    int64_t start = 0;
    for (int i = 0; i < index; i++) start += c.xs[i + 1].arr.shape[ax];
    int64_t len = c.xs[index + 1].arr.shape[ax];
    c.append_output_view(gy.sliced(ax, start, len));
  }
  void grad(GradientContext& c) override { for (int i = 0; i < c.num_inputs(); i++) c.append_none(); }
};
struct Concat : Op {                   // Concat / Tile, array_ops.rs:576-620,679-696
  int axis; int tile_num = 0;
  const char* name() const override { return tile_num ? REFNAME("array_ops", "Tile") : REFNAME("array_ops", "Concat"); }
  void compute(ComputeContext& c) override {
    std::vector<NdArray> xs;
    if (tile_num) { NdArray x = on_dev(c.dev, c.input(0)); for (int i = 0; i < tile_num; i++) xs.push_back(x); }
    else for (int i = 0; i < c.num_inputs(); i++) xs.push_back(on_dev(c.dev, c.input(i)));
    int ax = normalize_negative_axis(axis, xs[0].ndim());
    if (ax < 0 || ax >= xs[0].ndim()) throw OpError(AGB_ERR_NDARRAY, "concat: axis out of bounds");
    Shape out = xs[0].shape; out[ax] = 0;
    for (auto& x : xs) {
      if (x.ndim() != xs[0].ndim()) throw OpError(AGB_ERR_NDARRAY, "concat: incompatible shapes");
      for (int k = 0; k < x.ndim(); k++) if (k != ax && x.shape[k] != xs[0].shape[k]) throw OpError(AGB_ERR_NDARRAY, "concat: incompatible shapes");
      out[ax] += x.shape[ax];
    }
    NdArray y = c.dev->empty(out); int64_t off = 0;
    for (auto& x : xs) {
      NdArray region = y.sliced(ax, off, x.shape[ax]); off += x.shape[ax];
      agb_tensor ts = x.desc(), td = region.desc();
      check_status(agb_copy_strided(c.dev->ctx, &ts, &td));
    }
    c.append_output(y);
  }
  void grad(GradientContext& c) override {
    if (tile_num) { c.append_input_grad(T::reduce("sum", c.output_grad(), T::as_tensor(c.graph(), {axis}), true)); return; }   // (sic) :693-695
    std::vector<Tensor> ins = c.inputs();
    for (int i = 0; i < c.num_inputs(); i++) {
      TensorBuilder b(c.graph()); b.set_shape(T::shape(c.input(0))).append_input(c.output_grad(), false);
      for (auto& in : ins) b.append_input(in, false);
      auto* op = new ConcatGrad(); op->index = i; op->axis = axis;
      c.append_input_grad(b.build(op));
    }
  }
};
Tensor T::concat(const std::vector<Tensor>& xs, int axis) {
  if (xs.empty()) throw Panic("concat: empty input");
  auto* op = new Concat(); op->axis = axis; TensorBuilder b(xs[0].graph);
  for (auto& x : xs) b.append_input(x, false);
  return b.build(op);
}
Tensor T::tile(Tensor x, int axis, int num) { auto* op = new Concat(); op->axis = axis; op->tile_num = num; return TensorBuilder(x.graph).append_input(x, false).build(op); }

struct GatherGrad : Op {               // array_ops.rs:401-474
  int axis;
  const char* name() const override { return REFNAME("array_ops", "GatherGrad"); }
  void compute(ComputeContext& c) override {
    NdArray idx_view = on_dev(c.dev, c.input(0)), param = c.input(1), gy = c.dev->contiguous(on_dev(c.dev, c.input(2)));
    int ax = normalize_negative_axis(axis, param.ndim());
    if (c.run->fuse && c.run->sole_consumer_sums(c.node)) { NdArray r = expr_scatter(c, param.shape, ax, idx_view, gy); if (r.expr) { c.append_output(r); return; } }
    NdArray idx = c.dev->contiguous(idx_view);
    int64_t pre = 1, post = 1; for (int k = 0; k < ax; k++) pre *= param.shape[k]; for (int k = ax + 1; k < param.ndim(); k++) post *= param.shape[k];
    NdArray gx = c.dev->empty(param.shape);
    check_status(agb_gather_grad(c.dev->ctx, gy.dptr, idx.dptr, gx.dptr, pre, param.shape[ax], post, idx.size()));
    c.append_output(gx);
  }
  void grad(GradientContext& c) override { c.append_none(); c.append_none(); c.append_none(); }
};
struct Gather : Op {                   // array_ops.rs:353-399; inputs are (indices, param)
  int axis; bool normalize;
  const char* name() const override { return REFNAME("array_ops", "Gather"); }
  void compute(ComputeContext& c) override {
    NdArray idx = c.input(0), param = c.input(1);
    int ax = normalize_negative_axis(axis, param.ndim());
    if (ax < 0 || ax >= param.ndim()) throw Panic("gather: axis out of range");
    Shape out(param.shape.begin(), param.shape.begin() + ax);
    out.insert(out.end(), idx.shape.begin(), idx.shape.end());
    out.insert(out.end(), param.shape.begin() + ax + 1, param.shape.end());
    if (all_meta(param) && idx.has_host() && param.ndim() == 1) {       // shape-vector lookup (ReduceMean::grad), host metadata
      std::vector<float> r;
      for (auto f : *idx.host) { int64_t k = (int64_t)f; if (k < 0 && normalize) k += param.shape[0]; if (k < 0 || k >= param.shape[0]) throw Panic("Invalid index value"); r.push_back((*param.host)[k]); }
      c.append_output(NdArray::from_host(out, r, true)); return;
    }
    idx = c.dev->contiguous(on_dev(c.dev, idx)); param = c.dev->contiguous(on_dev(c.dev, param));
    int64_t pre = 1, post = 1; for (int k = 0; k < ax; k++) pre *= param.shape[k]; for (int k = ax + 1; k < param.ndim(); k++) post *= param.shape[k];
    NdArray y = c.dev->empty(out);
    check_status(agb_gather(c.dev->ctx, param.dptr, idx.dptr, y.dptr, pre, param.shape[ax], post, idx.size(), normalize ? 1 : 0));
    c.append_output(y);
  }
  // the embedding lookups of an unrolled RNN: one gather with the stacked token ids; its row blocks are already the stacked lhs of x_t * wx
  const char* stack_key() const override { return axis == 0 ? (normalize ? "gather0n" : "gather0") : nullptr; }
  bool compute_stacked(Device* dev, Evaluation& run, const std::vector<std::vector<NdArray>>& ins, std::vector<std::vector<NdArray>>* outs) override {
    std::vector<NdArray> ids; const NdArray& p0 = ins[0][1];
    if (!p0.on_device() || p0.ndim() < 2 || !p0.is_contiguous() || p0.meta) return false;
    for (auto& in : ins) {
      if (in.size() != 2 || !in[0].on_device() || in[0].meta || in[0].shape != ins[0][0].shape || in[0].size() == 0) return false;
      if (in[1].dptr != p0.dptr || in[1].shape != p0.shape || in[1].stride != p0.stride) return false;      // one table for all
      int nontrivial = 0; for (auto d : in[0].shape) if (d != 1) nontrivial++;
      if (nontrivial > 1) return false;                                                                     // ids as a vector ([B] / [B, 1]) only
      ids.push_back(in[0]);
    }
    const int64_t n = (int64_t)ins.size(), B = ids[0].size();
    int64_t post = 1; for (int k = 1; k < p0.ndim(); k++) post *= p0.shape[k];
    NdArray L = stack_vectors(run, dev, ids), y = dev->empty({n * B, post});
    check_status(agb_gather(dev->ctx, p0.dptr, L.dptr, y.dptr, 1, p0.shape[0], post, n * B, normalize ? 1 : 0));
    Shape out(ids[0].shape); out.insert(out.end(), p0.shape.begin() + 1, p0.shape.end());
    for (int64_t i = 0; i < n; i++) outs->push_back({y.sliced(0, i * B, B).reshaped(out)});
    return true;
  }
  void grad(GradientContext& c) override {
    Tensor x = c.input(0), x1 = c.input(1);
    auto* op = new GatherGrad(); op->axis = axis;
    Tensor gx = TensorBuilder(c.graph()).append_input(x, false).append_input(x1, false).append_input(c.output_grad(), false).set_shape(T::shape(x)).build(op);
    c.append_none(); c.append_input_grad(gx);
  }
};
Tensor T::gather_common(Tensor param, Tensor indices, int axis) { auto* op = new Gather(); op->axis = axis; op->normalize = true; return TensorBuilder(param.graph).append_input(indices, false).append_input(param, false).build(op); }
Tensor T::gather(Tensor param, Tensor indices, int axis) { auto* op = new Gather(); op->axis = axis; op->normalize = false; return TensorBuilder(param.graph).append_input(indices, false).append_input(param, false).build(op); }

struct IndexOpGrad : Op {              // array_ops.rs:313-351
  int64_t index;
  const char* name() const override { return REFNAME("array_ops", "IndexOpGrad"); }
  void compute(ComputeContext& c) override {
    NdArray x = c.input(0), gy = on_dev(c.dev, c.input(1));
    int64_t i = index < 0 ? x.size() + index : index;
    if (i < 0 || i >= x.size()) throw OpError(AGB_ERR_OUT_OF_BOUNDS, "access_elem: tried to access an index outside the tensor");
    NdArray gx = c.dev->zeros(x.shape);
    NdArray cell = gx.reshaped({x.size()}).sliced(0, i, 1).reshaped({});
    agb_tensor ts = gy.reshaped({}).desc(), td = cell.desc();
    check_status(agb_copy_strided(c.dev->ctx, &ts, &td));
    c.append_output(gx);
  }
  void grad(GradientContext& c) override { c.append_none(); }
};
struct IndexOp : Op {                  // array_ops.rs:281-311
  int64_t index;
  const char* name() const override { return REFNAME("array_ops", "IndexOp"); }
  void compute(ComputeContext& c) override {
    NdArray x = c.dev->contiguous(on_dev(c.dev, c.input(0)));
    int64_t i = index < 0 ? x.size() + index : index;
    if (i < 0 || i >= x.size()) throw OpError(AGB_ERR_OUT_OF_BOUNDS, "access_elem: tried to access an index outside the tensor");
    c.append_output(c.dev->copy(x.reshaped({x.size()}).sliced(0, i, 1).reshaped({})));
  }
  void grad(GradientContext& c) override {
    auto* op = new IndexOpGrad(); op->index = index; Tensor x = c.input(0);
    c.append_input_grad(TensorBuilder(c.graph()).set_shape(T::shape(x)).append_input(x, false).append_input(c.output_grad(), false).build(op));
  }
};
Tensor T::access_elem(Tensor x, int64_t i) { auto* op = new IndexOp(); op->index = i; return TensorBuilder(x.graph).append_input(x, false).build(op); }

struct Assign : Op {                   // array_ops.rs:94-105: device copy into the variable
  const char* name() const override { return REFNAME("array_ops", "Assign"); }
  void compute(ComputeContext& c) override {
    c.dev->small_copies.clear();
    NdArray dst = c.input_mut(0), src = on_dev(c.dev, c.input(1));
    if (dst.shape != src.shape) src = dev_broadcast_to(c.dev, src, dst.shape);
    agb_tensor ts = src.desc(), td = dst.desc();
    check_status(agb_copy_strided(c.dev->ctx, &ts, &td));
    c.append_empty_output();
  }
  void grad(GradientContext& c) override { c.append_none(); c.append_none(); }
  bool mutates_now() const override { return true; }
};
Tensor T::assign(Tensor x, Tensor y) { return TensorBuilder(x.graph).append_input(x, true).append_input(y, false).build(new Assign()); }

// ================================================================================================ composites (tensor_ops/mod.rs)
Tensor T::reduce_variance(Tensor x, Tensor axes, bool keep) { return reduce("mean", unary("square", sub(x, reduce("mean", x, axes, true))), axes, keep); }   // :1291
Tensor T::leaky_relu(Tensor x, float alpha) { return cmp("maximum", x, mul(scalar(x.graph, alpha), x)); }                                                     // :1695
Tensor T::mean_squared_error(Tensor y, Tensor t) { return reduce("mean", unary("square", sub(y, t)), as_tensor(y.graph, {-1}), false); }                      // :1845
Tensor T::normalize(Tensor x, Tensor axes) {                                                                                                                   // :2325
  Tensor mean = reduce("mean", x, axes, true), centered = sub(x, mean);
  Tensor variance = reduce("mean", unary("square", centered), axes, true);
  return mul(centered, unary("inv_sqrt", add(variance, scalar(x.graph, 1e-5f))));
}
Tensor T::batch_norm(Tensor x, Tensor scale, Tensor shift) { return add(mul(normalize(x, as_tensor(x.graph, {0})), scale), shift); }                          // :2362

// ================================================================================================ grad entry points
std::vector<Tensor> T::grad(const std::vector<Tensor>& ys_, const std::vector<Tensor>& xs) {       // mod.rs:94-114
  if (ys_.empty()) throw Panic("grad: ys is empty");
  Graph* g = ys_[0].graph;
  std::vector<Tensor> ys; for (auto& y : ys_) ys.push_back(sum_all(y));
  std::vector<Tensor> gs = compute_gradients(ys, xs, nullptr, g), ret;
  for (size_t i = 0; i < xs.size(); i++) ret.push_back(gs[i].valid() ? gs[i] : zeros(g, shape(xs[i])));
  return ret;
}
std::vector<Tensor> T::grad_with_default(const std::vector<Tensor>& ys, const std::vector<Tensor>& xs, const std::vector<Tensor>& gys) {
  Graph* g = ys[0].graph;
  std::vector<Tensor> gs = compute_gradients(ys, xs, &gys, g), ret;
  for (size_t i = 0; i < xs.size(); i++) ret.push_back(gs[i].valid() ? gs[i] : zeros(g, shape(xs[i])));
  return ret;
}

}  // namespace agx
