// core.cc — arrays, graph, reverse-mode gradient builder, evaluator, variable environment (see agx.h for the reference map).
#include "agx.h"
#include <cmath>
#include <limits>
#include <map>
#include <tuple>
#include <algorithm>
#include <queue>
#include <sstream>
#include <stdio.h>
#include <string.h>

namespace agx {

void check_status(int status) {
  if (status != AGB_OK) throw OpError(status, agb_last_error());
}

// ================================================================================================ NdArray / Device
Buffer::Buffer(agb_ctx* c, size_t b) : ctx(c), ptr(nullptr), bytes(b) {
  void* p = nullptr; check_status(agb_alloc(c, b ? b : 4, &p)); ptr = (float*)p;
}
Buffer::~Buffer() { if (ptr) agb_free(ctx, ptr); }

Shape NdArray::contiguous_strides(const Shape& s) {
  Shape st(s.size()); int64_t acc = 1;
  for (int i = (int)s.size() - 1; i >= 0; i--) { st[i] = acc; acc *= s[i]; }
  return st;
}
bool NdArray::is_contiguous() const {
  if (!on_device()) return true;
  int64_t acc = 1;
  for (int i = ndim() - 1; i >= 0; i--) { if (shape[i] != 1 && stride[i] != acc) return false; acc *= shape[i]; }
  return true;
}
bool NdArray::dense_order(std::vector<int>& order) const {
  const int n = ndim();
  order.resize(n);
  for (int i = 0; i < n; i++) order[i] = i;
  if (!on_device()) return true;
  std::vector<int> big;                       // non-unit axes by descending stride; unit axes keep their logical slot
  for (int i = 0; i < n; i++) if (shape[i] != 1) big.push_back(i);
  std::sort(big.begin(), big.end(), [&](int a, int b) { return stride[a] != stride[b] ? stride[a] > stride[b] : a < b; });
  int bi = 0;
  for (int i = 0; i < n; i++) order[i] = shape[i] == 1 ? i : big[bi++];
  int64_t acc = 1;
  for (int i = n - 1; i >= 0; i--) { int a = order[i]; if (shape[a] != 1 && stride[a] != acc) return false; acc *= shape[a]; }
  return true;
}
agb_tensor NdArray::desc() const {
  agb_tensor t; t.ptr = dptr; t.rank = ndim();
  if (t.rank > AGB_MAX_RANK) throw OpError(AGB_ERR_INVALID_DIMS, "tensor rank exceeds AGB_MAX_RANK");
  for (int i = 0; i < t.rank; i++) { t.shape[i] = shape[i]; t.stride[i] = stride[i]; }
  return t;
}
NdArray NdArray::from_host(const Shape& shape, std::vector<float> v, bool meta) {
  NdArray a; a.shape = shape; a.stride = contiguous_strides(shape);
  a.host = std::make_shared<std::vector<float>>(std::move(v)); a.meta = meta;
  return a;
}
NdArray NdArray::reshaped(const Shape& s) const {
  NdArray r = *this; r.shape = s; r.stride = contiguous_strides(s); r.virt.reset(); r.chan_sum.reset(); r.pool.reset();
  return r;
}
NdArray NdArray::permuted(const std::vector<int>& perm) const {
  NdArray r = *this; r.host.reset(); r.virt.reset(); r.chan_sum.reset(); r.pool.reset();
  for (size_t i = 0; i < perm.size(); i++) { r.shape[i] = shape[perm[i]]; r.stride[i] = stride[perm[i]]; }
  return r;
}
NdArray NdArray::sliced(int axis, int64_t start, int64_t len) const {
  NdArray r = *this; r.host.reset(); r.virt.reset(); r.chan_sum.reset(); r.pool.reset();
  r.dptr = dptr + start * stride[axis]; r.shape[axis] = len;
  return r;
}

Device::Device(int index) { check_status(agb_init(index, &ctx)); stream_cells = std::make_shared<StreamCellPool>(); stream_cells->ctx = ctx; }
Device::~Device() { if (stream_cells) { stream_cells->alive = false; stream_cells->free_cells.clear(); } if (ctx) agb_destroy(ctx); }      // agb_destroy releases the pool's arena blocks with everything else
std::shared_ptr<StreamCell> Device::new_stream_cell() {
  StreamCellPool& p = *stream_cells;
  if (p.free_cells.empty()) {
    const size_t cell = (size_t)agb_stream_cell_bytes(), n = 512;
    void* blk = nullptr; check_status(agb_alloc(ctx, cell * n, &blk)); p.blocks.push_back(blk);
    check_status(agb_memset0(ctx, blk, cell * n));
    for (size_t i = 0; i < n; i++) p.free_cells.push_back((uint32_t*)((char*)blk + i * cell));
  }
  auto c = std::make_shared<StreamCell>(); c->pool = stream_cells; c->ptr = p.free_cells.back(); p.free_cells.pop_back();
  return c;
}
NdArray Device::empty(const Shape& s) {
  NdArray a; a.shape = s; a.stride = NdArray::contiguous_strides(s);
  a.buf = std::make_shared<Buffer>(ctx, (size_t)std::max<int64_t>(a.size(), 1) * sizeof(float)); a.dptr = a.buf->ptr;
  return a;
}
NdArray Device::empty_ordered(const Shape& s, const std::vector<int>& order) {
  NdArray a = empty(s); int64_t acc = 1;
  for (int i = (int)s.size() - 1; i >= 0; i--) { a.stride[order[i]] = acc; acc *= s[order[i]]; }
  return a;
}
NdArray Device::zeros(const Shape& s) {
  NdArray a = empty(s);
  check_status(agb_memset0(ctx, a.dptr, (size_t)a.size() * sizeof(float)));
  return a;
}
NdArray Device::full(const Shape& s, float v) {
  NdArray a = empty(s); agb_tensor t = a.desc();
  check_status(agb_fill(ctx, &t, v));
  return a;
}
void Device::ensure_device(NdArray& a) {
  if (a.on_device()) return;
  if (a.virt) throw Panic("virtual im2col tensor used where a materialised array is required");
  if (!a.host) throw Panic("array has neither device nor host storage");
  NdArray d = empty(a.shape);
  if (a.size() == 1) { agb_tensor t = d.desc(); check_status(agb_fill(ctx, &t, (*a.host)[0])); }   // scalars: no H2D at all
  else if (a.size() > 0) check_status(agb_h2d(ctx, d.dptr, a.host->data(), (size_t)a.size() * sizeof(float)));
  a.buf = d.buf; a.dptr = d.dptr; a.stride = d.stride;
}
const std::vector<float>& Device::ensure_host(NdArray& a) {
  if (a.host) return *a.host;
  if (a.i32) { NdArray f = i32_to_f32(a); a.buf = f.buf; a.dptr = f.dptr; a.stride = f.stride; a.i32 = false; }
  NdArray c = a.is_contiguous() ? a : contiguous(a);
  auto h = std::make_shared<std::vector<float>>((size_t)c.size());
  check_status(agb_d2h(ctx, h->data(), c.dptr, (size_t)c.size() * sizeof(float)));
  check_status(agb_sync(ctx));
  a.host = h;
  return *a.host;
}
NdArray Device::contiguous(const NdArray& a) {
  if (a.is_contiguous()) return a;
  if (a.on_device() && a.size() <= (1 << 14)) {
    for (auto& e : small_copies) if (e.first.dptr == a.dptr && e.first.buf == a.buf && e.first.shape == a.shape && e.first.stride == a.stride && e.first.i32 == a.i32) return e.second;
    NdArray c = copy(a);
    if (small_copies.size() < 4096) small_copies.push_back({a, c});
    return c;
  }
  return copy(a);
}
NdArray Device::copy(const NdArray& a_) {
  NdArray a = a_; ensure_device(a);
  NdArray c = empty(a.shape);
  agb_tensor s = a.desc(), d = c.desc();
  check_status(agb_copy_strided(ctx, &s, &d));
  c.meta = a.meta; c.i32 = a.i32;
  return c;
}
NdArray Device::i32_to_f32(const NdArray& a) {
  if (!a.i32) return a;
  NdArray s = a; s.i32 = false; s = contiguous(s);       // bit patterns are copied verbatim by the strided copy
  NdArray d = empty(a.shape);
  check_status(agb_convert_i32_f32(ctx, (const int32_t*)s.dptr, d.dptr, s.size()));
  return d;
}
void Device::sync() { check_status(agb_sync(ctx)); }

Shape as_shape(Device* dev, NdArray& a) {
  const std::vector<float>& h = dev->ensure_host(a);
  Shape s(h.size()); for (size_t i = 0; i < h.size(); i++) s[i] = (int64_t)h[i];
  return s;
}
std::vector<int64_t> as_ints(Device* dev, NdArray& a) { return as_shape(dev, a); }

// ================================================================================================ Graph / builder
TensorID Graph::install(std::unique_ptr<TensorInternal> node) {
  size_t id = node_set.size();
  if (id == NUM_NODES_WARN)
    fprintf(stderr, "Too many tensors in this graph: %zu. Use Graph::clear, or move the training loop out of the `run` block\n", NUM_NODES_WARN);
  if (id > NUM_NODES_CRITICAL)
    throw Panic("Maximum graph size exceeded: 500000. Use Graph::clear, or move the training loop out of the `run` block");   // graph.rs:36-41
  node->id = (TensorID)id;
  node_set.push_back(std::move(node));
  return (TensorID)id;
}

TensorBuilder& TensorBuilder::append_input_with_selector(Tensor t, bool allow_mut, int sel) {
  if (t.graph != graph) throw Panic("Detected tensors belonging to different graphs");     // graph.rs:229-236
  in_nodes.push_back(IncomingTensor{t.id, allow_mut, sel});
  return *this;
}
TensorBuilder& TensorBuilder::append_backprop_input(Tensor t) {
  if (t.graph != graph) throw Panic("Detected tensors belonging to different graphs");
  has_bp = true; bp.push_back(IncomingTensor{t.id, false, 0});
  return *this;
}
TensorBuilder& TensorBuilder::set_known_shape(const std::vector<int64_t>& s) {
  for (auto a : s) if (a != -1 && a <= 0) throw Panic("Given shape contains invalid dim size(s)");   // tensor.rs:629-640
  has_known = true; known = s; return *this;
}
Tensor TensorBuilder::build(Op* op) {
  auto n = std::make_unique<TensorInternal>();
  int rank = 0;
  for (auto& in : in_nodes) rank = std::max(rank, graph->inner(in.id).topo_rank + 1);   // tensor.rs:773-783
  n->op.reset(op); n->incoming_nodes = in_nodes; n->topo_rank = rank; n->shape = shape;
  n->is_differentiable = differentiable; n->has_backprop_inputs = has_bp; n->backprop_inputs = bp;
  n->has_known_shape = has_known; n->known_shape = known; n->variable_id = variable_id;
  n->placeholder_name = placeholder; n->is_placeholder = is_ph;
  return graph->tensor(graph->install(std::move(n)));
}

// marker ops whose compute is unreachable (basic_source_ops.rs:3-23)
struct SourceOp : Op {
  const char* nm; explicit SourceOp(const char* n) : nm(n) {}
  const char* name() const override { return nm; }
  void compute(ComputeContext&) override { throw Panic("unreachable: source op computed"); }
  void grad(GradientContext&) override {}
};

Tensor Graph::placeholder(const std::string& name, const std::vector<int64_t>& shp) {
  TensorBuilder b(this);
  b.set_placeholder_name(name);
  if (shp.empty() || shp[0] != -1) b.set_shape(T::as_tensor(this, shp));          // graph.rs:182-196
  b.set_known_shape(shp);
  return b.build(new SourceOp("autograd::tensor_ops::basic_source_ops::Placeholder"));
}
Tensor Graph::variable_by_id(VariableID vid) {
  auto it = variable2node.find(vid.v);
  if (it != variable2node.end()) return tensor(it->second);
  if (!env || vid.v < 0 || vid.v >= (int)env->array_list.size()) throw Panic("variable array not found");
  const NdArray& arr = env->array_list[vid.v];
  TensorBuilder b(this);
  b.set_shape(T::as_tensor(this, arr.shape)).set_variable(vid);
  Tensor t = b.build(new SourceOp("autograd::tensor_ops::basic_source_ops::Variable"));
  variable2node[vid.v] = t.id;
  return t;
}
Tensor Graph::variable_by_name(const std::string& name, const std::string& ns) {
  VariableID v = env->find(ns, name);
  if (!v.valid()) throw Panic("variable array not found in `" + ns + "`: " + name);        // variable.rs:742-752
  return variable_by_id(v);
}

// ================================================================================================ contexts
NdArray ComputeContext::input(int i) {
  if (i < 0 || i >= (int)xs.size()) throw Panic("Bad op impl: input index out of range.");
  OpInput& x = xs[i];
  if (x.kind == InputKind::RdWrVariable) throw Panic("Bad op impl: cannot perform mutable borrowing for input. Use input_mut() instead.");
  if (x.taken) throw Panic("Bad op impl: input()/input_mut() cannot be called twice");
  x.taken = true;
  if (x.arr.expr && (!accept_expr || expr_has_value(x.arr))) return expr_materialize(dev, x.arr);
  if (x.arr.lazy && !accept_lazy) return materialize_lazy(dev, x.arr);
  if (x.arr.i32 && !accept_i32) return dev->i32_to_f32(x.arr);
  return x.arr;
}
NdArray ComputeContext::input_mut(int i) {
  if (i < 0 || i >= (int)xs.size()) throw Panic("Bad op impl: input doesn't exist.");
  OpInput& x = xs[i];
  if (x.kind != InputKind::RdWrVariable) throw Panic("Bad op impl: cannot perform mutable borrowing for input");
  if (x.taken) throw Panic("Bad op impl: input()/input_mut() cannot be called twice");
  x.taken = true; return x.arr;
}
void ComputeContext::append_output_view(NdArray y) {
  bool has_var = false;
  for (auto& x : xs) if (x.kind != InputKind::NonVariable) has_var = true;
  if (has_var && y.on_device()) ys.push_back(dev->copy(y));      // copy beforehand, like op.rs:273-287
  else ys.push_back(std::move(y));
}
void ComputeContext::append_empty_output() { ys.push_back(NdArray::scalar_host(0.0f)); }

Tensor GradientContext::input(int i) const {
  auto& in = g->inner(y.id).incoming_nodes;
  if (i < 0 || i >= (int)in.size()) throw Panic("bad Op::grad impl");
  return g->tensor(in[i].id);
}
std::vector<Tensor> GradientContext::inputs() const {
  std::vector<Tensor> r; for (auto& in : g->inner(y.id).incoming_nodes) r.push_back(g->tensor(in.id)); return r;
}
int GradientContext::num_inputs() const { return (int)g->inner(y.id).incoming_nodes.size(); }

// ================================================================================================ gradients (gradient.rs)
namespace {
struct GradInfo { std::vector<Tensor> gradients; bool on_path = false; bool present = false; };
struct HeapNode { TensorID id; int rank; bool operator<(const HeapNode& o) const { return rank < o.rank; } };
}

std::vector<Tensor> compute_gradients(const std::vector<Tensor>& ys, const std::vector<Tensor>& xs, const std::vector<Tensor>* gys, Graph* g) {
  std::unordered_map<TensorID, GradInfo> map;
  auto is_xs = [&](TensorID id) { for (auto& x : xs) if (x.id == id) return true; return false; };
  // init_gradient_map (gradient.rs:203-248): DFS from ys, mark nodes between ys and xs
  std::vector<std::pair<TensorID, bool>> st;
  for (auto& y : ys) st.push_back({y.id, false});
  while (!st.empty()) {
    auto [cur, visit] = st.back(); st.pop_back();
    TensorInternal& n = g->inner(cur);
    if (visit) {
      bool child_on = false;
      for (auto& c : n.get_backprop_inputs()) { auto it = map.find(c.id); if (it != map.end() && it->second.on_path) child_on = true; }
      GradInfo gi; gi.present = true; gi.on_path = n.is_differentiable && (is_xs(cur) || child_on);
      map[cur] = gi;
    } else {
      st.push_back({cur, true});
      for (auto& c : n.get_backprop_inputs()) {
        if (map.find(cur) != map.end()) continue;
        TensorInternal& ch = g->inner(c.id);
        if (ch.is_source() || !ch.is_differentiable) {
          GradInfo gi; gi.present = true; gi.on_path = ch.is_differentiable && is_xs(c.id);
          map[c.id] = gi;
        } else st.push_back({c.id, false});
      }
    }
  }
  auto gradient_of = [&](GradInfo& gi) -> Tensor {
    if (gi.gradients.size() > 1) { Tensor s = T::add_n(gi.gradients); gi.gradients.clear(); gi.gradients.push_back(s); }   // :168-173
    return gi.gradients[0];
  };
  if (gys) {
    if (gys->size() != ys.size()) throw Panic("`ys.len()` must match `gys.len()`");
    for (size_t i = 0; i < ys.size(); i++) map[ys[i].id].gradients.push_back((*gys)[i]);
  } else {
    Tensor one = T::scalar(g, 1.0f);
    for (auto& y : ys) map[y.id].gradients.push_back(one);
  }
  std::priority_queue<HeapNode> heap;
  for (auto& y : ys) heap.push(HeapNode{y.id, g->inner(y.id).topo_rank});
  while (!heap.empty()) {
    HeapNode y = heap.top(); heap.pop();
    Tensor gy = gradient_of(map[y.id]);
    GradientContext ctx; ctx.gy = gy; ctx.y = g->tensor(y.id); ctx.g = g;
    Op* op = g->inner(y.id).op.get();            // (the reference temporarily steals the boxed op, op.rs:365-380)
    op->grad(ctx);
    // copy: Op::grad may have appended nodes and reallocated node storage
    std::vector<IncomingTensor> bins = g->inner(y.id).get_backprop_inputs();
    size_t n = std::min(bins.size(), ctx.gxs.size());
    for (size_t i = 0; i < n; i++) {
      TensorID xid = bins[i].id;
      auto it = map.find(xid);
      if (it == map.end() || !it->second.on_path) continue;
      if (!ctx.gxs[i].valid()) continue;
      bool not_visited = it->second.gradients.empty();
      it->second.gradients.push_back(ctx.gxs[i]);
      if (!g->inner(xid).is_source() && not_visited) heap.push(HeapNode{xid, g->inner(xid).topo_rank});
    }
  }
  std::vector<Tensor> ret;
  for (auto& x : xs) {
    auto it = map.find(x.id);
    if (it != map.end() && it->second.on_path && !it->second.gradients.empty()) ret.push_back(gradient_of(it->second));
    else ret.push_back(Tensor{});     // None: not differentiable
  }
  return ret;
}

// ================================================================================================ evaluation (evaluation.rs)
namespace {
struct Stored { bool ok = true; int code = 0; std::string msg; std::vector<NdArray> ys; };

NdArray find_placeholder_value(const std::vector<Feed>& feeds, Graph* g, TensorID id) {
  TensorInternal& n = g->inner(id);
  for (auto& f : feeds) {
    bool hit = f.by_name ? (f.name == n.placeholder_name) : (f.id == id);
    if (!hit) continue;
    if (n.has_known_shape) {              // validate_using_known_shape, tensor.rs:374-386
      bool ok = n.known_shape.size() == f.value.shape.size();
      for (size_t i = 0; ok && i < n.known_shape.size(); i++) if (n.known_shape[i] > 0 && n.known_shape[i] != f.value.shape[i]) ok = false;
      if (!ok) throw Panic("Shape error: placeholder required a different shape than the value fed to `" + n.placeholder_name + "`");
    }
    return f.value;
  }
  throw Panic("Placeholder unfilled");    // evaluation.rs:248
}
}  // namespace

bool Evaluation::sole_consumer_sums(TensorID id) const {
  if (id < 0 || id >= (int)sole_consumer.size() || sole_consumer[id] < 0) return false;
  TensorInternal& n = graph->inner(sole_consumer[id]);
  return n.op && n.op->sums_inputs();
}

std::vector<EvalResult> eval(Graph* g, const std::vector<Tensor>& targets, const std::vector<Feed>& feeds, bool fetch_to_host) {
  VariableEnvironment* env = g->env; Device* dev = env->dev;
  std::unordered_map<TensorID, Stored> storage;
  Evaluation run; run.graph = g; run.dev = dev;
  dev->small_copies.clear();
  struct ClearCopies { Device* d; ~ClearCopies() { d->small_copies.clear(); } } clear_copies{dev};
  auto would_not_visit = [&](TensorID id) {
    TensorInternal& n = g->inner(id);
    return n.is_placeholder || n.is_variable() || storage.count(id) > 0;
  };
  std::vector<std::pair<TensorID, bool>> st; st.reserve(1 << 10);
  for (auto& t : targets) { if (t.graph != g) throw Panic("Detected tensors belonging to different graphs"); st.push_back({t.id, false}); }
  run.fuse = env->fuse_elementwise;
  std::vector<char> seen;      // nodes this evaluation will compute
  if (run.fuse) {         // pre-pass for fuse.cc: how many consumers will read each node's VALUES in this evaluation
    run.consumers.assign(g->node_set.size(), 0); run.sole_consumer.assign(g->node_set.size(), -1);
    for (auto& t : targets) run.sole_consumer[t.id] = -2;
    seen.assign(g->node_set.size(), 0); std::vector<TensorID> todo;
    for (auto& t : targets) { run.consumers[t.id]++; if (!seen[t.id]) { seen[t.id] = 1; todo.push_back(t.id); } }
    while (!todo.empty()) {
      TensorInternal& n = g->inner(todo.back()); todo.pop_back();
      if (n.is_placeholder || n.is_variable()) continue;
      const bool meta_only = n.op && n.op->metadata_only();
      if (n.op && n.op->mutates_now()) run.fuse = false;      // a pending expression must never read a variable after an Assign of the same run
      for (auto& c : n.incoming_nodes) {
        if (!meta_only) { run.consumers[c.id]++; run.sole_consumer[c.id] = run.sole_consumer[c.id] == -1 ? n.id : -2; }
        if (!seen[c.id]) { seen[c.id] = 1; todo.push_back(c.id); }
      }
    }
  }
  // ---- row-stacked MatMuls (the contraction side): an unrolled RNN multiplies T different [B, k] inputs by the SAME
  // weight (lstm_lm.rs:30-35,48: x_t * wx, h_t * w_pred, and their MatMul::grad counterparts gy_t * W^T).  Members whose lhs does not depend on
  // another member's product are evaluated together: when the traversal first reaches one, the lhs of ALL of them are scheduled first, the
  // rows are stacked (agb_concat_rows) and ONE [T*B, k] x W GEMM fills every member's output (row blocks of one buffer).  Per-row
  // arithmetic is unchanged; a 128-row GEMM cannot fill 148 SMs, a 8064-row one runs at the large-GEMM rate.
  // The same scheduling serves row-wise ops that declare a stack_key (the per-step softmax cross-entropy and its gradient): one launch over
  // the stacked rows instead of T (Op::compute_stacked).
  struct RowBatch { std::vector<TensorID> members; TensorID weight; bool tb; bool generic = false; bool expanded = false, done = false; };
  std::vector<RowBatch> batches; std::vector<int> batch_of;
  if (run.fuse) {
    const size_t N = g->node_set.size();
    std::map<std::tuple<std::string, TensorID, bool>, int> key2group; std::vector<RowBatch> groups; std::vector<int> group_of(N, -1);
    bool ordered = true;
    for (size_t id = 0; id < N; id++) {
      if (!seen[id]) continue;                                    // not part of this evaluation
      TensorInternal& n = g->inner((TensorID)id); bool tb = false;
      if (n.is_placeholder || n.is_variable() || !n.op) continue;
      for (auto& c : n.incoming_nodes) if (c.id >= (TensorID)id) ordered = false;      // (control_dependencies rewiring) no stacking then
      std::tuple<std::string, TensorID, bool> key; bool generic = false;
      if (n.op->plain_matmul(&tb) && n.incoming_nodes.size() == 2 && g->inner(n.incoming_nodes[1].id).is_variable()) key = std::make_tuple(std::string(), n.incoming_nodes[1].id, tb);
      else if (n.op->stack_key() && !n.incoming_nodes.empty()) { key = std::make_tuple(std::string(n.op->stack_key()), (TensorID)-1, false); generic = true; }
      else continue;
      auto it = key2group.find(key);
      if (it == key2group.end()) { if (groups.size() >= 64) continue; it = key2group.insert({key, (int)groups.size()}).first; RowBatch b; b.weight = std::get<1>(key); b.tb = tb; b.generic = generic; groups.push_back(b); }
      groups[it->second].members.push_back((TensorID)id); group_of[id] = it->second;
    }
    if (ordered && !groups.empty()) {
      std::vector<uint64_t> dep(N, 0);                            // bit q: the node's value depends on a product of group q
      for (size_t id = 0; id < N; id++) {
        if (!seen[id]) continue;
        TensorInternal& n = g->inner((TensorID)id);
        if (n.is_placeholder || n.is_variable()) continue;
        uint64_t d = 0;
        for (auto& c : n.incoming_nodes) { d |= dep[c.id]; if (group_of[c.id] >= 0) d |= 1ull << group_of[c.id]; }
        dep[id] = d;
      }
      batch_of.assign(N, -1);
      for (size_t q = 0; q < groups.size(); q++) {
        RowBatch b; b.weight = groups[q].weight; b.tb = groups[q].tb; b.generic = groups[q].generic;
        for (TensorID m : groups[q].members) {
          uint64_t d = 0; auto& ins = g->inner(m).incoming_nodes;
          for (size_t k = 0; k < (b.generic ? ins.size() : 1); k++) { TensorID a = ins[k].id; d |= dep[a] | (group_of[a] >= 0 ? 1ull << group_of[a] : 0); }
          if (!((d >> q) & 1)) b.members.push_back(m);
        }
        if (b.members.size() >= 2) { for (TensorID m : b.members) batch_of[m] = (int)batches.size(); batches.push_back(b); }
      }
    }
  }
  auto fetch = [&](const IncomingTensor& in, NdArray* out, bool need_device = true) {      // the value an op would receive through ComputeContext::input
    TensorInternal& x = g->inner(in.id);
    if (x.is_placeholder) *out = find_placeholder_value(feeds, g, in.id);
    else if (x.is_variable()) *out = env->array_list[x.variable_id.v];
    else {
      auto it = storage.find(in.id);
      if (it == storage.end() || !it->second.ok || in.array_selector >= (int)it->second.ys.size()) return false;
      *out = it->second.ys[in.array_selector];
    }
    if (out->expr) *out = expr_materialize(dev, *out);
    if (out->lazy) *out = materialize_lazy(dev, *out);
    if (out->i32 || out->virt) return false;
    if (!need_device) return true;                    // (host-known scalars stay on the host, like ComputeContext::input)
    dev->ensure_device(*out);
    return out->on_device();
  };
  auto run_row_batch = [&](RowBatch& b) {
    std::vector<TensorID> ms; std::vector<NdArray> as; NdArray w;
    if (!fetch(IncomingTensor{b.weight, false, 0}, &w) || w.ndim() != 2) return false;
    for (TensorID m : b.members) {
      if (storage.count(m)) continue;
      NdArray a;
      if (!fetch(g->inner(m).incoming_nodes[0], &a) || a.ndim() != 2) continue;
      if (!as.empty() && a.shape != as[0].shape) continue;
      if (!stackable(a)) continue;
      ms.push_back(m); as.push_back(a);
    }
    if (ms.size() < 2) return false;
    const int64_t rows = as[0].shape[0], k = as[0].shape[1], ncol = b.tb ? w.shape[0] : w.shape[1];
    if ((b.tb ? w.shape[1] : w.shape[0]) != k || rows == 0 || k == 0 || ncol == 0) return false;      // the members raise their own shape errors
    const int n = (int)ms.size();
    NdArray A = stack_rows(run, dev, as), Y = dev->empty({n * rows, ncol});
    agb_tensor da = A.desc(), dw = w.desc(), dy = Y.desc();
    check_status(agb_gemm_f32(dev->ctx, 0, b.tb ? 1 : 0, &da, &dw, &dy, 0.0f));
    for (int i = 0; i < n; i++) { Stored o; o.ys.push_back(Y.sliced(0, i * rows, rows)); storage[ms[i]] = std::move(o); }
    return true;
  };
  auto run_stacked_ops = [&](RowBatch& b) {
    std::vector<TensorID> ms; std::vector<std::vector<NdArray>> ins;
    for (TensorID m : b.members) {
      if (storage.count(m)) continue;
      std::vector<NdArray> xs; bool ok = true;
      for (auto& in : g->inner(m).incoming_nodes) { NdArray a; if (!fetch(in, &a, false)) { ok = false; break; } xs.push_back(a); }
      if (ok) { ms.push_back(m); ins.push_back(std::move(xs)); }
    }
    if (ms.size() < 2) return false;
    std::vector<std::vector<NdArray>> outs;
    if (!g->inner(ms[0]).op->compute_stacked(dev, run, ins, &outs) || outs.size() != ms.size()) return false;
    for (size_t i = 0; i < ms.size(); i++) { Stored o; o.ys = std::move(outs[i]); storage[ms[i]] = std::move(o); }
    return true;
  };
  while (!st.empty()) {
    auto [id, visit] = st.back(); st.pop_back();
    if (visit) {
      if (would_not_visit(id)) continue;
      if (!batch_of.empty() && batch_of[id] >= 0 && !batches[batch_of[id]].done) {
        RowBatch& b = batches[batch_of[id]]; b.done = true;
        if ((b.generic ? run_stacked_ops(b) : run_row_batch(b)) && storage.count(id)) continue;
      }
      TensorInternal& n = g->inner(id);
      Stored out;
      ComputeContext ctx; ctx.dev = dev; ctx.run = &run; ctx.node = id;
      for (auto& in : n.incoming_nodes) {
        TensorInternal& x = g->inner(in.id);
        if (x.is_placeholder) ctx.xs.push_back(OpInput{find_placeholder_value(feeds, g, in.id), InputKind::NonVariable});
        else if (x.is_variable()) ctx.xs.push_back(OpInput{env->array_list[x.variable_id.v], in.allow_mut ? InputKind::RdWrVariable : InputKind::RdOnlyVariable});
        else {
          Stored& s = storage[in.id];
          if (!s.ok) { out.ok = false; out.code = s.code; out.msg = s.msg; break; }     // errors propagate to dependents (:202-211)
          if (in.array_selector >= (int)s.ys.size()) throw Panic("Bad op implementation: output selector out of range");
          ctx.xs.push_back(OpInput{s.ys[in.array_selector], InputKind::NonVariable});
        }
      }
      if (out.ok) {
        try {
          n.op->compute(ctx);
          if (ctx.ys.empty()) throw Panic("Bad op implementation: empty return value");
          out.ys = std::move(ctx.ys);
        } catch (const OpError& e) { out.ok = false; out.code = e.code; out.msg = e.msg; }
      }
      storage[id] = std::move(out);
    } else {
      st.push_back({id, true});
      for (auto& c : g->inner(id).incoming_nodes) if (!would_not_visit(c.id)) st.push_back({c.id, false});
      if (!batch_of.empty() && batch_of[id] >= 0 && !batches[batch_of[id]].expanded) {       // schedule the lhs of every member of the stack first
        RowBatch& b = batches[batch_of[id]]; b.expanded = true;
        for (TensorID m : b.members) {
          if (m == id) continue;
          auto& ins = g->inner(m).incoming_nodes;
          for (size_t k = 0; k < (b.generic ? ins.size() : 1); k++) if (!would_not_visit(ins[k].id)) st.push_back({ins[k].id, false});
        }
      }
    }
  }
  // pending elementwise expressions among the targets read the variables' CURRENT values: compute them before the optimizer writes
  for (auto& t : targets) { auto it = storage.find(t.id); if (it != storage.end() && it->second.ok) for (auto& y : it->second.ys) if (y.expr) y = expr_materialize(dev, y); }
  // all gradients of this run exist now: (allreduce +) ONE fused multi-tensor optimizer launch (north_star item 5)
  flush_pending_updates(run, env);

  std::vector<EvalResult> ret;
  for (auto& t : targets) {
    EvalResult r; TensorInternal& n = g->inner(t.id);
    if (n.is_variable()) r.value = env->array_list[n.variable_id.v];                      // case 1 (:347-349) (cloned on fetch)
    else if (n.is_placeholder) r.value = find_placeholder_value(feeds, g, t.id);          // case 2
    else {
      auto it = storage.find(t.id);
      if (it == storage.end()) throw Panic("eval: the same tensor was requested twice");   // storage.take(..).unwrap()
      if (!it->second.ok) { r.ok = false; r.err_code = it->second.code; r.err_msg = it->second.msg; }
      else r.value = it->second.ys[0];
      storage.erase(it);
    }
    if (r.ok && r.value.lazy) r.value = materialize_lazy(dev, r.value);
    if (r.ok && r.value.virt && !r.value.on_device() && !r.value.has_host()) {
      extern NdArray materialize_im2col(Device*, const NdArray&);
      r.value = materialize_im2col(dev, r.value);
    }
    if (r.ok && fetch_to_host) dev->ensure_host(r.value);
    ret.push_back(std::move(r));
  }
  if (!fetch_to_host) return ret;
  dev->sync();      // surfaces device-side index errors (bad labels / gather ids) as OutOfBounds
  return ret;
}

// ================================================================================================ variables
VariableEnvironment::VariableEnvironment(int device_index) { dev = new Device(device_index); owns_dev = true; }
VariableEnvironment::~VariableEnvironment() { array_list.clear(); if (owns_dev) delete dev; }

VariableID VariableEnvironment::set(const std::string& ns, const std::string& name, const Shape& shape, const float* data) {
  NdArray a = dev->empty(shape);
  if (a.size() > 0) { check_status(agb_h2d(dev->ctx, a.dptr, data, (size_t)a.size() * sizeof(float))); dev->sync(); }
  VariableID id{(int)array_list.size()};
  name_to_id[{ns, name}] = id.v;               // register_variable, variable.rs:347-358 (a re-used name re-points to the new slot)
  names.push_back({ns, name});
  array_list.push_back(a);
  return id;
}
VariableID VariableEnvironment::find(const std::string& ns, const std::string& name) const {
  auto it = name_to_id.find({ns, name});
  return it == name_to_id.end() ? VariableID{} : VariableID{it->second};
}
std::vector<VariableID> VariableEnvironment::current_var_ids(const std::string& ns) const {
  std::vector<VariableID> r;
  for (auto& kv : name_to_id) if (kv.first.first == ns) r.push_back(VariableID{kv.second});
  std::sort(r.begin(), r.end(), [](VariableID a, VariableID b) { return a.v < b.v; });   // deterministic (the reference's FxHashMap order is unspecified)
  return r;
}
std::vector<float> VariableEnvironment::get(VariableID v) {
  NdArray a = array_list.at(v.v); a.host.reset();
  return dev->ensure_host(a);
}
void VariableEnvironment::put(VariableID v, const float* data, size_t n) {
  NdArray& a = array_list.at(v.v);
  if ((int64_t)n != a.size()) throw OpError(AGB_ERR_INCOMPATIBLE_SHAPE, "VariableEnvironment::put: size mismatch");
  if (n) { check_status(agb_h2d(dev->ctx, a.dptr, data, n * sizeof(float))); dev->sync(); }
}

// JSON checkpoint in the reference's serde layout (variable.rs:549-598; ndarray's {"v":1,"dim":[..],"data":[..]}),
// names serialised as "namespacename" (variable.rs:225-229).
static std::string json_escape(const std::string& in) {
  std::string r; char buf[8];
  for (unsigned char c : in) {
    if (c == '"' || c == '\\') { r.push_back('\\'); r.push_back((char)c); }
    else if (c < 0x20) { snprintf(buf, sizeof(buf), "\\u%04x", (unsigned)c); r += buf; }
    else r.push_back((char)c);
  }
  return r;
}
std::string VariableEnvironment::save_json() {
  std::ostringstream o; o.precision(9);
  o << "{\"array_list\":[";
  for (size_t i = 0; i < array_list.size(); i++) {
    std::vector<float> h = get(VariableID{(int)i});
    if (i) o << ",";
    o << "{\"v\":1,\"dim\":[";
    for (size_t d = 0; d < array_list[i].shape.size(); d++) { if (d) o << ","; o << array_list[i].shape[d]; }
    o << "],\"data\":[";
    for (size_t k = 0; k < h.size(); k++) { if (k) o << ","; if (std::isfinite(h[k])) o << h[k]; else o << "null"; }      // serde_json writes non-finite floats as null
    o << "]}";
  }
  o << "],\"name_to_id\":{";
  bool first = true;
  for (auto& kv : name_to_id) {
    if (!first) o << ","; first = false;
    o << "\"" << json_escape(kv.first.first) << "\\u0001" << json_escape(kv.first.second) << "\":" << kv.second;
  }
  o << "}}";
  return o.str();
}

namespace {
struct JsonCur {
  const std::string& s; size_t i = 0;
  explicit JsonCur(const std::string& str) : s(str) {}
  void ws() { while (i < s.size() && (s[i] == ' ' || s[i] == '\n' || s[i] == '\t' || s[i] == '\r')) i++; }
  bool eat(char c) { ws(); if (i < s.size() && s[i] == c) { i++; return true; } return false; }
  void expect(char c) { if (!eat(c)) throw OpError(AGB_ERR_NDARRAY, std::string("load: malformed checkpoint near '") + c + "'"); }
  std::string str() {
    expect('"'); std::string r;
    while (i < s.size() && s[i] != '"') {
      if (s[i] == '\\' && i + 5 < s.size() && s[i + 1] == 'u') { r.push_back((char)strtol(s.substr(i + 2, 4).c_str(), nullptr, 16)); i += 6; }
      else if (s[i] == '\\' && i + 1 < s.size()) { r.push_back(s[i + 1]); i += 2; }
      else r.push_back(s[i++]);
    }
    expect('"'); return r;
  }
  double num() { ws(); if (s.compare(i, 4, "null") == 0) { i += 4; return std::numeric_limits<double>::quiet_NaN(); } size_t j = i; while (j < s.size() && (isdigit((unsigned char)s[j]) || strchr("+-.eE", s[j]))) j++; double v = atof(s.substr(i, j - i).c_str()); i = j; return v; }
};
}  // namespace

void VariableEnvironment::load_json(const std::string& js) {
  JsonCur c(js);
  std::vector<std::pair<Shape, std::vector<float>>> arrays; std::map<std::string, int> ids;
  c.expect('{');
  do {
    std::string key = c.str(); c.expect(':');
    if (key == "array_list") {
      c.expect('[');
      if (!c.eat(']')) {
        do {
          Shape dim; std::vector<float> data;
          c.expect('{');
          do {
            std::string k = c.str(); c.expect(':');
            if (k == "dim") { c.expect('['); if (!c.eat(']')) { do dim.push_back((int64_t)c.num()); while (c.eat(',')); c.expect(']'); } }
            else if (k == "data") { c.expect('['); if (!c.eat(']')) { do data.push_back((float)c.num()); while (c.eat(',')); c.expect(']'); } }
            else c.num();
          } while (c.eat(','));
          c.expect('}');
          arrays.push_back({dim, data});
        } while (c.eat(','));
        c.expect(']');
      }
    } else if (key == "name_to_id") {
      c.expect('{');
      if (!c.eat('}')) { do { std::string k = c.str(); c.expect(':'); ids[k] = (int)c.num(); } while (c.eat(',')); c.expect('}'); }
    }
  } while (c.eat(','));
  array_list.clear(); names.assign(arrays.size(), {"", ""}); name_to_id.clear();
  for (auto& a : arrays) {
    NdArray d = dev->empty(a.first);
    if (d.size() != (int64_t)a.second.size()) throw OpError(AGB_ERR_NDARRAY, "load: dim/data mismatch");
    if (d.size()) check_status(agb_h2d(dev->ctx, d.dptr, a.second.data(), a.second.size() * sizeof(float)));
    array_list.push_back(d);
  }
  dev->sync();
  for (auto& kv : ids) {
    if (kv.second < 0 || kv.second >= (int)arrays.size()) throw OpError(AGB_ERR_NDARRAY, "load: name_to_id refers to a variable id outside array_list");
    size_t p = kv.first.find('\x01');
    std::string ns = p == std::string::npos ? "" : kv.first.substr(0, p), nm = p == std::string::npos ? kv.first : kv.first.substr(p + 1);
    name_to_id[{ns, nm}] = kv.second;
    if (kv.second >= 0 && kv.second < (int)names.size()) names[kv.second] = {ns, nm};
  }
}

}  // namespace agx
