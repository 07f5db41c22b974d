// fuse.cc — deferred elementwise expressions ("elementwise fusion of backward chains produced by Op::grad
// compositions ... must keep unfused intermediates available when a test evaluates them").
//
// The reference evaluates every node of a composition such as  gy * (y - square(y))  (Sigmoid::grad, activation_ops.rs:150),
// gy * (1 - square(y))  (Tanh::grad, math_ops.rs:854-858)  or the LSTM cell  sigmoid(f) * c + sigmoid(i) * tanh(g)
// (examples/lstm_lm.rs:36-45) as its own ndarray pass.  Here the unary / binary / compare / small AddN ops return an EXPRESSION
// array (shape only, `NdArray::expr`) instead of launching; the first consumer that needs memory (a GEMM, a reduction, a slice, the
// user) compiles the pending DAG into one `agb_fused_ewise` program:
//   * leaves are the device arrays the DAG reads — sliced and row/column-broadcast views are read in place;
//   * every node with more than one consumer in this evaluation (counted by eval()'s pre-pass; the forward values the backward
//     pass re-reads) is stored as an extra output of the same launch, so nothing is ever recomputed and every intermediate of the
//     reference graph stays observable;
//   * each instruction is the functor the single-op kernel applies, so fused and unfused runs are bit-identical.
// Programs are bounded (AGB_FUSE_MAX_*); an operand sub-DAG that would overflow is materialised first and becomes a leaf.
#include "agx.h"
#include <algorithm>
#include <math.h>
#include <functional>
#include <unordered_map>

namespace agx {

struct ExprNode {
  int kind = AGB_F_UNARY, op = 0; float p0 = 0.f;
  NdArray a, b;                 // operands: an expression (a.expr) of this node's shape, or a device array broadcastable to it
  Shape shape;
  int consumers = 1;            // consuming edges in this evaluation (+1 when the node is a requested target)
  int n_instr = 1, n_leaves = 0, n_multi = 0;     // upper bounds over the not-yet-computed part of the DAG below (shared nodes count twice)
  NdArray value; bool has_value = false;
  std::vector<int64_t> pad_start;     // kind == kPad (SliceGrad / SplitGrad): `a` (any shape-compatible array or expression) placed at this offset of a
                                      // zero array of `shape`.  Never part of a program: materialised on its own, or summed in place by AddN
};

namespace {
const int kPad = 100, kGemmTA = 101, kScatter = 102, kColSum = 103;     // nodes that are memory, not instructions: never part of a program
void materialize_node(Device* dev, ExprNode* n, NdArray* dest = nullptr);
const int64_t kMaxFusedElems = (int64_t)1 << 21;     // the interpreter sustains ~1.5 TB/s of operand traffic: measured break-even with the vectorised
                                                     // single-op kernels is ~2^22 elements for a 3-instruction chain (profiles/ops_r1.jsonl), above it they win

bool unvalued(const NdArray& x) { return x.expr && !x.expr->has_value; }
const NdArray& resolved(const NdArray& x) { return x.expr ? x.expr->value : x; }

// a leaf read as ptr + r * pitch + c * cstride over [rows, cols] = [prod(shape[:-1]), shape[-1]]
bool as_2d(const NdArray& leaf, const Shape& out, int64_t& pitch, int64_t& cs) {
  pitch = 0; cs = 0;
  if (leaf.size() == 1) return true;
  if (leaf.ndim() != (int)out.size()) return false;
  const int n = (int)out.size();
  for (int d = 0; d < n; d++) if (leaf.shape[d] != out[d] && leaf.shape[d] != 1) return false;
  auto eff = [&](int d) { return leaf.shape[d] == 1 ? (int64_t)0 : leaf.stride[d]; };
  if (out[n - 1] != 1) cs = eff(n - 1);
  bool first = true; int64_t expected = 0;
  for (int d = n - 2; d >= 0; d--) {
    if (out[d] == 1) continue;
    if (first) { pitch = eff(d); expected = pitch * out[d]; first = false; }
    else { if (eff(d) != expected) return false; expected *= out[d]; }
  }
  return true;
}

bool leaf_ok(Device* dev, NdArray& x, const Shape& out) {
  if (x.lazy || x.i32 || x.virt) return false;
  if (!x.on_device()) { if (!x.has_host()) return false; dev->ensure_device(x); }
  int64_t p, c; return as_2d(x, out, p, c);
}

// value ids: [0, n_leaves) leaves, then one per instruction
struct Compiler {
  Device* dev; ExprNode* root; const Shape& shape;
  std::vector<ExprNode*> order; std::unordered_map<ExprNode*, int> instr_of;
  struct Leaf { NdArray arr; int64_t pitch, cs; };
  std::vector<Leaf> leaves;
  bool ok = true;
  Compiler(Device* d, ExprNode* r) : dev(d), root(r), shape(r->shape) {}
  void visit(ExprNode* n) {
    if (instr_of.count(n)) return;
    if (unvalued(n->a)) visit(n->a.expr.get());
    if (n->kind == AGB_F_BINARY && unvalued(n->b)) visit(n->b.expr.get());
    instr_of[n] = (int)order.size(); order.push_back(n);
  }
  int leaf_id(const NdArray& x) {
    int64_t p, c;
    if (!as_2d(x, shape, p, c)) { ok = false; return 0; }
    for (size_t i = 0; i < leaves.size(); i++) if (leaves[i].arr.dptr == x.dptr && leaves[i].pitch == p && leaves[i].cs == c) return (int)i;
    leaves.push_back(Leaf{x, p, c}); return (int)leaves.size() - 1;
  }
};

struct Compiled {
  std::vector<ExprNode*> order;            // instruction k computes order[k]
  std::vector<NdArray> leaf_arrays; std::vector<agb_fuse_leaf> leaves;     // leaves[l].pitch is the 2-D pitch (the launcher may flatten)
  std::vector<agb_fuse_instr> code;
  std::vector<int> outs;                   // instruction indices stored: the roots first, then the other multi-consumer nodes
};
// host-only half: DAG -> leaves, instructions, registers, outputs (no device call; exercised on the CPU by agx_fuse_selftest)
bool compile_program(const std::vector<ExprNode*>& roots, Compiled* out) {
  ExprNode* root = roots[0];
  if ((int)roots.size() > AGB_FUSE_MAX_OUT) return false;
  Compiler C(nullptr, root);
  for (ExprNode* r : roots) { if (r->shape != root->shape || r->has_value || r->kind >= kPad) return false; C.visit(r); }
  const int I = (int)C.order.size();
  if (I > AGB_FUSE_MAX_INSTR) return false;
  // operands -> value ids (leaves first; instruction k is value L + k, fixed up once L is known)
  struct Opnd { bool leaf; int id; };
  std::vector<Opnd> oa(I), ob(I);
  auto operand = [&](const NdArray& x) { if (unvalued(x)) return Opnd{false, C.instr_of[x.expr.get()]}; return Opnd{true, C.leaf_id(resolved(x))}; };
  for (int k = 0; k < I; k++) {
    ExprNode* n = C.order[k];
    if (n->kind != AGB_F_BINARY_IMM_A) oa[k] = operand(n->a); else oa[k] = Opnd{true, -1};
    if (n->kind == AGB_F_BINARY) ob[k] = operand(n->b); else if (n->kind == AGB_F_BINARY_IMM_A) ob[k] = operand(n->a); else ob[k] = Opnd{true, -1};
  }
  const int L = (int)C.leaves.size();
  if (!C.ok || L > AGB_FUSE_MAX_LEAVES) return false;
  // outputs: the root + every other node somebody else will read
  std::vector<int> outs; std::vector<char> is_out(I, 0);
  for (size_t i = 0; i < roots.size(); i++) { int k = C.instr_of[roots[i]]; if (is_out[k]) return false; is_out[k] = 1; outs.push_back(k); }
  for (int k = 0; k < I && (int)outs.size() < AGB_FUSE_MAX_OUT; k++) if (!is_out[k] && C.order[k]->consumers > 1) { is_out[k] = 1; outs.push_back(k); }
  // linear-scan register allocation
  const int V = L + I;
  auto vid = [&](const Opnd& o) { return o.id < 0 ? -1 : (o.leaf ? o.id : L + o.id); };
  std::vector<int> last(V, -1), reg(V, -1); std::vector<char> released(V, 0);
  for (int k = 0; k < I; k++) { int x = vid(oa[k]), y = vid(ob[k]); if (x >= 0) last[x] = k; if (y >= 0) last[y] = k; }
  std::vector<int> free_regs; for (int r = AGB_FUSE_REGS - 1; r >= 0; r--) free_regs.push_back(r);
  for (int l = 0; l < L; l++) { reg[l] = free_regs.back(); free_regs.pop_back(); }
  std::vector<agb_fuse_instr> code(I);
  for (int k = 0; k < I; k++) {
    ExprNode* n = C.order[k];
    int x = vid(oa[k]), y = vid(ob[k]);
    agb_fuse_instr& ins = code[k];
    ins.kind = n->kind; ins.op = n->op; ins.p0 = n->p0; ins.a = x >= 0 ? reg[x] : 0; ins.b = y >= 0 ? reg[y] : 0;
    auto release = [&](int v) { if (v >= 0 && last[v] == k && !released[v] && !(v >= L && is_out[v - L])) { free_regs.push_back(reg[v]); released[v] = 1; } };
    release(x); if (y != x) release(y);
    if (free_regs.empty()) return false;
    reg[L + k] = free_regs.back(); free_regs.pop_back();
    ins.dst = reg[L + k];
    if (last[L + k] < 0 && !is_out[k]) { free_regs.push_back(reg[L + k]); }      // dead value (cannot happen for a DAG reachable from the root)
  }
  out->order = C.order; out->code = code; out->outs = outs;
  out->leaves.resize(L); out->leaf_arrays.resize(L);
  for (int l = 0; l < L; l++) { out->leaf_arrays[l] = C.leaves[l].arr; out->leaves[l].ptr = C.leaves[l].arr.dptr; out->leaves[l].pitch = C.leaves[l].pitch; out->leaves[l].cstride = C.leaves[l].cs; out->leaves[l].reg = reg[l]; }
  return true;
}

// One launch for the DAG below `roots` (all of one shape).  dests[i] with a device pointer = root i is written there (a strided region of
// a larger buffer); otherwise a fresh array is allocated.
bool run_program(Device* dev, const std::vector<ExprNode*>& roots, const std::vector<NdArray>& dests) {
  Compiled P;
  if (!compile_program(roots, &P)) return false;
  ExprNode* root = roots[0];
  const int L = (int)P.leaves.size(), I = (int)P.code.size();
  const std::vector<int>& outs = P.outs; const std::vector<agb_fuse_instr>& code = P.code;
  bool any_dest = false; for (auto& d : dests) if (d.on_device()) any_dest = true;
  // launch
  const int nd = (int)root->shape.size();
  int64_t cols = nd == 0 ? 1 : root->shape[nd - 1], total = 1; for (auto d : root->shape) total *= d;
  int64_t rows = cols == 0 ? 0 : total / cols;
  bool flat = !any_dest;
  for (auto& lf : P.leaves) if (!((lf.pitch == 0 && lf.cstride == 0) || (lf.cstride == 1 && (lf.pitch == cols || rows == 1)))) flat = false;
  std::vector<agb_fuse_leaf> lv = P.leaves;
  if (flat) for (auto& lf : lv) lf.pitch = 0;
  std::vector<agb_fuse_out> ov(outs.size()); std::vector<NdArray> values(outs.size());
  for (size_t o = 0; o < outs.size(); o++) {
    if (o < roots.size() && dests[o].on_device()) { int64_t p, cs; as_2d(dests[o], root->shape, p, cs); values[o] = dests[o]; ov[o].pitch = p; }
    else { values[o] = dev->empty(root->shape); ov[o].pitch = flat ? 0 : cols; }
    ov[o].ptr = values[o].dptr; ov[o].reg = code[outs[o]].dst;
  }
  check_status(agb_fused_ewise(dev->ctx, flat ? 1 : rows, flat ? total : cols, L, lv.data(), I, code.data(), (int)ov.size(), ov.data()));
  for (size_t o = 0; o < outs.size(); o++) {
    ExprNode* n = P.order[outs[o]];
    n->value = values[o]; n->has_value = true; n->a = NdArray(); n->b = NdArray();      // operands are no longer needed
  }
  return true;
}

NdArray pad_region(const NdArray& full, const std::vector<int64_t>& start, const Shape& part) {
  NdArray r = full;
  for (int k = 0; k < full.ndim(); k++) r = r.sliced(k, start[k], part[k]);
  return r;
}
void write_region(Device* dev, NdArray src, NdArray region) {      // region <- src: a pending expression is computed straight into it
  if (unvalued(src) && expr_materialize_into(dev, src, region)) return;
  if (src.expr) src = expr_materialize(dev, src);
  dev->ensure_device(src);
  agb_tensor ts = src.desc(), td = region.desc();
  check_status(agb_copy_strided(dev->ctx, &ts, &td));
}

void materialize_node(Device* dev, ExprNode* n, NdArray* dest) {
  if (n->has_value) return;
  if (n->kind == kColSum) {       // MaybeReduceSum on its own: sum over the rows
    NdArray y = dev->empty(n->shape);
    check_status(agb_reduce(dev->ctx, AGB_R_SUM, n->a.dptr, y.dptr, 1, n->a.shape[0], n->a.shape[1]));
    n->value = y; n->has_value = true; n->a = NdArray();
    return;
  }
  if (n->kind == kScatter) {      // GatherGrad on its own: zero table + scatter-add (array_ops.rs:401-466)
    NdArray gx = dev->empty(n->shape);
    int64_t pre = 1, post = 1; for (int k = 0; k < n->op; k++) pre *= n->shape[k]; for (int k = n->op + 1; k < (int)n->shape.size(); k++) post *= n->shape[k];
    NdArray idx = dev->contiguous(n->a);
    check_status(agb_gather_grad(dev->ctx, n->b.dptr, idx.dptr, gx.dptr, pre, n->shape[n->op], post, idx.size()));
    n->value = gx; n->has_value = true; n->a = NdArray(); n->b = NdArray();
    return;
  }
  if (n->kind == kGemmTA) {
    NdArray y = dev->empty(n->shape);
    agb_tensor da = n->a.desc(), db = n->b.desc(), dy = y.desc();
    check_status(agb_gemm_f32(dev->ctx, 1, 0, &da, &db, &dy, 0.0f));
    n->value = y; n->has_value = true; n->a = NdArray(); n->b = NdArray();
    return;
  }
  if (n->kind == kPad) {
    NdArray gx = dev->zeros(n->shape);
    write_region(dev, n->a, pad_region(gx, n->pad_start, n->a.shape));
    n->value = gx; n->has_value = true; n->a = NdArray();
    return;
  }
  const std::vector<ExprNode*> roots{n}; const std::vector<NdArray> dests{dest ? *dest : NdArray()};
  if (run_program(dev, roots, dests)) return;
  // the DAG does not fit one program (registers / leaves / instructions): compute the operands first, then this node alone
  if (unvalued(n->a)) materialize_node(dev, n->a.expr.get());
  if (n->kind == AGB_F_BINARY && unvalued(n->b)) materialize_node(dev, n->b.expr.get());
  if (!run_program(dev, roots, dests)) throw Panic("fused elementwise: a single instruction does not fit a program");
}

void account(ExprNode* n) {
  auto add = [&](const NdArray& x) {
    if (unvalued(x)) { n->n_instr += x.expr->n_instr; n->n_leaves += x.expr->n_leaves; n->n_multi += x.expr->n_multi; }
    else n->n_leaves += 1;
  };
  n->n_instr = 1; n->n_leaves = 0; n->n_multi = n->consumers > 1 ? 1 : 0;
  add(n->a); if (n->kind == AGB_F_BINARY) add(n->b);
}

NdArray finish(ComputeContext& c, std::shared_ptr<ExprNode> n) {
  n->consumers = c.run->consumers_of(c.node);
  account(n.get());
  // keep the pending DAG inside one program: materialise the larger operand first when the bounds would overflow
  for (int round = 0; round < 2; round++) {
    if (n->n_instr <= AGB_FUSE_MAX_INSTR - 8 && n->n_leaves <= AGB_FUSE_MAX_LEAVES - 1 && n->n_multi <= AGB_FUSE_MAX_OUT - 1) break;
    NdArray* big = nullptr;
    if (unvalued(n->a)) big = &n->a;
    if (n->kind == AGB_F_BINARY && unvalued(n->b) && (!big || n->b.expr->n_instr > n->a.expr->n_instr)) big = &n->b;
    if (!big) break;
    materialize_node(c.dev, big->expr.get());
    account(n.get());
  }
  NdArray r; r.shape = n->shape; r.stride = NdArray::contiguous_strides(r.shape); r.expr = n;
  return r;
}

bool operand_ok(Device* dev, NdArray& x, const Shape& out) {
  if (unvalued(x) && x.expr->kind >= kPad) materialize_node(dev, x.expr.get());       // a padded slice gradient is memory, not an instruction
  if (x.expr) { if (x.expr->has_value) { NdArray v = x.expr->value; return leaf_ok(dev, v, out); } return x.shape == out; }
  return leaf_ok(dev, x, out);
}
bool size_ok(const Shape& s) { int64_t n = 1; for (auto d : s) n *= d; return n >= 1 && n <= kMaxFusedElems && s.size() <= 6; }
}  // namespace

NdArray expr_unary(ComputeContext& c, int op, float p0, NdArray x) {
  if (!c.run->fuse || op == AGB_U_CLIP || op < 0 || op >= AGB_U_COUNT || !size_ok(x.shape) || !operand_ok(c.dev, x, x.shape)) return NdArray();
  auto n = std::make_shared<ExprNode>(); n->kind = AGB_F_UNARY; n->op = op; n->p0 = p0; n->a = x; n->shape = x.shape;
  return finish(c, n);
}
NdArray expr_binary(ComputeContext& c, int op, NdArray a, NdArray b) {
  if (!c.run->fuse || op < 0 || op > AGB_B_MIN || a.ndim() != b.ndim()) return NdArray();
  Shape out(a.shape.size());
  for (size_t i = 0; i < out.size(); i++) {
    if (a.shape[i] != b.shape[i] && a.shape[i] != 1 && b.shape[i] != 1) return NdArray();
    out[i] = a.shape[i] == 1 ? b.shape[i] : a.shape[i];
  }
  if (!size_ok(out) || !operand_ok(c.dev, a, out) || !operand_ok(c.dev, b, out)) return NdArray();
  auto n = std::make_shared<ExprNode>(); n->kind = AGB_F_BINARY; n->op = op; n->a = a; n->b = b; n->shape = out;
  return finish(c, n);
}
NdArray expr_binary_imm(ComputeContext& c, int op, NdArray x, float imm, bool imm_is_lhs) {
  if (!c.run->fuse || op < 0 || op > AGB_B_MIN || !size_ok(x.shape) || !operand_ok(c.dev, x, x.shape)) return NdArray();
  auto n = std::make_shared<ExprNode>(); n->kind = imm_is_lhs ? AGB_F_BINARY_IMM_A : AGB_F_BINARY_IMM_B; n->op = op; n->p0 = imm; n->a = x; n->shape = x.shape;
  return finish(c, n);
}
// a view op (MaybeReduceSum / Identity with nothing to do) hands the same pending node to ITS consumers
NdArray expr_passthrough(ComputeContext& c, const NdArray& x) {
  if (x.expr && !x.expr->has_value) { x.expr->consumers += c.run->consumers_of(c.node) - 1; if (x.expr->consumers > 1 && x.expr->n_multi == 0) x.expr->n_multi = 1; }
  return x;
}
// SliceGrad / SplitGrad (array_ops.rs:726-749,803-825): zeros(full) with `gy` assigned at `start`, deferred so that AddN can sum
// the disjoint pieces of one gradient (the 4 gate slices of the LSTM pre-activation) by writing them side by side
NdArray expr_pad(ComputeContext& c, const Shape& full, const std::vector<int64_t>& start, NdArray gy) {
  if (!c.run->fuse || full.size() != gy.shape.size() || full.empty() || gy.lazy || gy.i32) return NdArray();
  if (!gy.expr && !gy.on_device()) { if (!gy.has_host()) return NdArray(); c.dev->ensure_device(gy); }
  auto n = std::make_shared<ExprNode>(); n->kind = kPad; n->a = gy; n->shape = full; n->pad_start = start;
  n->consumers = c.run->consumers_of(c.node); n->n_instr = 0; n->n_leaves = 1; n->n_multi = 0;
  NdArray r; r.shape = full; r.stride = NdArray::contiguous_strides(full); r.expr = n;
  return r;
}
// AddN over padded pieces that differ along ONE axis and do not overlap: one buffer, every piece written in place
bool expr_sum_pads(ComputeContext& c, const std::vector<NdArray>& xs, NdArray* out) {
  if (xs.size() < 2) return false;
  for (auto& x : xs) if (!unvalued(x) || x.expr->kind != kPad || x.shape != xs[0].shape) return false;
  const Shape& full = xs[0].shape; const int nd = (int)full.size();
  int axis = -1;
  for (auto& x : xs) for (int k = 0; k < nd; k++) {
    if (x.expr->pad_start[k] == 0 && x.expr->a.shape[k] == full[k]) continue;
    if (axis >= 0 && axis != k) return false;
    axis = k;
  }
  if (axis < 0) return false;
  std::vector<std::pair<int64_t, int64_t>> spans;
  for (auto& x : xs) spans.push_back({x.expr->pad_start[axis], x.expr->a.shape[axis]});
  std::sort(spans.begin(), spans.end());
  int64_t at = 0; bool covered = true;
  for (auto& sp : spans) { if (sp.first < at) return false; if (sp.first > at) covered = false; at = sp.first + sp.second; }
  if (at != full[axis]) covered = false;
  NdArray y = covered ? c.dev->empty(full) : c.dev->zeros(full);
  // pieces that are pending expressions of one shape (the four gate gradients share most of their DAG): ONE program with several roots,
  // each stored into its own region
  std::vector<ExprNode*> roots; std::vector<NdArray> dests; std::vector<char> in_prog(xs.size(), 0);
  for (size_t i = 0; i < xs.size(); i++) {
    const NdArray& src = xs[i].expr->a;
    NdArray region = pad_region(y, xs[i].expr->pad_start, src.shape);
    int64_t p, cs; const int nd = region.ndim();
    if (unvalued(src) && src.expr->kind < kPad && (roots.empty() || src.expr->shape == roots[0]->shape) && nd > 0 && (region.shape[nd - 1] == 1 || region.stride[nd - 1] == 1) &&
        as_2d(region, region.shape, p, cs) && std::find(roots.begin(), roots.end(), src.expr.get()) == roots.end()) { roots.push_back(src.expr.get()); dests.push_back(region); in_prog[i] = 1; }
  }
  if (roots.size() < 2 || !run_program(c.dev, roots, dests)) std::fill(in_prog.begin(), in_prog.end(), 0);
  for (size_t i = 0; i < xs.size(); i++) if (!in_prog[i]) write_region(c.dev, xs[i].expr->a, pad_region(y, xs[i].expr->pad_start, xs[i].expr->a.shape));
  *out = y;
  return true;
}
// rows of equally-shaped 2-D blocks stacked into one matrix (one launch per 64 blocks); the stack is remembered for the rest of the run
bool stackable(const NdArray& t) { return t.ndim() == 2 && t.stride[1] == 1 && t.stride[0] % 4 == 0 && t.shape[1] % 4 == 0 && (((uintptr_t)t.dptr) & 15) == 0; }
NdArray stack_rows(Evaluation& run, Device* dev, const std::vector<NdArray>& parts) {
  std::vector<const float*> ps(parts.size()); std::vector<int64_t> pitch(parts.size());
  for (size_t i = 0; i < parts.size(); i++) { ps[i] = parts[i].dptr; pitch[i] = parts[i].stride[0]; }
  {   // consecutive row blocks of ONE buffer (slices of a stacked GEMM / stacked op output): the stack is a view, nothing is copied
    bool consecutive = parts[0].buf != nullptr && parts[0].stride[0] == parts[0].shape[1];
    const int64_t block = parts[0].shape[0] * parts[0].shape[1];
    for (size_t i = 1; consecutive && i < parts.size(); i++) consecutive = parts[i].buf == parts[0].buf && parts[i].dptr == parts[0].dptr + (int64_t)i * block && parts[i].stride[0] == parts[i].shape[1];
    if (consecutive) { NdArray v = parts[0]; v.shape[0] = (int64_t)parts.size() * parts[0].shape[0]; v.host.reset(); v.chan_sum.reset(); return v; }
  }
  for (auto& e : run.row_stacks) if (e.key == ps && e.stacked.shape[1] == parts[0].shape[1] && e.parts[0].shape == parts[0].shape) return e.stacked;
  NdArray S = dev->empty({(int64_t)parts.size() * parts[0].shape[0], parts[0].shape[1]});
  check_status(agb_concat_rows(dev->ctx, (int)parts.size(), ps.data(), pitch.data(), parts[0].shape[0], parts[0].shape[1], S.dptr));
  run.row_stacks.push_back(Evaluation::RowStack{ps, parts, S});
  return S;
}
// per-member vectors (labels [B] or [B, 1], any strides) -> one contiguous [n * B] vector; evenly spaced column views of one array (the
// token-id columns of one feed) take a single strided copy
NdArray stack_vectors(Evaluation& run, Device* dev, const std::vector<NdArray>& parts) {
  const int n = (int)parts.size(); const int64_t B = parts[0].size();
  std::vector<const float*> ps(n); for (int i = 0; i < n; i++) ps[i] = parts[i].dptr;
  for (auto& e : run.row_stacks) if (e.key == ps && e.stacked.ndim() == 1 && e.stacked.shape[0] == n * B && e.parts[0].shape == parts[0].shape && e.parts[0].stride == parts[0].stride) return e.stacked;
  auto step_of = [](const NdArray& t) { for (int k = 0; k < t.ndim(); k++) if (t.shape[k] != 1) return t.stride[k]; return (int64_t)1; };
  const int64_t s0 = step_of(parts[0]), delta = n > 1 ? parts[1].dptr - parts[0].dptr : 0;
  bool even = true;
  for (int i = 0; i < n && even; i++) even = parts[i].size() == B && step_of(parts[i]) == s0 && parts[i].dptr - parts[0].dptr == (int64_t)i * delta && parts[i].buf == parts[0].buf;
  NdArray L = dev->empty({n * B});
  if (even) {
    agb_tensor ts, td; ts.ptr = parts[0].dptr; ts.rank = 2; ts.shape[0] = n; ts.shape[1] = B; ts.stride[0] = delta; ts.stride[1] = s0;
    td.ptr = L.dptr; td.rank = 2; td.shape[0] = n; td.shape[1] = B; td.stride[0] = B; td.stride[1] = 1;
    check_status(agb_copy_strided(dev->ctx, &ts, &td));
  } else {
    for (int i = 0; i < n; i++) {
      agb_tensor ts, td; ts.ptr = parts[i].dptr; ts.rank = 1; ts.shape[0] = B; ts.stride[0] = step_of(parts[i]);
      td.ptr = L.dptr + (int64_t)i * B; td.rank = 1; td.shape[0] = B; td.stride[0] = 1;
      check_status(agb_copy_strided(dev->ctx, &ts, &td));
    }
  }
  run.row_stacks.push_back(Evaluation::RowStack{ps, parts, L});
  return L;
}
NdArray expr_colsum(ComputeContext& c, NdArray gy, const Shape& target) {
  if (gy.ndim() != 2 || target.size() != 2 || target[0] != 1 || target[1] != gy.shape[1] || gy.shape[0] < 2 || !gy.on_device() || !gy.is_contiguous() || gy.lazy || gy.i32) return NdArray();
  auto n = std::make_shared<ExprNode>(); n->kind = kColSum; n->a = gy; n->shape = target;
  n->consumers = 1; n->n_instr = 0; n->n_leaves = 1; n->n_multi = 0;
  NdArray r; r.shape = target; r.stride = NdArray::contiguous_strides(target); r.expr = n;
  return r;
}
// AddN over deferred row sums of equally-shaped blocks (the bias gradient of an unrolled RNN): ONE reduction over the stacked rows
bool expr_sum_colsums(ComputeContext& c, const std::vector<NdArray>& xs, NdArray* out) {
  if (xs.size() < 2) return false;
  std::vector<NdArray> parts;
  for (auto& x : xs) { if (!unvalued(x) || x.expr->kind != kColSum || x.expr->a.shape != xs[0].expr->a.shape || !stackable(x.expr->a)) return false; parts.push_back(x.expr->a); }
  NdArray S = stack_rows(*c.run, c.dev, parts);
  NdArray y = c.dev->empty(xs[0].expr->shape);
  check_status(agb_reduce(c.dev->ctx, AGB_R_SUM, S.dptr, y.dptr, 1, S.shape[0], S.shape[1]));
  *out = y;
  return true;
}
NdArray expr_gemm_ta(ComputeContext& c, NdArray a, NdArray b) {
  if (a.ndim() != 2 || b.ndim() != 2 || !a.on_device() || !b.on_device() || a.shape[0] != b.shape[0] || a.lazy || b.lazy || a.i32 || b.i32) return NdArray();
  auto n = std::make_shared<ExprNode>(); n->kind = kGemmTA; n->a = a; n->b = b; n->shape = {a.shape[1], b.shape[1]};
  n->consumers = 1; n->n_instr = 0; n->n_leaves = 1; n->n_multi = 0;
  NdArray r; r.shape = n->shape; r.stride = NdArray::contiguous_strides(r.shape); r.expr = n;
  return r;
}
// AddN over deferred A_t^T * G_t terms of one shape: stack the A_t and the G_t (one launch each per 64 terms) and run one GEMM with
// K = sum of the terms' K.  The result differs from the term-by-term sum only by fp32 reassociation.
bool expr_sum_gemms(ComputeContext& c, const std::vector<NdArray>& xs, NdArray* out) {
  if (xs.size() < 2) return false;
  for (auto& x : xs) if (!unvalued(x) || x.expr->kind != kGemmTA || x.expr->a.shape != xs[0].expr->a.shape || x.expr->b.shape != xs[0].expr->b.shape) return false;
  std::vector<NdArray> as, bs;
  for (auto& x : xs) { if (!stackable(x.expr->a) || !stackable(x.expr->b)) return false; as.push_back(x.expr->a); bs.push_back(x.expr->b); }
  const int64_t M = xs[0].expr->a.shape[1], N = xs[0].expr->b.shape[1];
  NdArray A = stack_rows(*c.run, c.dev, as), B = stack_rows(*c.run, c.dev, bs);
  NdArray y = c.dev->empty({M, N});
  agb_tensor da = A.desc(), db = B.desc(), dy = y.desc();
  check_status(agb_gemm_f32(c.dev->ctx, 1, 0, &da, &db, &dy, 0.0f));
  *out = y;
  return true;
}
NdArray expr_scatter(ComputeContext& c, const Shape& table, int axis, NdArray idx, NdArray gy) {
  if (!idx.on_device() || !gy.on_device() || !gy.is_contiguous() || axis < 0 || axis >= (int)table.size()) return NdArray();      // idx: any strided view
  auto n = std::make_shared<ExprNode>(); n->kind = kScatter; n->op = axis; n->a = idx; n->b = gy; n->shape = table;
  n->consumers = 1; n->n_instr = 0; n->n_leaves = 1; n->n_multi = 0;
  NdArray r; r.shape = table; r.stride = NdArray::contiguous_strides(table); r.expr = n;
  return r;
}
// AddN over deferred GatherGrads of one table: ONE zero fill, every term scatter-adds into the same buffer
bool expr_sum_scatters(ComputeContext& c, const std::vector<NdArray>& xs, NdArray* out) {
  if (xs.size() < 2) return false;
  for (auto& x : xs) if (!unvalued(x) || x.expr->kind != kScatter || x.expr->shape != xs[0].expr->shape || x.expr->op != xs[0].expr->op) return false;
  const Shape& table = xs[0].expr->shape; const int ax = xs[0].expr->op;
  int64_t pre = 1, post = 1; for (int k = 0; k < ax; k++) pre *= table[k]; for (int k = ax + 1; k < (int)table.size(); k++) post *= table[k];
  NdArray gx = c.dev->zeros(table);
  {   // token-id vectors and gradient row blocks that stack (the gy_t are slices of one stacked GEMM output): ONE scatter-add for all terms
    bool vec = ax == 0 && pre == 1 && post % 4 == 0;
    std::vector<NdArray> ids, gys;
    for (auto& x : xs) {
      const NdArray& id = x.expr->a; int nontrivial = 0; for (auto d : id.shape) if (d != 1) nontrivial++;
      if (nontrivial > 1 || id.shape != xs[0].expr->a.shape || x.expr->b.size() != id.size() * post) vec = false;
      if (!vec) break;
      NdArray g2 = x.expr->b; g2.shape = {id.size(), post}; g2.stride = {post, 1};       // contiguous [B, ..., post] viewed as [B, post]
      if (!stackable(g2)) { vec = false; break; }
      ids.push_back(id); gys.push_back(g2);
    }
    if (vec) {
      NdArray L = stack_vectors(*c.run, c.dev, ids), G = stack_rows(*c.run, c.dev, gys);
      check_status(agb_scatter_add(c.dev->ctx, G.dptr, L.dptr, gx.dptr, 1, table[0], post, L.size()));
      *out = gx;
      return true;
    }
  }
  for (auto& x : xs) { NdArray idx = c.dev->contiguous(x.expr->a); check_status(agb_scatter_add(c.dev->ctx, x.expr->b.dptr, idx.dptr, gx.dptr, pre, table[ax], post, idx.size())); }
  *out = gx;
  return true;
}
// Slice of a pending expression: elementwise ops commute with slicing, so the slice is the same (small) expression over sliced leaves and
// the full-size value (the LSTM pre-activation x*wx + h*wh + b, consumed only through its four gate slices) is never written.
NdArray expr_slice(ComputeContext& c, const NdArray& x, const std::vector<int64_t>& start, const std::vector<int64_t>& len) {
  if (!c.run->fuse || !unvalued(x) || x.expr->kind >= kPad || x.expr->n_instr > 8 || (int)start.size() != x.ndim()) return NdArray();
  for (auto l : len) if (l <= 0) return NdArray();
  Shape part(len.begin(), len.end());
  bool ok = true; ExprNode* top = x.expr.get();
  std::function<NdArray(const NdArray&)> clone = [&](const NdArray& o) -> NdArray {
    if (!ok) return NdArray();
    if (unvalued(o)) {
      ExprNode* s = o.expr.get();
      if (s->kind >= kPad || s->shape != x.shape || (s != top && s->consumers > 1)) { ok = false; return NdArray(); }      // an interior node somebody else reads: keep it whole
      auto n = std::make_shared<ExprNode>(); n->kind = s->kind; n->op = s->op; n->p0 = s->p0; n->shape = part; n->consumers = 1;
      n->a = clone(s->a); if (s->kind == AGB_F_BINARY) n->b = clone(s->b);
      if (!ok) return NdArray();
      account(n.get());
      NdArray r; r.shape = part; r.stride = NdArray::contiguous_strides(part); r.expr = n;
      return r;
    }
    NdArray leaf = resolved(o);
    if (leaf.size() == 1) return leaf;
    if (leaf.ndim() != x.ndim()) { ok = false; return NdArray(); }
    for (int k = 0; k < leaf.ndim(); k++) if (leaf.shape[k] != 1) leaf = leaf.sliced(k, start[k], len[k]);
    leaf.host.reset(); leaf.chan_sum.reset();
    return leaf;
  };
  NdArray r = clone(x);
  if (!ok || !r.expr) return NdArray();
  r.expr->consumers = c.run->consumers_of(c.node);
  account(r.expr.get());
  return r;
}
bool expr_has_value(const NdArray& x) { return x.expr && x.expr->has_value; }
NdArray expr_materialize(Device* dev, const NdArray& x) {
  if (!x.expr) return x;
  materialize_node(dev, x.expr.get());
  return x.expr->value;
}
// SliceGrad: the pending value is written straight into its region of the zero-filled gradient
bool expr_materialize_into(Device* dev, const NdArray& x, NdArray dest) {
  if (!x.expr || x.expr->has_value || x.expr->kind >= kPad || dest.shape != x.shape) return false;      // (memory nodes produce their own buffer)
  const int nd = dest.ndim();
  if (nd == 0 || (dest.shape[nd - 1] != 1 && dest.stride[nd - 1] != 1)) return false;
  int64_t p, cs; if (!as_2d(dest, dest.shape, p, cs)) return false;
  materialize_node(dev, x.expr.get(), &dest);
  return true;
}

// ---- host-only self test of the program compiler (tests/test_abi.py, no device needed): random DAGs over fake leaves are compiled and the
// instruction stream is interpreted on the host for one element; every stored register must equal the direct evaluation of its node.
int fuse_selftest(int n_cases, uint32_t seed, int* n_compiled) {
  uint32_t st = seed ? seed : 1u;
  auto rnd = [&]() { st ^= st << 13; st ^= st >> 17; st ^= st << 5; return st; };
  auto frand = [&]() { return (float)((int)(rnd() % 2001) - 1000) / 500.0f; };
  const int u_ops[] = {AGB_U_NEG, AGB_U_SQUARE, AGB_U_ABS, AGB_U_SCALE, AGB_U_ADD_SCALAR, AGB_U_RSUB_SCALAR};
  const int b_ops[] = {AGB_B_ADD, AGB_B_SUB, AGB_B_MUL, AGB_B_MAX, AGB_B_GT};
  auto un = [](int op, float a, float p) { return op == AGB_U_NEG ? -a : op == AGB_U_SQUARE ? a * a : op == AGB_U_ABS ? fabsf(a) : op == AGB_U_SCALE ? a * p : op == AGB_U_ADD_SCALAR ? a + p : p - a; };
  auto bi = [](int op, float a, float b) { return op == AGB_B_ADD ? a + b : op == AGB_B_SUB ? a - b : op == AGB_B_MUL ? a * b : op == AGB_B_MAX ? (a > b ? a : b) : (a > b ? 1.0f : 0.0f); };
  int compiled = 0;
  for (int cs = 0; cs < n_cases; cs++) {
    const Shape shape{4, 8};
    const int n_leaf = 1 + (int)(rnd() % 24), n_node = 1 + (int)(rnd() % 90);     // some cases exceed the program bounds and must be refused
    std::vector<NdArray> leaves(n_leaf); std::vector<float> leaf_val(n_leaf);
    for (int l = 0; l < n_leaf; l++) { leaves[l].shape = shape; leaves[l].stride = NdArray::contiguous_strides(shape); leaves[l].dptr = (float*)(uintptr_t)(0x10000 * (l + 1)); leaf_val[l] = frand(); }
    std::vector<NdArray> nodes; std::vector<float> node_val;          // expression arrays in creation order
    auto pick = [&](float* v) -> NdArray {                            // an earlier node (preferably a recent one) or a leaf
      if (!nodes.empty() && rnd() % 5 != 0) { size_t k = nodes.size() - 1 - rnd() % std::min<size_t>(nodes.size(), 6); *v = node_val[k]; return nodes[k]; }
      int l = (int)(rnd() % n_leaf); *v = leaf_val[l]; return leaves[l];
    };
    for (int k = 0; k < n_node; k++) {
      auto n = std::make_shared<ExprNode>(); n->shape = shape; n->consumers = rnd() % 4 == 0 ? 2 : 1;
      float va, vb, val; const int f8 = (int)(rnd() % 8), form = f8 < 4 ? 1 : f8 < 6 ? 0 : f8 - 4;      // half of the nodes are binary
      if (form == 0) { n->kind = AGB_F_UNARY; n->op = u_ops[rnd() % 6]; n->p0 = frand(); n->a = pick(&va); val = un(n->op, va, n->p0); }
      else if (form == 1) { n->kind = AGB_F_BINARY; n->op = b_ops[rnd() % 5]; n->a = pick(&va); n->b = pick(&vb); val = bi(n->op, va, vb); }
      else if (form == 2) { n->kind = AGB_F_BINARY_IMM_B; n->op = b_ops[rnd() % 5]; n->p0 = frand(); n->a = pick(&va); val = bi(n->op, va, n->p0); }
      else { n->kind = AGB_F_BINARY_IMM_A; n->op = b_ops[rnd() % 5]; n->p0 = frand(); n->a = pick(&va); val = bi(n->op, n->p0, va); }
      NdArray r; r.shape = shape; r.stride = NdArray::contiguous_strides(shape); r.expr = n;
      nodes.push_back(r); node_val.push_back(val);
    }
    std::vector<ExprNode*> roots{nodes.back().expr.get()};
    if (n_node > 3 && rnd() % 2) roots.push_back(nodes[n_node - 2].expr.get());       // a second root (it may also be an operand of the first)
    Compiled P;
    if (!compile_program(roots, &P)) continue;                        // too many leaves / registers / outputs: the evaluator would split
    compiled++;
    if ((int)P.leaves.size() > AGB_FUSE_MAX_LEAVES || (int)P.code.size() > AGB_FUSE_MAX_INSTR || (int)P.outs.size() > AGB_FUSE_MAX_OUT) return 10;
    float regs[AGB_FUSE_REGS]; bool live[AGB_FUSE_REGS] = {false};
    for (size_t l = 0; l < P.leaves.size(); l++) {
      int idx = (int)((uintptr_t)P.leaves[l].ptr / 0x10000) - 1;
      if (P.leaves[l].reg < 0 || P.leaves[l].reg >= AGB_FUSE_REGS || live[P.leaves[l].reg]) return 11;          // two leaves in one register
      regs[P.leaves[l].reg] = leaf_val[idx]; live[P.leaves[l].reg] = true;
    }
    std::unordered_map<ExprNode*, float> expect;
    for (int k = 0; k < n_node; k++) expect[nodes[k].expr.get()] = node_val[k];
    for (size_t k = 0; k < P.code.size(); k++) {
      const agb_fuse_instr& in = P.code[k];
      if (in.dst < 0 || in.dst >= AGB_FUSE_REGS) return 12;
      float a = regs[in.a], b = regs[in.b], y;
      if (in.kind == AGB_F_UNARY) y = un(in.op, a, in.p0);
      else if (in.kind == AGB_F_BINARY) y = bi(in.op, a, b);
      else if (in.kind == AGB_F_BINARY_IMM_B) y = bi(in.op, a, in.p0);
      else y = bi(in.op, in.p0, b);
      regs[in.dst] = y;
      if (y != expect[P.order[k]] && !(y != y && expect[P.order[k]] != expect[P.order[k]])) return 13;           // an operand register was clobbered
    }
    for (size_t o = 0; o < P.outs.size(); o++) {
      const float got = regs[P.code[P.outs[o]].dst], want = expect[P.order[P.outs[o]]];
      if (got != want && !(got != got && want != want)) return 14;                                               // a stored register was reused before the end
    }
    for (size_t i = 0; i < roots.size(); i++) if (P.order[P.outs[i]] != roots[i]) return 15;
    for (size_t k = 0; k + 1 < P.order.size(); k++) {                 // every other multi-consumer node is stored (while there is room)
      bool stored = false; for (int o : P.outs) if (o == (int)k) stored = true;
      if (P.order[k]->consumers > 1 && !stored && (int)P.outs.size() < AGB_FUSE_MAX_OUT) return 16;
    }
  }
  // leaf addressing: whenever as_2d accepts a (sliced / broadcast / squeezed) view, ptr + r * pitch + c * cstride must hit exactly the element
  // the strided view holds at the output's multi-index
  for (int cs = 0; cs < n_cases; cs++) {
    const int nd = 1 + (int)(rnd() % 4);
    Shape base(nd), out(nd); NdArray leaf;
    for (int k = 0; k < nd; k++) { base[k] = 1 + (int64_t)(rnd() % 5); out[k] = base[k]; }
    leaf.shape = base; leaf.stride = NdArray::contiguous_strides(base); leaf.dptr = (float*)(uintptr_t)0x100000;
    int64_t off0 = 0;
    for (int k = 0; k < nd; k++) {
      const uint32_t what = rnd() % 4;
      if (what == 0 && base[k] > 1) { int64_t s0 = (int64_t)(rnd() % base[k]), ln = 1 + (int64_t)(rnd() % (base[k] - s0)); off0 += s0 * leaf.stride[k]; leaf.shape[k] = ln; out[k] = ln; }   // slice
      else if (what == 1) { leaf.shape[k] = 1; out[k] = 1 + (int64_t)(rnd() % 4); }                                                                                        // broadcast axis
    }
    leaf.dptr += off0;
    int64_t pitch, cst;
    if (!as_2d(leaf, out, pitch, cst)) continue;
    const int64_t cols = out[nd - 1]; int64_t total = 1; for (auto d : out) total *= d;
    for (int64_t i = 0; i < total; i++) {
      int64_t rem = i, want = 0;
      for (int k = nd - 1; k >= 0; k--) { const int64_t idx = rem % out[k]; rem /= out[k]; if (leaf.shape[k] != 1) want += idx * leaf.stride[k]; }
      if (want != (i / cols) * pitch + (i % cols) * cst) return 20;
    }
  }
  if (n_compiled) *n_compiled = compiled;
  return 0;
}

}  // namespace agx
