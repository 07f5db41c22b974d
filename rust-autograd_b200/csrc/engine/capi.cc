// capi.cc — flat C ABI (include/agx200.h) over the C++ host engine; bound from Python with ctypes for the re-hosted
// reference test-suite, the benchmarks and smoke().  No logic lives here beyond argument marshalling and the
// name -> tensor_ops constructor table.
#include "agx.h"
#include "../../../include/agx200.h"
#include <fstream>
#include <memory>
#include <stdlib.h>
#include <sstream>
#include <string.h>

using namespace agx;
namespace agx { NdArray materialize_im2col(Device*, const NdArray&); }

// ---- automatic step-plan cache -------------------------------------------------------------------------------------
// The reference rebuilds its graph and re-walks it on every training step (examples/mlp_mnist.rs:74, cnn_mnist.rs:96).  Every call that
// builds a node is folded into a running 128-bit signature of the graph; an evaluation is keyed by (signature, targets, feed keys and
// shapes, math mode, fusion).  First sight of a key: eager.  Second sight: the feeds are staged into plan-owned device buffers, the
// evaluation runs eagerly on them (that IS this step) and is then captured into a CUDA graph without executing.  From the third sight on
// the plan is replayed: feeds are copied into the staging buffers (H2D or D2D), one graph launch, results are copied out of the graph's
// private memory.  Graphs with host callbacks (custom ops, hooks), evaluations that need the host in the middle (capture fails: the plan
// is marked bad and stays eager) and data-parallel runs are never cached.  agx_env_set_plan_cache(env, 0) / AGX_PLAN_CACHE=0 turn it off.
struct PlanKey { uint64_t a, b; bool operator<(const PlanKey& o) const { return a != o.a ? a < o.a : b < o.b; } };
struct Plan {
  int sights = 0, failures = 0; bool bad = false; void* exec = nullptr; uint64_t last_use = 0, graph_serial = 0;
  std::vector<NdArray> inputs; std::vector<Feed> feeds; std::vector<EvalResult> outs;
  std::vector<std::shared_ptr<StreamCell>> cells;      // the captured ops' stream positions: alive (and not re-assigned) while the plan lives
};
struct agx_env { VariableEnvironment* env; bool plan_cache = true; uint64_t tick = 0; std::map<PlanKey, Plan> plans; int64_t replays = 0, captures = 0; };
struct SigHash {
  uint64_t a = 1469598103934665603ull, b = 0x9E3779B97F4A7C15ull;
  void bytes(const void* p, size_t n) { const unsigned char* c = (const unsigned char*)p; for (size_t i = 0; i < n; i++) { a = (a ^ c[i]) * 1099511628211ull; b = (b + c[i] + 1) * 0xD6E8FEB86659FD93ull; b ^= b >> 32; } }
  template <class T> void pod(const T& v) { bytes(&v, sizeof(T)); }
  void str(const char* s) { size_t n = s ? strlen(s) : 0; pod(n); if (n) bytes(s, n); }
};
static uint64_t g_graph_serial = 0, g_opt_serial = 0;
struct agx_graph { Graph g; agx_env* owner = nullptr; SigHash sig; bool cacheable = true; uint64_t serial = ++g_graph_serial; };
struct agx_opt { Optimizer* o; uint64_t serial = ++g_opt_serial; };
static void plan_free(Plan& p) { if (p.exec) { agb_graph_destroy(p.exec); p.exec = nullptr; } p.outs.clear(); p.feeds.clear(); p.inputs.clear(); p.cells.clear(); }
static void plans_clear(agx_env* e) { for (auto& kv : e->plans) plan_free(kv.second); e->plans.clear(); }

struct agx_results {
  std::vector<EvalResult> rs;
  // deferred fetch (agx_eval_launch / agx_results_fetch): pinned staging blocks of the in-flight D2H copies
  Device* dev = nullptr; std::vector<void*> pinned, events; std::vector<size_t> bytes; std::vector<NdArray> keep;
};
// small pinned-block pool: cudaHostAlloc / cudaFreeHost synchronise the device, so blocks are recycled
static std::multimap<size_t, void*> g_pinned_pool;
static void* pinned_get(size_t bytes) {
  auto it = g_pinned_pool.lower_bound(bytes);
  if (it != g_pinned_pool.end() && it->first <= 4 * bytes + 4096) { void* p = it->second; g_pinned_pool.erase(it); return p; }
  void* p = nullptr; check_status(agb_host_alloc(bytes ? bytes : 4, &p)); return p;
}

static thread_local std::string g_err;
extern "C" const char* agx_last_error(void) { return g_err.c_str(); }

#define AGX_TRY try {
#define AGX_CATCH                                                                                   \
  return 0; }                                                                                       \
  catch (const OpError& e) { g_err = e.msg; return e.code ? e.code : AGB_ERR_NDARRAY; }             \
  catch (const Panic& e) { g_err = std::string("panic: ") + e.msg; return AGX_ERR_PANIC; }          \
  catch (const std::exception& e) { g_err = std::string("internal error: ") + e.what(); return AGX_ERR_PANIC; }

static std::vector<Tensor> tv(agx_graph* g, const int* ids, int n) {
  std::vector<Tensor> r;
  for (int i = 0; i < n; i++) {
    if (ids[i] < 0 || ids[i] >= (int)g->g.node_set.size()) throw Panic("tensor id out of range");
    r.push_back(g->g.tensor(ids[i]));
  }
  return r;
}
static std::vector<Feed> fv(agx_graph* g, const agx_feed* feeds, int n) {
  std::vector<Feed> r; Device* dev = g->g.env->dev;
  for (int i = 0; i < n; i++) {
    Feed f; f.by_name = feeds[i].name != nullptr; if (f.by_name) f.name = feeds[i].name; f.id = feeds[i].tensor_id;
    Shape shp(feeds[i].shape, feeds[i].shape + feeds[i].rank);
    if (feeds[i].on_device) {          // value already resident in HBM (owned by the caller): wrap, no copy
      f.value.shape = shp; f.value.stride = NdArray::contiguous_strides(shp); f.value.dptr = const_cast<float*>(feeds[i].data);
    } else {                           // evaluation.rs:296: the feed view enters the graph -> one H2D copy on the run's stream
      f.value = dev->empty(shp);
      if (f.value.size() > 0) check_status(agb_h2d(dev->ctx, f.value.dptr, feeds[i].data, (size_t)f.value.size() * sizeof(float)));
    }
    r.push_back(std::move(f));
  }
  return r;
}

// ================================================================================================ environment
extern "C" int agx_env_new(int device, agx_env** out) { AGX_TRY *out = nullptr; auto* e = new agx_env(); e->env = new VariableEnvironment(device); *out = e; AGX_CATCH }
extern "C" int agx_env_free(agx_env* env) { AGX_TRY if (env) { plans_clear(env); delete env->env; delete env; } AGX_CATCH }
extern "C" int agx_env_set_plan_cache(agx_env* env, int on) { AGX_TRY env->plan_cache = on != 0; if (!on) plans_clear(env); AGX_CATCH }
extern "C" int agx_env_plan_stats(agx_env* env, int64_t* captures, int64_t* replays, int* live_plans) {
  AGX_TRY if (captures) *captures = env->captures; if (replays) *replays = env->replays; if (live_plans) { int n = 0; for (auto& kv : env->plans) n += kv.second.exec != nullptr; *live_plans = n; } AGX_CATCH
}
extern "C" int agx_env_ctx(agx_env* env, agb_ctx** out) { AGX_TRY *out = env->env->dev->ctx; AGX_CATCH }
extern "C" int agx_env_set(agx_env* env, const char* ns, const char* name, const float* data, const int64_t* shape, int rank, int* vid) {
  AGX_TRY *vid = env->env->set(ns ? ns : "", name, Shape(shape, shape + rank), data).v; AGX_CATCH
}
extern "C" int agx_env_find(agx_env* env, const char* ns, const char* name, int* vid) { AGX_TRY *vid = env->env->find(ns ? ns : "", name).v; AGX_CATCH }
extern "C" int agx_env_var_count(agx_env* env, int* n) { AGX_TRY *n = (int)env->env->array_list.size(); AGX_CATCH }
extern "C" int agx_env_var_ids(agx_env* env, const char* ns, int* out, int cap, int* n) {
  AGX_TRY auto v = env->env->current_var_ids(ns ? ns : ""); *n = (int)v.size(); for (int i = 0; i < (int)v.size() && i < cap; i++) out[i] = v[i].v; AGX_CATCH
}
extern "C" int agx_env_var_shape(agx_env* env, int vid, int64_t* shape, int* rank) {
  AGX_TRY const NdArray& a = env->env->array_list.at(vid); *rank = a.ndim(); for (int i = 0; i < a.ndim(); i++) shape[i] = a.shape[i]; AGX_CATCH
}
extern "C" int agx_env_get(agx_env* env, int vid, float* out, int64_t cap) {
  AGX_TRY std::vector<float> h = env->env->get(VariableID{vid}); if ((int64_t)h.size() > cap) throw OpError(AGB_ERR_INVALID_DIMS, "agx_env_get: buffer too small"); memcpy(out, h.data(), h.size() * sizeof(float)); AGX_CATCH
}
extern "C" int agx_env_put(agx_env* env, int vid, const float* data, int64_t n) { AGX_TRY env->env->put(VariableID{vid}, data, (size_t)n); AGX_CATCH }
extern "C" int agx_env_var_ptr(agx_env* env, int vid, float** dptr) { AGX_TRY *dptr = env->env->array_list.at(vid).dptr; AGX_CATCH }
extern "C" int agx_env_save(agx_env* env, const char* path) {
  AGX_TRY std::ofstream f(path); if (!f) throw OpError(AGB_ERR_NDARRAY, std::string("save: cannot open ") + path); f << env->env->save_json(); f.flush(); if (!f.good()) throw OpError(AGB_ERR_NDARRAY, std::string("save: write failed: ") + path); AGX_CATCH
}
extern "C" int agx_env_load(agx_env* env, const char* path) {
  AGX_TRY std::ifstream f(path); if (!f) throw OpError(AGB_ERR_NDARRAY, std::string("load: cannot open ") + path); std::stringstream ss; ss << f.rdbuf(); plans_clear(env); env->env->load_json(ss.str()); AGX_CATCH
}
extern "C" int agx_fuse_selftest(int n_cases, unsigned seed, int* n_compiled) { AGX_TRY return fuse_selftest(n_cases, seed, n_compiled); AGX_CATCH }
extern "C" int agx_env_set_fusion(agx_env* env, int on) { AGX_TRY env->env->fuse_elementwise = on != 0; AGX_CATCH }
extern "C" int agx_env_set_data_parallel(agx_env* env, int rank, int world, const void* id) {
  AGX_TRY if (world > 1) check_status(agb_nccl_init(env->env->dev->ctx, rank, world, id)); env->env->rank = rank; env->env->world = world; AGX_CATCH
}

// ================================================================================================ graph
extern "C" int agx_graph_new(agx_env* env, agx_graph** out) { AGX_TRY auto* g = new agx_graph(); g->g.env = env->env; g->owner = env; g->g.node_set.reserve(512); *out = g; AGX_CATCH }
extern "C" int agx_graph_free(agx_graph* g) { AGX_TRY delete g; AGX_CATCH }
extern "C" int agx_graph_clear(agx_graph* g) { AGX_TRY g->g.clear(); g->sig = SigHash(); g->cacheable = true; g->serial = ++g_graph_serial; AGX_CATCH }
extern "C" int agx_graph_size(agx_graph* g, int* n) { AGX_TRY *n = (int)g->g.node_set.size(); AGX_CATCH }
extern "C" int agx_placeholder(agx_graph* g, const char* name, const int64_t* shape, int rank, int* tid) {
  AGX_TRY *tid = g->g.placeholder(name, std::vector<int64_t>(shape, shape + rank)).id;
  g->sig.str("placeholder"); g->sig.str(name); g->sig.pod(rank); g->sig.bytes(shape, sizeof(int64_t) * rank); AGX_CATCH
}
extern "C" int agx_variable(agx_graph* g, int vid, int* tid) { AGX_TRY *tid = g->g.variable_by_id(VariableID{vid}).id; g->sig.str("variable"); g->sig.pod(vid); g->sig.pod(*tid); AGX_CATCH }
extern "C" int agx_variable_by_name(agx_graph* g, const char* ns, const char* name, int* tid) {
  AGX_TRY *tid = g->g.variable_by_name(name, ns ? ns : "").id; g->sig.str("variable"); g->sig.pod(g->g.inner(*tid).variable_id.v); g->sig.pod(*tid); AGX_CATCH
}
extern "C" int agx_convert_to_tensor(agx_graph* g, const float* data, const int64_t* shape, int rank, int* tid) {
  AGX_TRY Shape s(shape, shape + rank); int64_t n = 1; for (auto d : s) n *= d;
  *tid = T::convert_to_tensor(&g->g, s, std::vector<float>(data, data + n)).id;
  g->sig.str("const"); g->sig.pod(rank); g->sig.bytes(shape, sizeof(int64_t) * rank); g->sig.bytes(data, sizeof(float) * (size_t)n); AGX_CATCH
}
extern "C" int agx_tensor_op_name(agx_graph* g, int tid, char* buf, int cap) { AGX_TRY snprintf(buf, cap, "%s", g->g.inner(tid).op->name()); AGX_CATCH }
extern "C" int agx_tensor_variable_id(agx_graph* g, int tid, int* vid) { AGX_TRY *vid = g->g.inner(tid).variable_id.v; AGX_CATCH }

extern "C" int agx_call(agx_graph* gg, const char* fn_, const int* tensors, int nt, const int64_t* I, int ni, const double* F, int nf, int* out, int cap, int* nout) {
  AGX_TRY
  Graph* g = &gg->g; std::string fn = fn_;
  std::vector<Tensor> t = tv(gg, tensors, nt); std::vector<Tensor> r;
  gg->sig.str(fn_); gg->sig.pod(nt); gg->sig.bytes(tensors, sizeof(int) * nt); gg->sig.pod(ni); gg->sig.bytes(I, sizeof(int64_t) * ni); gg->sig.pod(nf); gg->sig.bytes(F, sizeof(double) * nf);
  auto need = [&](int a, int b, int c) { if (nt < a || ni < b || nf < c) throw Panic("agx_call(" + fn + "): expected at least " + std::to_string(a) + " tensors, " + std::to_string(b) + " ints, " + std::to_string(c) + " floats"); };
  auto ints = [&](int from) { return std::vector<int64_t>(I + from, I + ni); };
  static const char* unary_names[] = {"sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh", "asinh", "acosh", "atanh", "exp", "exp2", "exp10", "ln", "log2",
                                      "log10", "sqrt", "neg", "abs", "sign", "floor", "ceil", "inv", "inv_sqrt", "square", "sigmoid", "relu", "softplus", "lgamma", "digamma"};
  static const char* cmp_names[] = {"equal", "not_equal", "greater", "lesser", "greater_equal", "lesser_equal", "maximum", "minimum"};
  bool done = false;
  for (auto u : unary_names) if (fn == u) { need(1, 0, 0); r = {T::unary(fn, t[0])}; done = true; }
  for (auto u : cmp_names) if (fn == u) { need(2, 0, 0); r = {T::cmp(fn, t[0], t[1])}; done = true; }
  if (done) {}
  else if (fn == "pow") { need(1, 0, 1); r = {T::unary("pow", t[0], (float)F[0])}; }
  else if (fn == "elu") { need(1, 0, 1); r = {T::unary("elu", t[0], (float)F[0])}; }
  else if (fn == "leaky_relu") { need(1, 0, 1); r = {T::leaky_relu(t[0], (float)F[0])}; }
  else if (fn == "clip") { need(1, 0, 2); r = {T::clip(t[0], (float)F[0], (float)F[1])}; }
  else if (fn == "add") { need(2, 0, 0); r = {T::add(t[0], t[1])}; }
  else if (fn == "sub") { need(2, 0, 0); r = {T::sub(t[0], t[1])}; }
  else if (fn == "mul") { need(2, 0, 0); r = {T::mul(t[0], t[1])}; }
  else if (fn == "div") { need(2, 0, 0); r = {T::div(t[0], t[1])}; }
  else if (fn == "add_n") { need(1, 0, 0); r = {T::add_n(t)}; }
  else if (fn == "scalar") { need(0, 0, 1); r = {T::scalar(g, (float)F[0])}; }
  else if (fn == "as_tensor") { r = {T::as_tensor(g, ints(0))}; }                    // integer literal arrays (axes / shapes)
  else if (fn == "zeros") { need(1, 0, 0); r = {T::zeros(g, t[0])}; }
  else if (fn == "ones") { need(1, 0, 0); r = {T::ones(g, t[0])}; }
  else if (fn == "shape") { need(1, 0, 0); r = {T::shape(t[0])}; }
  else if (fn == "rank") { need(1, 0, 0); r = {T::rank(t[0])}; }
  else if (fn == "size") { need(1, 0, 0); r = {T::size(t[0])}; }
  else if (fn == "identity") { need(1, 0, 0); r = {T::identity(t[0])}; }
  else if (fn == "nth_tensor") { need(1, 1, 0); r = {T::nth_tensor(t[0], (int)I[0])}; }
  else if (fn == "stop_gradient") { need(1, 0, 0); r = {T::stop_gradient(t[0])}; }
  else if (fn == "reduce_sum" || fn == "reduce_mean" || fn == "reduce_prod" || fn == "reduce_min" || fn == "reduce_max") { need(2, 1, 0); r = {T::reduce(fn.substr(7), t[0], t[1], I[0] != 0)}; }
  else if (fn == "reduce_variance") { need(2, 1, 0); r = {T::reduce_variance(t[0], t[1], I[0] != 0)}; }
  else if (fn == "sum_all") { need(1, 0, 0); r = {T::sum_all(t[0])}; }
  else if (fn == "mean_all") { need(1, 0, 0); r = {T::mean_all(t[0])}; }
  else if (fn == "argmax") { need(1, 2, 0); r = {T::argmax(t[0], (int)I[0], I[1] != 0)}; }
  else if (fn == "argmin") { need(1, 2, 0); r = {T::argmin(t[0], (int)I[0], I[1] != 0)}; }
  else if (fn == "reduce_logsumexp") { need(1, 2, 0); r = {T::reduce_logsumexp(t[0], (int)I[0], I[1] != 0)}; }
  else if (fn == "softmax") { need(1, 1, 0); r = {T::softmax(t[0], (int)I[0])}; }
  else if (fn == "log_softmax") { need(1, 1, 0); r = {T::log_softmax(t[0], (int)I[0])}; }
  else if (fn == "sigmoid_cross_entropy") { need(2, 0, 0); r = {T::sigmoid_cross_entropy(t[0], t[1])}; }
  else if (fn == "softmax_cross_entropy") { need(2, 0, 0); r = {T::softmax_cross_entropy(t[0], t[1])}; }
  else if (fn == "sparse_softmax_cross_entropy") { need(2, 0, 0); r = {T::sparse_softmax_cross_entropy(t[0], t[1])}; }
  else if (fn == "mean_squared_error") { need(2, 0, 0); r = {T::mean_squared_error(t[0], t[1])}; }
  else if (fn == "matmul") { need(2, 0, 0); r = {T::matmul(t[0], t[1])}; }
  else if (fn == "batch_matmul") { need(2, 0, 0); r = {T::batch_matmul_t(t[0], t[1], false, false)}; }
  else if (fn == "batch_matmul_t") { need(2, 2, 0); r = {T::batch_matmul_t(t[0], t[1], I[0] != 0, I[1] != 0)}; }
  else if (fn == "tensordot") { need(4, 0, 0); r = {T::tensordot(t[0], t[1], t[2], t[3])}; }
  else if (fn == "reshape") { need(2, 0, 0); r = {T::reshape(t[0], t[1])}; }
  else if (fn == "flatten") { need(1, 0, 0); r = {T::flatten(t[0])}; }
  else if (fn == "transpose") { need(2, 0, 0); r = {T::transpose(t[0], t[1])}; }
  else if (fn == "squeeze") { need(2, 0, 0); r = {T::squeeze(t[0], t[1])}; }
  else if (fn == "expand_dims") { need(2, 0, 0); r = {T::expand_dims(t[0], t[1])}; }
  else if (fn == "slice") { need(1, 2, 0); std::vector<int64_t> all = ints(0); size_t h = all.size() / 2; r = {T::slice(t[0], std::vector<int64_t>(all.begin(), all.begin() + h), std::vector<int64_t>(all.begin() + h, all.end()))}; }
  else if (fn == "split") { need(1, 2, 0); r = T::split(t[0], ints(1), (int)I[0]); }                         // ints: axis, sizes...
  else if (fn == "concat") { need(1, 1, 0); r = {T::concat(t, (int)I[0])}; }
  else if (fn == "tile") { need(1, 2, 0); r = {T::tile(t[0], (int)I[0], (int)I[1])}; }
  else if (fn == "gather_common") { need(2, 1, 0); r = {T::gather_common(t[0], t[1], (int)I[0])}; }
  else if (fn == "gather") { need(2, 1, 0); r = {T::gather(t[0], t[1], (int)I[0])}; }
  else if (fn == "access_elem") { need(1, 1, 0); r = {T::access_elem(t[0], I[0])}; }
  else if (fn == "setdiff1d") { need(2, 0, 0); r = {T::setdiff1d(t[0], t[1])}; }
  else if (fn == "conv2d") { need(2, 2, 0); r = {T::conv2d(t[0], t[1], (int)I[0], (int)I[1], 1)}; }
  else if (fn == "dilated_conv2d") { need(2, 3, 0); r = {T::conv2d(t[0], t[1], (int)I[0], (int)I[1], (int)I[2])}; }
  else if (fn == "conv2d_transpose") { need(2, 2, 0); r = {T::conv2d_transpose(t[0], t[1], (int)I[0], (int)I[1], 1)}; }
  else if (fn == "dilated_conv2d_transpose") { need(2, 3, 0); r = {T::conv2d_transpose(t[0], t[1], (int)I[0], (int)I[1], (int)I[2])}; }
  else if (fn == "max_pool2d") { need(1, 3, 0); r = {T::max_pool2d(t[0], (int)I[0], (int)I[1], (int)I[2])}; }
  else if (fn == "dropout") { need(1, 1, 1); r = {T::dropout(t[0], (float)F[0], I[0] != 0, ni > 1 ? (uint64_t)I[1] : 0)}; }
  else if (fn == "random") { need(1, 2, 2); r = {T::random(g, (int)I[0], t[0], (float)F[0], (float)F[1], (uint64_t)I[1])}; }      // ints: agb_rand_kind, seed (0 = the default rng)
  else if (fn == "normalize") { need(2, 0, 0); r = {T::normalize(t[0], t[1])}; }
  else if (fn == "batch_norm") { need(3, 0, 0); r = {T::batch_norm(t[0], t[1], t[2])}; }
  else if (fn == "assign") { need(2, 0, 0); r = {T::assign(t[0], t[1])}; }
  else if (fn == "control_dependencies") { need(1, 0, 0); r = {T::control_dependencies(t[0], std::vector<Tensor>(t.begin() + 1, t.end()))}; }
  else throw Panic("agx_call: unknown tensor_ops function `" + fn + "`");
  *nout = (int)r.size();
  for (int i = 0; i < (int)r.size() && i < cap; i++) out[i] = r[i].id;
  AGX_CATCH
}

// ================================================================================================ host-callback ops
struct agx_out { std::vector<NdArray> ys; std::string err; };
extern "C" int agx_out_append(agx_out* out, const float* data, const int64_t* shape, int rank) {
  AGX_TRY Shape s(shape, shape + rank); int64_t n = 1; for (auto d : s) n *= d;
  out->ys.push_back(NdArray::from_host(s, std::vector<float>(data, data + n))); AGX_CATCH
}
extern "C" int agx_out_error(agx_out* out, const char* message) { AGX_TRY out->err = message ? message : ""; AGX_CATCH }
namespace {
struct HostView { std::vector<float> data; Shape shape; agx_host_array view() const { return agx_host_array{data.data(), shape.data(), (int)shape.size()}; } };
HostView host_view(Device* dev, NdArray x) {          // D2H (+ stream sync) of a possibly strided device view, in logical (C) order
  HostView h; h.shape = x.shape;
  if (x.on_device() && !x.has_host() && !x.is_contiguous()) x = dev->contiguous(x);
  h.data = dev->ensure_host(x);
  return h;
}
struct CustomOp : Op {                 // a user-defined Op (src/op.rs:90-101) whose compute / grad live behind C callbacks
  std::string nm; agx_compute_fn fc; agx_grad_fn fg; void* user; agx_graph* owner;
  const char* name() const override { return nm.c_str(); }
  void compute(ComputeContext& c) override {
    std::vector<HostView> hs; std::vector<agx_host_array> views;
    for (int i = 0; i < c.num_inputs(); i++) hs.push_back(host_view(c.dev, c.input(i)));
    for (auto& h : hs) views.push_back(h.view());
    agx_out out;
    int code = fc(user, views.data(), (int)views.size(), &out);
    if (code != 0) throw OpError(code >= 1 && code <= 5 ? code : AGB_ERR_NDARRAY, nm + ": " + (out.err.empty() ? "compute failed" : out.err));
    if (out.ys.empty()) throw Panic("Bad op implementation: empty return value");
    for (auto& y : out.ys) c.append_output(y);
  }
  void grad(GradientContext& c) override {
    const int n = c.num_inputs();
    if (!fg) { for (int i = 0; i < n; i++) c.append_none(); return; }
    std::vector<int> ins, gxs(n, -1);
    for (int i = 0; i < n; i++) ins.push_back(c.input(i).id);
    fg(user, owner, ins.data(), n, c.output().id, c.output_grad().id, gxs.data());
    for (int i = 0; i < n; i++) { if (gxs[i] < 0) c.append_none(); else c.append_input_grad(c.graph()->tensor(gxs[i])); }
  }
};
}  // namespace
extern "C" int agx_custom_op(agx_graph* g, const char* name, const int* inputs, int n_inputs, agx_compute_fn compute, agx_grad_fn grad, void* user, int* tid) {
  AGX_TRY
  if (!compute) throw Panic("agx_custom_op: compute callback is NULL");
  g->cacheable = false;                   // host callback inside the evaluation: never replayed from a captured plan
  auto* op = new CustomOp(); op->nm = name ? name : "CustomOp"; op->fc = compute; op->fg = grad; op->user = user; op->owner = g;
  TensorBuilder b(&g->g);
  for (auto& t : tv(g, inputs, n_inputs)) b.append_input(t, false);
  *tid = b.build(op).id;
  AGX_CATCH
}
extern "C" int agx_hook(agx_graph* g, int tensor, int kind, const char* text, agx_hook_fn fn, void* user, int* tid) {
  AGX_TRY
  if (kind < 0 || kind > 3 || (kind == 0 && !fn)) throw Panic("agx_hook: bad hook kind / missing callback");
  std::string prefix = text ? text : "";
  g->cacheable = false;                   // hooks run on the host in the middle of the evaluation
  Tensor x = tv(g, &tensor, 1)[0];
  *tid = T::hook(x, [kind, prefix, fn, user](const NdArray& a, const std::vector<float>& h) {
    if (kind == 0) { agx_host_array v{h.data(), a.shape.data(), (int)a.shape.size()}; fn(user, &v); return; }
    if (kind == 3) fprintf(stderr, "%s\n", prefix.c_str());
    fprintf(stderr, "[");
    for (size_t i = 0; i < a.shape.size(); i++) fprintf(stderr, "%s%lld", i ? ", " : "", (long long)a.shape[i]);
    fprintf(stderr, "]");
    if (kind != 2) { fprintf(stderr, " ["); for (size_t i = 0; i < h.size() && i < 64; i++) fprintf(stderr, "%s%g", i ? ", " : "", h[i]); fprintf(stderr, h.size() > 64 ? ", ...]" : "]"); }
    fprintf(stderr, "\n");
  }).id;
  AGX_CATCH
}

extern "C" int agx_grad(agx_graph* g, const int* ys, int ny, const int* xs, int nx, const int* gys, int* out) {
  AGX_TRY
  std::vector<Tensor> r = gys ? T::grad_with_default(tv(g, ys, ny), tv(g, xs, nx), tv(g, gys, ny)) : T::grad(tv(g, ys, ny), tv(g, xs, nx));
  for (int i = 0; i < nx; i++) out[i] = r[i].id;
  g->sig.str("grad"); g->sig.pod(ny); g->sig.bytes(ys, sizeof(int) * ny); g->sig.pod(nx); g->sig.bytes(xs, sizeof(int) * nx); if (gys) g->sig.bytes(gys, sizeof(int) * ny);
  AGX_CATCH
}
extern "C" int agx_grad_helper(agx_graph* g, const int* losses, int n, const char* ns, int* vars, int* grads, int cap, int* nout) {
  AGX_TRY
  std::vector<Tensor> vs, gs; grad_helper(tv(g, losses, n), ns ? ns : "", &g->g, vs, gs);
  g->sig.str("grad_helper"); g->sig.pod(n); g->sig.bytes(losses, sizeof(int) * n); g->sig.str(ns);
  *nout = (int)vs.size();
  for (int i = 0; i < (int)vs.size() && i < cap; i++) { vars[i] = vs[i].id; grads[i] = gs[i].id; }
  AGX_CATCH
}

// ================================================================================================ evaluation
static bool plan_cache_enabled(agx_env* e) {
  static const int env_on = [] { const char* v = getenv("AGX_PLAN_CACHE"); return (v && v[0] == '0') ? 0 : 1; }();
  return e && e->plan_cache && env_on && e->env->world == 1;
}
// device-resident results of one evaluation, through the plan cache when the graph allows it (see the comment at struct Plan)
static std::vector<EvalResult> eval_cached(agx_graph* g, const int* targets, int n, const agx_feed* feeds, int nfeeds) {
  agx_env* owner = g->owner; VariableEnvironment* env = g->g.env; Device* dev = env->dev;
  if (!plan_cache_enabled(owner) || !g->cacheable || n <= 0) return eval(&g->g, tv(g, targets, n), fv(g, feeds, nfeeds), false);
  SigHash k = g->sig;
  k.str("eval"); k.pod(g->g.node_set.size()); k.pod(n); k.bytes(targets, sizeof(int) * n); k.pod(nfeeds);
  for (int i = 0; i < nfeeds; i++) { k.str(feeds[i].name); k.pod(feeds[i].tensor_id); k.pod(feeds[i].rank); k.bytes(feeds[i].shape, sizeof(int64_t) * feeds[i].rank); }
  int mode = 0; agb_get_math_mode(dev->ctx, &mode); k.pod(mode); k.pod(env->fuse_elementwise);
  Plan& p = owner->plans[PlanKey{k.a, k.b}];
  p.last_use = ++owner->tick; p.sights++;
  if (p.bad || p.sights == 1) return eval(&g->g, tv(g, targets, n), fv(g, feeds, nfeeds), false);
  auto stage = [&]() {
    for (int i = 0; i < nfeeds; i++) {
      const size_t nb = (size_t)p.inputs[i].size() * sizeof(float);
      if (!nb) continue;
      if (feeds[i].on_device) check_status(agb_d2d(dev->ctx, p.inputs[i].dptr, feeds[i].data, nb));
      else check_status(agb_h2d(dev->ctx, p.inputs[i].dptr, feeds[i].data, nb));
    }
  };
  if (p.exec == nullptr) {
    // ---- second sight: stage the feeds and CAPTURE the evaluation on the staged copies (nothing executes; every allocation must hit the
    // arena's free list, which the first sight's eager run left holding exactly these block sizes); the replay below then IS this step.
    // A capture that fails (arena miss, an op that needs the host mid-evaluation) executes nothing either: the step runs eagerly and the
    // plan gets one more attempt on a warmer arena before it is marked bad.
    if (p.inputs.empty()) {
      for (int i = 0; i < nfeeds; i++) {
        Feed f; f.by_name = feeds[i].name != nullptr; if (f.by_name) f.name = feeds[i].name; f.id = feeds[i].tensor_id;
        f.value = dev->empty(Shape(feeds[i].shape, feeds[i].shape + feeds[i].rank));
        p.inputs.push_back(f.value); p.feeds.push_back(std::move(f));
      }
    }
    bool captured = false;
    try {
      check_status(agb_graph_begin(dev->ctx));
      try {
        p.outs = eval(&g->g, tv(g, targets, n), p.feeds, false);
        for (auto& r : p.outs) if (!r.ok) throw OpError(r.err_code, r.err_msg);
      } catch (...) { agb_graph_end(dev->ctx, nullptr); throw; }
      check_status(agb_graph_end(dev->ctx, &p.exec));
      for (auto& nd : g->g.node_set) if (nd && nd->op) { auto c = nd->op->stream_cell(); if (c) p.cells.push_back(c); }
      p.graph_serial = g->serial; owner->captures++; captured = true;
    } catch (const std::exception&) { p.outs.clear(); p.cells.clear(); if (p.exec) { agb_graph_destroy(p.exec); p.exec = nullptr; } }
    while (owner->plans.size() > 8) {            // least-recently-used plans give their private memory back
      auto victim = owner->plans.end();
      for (auto it = owner->plans.begin(); it != owner->plans.end(); ++it) if (&it->second != &p && (victim == owner->plans.end() || it->second.last_use < victim->second.last_use)) victim = it;
      if (victim == owner->plans.end()) break;
      plan_free(victim->second); owner->plans.erase(victim);
    }
    if (!captured) {
      if (++p.failures >= 2) { p.bad = true; plan_free(p); }
      return eval(&g->g, tv(g, targets, n), fv(g, feeds, nfeeds), false);
    }
  }
  // ---- replay
  if (p.graph_serial != g->serial) {             // a graph REBUILT since the last replay: its random ops start their streams over, like freshly
    for (auto& c : p.cells) check_status(agb_memset0(dev->ctx, c->ptr, (size_t)agb_stream_cell_bytes()));      // constructed ops do (mod.rs:2895-2905)
    p.graph_serial = g->serial;
  }
  stage();
  check_status(agb_graph_launch(dev->ctx, p.exec));
  owner->replays++;
  std::vector<EvalResult> rs;
  for (auto& o : p.outs) {                       // results leave the graph's private memory before the next replay can overwrite them
    EvalResult r = o;
    if (r.ok && r.value.on_device() && !r.value.has_host() && !r.value.lazy && !r.value.expr && !r.value.virt) r.value = dev->copy(r.value);
    rs.push_back(std::move(r));
  }
  return rs;
}
static void fetch_results(Device* dev, std::vector<EvalResult>& rs) {       // the tail of eval(.., fetch_to_host = true)
  for (auto& r : rs) {
    if (!r.ok) continue;
    if (r.value.lazy) r.value = materialize_lazy(dev, r.value);
    if (r.value.virt && !r.value.on_device() && !r.value.has_host()) r.value = materialize_im2col(dev, r.value);
    dev->ensure_host(r.value);
  }
  dev->sync();      // surfaces device-side index errors (bad labels / gather ids) as OutOfBounds
}
extern "C" int agx_eval(agx_graph* g, const int* targets, int n, const agx_feed* feeds, int nfeeds, agx_results** out) {
  AGX_TRY *out = nullptr; auto* r = new agx_results(); std::unique_ptr<agx_results> guard(r);
  r->rs = eval_cached(g, targets, n, feeds, nfeeds); fetch_results(g->g.env->dev, r->rs); *out = guard.release(); AGX_CATCH
}
// launch-only evaluation: every kernel of the run is enqueued, each result's D2H is queued on the copy stream into pinned memory,
// and the call returns without a host sync; agx_results_fetch completes it later (after more work has been enqueued)
extern "C" int agx_eval_launch(agx_graph* g, const int* targets, int n, const agx_feed* feeds, int nfeeds, agx_results** out) {
  AGX_TRY
  *out = nullptr; auto* r = new agx_results(); r->dev = g->g.env->dev;
  r->rs = eval_cached(g, targets, n, feeds, nfeeds);
  for (auto& e : r->rs) {
    void* pin = nullptr; void* ev = nullptr; size_t nb = 0;
    if (e.ok && !e.value.host && e.value.on_device()) {
      NdArray c = e.value;
      if (c.i32) c = r->dev->i32_to_f32(c);
      c = r->dev->contiguous(c);
      nb = (size_t)c.size() * sizeof(float); pin = pinned_get(nb);
      check_status(agb_stage_d2h(r->dev->ctx, pin, c.dptr, nb, &ev));
      r->keep.push_back(c);
    }
    r->pinned.push_back(pin); r->events.push_back(ev); r->bytes.push_back(nb);
  }
  *out = r;
  AGX_CATCH
}
extern "C" int agx_results_fetch(agx_results* r) {
  AGX_TRY
  for (size_t i = 0; i < r->rs.size(); i++) {
    if (i < r->pinned.size() && r->pinned[i]) {
      check_status(agb_event_sync(r->events[i])); agb_event_destroy(r->events[i]); r->events[i] = nullptr;
      auto h = std::make_shared<std::vector<float>>(r->bytes[i] / sizeof(float));
      memcpy(h->data(), r->pinned[i], r->bytes[i]);
      r->rs[i].value.host = h;
      g_pinned_pool.insert({r->bytes[i] ? r->bytes[i] : 4, r->pinned[i]}); r->pinned[i] = nullptr;
    }
  }
  r->keep.clear();
  AGX_CATCH
}
// ---- step graphs: the kernels of one evaluation captured into a CUDA graph and replayed without walking the
// graph on the host.  Every feed must be device-resident (the replay reads the same addresses: refill them in place), the step must
// not need a host decision that depends on device data, and nothing is read back (fetch results from variables / fed buffers).
struct agx_step { Device* dev; void* exec; };
extern "C" int agx_step_capture(agx_graph* g, const int* targets, int n, const agx_feed* feeds, int nfeeds, agx_step** out) {
  AGX_TRY
  *out = nullptr;
  for (int i = 0; i < nfeeds; i++) if (!feeds[i].on_device) throw OpError(AGB_ERR_UNSUPPORTED, "agx_step_capture: every feed must be device-resident");
  Device* dev = g->g.env->dev;
  // two eager runs: the first grows the arena / scratch / descriptor tables, the second starts from the free-list state the captured
  // run will start from, so cached tables keyed by buffer addresses hit during capture
  for (int w = 0; w < 2; w++) { auto rs = eval(&g->g, tv(g, targets, n), fv(g, feeds, nfeeds), false); for (auto& r : rs) if (!r.ok) throw OpError(r.err_code, r.err_msg); }
  dev->sync();
  check_status(agb_graph_begin(dev->ctx));
  void* exec = nullptr;
  try {
    auto rs = eval(&g->g, tv(g, targets, n), fv(g, feeds, nfeeds), false);
    for (auto& r : rs) if (!r.ok) throw OpError(r.err_code, r.err_msg);
  } catch (...) { agb_graph_end(dev->ctx, nullptr); throw; }
  check_status(agb_graph_end(dev->ctx, &exec));
  agb_arena_pin(dev->ctx, +1);
  *out = new agx_step{dev, exec};
  AGX_CATCH
}
extern "C" int agx_step_launch(agx_step* s) { AGX_TRY check_status(agb_graph_launch(s->dev->ctx, s->exec)); AGX_CATCH }
extern "C" int agx_step_free(agx_step* s) { if (s) { agb_graph_destroy(s->exec); agb_arena_pin(s->dev->ctx, -1); delete s; } return 0; }

extern "C" int agx_run(agx_graph* g, const int* targets, int n, const agx_feed* feeds, int nfeeds) {
  AGX_TRY
  std::vector<EvalResult> rs = eval_cached(g, targets, n, feeds, nfeeds);
  for (auto& r : rs) if (!r.ok) throw OpError(r.err_code, r.err_msg);
  AGX_CATCH
}
extern "C" int agx_results_count(agx_results* r, int* n) { *n = (int)r->rs.size(); return 0; }
extern "C" int agx_results_status(agx_results* r, int i, int* code, const char** msg) { *code = r->rs[i].ok ? 0 : r->rs[i].err_code; *msg = r->rs[i].err_msg.c_str(); return 0; }
extern "C" int agx_results_shape(agx_results* r, int i, int64_t* shape, int* rank) {
  const NdArray& a = r->rs[i].value; *rank = a.ndim(); for (int k = 0; k < a.ndim(); k++) shape[k] = a.shape[k]; return 0;
}
extern "C" int agx_results_data(agx_results* r, int i, const float** data, int64_t* n) {
  const NdArray& a = r->rs[i].value;
  if (!a.host) { g_err = "result has no host copy"; return AGX_ERR_PANIC; }
  *data = a.host->data(); *n = (int64_t)a.host->size(); return 0;
}
extern "C" int agx_results_free(agx_results* r) {
  if (r) { for (size_t i = 0; i < r->pinned.size(); i++) if (r->pinned[i]) { if (r->events[i]) { agb_event_sync(r->events[i]); agb_event_destroy(r->events[i]); } g_pinned_pool.insert({r->bytes[i] ? r->bytes[i] : 4, r->pinned[i]}); } }
  delete r; return 0;
}

// ================================================================================================ optimizers
static std::vector<VariableID> vids(const int* v, int n) { std::vector<VariableID> r; for (int i = 0; i < n; i++) r.push_back(VariableID{v[i]}); return r; }
extern "C" int agx_opt_adam(agx_env* env, const int* v, int n, const char* ns, float alpha, float eps, float b1, float b2, agx_opt** out) {
  AGX_TRY auto* o = new agx_opt(); o->o = make_adam(env->env, vids(v, n), ns, alpha, eps, b1, b2); *out = o; AGX_CATCH
}
extern "C" int agx_opt_sgd(float lr, agx_opt** out) { AGX_TRY auto* o = new agx_opt(); o->o = make_sgd(lr); *out = o; AGX_CATCH }
extern "C" int agx_opt_momentum_sgd(agx_env* env, const int* v, int n, const char* ns, float lr, float momentum, agx_opt** out) {
  AGX_TRY auto* o = new agx_opt(); o->o = make_momentum_sgd(env->env, vids(v, n), ns, lr, momentum); *out = o; AGX_CATCH
}
extern "C" int agx_opt_adagrad(agx_env* env, const int* v, int n, const char* ns, float lr, agx_opt** out) {
  AGX_TRY auto* o = new agx_opt(); o->o = make_adagrad(env->env, vids(v, n), ns, lr); *out = o; AGX_CATCH
}
extern "C" int agx_opt_compute_updates(agx_opt* o, agx_graph* g, const int* params, const int* grads, int n, int* out) {
  AGX_TRY std::vector<Tensor> r = o->o->compute_updates(tv(g, params, n), tv(g, grads, n), &g->g); for (int i = 0; i < n; i++) out[i] = r[i].id;
  g->sig.str("opt_updates"); g->sig.pod(o->serial); g->sig.pod(n); g->sig.bytes(params, sizeof(int) * n); g->sig.bytes(grads, sizeof(int) * n); AGX_CATCH
}
extern "C" int agx_opt_get_update_op(agx_opt* o, agx_graph* g, const int* params, const int* grads, int n, int* tid) {
  AGX_TRY *tid = o->o->get_update_op(tv(g, params, n), tv(g, grads, n), &g->g).id;
  g->sig.str("opt_update_op"); g->sig.pod(o->serial); g->sig.pod(n); g->sig.bytes(params, sizeof(int) * n); g->sig.bytes(grads, sizeof(int) * n); AGX_CATCH
}
extern "C" int agx_opt_update(agx_opt* o, agx_graph* g, const int* params, const int* grads, int n, const agx_feed* feeds, int nfeeds) {
  AGX_TRY o->o->update(tv(g, params, n), tv(g, grads, n), &g->g, fv(g, feeds, nfeeds));
  g->sig.str("opt_update"); g->sig.pod(o->serial); g->sig.pod(n); g->sig.bytes(params, sizeof(int) * n); g->sig.bytes(grads, sizeof(int) * n); AGX_CATCH
}
extern "C" int agx_opt_free(agx_opt* o) { AGX_TRY if (o) { delete o->o; delete o; } AGX_CATCH }
