// agx.h — C++ host engine of the B200 backend: the device-resident counterpart of rust-autograd's graph / evaluator layer.
//
// It mirrors the reference's host-side interfaces by name and meaning (the Rust crate cannot be compiled in this image, so
// the host side above the kernel C ABI is written in C++):
//   Op / ComputeContext / GradientContext / OpError      reference src/op.rs:67-73,90-101,186-309,342-434
//   Graph / Tensor / TensorBuilder / IncomingTensor       reference src/graph.rs:17-100, src/tensor.rs:408-803
//   compute_gradients (reverse-mode graph builder)        reference src/gradient.rs:20-248
//   Graph::eval / Evaluator / Feeder                      reference src/evaluation.rs:58-371
//   VariableEnvironment / namespaces                      reference src/variable.rs:152-358,670-781
//   Optimizer / Adam / SGD / MomentumSGD / AdaGrad        reference src/optimizers/*.rs
// What changes relative to the reference (BASELINE.json north_star): NdArray storage lives in HBM (agb_alloc arena),
// every Op::compute launches sm_100a kernels through include/agb200.h on the context's CUDA stream, values reach the host
// only at eval/run boundaries, hooks and MapOp.  Shape/axes vectors ("meta" arrays) stay on the host.
#pragma once
#include <stdint.h>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>
#include "../../../include/agb200.h"

namespace agx {

// ---------------------------------------------------------------------------------------------- errors
// OpError (src/op.rs:67-73).  code = the C ABI status (1..5 = the five variants, >= 100 device failures).
struct OpError : public std::exception {
  int code; std::string msg;
  OpError(int c, std::string m) : code(c), msg(std::move(m)) {}
  const char* what() const noexcept override { return msg.c_str(); }
};
// API misuse that panics in the reference ("Bad op impl", unfilled placeholder, rank mismatch in compare ops, ...)
struct Panic : public std::exception {
  std::string msg;
  explicit Panic(std::string m) : msg(std::move(m)) {}
  const char* what() const noexcept override { return msg.c_str(); }
};
void check_status(int status);   // throws OpError with agb_last_error()

// ---------------------------------------------------------------------------------------------- arrays
struct Buffer {                   // one arena block in HBM
  agb_ctx* ctx; float* ptr; size_t bytes;
  Buffer(agb_ctx* c, size_t b);
  ~Buffer();
};
typedef std::shared_ptr<Buffer> BufferP;
typedef std::vector<int64_t> Shape;

struct NdArray;
struct Im2colRef;                 // virtual `cols` tensor (Conv2D output #1), see ops_nn.cc
struct PoolRef;                   // (max-pool index buffers) the pooled forward output + identity of the pooled input, see ops_nn.cc
struct ExprNode;                  // deferred elementwise expression (fuse.cc): value = one instruction of a fused program
struct Lazy;                      // deferred epilogue (conv [+bias] awaiting a ReLU, "x > 0" mask awaiting a multiply), see ops_nn.cc

// f32 array: a strided view on an HBM block and/or a small contiguous host vector.
//   - data arrays live on the device (dptr != null);
//   - constants / meta arrays (shapes, axes: "everything is a float tensor", src/ndarray_ext.rs:29-31) carry `host`
//     and get a device copy lazily, only if a device kernel consumes them.
struct NdArray {
  Shape shape, stride;            // stride in elements (device view); host copy is always C-contiguous
  BufferP buf; float* dptr = nullptr;
  std::shared_ptr<std::vector<float>> host;
  bool meta = false;              // shape-derived value: arithmetic on it stays on the host
  bool i32 = false;               // the buffer holds int32 indices (max-pool argmax): exact beyond 2^24, converted to f32 only when a
                                  // float consumer or the user asks (the reference stores indices as floats, max_pool2d.rs:74-75)
  std::shared_ptr<Im2colRef> virt;
  std::shared_ptr<PoolRef> pool;
  std::shared_ptr<NdArray> chan_sum;   // (4-D activations gradients) per-channel sums over (b, h, w), produced for free by the fused dgrad /
                                       // pool-backward epilogues; MaybeReduceSum (the bias gradient) takes it instead of re-reading the tensor
  std::shared_ptr<NdArray> relu_bits;  // (ReLU activations written by a fused conv kernel) the sign bits of exactly this buffer, numel / 32 words: the fused dgrad of
                                       // the next layer reads them instead of the activation (agb_conv2d_*_bits_f32); `relu_bits_of` = the dptr they describe
  const float* relu_bits_of = nullptr;
  std::shared_ptr<Lazy> lazy;     // value not computed yet: only `shape` is valid.  ComputeContext::input() materialises it unless the
                                  // consuming op declared accept_lazy (the ops that can fuse it into their own kernel)

  std::shared_ptr<ExprNode> expr; // pending elementwise expression (fuse.cc): only `shape` is valid until a consumer that needs memory compiles
                                  // the DAG into one fused launch.  ComputeContext::input() materialises it unless the op declared accept_expr

  int ndim() const { return (int)shape.size(); }
  int64_t size() const { int64_t n = 1; for (auto d : shape) n *= d; return n; }
  bool on_device() const { return dptr != nullptr; }
  bool has_host() const { return host != nullptr; }
  bool is_contiguous() const;
  agb_tensor desc() const;        // requires on_device()
  static Shape contiguous_strides(const Shape& s);
  static NdArray from_host(const Shape& shape, std::vector<float> v, bool meta = false);
  static NdArray scalar_host(float v, bool meta = false) { return from_host({}, {v}, meta); }
  // Memory-order permutation of a dense array: order[0] is the slowest axis.  Returns false when the view is not dense
  // (slices, broadcasts).  A C-contiguous array yields the identity; a channels-last [B,C,H,W] yields {0,2,3,1}.
  bool dense_order(std::vector<int>& order) const;
  NdArray reshaped(const Shape& s) const;      // contiguous only (device) / always (host)
  NdArray permuted(const std::vector<int>& perm) const;
  NdArray sliced(int axis, int64_t start, int64_t len) const;
};

// Device-resident stream positions of the random ops (agb_*_stream): one 8-byte cell per op instance, carved from pooled arena blocks.
// The pool outlives neither the context nor its users' handles: `alive` goes false when the Device dies, handles then do nothing.
struct StreamCellPool { agb_ctx* ctx = nullptr; bool alive = true; std::vector<uint32_t*> free_cells; std::vector<void*> blocks; };
struct StreamCell {
  std::shared_ptr<StreamCellPool> pool; uint32_t* ptr = nullptr;
  ~StreamCell() { if (pool && pool->alive && ptr) { agb_memset0(pool->ctx, ptr, 8); pool->free_cells.push_back(ptr); } }   // cells are zero while they wait in the pool
};

struct Device {                   // thin C++ handle on the kernel C ABI context
  agb_ctx* ctx = nullptr;
  std::shared_ptr<StreamCellPool> stream_cells;
  std::shared_ptr<StreamCell> new_stream_cell();     // cells in the pool are zero (zeroed when their block is created and when they are returned): acquiring one launches nothing, so it is safe inside a graph capture
  explicit Device(int index);
  ~Device();
  NdArray empty(const Shape& s);
  NdArray empty_ordered(const Shape& s, const std::vector<int>& order);   // dense, memory order = `order` (e.g. channels-last)
  NdArray zeros(const Shape& s);
  NdArray full(const Shape& s, float v);
  void ensure_device(NdArray& a);               // upload the host copy if there is no device copy yet
  const std::vector<float>& ensure_host(NdArray& a);   // D2H (+ stream sync) if there is no host copy yet
  NdArray contiguous(const NdArray& a);         // materialise a strided view (reference: ndarray_ext::deep_copy)
  // deep copies of SMALL strided views made during the current evaluation (label / token-id columns sliced out of one feed are consumed by a
  // forward op and again by its gradient op): the second request reuses the first copy.  Cleared by eval() and by Assign.
  std::vector<std::pair<NdArray, NdArray>> small_copies;
  NdArray i32_to_f32(const NdArray& a);         // float copy of an int32 index buffer
  NdArray copy(const NdArray& a);
  void sync();
};

// ---------------------------------------------------------------------------------------------- graph
struct Graph; struct Context; struct VariableEnvironment; struct ComputeContext; struct GradientContext; struct Evaluation;
typedef int TensorID;
struct VariableID { int v = -1; bool valid() const { return v >= 0; } };

struct Tensor {                   // Copy handle {id, graph} (src/tensor.rs:22-30)
  TensorID id = -1; Graph* graph = nullptr;
  bool valid() const { return graph != nullptr; }
};

struct Op {                       // trait Op (src/op.rs:90-101)
  virtual ~Op() {}
  virtual const char* name() const = 0;
  virtual void compute(ComputeContext& ctx) = 0;     // throws OpError for Err(..), Panic for panics
  virtual void grad(GradientContext& ctx) = 0;
  virtual bool metadata_only() const { return false; }   // Shape / Rank / Size: read the input's shape, never its values
  virtual bool plain_matmul(bool* tb) const { return false; }   // MatMul (2-D, lhs not transposed): rows of several such products with one rhs can be stacked
  // Row-wise ops (every output row depends only on the same row of the inputs): nodes with the same non-null stack_key whose inputs do not
  // depend on each other may be computed together on stacked rows.  ins[m] / outs[m] = the inputs / outputs of member m; false = not
  // applicable (the members then run on their own and raise their own errors).
  virtual const char* stack_key() const { return nullptr; }
  virtual bool compute_stacked(Device* dev, Evaluation& run, const std::vector<std::vector<NdArray>>& ins, std::vector<std::vector<NdArray>>* outs) { return false; }
  virtual bool sums_inputs() const { return false; }     // AddN: lets a producer defer itself so that the sum can absorb it (fuse.cc)
  virtual bool mutates_now() const { return false; }     // Assign: writes a variable in the middle of the traversal (optimizer ops are deferred)
  virtual std::shared_ptr<StreamCell> stream_cell() const { return nullptr; }   // random ops: the device-resident stream position (kept alive by cached step plans)
};

struct IncomingTensor { TensorID id; bool allow_mut; int array_selector; };   // src/tensor.rs:542-550

struct TensorInternal {           // src/tensor.rs:408-441
  TensorID id = 0;
  std::unique_ptr<Op> op;
  std::vector<IncomingTensor> incoming_nodes;
  int topo_rank = 0;
  TensorID shape = -1;            // id of the tensor holding this tensor's shape, or -1
  std::string placeholder_name; bool is_placeholder = false;
  bool is_differentiable = true;
  bool has_backprop_inputs = false; std::vector<IncomingTensor> backprop_inputs;
  bool has_known_shape = false; std::vector<int64_t> known_shape;
  VariableID variable_id;
  const std::vector<IncomingTensor>& get_backprop_inputs() const { return has_backprop_inputs ? backprop_inputs : incoming_nodes; }
  bool is_source() const { return incoming_nodes.empty(); }
  bool is_variable() const { return variable_id.valid(); }
};

struct Graph {                    // src/graph.rs:17-100
  static const size_t NUM_NODES_WARN = 50000, NUM_NODES_CRITICAL = 500000;
  std::vector<std::unique_ptr<TensorInternal>> node_set;
  std::unordered_map<int, TensorID> variable2node;
  VariableEnvironment* env = nullptr;
  TensorID install(std::unique_ptr<TensorInternal> node);
  TensorInternal& inner(TensorID id) { return *node_set[id]; }
  Tensor tensor(TensorID id) { return Tensor{id, this}; }
  Tensor placeholder(const std::string& name, const std::vector<int64_t>& shape);   // Context::placeholder graph.rs:178-199
  Tensor variable_by_id(VariableID vid);                                            // variable.rs:709-726
  Tensor variable_by_name(const std::string& name, const std::string& ns);
  void clear() { node_set.clear(); variable2node.clear(); }
};

struct TensorBuilder {            // src/tensor.rs:609-803
  Graph* graph; TensorID shape = -1; std::vector<IncomingTensor> in_nodes; bool differentiable = true;
  bool has_bp = false; std::vector<IncomingTensor> bp; bool has_known = false; std::vector<int64_t> known;
  VariableID variable_id; std::string placeholder; bool is_ph = false;
  explicit TensorBuilder(Graph* g) : graph(g) {}
  TensorBuilder& append_input(Tensor t, bool allow_mut) { return append_input_with_selector(t, allow_mut, 0); }
  TensorBuilder& append_input_with_selector(Tensor t, bool allow_mut, int sel);
  TensorBuilder& append_backprop_input(Tensor t);
  TensorBuilder& set_shape(Tensor s) { shape = s.id; return *this; }
  TensorBuilder& set_differentiable(bool d) { differentiable = d; return *this; }
  TensorBuilder& set_known_shape(const std::vector<int64_t>& s);
  TensorBuilder& set_variable(VariableID v) { variable_id = v; return *this; }
  TensorBuilder& set_placeholder_name(const std::string& n) { placeholder = n; is_ph = true; return *this; }
  Tensor build(Op* op);            // takes ownership
};

// ---------------------------------------------------------------------------------------------- Op::compute side
enum class InputKind { NonVariable, RdOnlyVariable, RdWrVariable };   // OpInput, src/op.rs:126-130
struct OpInput { NdArray arr; InputKind kind; bool taken = false; };
struct Evaluation;                // per-run state (pending optimizer updates, dropout counters)

struct ComputeContext {           // src/op.rs:186-309
  std::vector<OpInput> xs; std::vector<NdArray> ys;
  Device* dev; Evaluation* run; TensorID node;
  bool accept_i32 = false;        // set by ops that consume int32 index buffers natively
  bool accept_lazy = false;       // set by ops that fuse a deferred producer (AddOp / ReLU / greater / MulOp / Shape)
  bool accept_expr = false;       // set by ops that extend (or pass through) a pending elementwise expression
  NdArray input(int i);           // each input may be taken once (:206-233)
  NdArray input_mut(int i);       // only RdWrVariable edges (:239-259)
  int num_inputs() const { return (int)xs.size(); }
  void append_output(NdArray y) { ys.push_back(std::move(y)); }
  void append_output_view(NdArray y);     // copied if any input is a variable (:273-287)
  void append_empty_output();             // 0-d zero (:289-294)
};

struct GradientContext {          // src/op.rs:342-434
  Tensor gy, y; Graph* g; std::vector<Tensor> gxs;   // invalid Tensor == None
  Tensor output_grad() const { return gy; }
  Tensor output() const { return y; }
  Tensor input(int i) const;
  std::vector<Tensor> inputs() const;
  int num_inputs() const;
  Graph* graph() const { return g; }
  void append_input_grad(Tensor gx) { gxs.push_back(gx); }
  void append_none() { gxs.push_back(Tensor{}); }
};

std::vector<Tensor> compute_gradients(const std::vector<Tensor>& ys, const std::vector<Tensor>& xs,
                                      const std::vector<Tensor>* gys, Graph* g);       // src/gradient.rs:20-84

// ---------------------------------------------------------------------------------------------- variables
struct VariableEnvironment {      // src/variable.rs:152-155: Vec<RefCell<NdArray>> -> arrays resident in HBM
  Device* dev; bool owns_dev = false;
  bool fuse_elementwise = true;   // deferred elementwise expressions (fuse.cc); off = one launch per op, same values
  std::vector<NdArray> array_list;
  std::vector<std::pair<std::string, std::string>> names;        // index = VariableID: (namespace, name)
  std::map<std::pair<std::string, std::string>, int> name_to_id;
  // data-parallel state
  int rank = 0, world = 1;
  explicit VariableEnvironment(int device_index);
  ~VariableEnvironment();
  VariableID set(const std::string& ns, const std::string& name, const Shape& shape, const float* data);  // slot().name(..).set(..)
  VariableID find(const std::string& ns, const std::string& name) const;
  std::vector<VariableID> current_var_ids(const std::string& ns) const;
  std::vector<float> get(VariableID v);                          // D2H
  void put(VariableID v, const float* data, size_t n);           // H2D (load / test perturbation)
  std::string save_json();                                       // variable.rs:549-598 format
  void load_json(const std::string& js);
};

// ---------------------------------------------------------------------------------------------- evaluation
struct Feed { bool by_name; std::string name; TensorID id; NdArray value; };   // src/evaluation.rs:174-180
struct EvalResult { bool ok = true; int err_code = 0; std::string err_msg; NdArray value; };

struct PendingUpdate { int kind; float h[4]; NdArray p, g, s0, s1, t; };
struct Evaluation {
  Graph* graph; Device* dev;
  std::vector<PendingUpdate> pending;     // optimizer ops of this run: flushed as ONE multi-tensor launch after all grads exist
  // data parallel: gradients are all-reduced in buckets on a communication stream as soon as enough of them exist, under the kernels that
  // still compute the remaining ones (ops_nn.cc reduce_bucket); `ar_next` = first pending update not yet handed to NCCL
  size_t ar_next = 0; std::vector<NdArray> ar_buckets;
  bool fuse = false;                      // elementwise fusion enabled for this run
  std::vector<int> consumers;             // per node id: consuming edges inside this evaluation (+1 per request as a target); metadata-only
                                          // consumers (Shape / Rank / Size) are not counted
  struct RowStack { std::vector<const float*> key; std::vector<NdArray> parts; NdArray stacked; };      // `parts` pins the blocks: their addresses cannot be recycled while the stack is cached
  std::vector<RowStack> row_stacks;       // stacked operands built in this run (the same G_t stack serves both weight gradients and the bias gradient)
  std::vector<int> sole_consumer;         // per node id: the one node that reads it (-1 none yet, -2 several / a target)
  bool sole_consumer_sums(TensorID id) const;    // true when the node's only reader in this evaluation is an AddN
  int consumers_of(TensorID id) const { return id >= 0 && id < (int)consumers.size() && consumers[id] > 0 ? consumers[id] : 1; }
};

// Graph::eval (src/evaluation.rs:252-362): DFS post-order, memo table, placeholders from feeds, variables from env.
std::vector<EvalResult> eval(Graph* g, const std::vector<Tensor>& targets, const std::vector<Feed>& feeds, bool fetch_to_host = true);

// ---------------------------------------------------------------------------------------------- optimizers
struct Optimizer {                // trait Optimizer, src/optimizers/mod.rs:49-99
  virtual ~Optimizer() {}
  virtual std::vector<Tensor> compute_updates(const std::vector<Tensor>& params, const std::vector<Tensor>& grads, Graph* g) = 0;
  void update(const std::vector<Tensor>& params, const std::vector<Tensor>& grads, Graph* g, const std::vector<Feed>& feeds);
  Tensor get_update_op(const std::vector<Tensor>& params, const std::vector<Tensor>& grads, Graph* g);
};
Optimizer* make_adam(VariableEnvironment* env, const std::vector<VariableID>& vars, const std::string& ns, float alpha, float eps, float b1, float b2);
Optimizer* make_sgd(float lr);
Optimizer* make_momentum_sgd(VariableEnvironment* env, const std::vector<VariableID>& vars, const std::string& ns, float lr, float momentum);
Optimizer* make_adagrad(VariableEnvironment* env, const std::vector<VariableID>& vars, const std::string& ns, float lr);
// optimizers::grad_helper (src/optimizers/mod.rs:21-46)
void grad_helper(const std::vector<Tensor>& losses, const std::string& ns, Graph* g, std::vector<Tensor>& vars, std::vector<Tensor>& grads);

// ---------------------------------------------------------------------------------------------- tensor_ops (src/tensor_ops/mod.rs)
namespace T {
Tensor convert_to_tensor(Graph* g, const Shape& shape, const std::vector<float>& data, bool meta = false);   // :2387
Tensor as_tensor(Graph* g, const std::vector<int64_t>& ints);   // AsTensor for [I; N], src/tensor.rs:925-940 (meta)
Tensor scalar(Graph* g, float v);                                // :2415
Tensor zeros(Graph* g, Tensor shape); Tensor ones(Graph* g, Tensor shape);
Tensor shape(Tensor x); Tensor rank(Tensor x); Tensor size(Tensor x);
Tensor nth_tensor(Tensor x, int n); Tensor identity(Tensor x); Tensor stop_gradient(Tensor x);
Tensor add(Tensor a, Tensor b); Tensor sub(Tensor a, Tensor b); Tensor mul(Tensor a, Tensor b); Tensor div(Tensor a, Tensor b);
Tensor unary(const std::string& name, Tensor x, float p0 = 0.f);   // sin .. atanh, exp.., sqrt, pow(p0), neg, abs, sign, floor, ceil, inv, inv_sqrt, square, sigmoid, relu, softplus, elu(p0)
Tensor clip(Tensor x, float lo, float hi);
Tensor cmp(const std::string& name, Tensor a, Tensor b);           // equal .. lesser_equal, maximum, minimum
Tensor add_n(const std::vector<Tensor>& xs);
Tensor reduce(const std::string& name, Tensor x, Tensor axes, bool keep_dims);   // sum mean prod min max
Tensor sum_all(Tensor x); Tensor mean_all(Tensor x);
Tensor argmax(Tensor x, int axis, bool keep_dim); Tensor argmin(Tensor x, int axis, bool keep_dim);
Tensor reduce_logsumexp(Tensor x, int axis, bool keep_dim); Tensor softmax(Tensor x, int axis); Tensor log_softmax(Tensor x, int axis);
Tensor sigmoid_cross_entropy(Tensor y, Tensor t); Tensor softmax_cross_entropy(Tensor y, Tensor t); Tensor sparse_softmax_cross_entropy(Tensor y, Tensor t);
Tensor matmul(Tensor a, Tensor b); Tensor batch_matmul_t(Tensor a, Tensor b, bool ta, bool tb);
Tensor tensordot(Tensor a, Tensor b, Tensor a_axes, Tensor b_axes);
Tensor reshape(Tensor x, Tensor shape); Tensor flatten(Tensor x); Tensor transpose(Tensor x, Tensor perm);
Tensor squeeze(Tensor x, Tensor axes); Tensor expand_dims(Tensor x, Tensor axes);
Tensor slice(Tensor x, const std::vector<int64_t>& starts, const std::vector<int64_t>& ends);
std::vector<Tensor> split(Tensor x, const std::vector<int64_t>& sizes, int axis);
Tensor concat(const std::vector<Tensor>& xs, int axis); Tensor tile(Tensor x, int axis, int num);
Tensor gather_common(Tensor param, Tensor indices, int axis); Tensor gather(Tensor param, Tensor indices, int axis);
Tensor access_elem(Tensor x, int64_t i); Tensor setdiff1d(Tensor a, Tensor b);
Tensor conv2d(Tensor x, Tensor w, int pad, int stride, int dilation);
Tensor conv2d_transpose(Tensor x, Tensor w, int pad, int stride, int dilation);
Tensor max_pool2d(Tensor x, int size, int pad, int stride);
Tensor dropout(Tensor x, float ratio, bool train, uint64_t seed);
Tensor random(Graph* g, int kind, Tensor shape, float p0, float p1, uint64_t seed);   // random_normal .. gamma (:2426-2676); kind = agb_rand_kind
Tensor assign(Tensor x, Tensor y); Tensor control_dependencies(Tensor x, const std::vector<Tensor>& deps);
Tensor hook(Tensor x, std::function<void(const NdArray&, const std::vector<float>&)> f);   // Tensor::raw_hook: D2H sync point
std::vector<Tensor> grad(const std::vector<Tensor>& ys, const std::vector<Tensor>& xs);          // :94-114
std::vector<Tensor> grad_with_default(const std::vector<Tensor>& ys, const std::vector<Tensor>& xs, const std::vector<Tensor>& gys);
// composites (:1173,1291,1695,1845,2325,2362)
Tensor reduce_variance(Tensor x, Tensor axes, bool keep_dims); Tensor leaky_relu(Tensor x, float alpha);
Tensor mean_squared_error(Tensor y, Tensor t); Tensor normalize(Tensor x, Tensor axes); Tensor batch_norm(Tensor x, Tensor scale, Tensor shift);
}  // namespace T

// helpers shared by the op files
Shape as_shape(Device* dev, NdArray& a);                         // ndarray_ext::as_shape: float vector -> usize vector
std::vector<int64_t> as_ints(Device* dev, NdArray& a);
inline bool is_scalar_shape(const Shape& s) { return s.empty() || (s.size() == 1 && s[0] == 0); }   // ndarray_ext.rs:120-122
inline int normalize_negative_axis(int64_t axis, int ndim) { return (int)(axis < 0 ? ndim + axis : axis); }
NdArray materialize_lazy(Device* dev, const NdArray& a);      // runs the deferred producer un-fused (always correct)
// fuse.cc: each returns an array with `expr` set, or an invalid array (no expr) when the op cannot join a fused program
NdArray expr_unary(ComputeContext& c, int op, float p0, NdArray x);
NdArray expr_binary(ComputeContext& c, int op, NdArray a, NdArray b);
NdArray expr_binary_imm(ComputeContext& c, int op, NdArray x, float imm, bool imm_is_lhs);
NdArray expr_passthrough(ComputeContext& c, const NdArray& x);
NdArray expr_materialize(Device* dev, const NdArray& x);
bool expr_has_value(const NdArray& x);
NdArray expr_slice(ComputeContext& c, const NdArray& x, const std::vector<int64_t>& start, const std::vector<int64_t>& len);   // slice of a pending expression = the expression over sliced leaves
NdArray expr_pad(ComputeContext& c, const Shape& full, const std::vector<int64_t>& start, NdArray gy);
bool expr_sum_pads(ComputeContext& c, const std::vector<NdArray>& xs, NdArray* out);
NdArray expr_gemm_ta(ComputeContext& c, NdArray a, NdArray b);                 // deferred A^T * B whose only reader is an AddN
bool expr_sum_gemms(ComputeContext& c, const std::vector<NdArray>& xs, NdArray* out);
NdArray expr_scatter(ComputeContext& c, const Shape& table, int axis, NdArray idx, NdArray gy);   // deferred GatherGrad whose only reader is an AddN
bool expr_sum_scatters(ComputeContext& c, const std::vector<NdArray>& xs, NdArray* out);
NdArray expr_colsum(ComputeContext& c, NdArray gy, const Shape& target);       // deferred MaybeReduceSum [R, N] -> [1, N] whose only reader is an AddN
bool expr_sum_colsums(ComputeContext& c, const std::vector<NdArray>& xs, NdArray* out);
int fuse_selftest(int n_cases, uint32_t seed, int* n_compiled);                                 // host-only check of the program compiler (0 = ok)
bool stackable(const NdArray& t);                                                                   // 2-D, unit column stride, 16-byte aligned rows
NdArray stack_rows(Evaluation& run, Device* dev, const std::vector<NdArray>& parts);
NdArray stack_vectors(Evaluation& run, Device* dev, const std::vector<NdArray>& parts);          // n vectors of B elements -> contiguous [n * B]              // agb_concat_rows, remembered for the rest of the run
bool expr_materialize_into(Device* dev, const NdArray& x, NdArray dest);
Op* make_optimizer_op(int kind, float h0, float h1, float h2, float h3);
void flush_pending_updates(Evaluation& run, VariableEnvironment* env);

}  // namespace agx
