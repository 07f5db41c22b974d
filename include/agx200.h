/*
 * agx200.h — graph-level C ABI of the B200 backend: the device-resident evaluator with the reference's host-side
 * interfaces (Graph/Context, Tensor, tensor_ops constructors, grad, Evaluator/Feeder, VariableEnvironment, Optimizer).
 *
 * The kernel-level boundary (what a Rust `Op::compute` binds) is include/agb200.h.  This header exports the C++ host
 * engine (rust-autograd_b200/csrc/engine) that stands in for the crate's L2 layer (reference src/evaluation.rs,
 * src/op.rs, src/variable.rs) plus the unchanged symbolic L3/L4 layers (src/graph.rs, src/gradient.rs,
 * src/tensor_ops/mod.rs, src/optimizers) so that the re-hosted test-suite and the benchmarks can drive whole graphs.
 *
 * Conventions: every function returns 0 or an error code: 1..5 = OpError variants (src/op.rs:67-73), >= 100 device
 * errors (agb200.h), 200 = a condition on which the reference panics.  `agx_last_error()` gives the message.
 * Tensors are addressed by integer id inside their graph (reference TensorID, src/graph.rs:11).
 */
#ifndef AGX200_H
#define AGX200_H
#include "agb200.h"
#ifdef __cplusplus
extern "C" {
#endif

#define AGX_ERR_PANIC 200

typedef struct agx_env agx_env;       /* VariableEnvironment (src/variable.rs:152-155) + its device context */
typedef struct agx_graph agx_graph;   /* Context / Graph (src/graph.rs:17-20,109-200) bound to an environment */
typedef struct agx_opt agx_opt;       /* Optimizer (src/optimizers/mod.rs:49-99) */
typedef struct agx_results agx_results;

typedef struct agx_feed {             /* Feed (src/evaluation.rs:174-180): by placeholder name (name != NULL) or by tensor id */
  const char* name; int tensor_id;
  const float* data;                  /* host memory (copied H2D inside the call) or, with on_device != 0, a device pointer */
  const int64_t* shape; int rank; int on_device;
} agx_feed;

const char* agx_last_error(void);

/* ---- VariableEnvironment ---- */
int agx_env_new(int device, agx_env** out);
int agx_env_free(agx_env* env);
/* Automatic step-plan cache: an evaluation seen for the third time with the same graph structure, targets and feed
 * shapes is replayed from a CUDA graph captured at its second sight — also when the graph object was rebuilt in between, as the
 * reference's training loops do (examples/mlp_mnist.rs:74).  On by default; results are the eager ones. */
int agx_env_set_plan_cache(agx_env* env, int on);
int agx_env_plan_stats(agx_env* env, int64_t* captures, int64_t* replays, int* live_plans);
int agx_env_ctx(agx_env* env, agb_ctx** out);
int agx_env_set(agx_env* env, const char* ns, const char* name, const float* data, const int64_t* shape, int rank, int* vid);   /* slot().name(..).set(..) */
int agx_env_find(agx_env* env, const char* ns, const char* name, int* vid);          /* -1 when absent */
int agx_env_var_count(agx_env* env, int* n);
int agx_env_var_ids(agx_env* env, const char* ns, int* out, int cap, int* n);        /* current_var_ids */
int agx_env_var_shape(agx_env* env, int vid, int64_t* shape, int* rank);
int agx_env_get(agx_env* env, int vid, float* out, int64_t cap);                     /* D2H */
int agx_env_put(agx_env* env, int vid, const float* data, int64_t n);                /* H2D */
int agx_env_var_ptr(agx_env* env, int vid, float** dptr);                            /* device address of the variable */
int agx_env_save(agx_env* env, const char* path);                                    /* JSON, src/variable.rs:549-598 */
int agx_env_load(agx_env* env, const char* path);
int agx_env_set_data_parallel(agx_env* env, int rank, int world, const void* nccl_id128);
/* Deferred elementwise expressions (engine/fuse.cc): on by default.  Chains of unary / binary / compare / small AddN
 * nodes run as ONE agb_fused_ewise launch, slice gradients are summed in place, a gradient accumulation sum_t A_t^T * G_t runs as one
 * long-K GEMM, and independent MatMuls against one weight (the time steps of an unrolled RNN) run as one row-stacked GEMM; off = one
 * launch per node.  Elementwise values are bit-identical either way; the regrouped GEMMs differ by fp32 reassociation. */
int agx_env_set_fusion(agx_env* env, int on);
/* ---- host-callback ops (class C): user-defined ops and hooks.  The reference's `Op` trait is public API (src/op.rs:1-48,90-101;
 * tests/test_core.rs:6-36) and its compute() works on ndarray views, so a user op here sees HOST copies of its inputs (an explicit D2H
 * synchronisation point, like HookOp / MapOp) and hands host arrays back; they re-enter HBM when a device kernel consumes them. */
typedef struct agx_host_array { const float* data; const int64_t* shape; int rank; } agx_host_array;
typedef struct agx_out agx_out;
/* Op::compute: append outputs with agx_out_append (several = a multi-output op, selected by nth_tensor); return 0, or 1..5 = an OpError
 * variant (op.rs:67-73) with the message set by agx_out_error. */
typedef int (*agx_compute_fn)(void* user, const agx_host_array* inputs, int n_inputs, agx_out* out);
int agx_out_append(agx_out* out, const float* data, const int64_t* shape, int rank);
int agx_out_error(agx_out* out, const char* message);
/* Op::grad: called while T::grad builds the backward graph; gxs[i] = tensor id of the gradient for input i, or -1 for None. */
typedef void (*agx_grad_fn)(void* user, agx_graph* g, const int* inputs, int n_inputs, int y, int gy, int* gxs);
/* Tensor::builder(g).append_input(..).build(op).  grad == NULL: every input gradient is None.  The callbacks must outlive the graph. */
int agx_custom_op(agx_graph* g, const char* name, const int* inputs, int n_inputs, agx_compute_fn compute, agx_grad_fn grad, void* user, int* tid);
/* HookOp (hook_ops.rs:5-31; Tensor::{raw_hook, show, show_shape, print}, tensor.rs:198-320): identity node that runs a callback on the
 * host value when it is evaluated.  kind 0 = raw_hook (fn), 1 = show, 2 = show_shape, 3 = print(text); 1..3 write to stderr. */
typedef void (*agx_hook_fn)(void* user, const agx_host_array* value);
int agx_hook(agx_graph* g, int tensor, int kind, const char* text, agx_hook_fn fn, void* user, int* tid);

/* Host-only self test of the fused-program compiler (no device needed; run by the CPU test suite): `n_cases` random expression DAGs are
 * compiled and their instruction streams interpreted on the host; returns 0 when every stored register reproduces its node's value. */
int agx_fuse_selftest(int n_cases, unsigned seed, int* n_compiled);

/* ---- Graph construction ---- */
int agx_graph_new(agx_env* env, agx_graph** out);
int agx_graph_free(agx_graph* g);
int agx_graph_clear(agx_graph* g);
int agx_graph_size(agx_graph* g, int* n);
int agx_placeholder(agx_graph* g, const char* name, const int64_t* shape, int rank, int* tid);
int agx_variable(agx_graph* g, int vid, int* tid);
int agx_variable_by_name(agx_graph* g, const char* ns, const char* name, int* tid);
int agx_convert_to_tensor(agx_graph* g, const float* data, const int64_t* shape, int rank, int* tid);
/* generic constructor: fn = the reference's tensor_ops function name ("matmul", "conv2d", "reduce_sum", ...);
 * tensor arguments, then integer and float attributes in the order of the reference signature (see capi.cc table) */
int agx_call(agx_graph* g, const char* fn, const int* tensors, int nt, const int64_t* ints, int ni, const double* floats, int nf,
             int* out, int cap, int* nout);
int agx_grad(agx_graph* g, const int* ys, int ny, const int* xs, int nx, const int* gys /* NULL or ny ids */, int* out /* nx */);
int agx_grad_helper(agx_graph* g, const int* losses, int n, const char* ns, int* vars, int* grads, int cap, int* nout);
int agx_tensor_op_name(agx_graph* g, int tid, char* buf, int cap);
int agx_tensor_variable_id(agx_graph* g, int tid, int* vid);

/* ---- Evaluation ---- */
int agx_eval(agx_graph* g, const int* targets, int n, const agx_feed* feeds, int nfeeds, agx_results** out);
/* the same evaluation without the host sync: kernels and the results' D2H copies (copy stream, pinned staging) are enqueued and the
 * call returns; agx_results_fetch(r) completes the copies later — e.g. after the NEXT step has been launched — and only then may
 * agx_results_data be read.  Device-side index errors of the run surface at the next synchronising call. */
int agx_eval_launch(agx_graph* g, const int* targets, int n, const agx_feed* feeds, int nfeeds, agx_results** out);
int agx_results_fetch(agx_results* r);
int agx_results_count(agx_results* r, int* n);
/* Step graphs (the reference rebuilds and re-walks its graph every step, mlp_mnist.rs:74, cnn_mnist.rs:96): one
 * evaluation for side effects (agx_run semantics) is captured into a CUDA graph after two eager warm-up runs; agx_step_launch replays
 * it with no host-side graph walk.  All feeds must be device-resident and are re-read at the same addresses on every launch. */
typedef struct agx_step agx_step;
int agx_step_capture(agx_graph* g, const int* targets, int n, const agx_feed* feeds, int nfeeds, agx_step** out);
int agx_step_launch(agx_step* s);
int agx_step_free(agx_step* s);
int agx_results_status(agx_results* r, int i, int* code, const char** msg);
int agx_results_shape(agx_results* r, int i, int64_t* shape, int* rank);
int agx_results_data(agx_results* r, int i, const float** data, int64_t* n);
int agx_results_free(agx_results* r);
/* evaluate without fetching values (training step): asynchronous; results stay in HBM */
int agx_run(agx_graph* g, const int* targets, int n, const agx_feed* feeds, int nfeeds);

/* ---- Optimizers ---- */
int agx_opt_adam(agx_env* env, const int* vids, int n, const char* ns, float alpha, float eps, float b1, float b2, agx_opt** out);
int agx_opt_sgd(float lr, agx_opt** out);
int agx_opt_momentum_sgd(agx_env* env, const int* vids, int n, const char* ns, float lr, float momentum, agx_opt** out);
int agx_opt_adagrad(agx_env* env, const int* vids, int n, const char* ns, float lr, agx_opt** out);
int agx_opt_compute_updates(agx_opt* o, agx_graph* g, const int* params, const int* grads, int n, int* out);
int agx_opt_get_update_op(agx_opt* o, agx_graph* g, const int* params, const int* grads, int n, int* tid);
int agx_opt_update(agx_opt* o, agx_graph* g, const int* params, const int* grads, int n, const agx_feed* feeds, int nfeeds);
int agx_opt_free(agx_opt* o);

#ifdef __cplusplus
}
#endif
#endif
