/*
 * agb200.h — C ABI of the B200-native (sm_100a) execution backend for rust-autograd's
 * op-evaluation hot path.
 *
 * This is the drop-in boundary: every function below is what a built-in
 * `Op::compute` (reference `src/op.rs:90-101`, called from `src/evaluation.rs:323-324`) binds through
 * a thin `extern "C"` module — the same seam where the reference declares its optional CPU
 * accelerator today (`src/tensor_ops/blas_ffi.rs:17-158`, `cblas_sgemm` call sites
 * `src/tensor_ops/dot_ops.rs:209-224`).  Plain pointers and sizes only: no torch types, no C++.
 *
 * Conventions
 *   - All tensors are f32 in device memory (HBM).  Index-valued tensors (argmax, pool indices,
 *     labels, gather ids) are f32 too, exactly as in the reference (`src/ndarray_ext.rs:29-31`).
 *   - `agb_tensor` = borrowed device view {ptr, rank, shape, stride}; strides are in ELEMENTS and
 *     may be 0 (broadcast) or permuted (transpose views), mirroring `NdArrayView`.
 *   - Outputs are caller-allocated, C-contiguous unless stated otherwise.
 *   - Every call is asynchronous on the context's CUDA stream; only agb_d2h / agb_sync block.
 *   - Return value: 0 = ok; 1..5 = the five `OpError` variants of `src/op.rs:67-73`;
 *     >= 100 = CUDA / NCCL / driver failure.  `agb_last_error()` gives the message.
 *   - There is no CPU fallback: without a CUDA device agb_init fails with AGB_ERR_CUDA.
 */
#ifndef AGB200_H
#define AGB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AGB_MAX_RANK 8

/* ---- status codes (1..5 mirror OpError, src/op.rs:67-73) ---- */
enum {
  AGB_OK = 0,
  AGB_ERR_NDARRAY = 1,            /* OpError::NdArrayError      */
  AGB_ERR_INCOMPATIBLE_SHAPE = 2, /* OpError::IncompatibleShape */
  AGB_ERR_TYPE_UNSUPPORTED = 3,   /* OpError::TypeUnsupported   */
  AGB_ERR_INVALID_DIMS = 4,       /* OpError::InvalidDims       */
  AGB_ERR_OUT_OF_BOUNDS = 5,      /* OpError::OutOfBounds       */
  AGB_ERR_CUDA = 100,
  AGB_ERR_NCCL = 101,
  AGB_ERR_UNSUPPORTED = 102       /* shape/stride combination the device path rejects */
};

typedef struct agb_ctx agb_ctx;

typedef struct agb_tensor {
  float*  ptr;
  int32_t rank;
  int64_t shape[AGB_MAX_RANK];
  int64_t stride[AGB_MAX_RANK]; /* elements */
} agb_tensor;

/* ---- GEMM / conv arithmetic modes (north_star: "f32-faithful 3xTF32 mode and a plain TF32 mode") ---- */
enum {
  AGB_MATH_3XTF32 = 0, /* tcgen05 kind::tf32, hi/lo split, 3 MMAs per product: ~1e-6 rel  */
  AGB_MATH_TF32   = 1, /* tcgen05 kind::tf32, single MMA: ~1e-3 rel                       */
  AGB_MATH_FP32   = 2  /* CUDA-core FMA path (also used automatically for shapes TMA rejects) */
};

/* ======================= runtime ======================= */
/* one host thread <-> one context <-> one CUDA stream (SURVEY §8b "Threading") */
int  agb_init(int device, agb_ctx** out);
int  agb_destroy(agb_ctx* ctx);
const char* agb_last_error(void);
int  agb_device_count(int* out);
int  agb_sm_count(agb_ctx* ctx, int* out);
int  agb_set_math_mode(agb_ctx* ctx, int mode);
int  agb_get_math_mode(agb_ctx* ctx, int* mode);
/* Deterministic reductions (default ON).  Split-K partial sums (filter gradients, skinny GEMMs) and the per-channel side sums of the
 * dgrad / pool-backward epilogues (bias gradients) are written to scratch and added in a fixed order instead of red.global.add in
 * arrival order, so a training step is bit-reproducible run to run and a captured step equals the eager one — the reference's
 * filter gradient is a sequential batch loop (conv2d.rs:631-734) and its reductions are sequential folds (reduction_ops.rs:54-108).
 * 0 restores the atomic form (saves the partial-sum traffic: ~1 % of a VGG step). */
int  agb_set_deterministic(agb_ctx* ctx, int on);
int  agb_get_deterministic(agb_ctx* ctx, int* on);
/* number of kernels this library has launched on ctx since agb_init (bench "gpu_launches") */
int  agb_launch_count(agb_ctx* ctx, int64_t* out);

/* Live per-entry-point timing with CUDA events on the context's own stream (bench.py roofline): while enabled, every call
 * of the profiled classes below is bracketed by an event pair and its algorithmic work (FLOPs for contractions, bytes for
 * bandwidth kernels) is accumulated.  agb_prof_collect synchronises and returns totals for one class. */
enum { AGB_PROF_GEMM = 0, AGB_PROF_CONV_FPROP, AGB_PROF_CONV_DGRAD, AGB_PROF_CONV_WGRAD, AGB_PROF_EWISE, AGB_PROF_REDUCE, AGB_PROF_SOFTMAX,
       AGB_PROF_POOL, AGB_PROF_OPTIM, AGB_PROF_CONV_SMALLC_FPROP, AGB_PROF_CONV_SMALLC_WGRAD, AGB_PROF_CONV_SIMT, AGB_PROF_COUNT };
int  agb_prof_enable(agb_ctx* ctx, int on);
int  agb_prof_collect(agb_ctx* ctx, int cls, double* total_ms, int64_t* calls, double* work /* FLOPs or bytes */);
int  agb_prof_reset(agb_ctx* ctx);

int  agb_alloc(agb_ctx* ctx, size_t bytes, void** out);    /* stream-ordered caching arena */
int  agb_free(agb_ctx* ctx, void* ptr);
int  agb_trim(agb_ctx* ctx);                                /* release cached blocks to the driver */
int  agb_mem_stats(agb_ctx* ctx, size_t* live_bytes, size_t* cached_bytes, size_t* peak_bytes);
int  agb_host_alloc(size_t bytes, void** out);              /* pinned host memory for feeds */
int  agb_host_free(void* ptr);
int  agb_h2d(agb_ctx* ctx, void* dst, const void* src, size_t bytes); /* async on ctx stream */
int  agb_d2h(agb_ctx* ctx, void* dst, const void* src, size_t bytes); /* async + stream sync  */
int  agb_d2d(agb_ctx* ctx, void* dst, const void* src, size_t bytes);
int  agb_memset0(agb_ctx* ctx, void* dst, size_t bytes);
int  agb_sync(agb_ctx* ctx);
/* Host-feed staging: the device analogue of handing the evaluator a host array per step (Feeder::push, evaluation.rs:296) without
 * serialising the copy in front of the step.  The caller keeps two device buffers per placeholder:
 *   agb_stage_wait  — the compute stream waits for every copy staged so far (call before launching the step that reads them);
 *   agb_stage_mark  — records "everything enqueued on the compute stream so far has been issued" (call before launching step i:
 *                     the buffers of step i-1 are free once the mark passes);
 *   agb_stage_h2d   — async copy from PINNED host memory on a second stream, ordered after the last mark. */
/* arena blocks referenced by instantiated step graphs must stay mapped: +1 / -1 pins (agb_trim and out-of-memory trimming are disabled while > 0) */
int  agb_arena_pin(agb_ctx* ctx, int delta);
int  agb_stage_mark(agb_ctx* ctx);
int  agb_stage_h2d(agb_ctx* ctx, void* dst, const void* src, size_t bytes);
int  agb_stage_wait(agb_ctx* ctx);
/* results: agb_stage_d2h copies src (device) to PINNED host memory on a third stream, ordered after the compute work enqueued so
 * far; *done_event is complete when the bytes are on the host (agb_event_sync, then agb_event_destroy).  Lets the host read step i's
 * loss while step i+1 runs. */
int  agb_stage_d2h(agb_ctx* ctx, void* dst_pinned, const void* src, size_t bytes, void** done_event);
int  agb_event_sync(void* ev);
int  agb_flush_l2(agb_ctx* ctx);                            /* writes a >L2 scratch buffer */

/* CUDA-event timing on the context's own stream (torch.cuda.Event would not see it) */
int  agb_event_create(void** ev);
int  agb_event_destroy(void* ev);
int  agb_event_record(agb_ctx* ctx, void* ev);
int  agb_event_elapsed_ms(void* start, void* stop, float* ms); /* synchronises on stop */

/* CUDA-graph capture of a launch sequence */
int  agb_graph_begin(agb_ctx* ctx);
/* graph_exec == NULL aborts the capture (nothing is instantiated) */
int  agb_graph_end(agb_ctx* ctx, void** graph_exec);
int  agb_graph_launch(agb_ctx* ctx, void* graph_exec);
int  agb_graph_destroy(void* graph_exec);

/* ======================= dense contractions ======================= */
/* replaces MatMul::compute / BatchMatMul::compute (src/tensor_ops/dot_ops.rs:565-606, 632-695) and the
 * cblas_sgemm / matrixmultiply::sgemm call sites (dot_ops.rs:209-224, 400).
 * a, b: rank-2 (or rank>=3 with identical leading batch dims) views, any strides;
 * c: C-contiguous [.., m, n].  beta in {0,1}.  Transposes are applied to the LAST TWO axes. */
int  agb_gemm_f32(agb_ctx* ctx, int trans_a, int trans_b,
                  const agb_tensor* a, const agb_tensor* b, agb_tensor* c, float beta);

/* replaces Conv2D::compute (conv_ops/conv2d.rs:532-554): x [B,C,H,W], w [O,C,kh,kw] -> y [B,O,yh,yw].
 * Implicit GEMM: the im2col buffer (conv_ops/mod.rs:73-124) is never materialised. */
int  agb_conv2d_fprop_f32(agb_ctx* ctx, const agb_tensor* x, const agb_tensor* w, agb_tensor* y,
                          int pad, int stride, int dilation);
/* fused form of Conv2D -> AddOp(bias) -> ReLU (reference graph: examples/cnn_mnist.rs:38-45; ops conv2d.rs:532, binary_ops.rs:147,
 * activation_ops.rs:154): y = [relu](conv(x, w) [+ bias[o]]).  bias = NULL or O contiguous floats; the epilogue runs on the accumulator
 * registers, the pre-activation tensors never reach HBM. */
int  agb_conv2d_fprop_fused_f32(agb_ctx* ctx, const agb_tensor* x, const agb_tensor* w, const float* bias, int relu, agb_tensor* y,
                                int pad, int stride, int dilation);
/* Conv2D -> AddOp(bias) -> ReLU -> MaxPool2D(size 2, pad 0, stride 2) (examples/cnn_mnist.rs:38-45) with the pooling in the conv
 * epilogue: y_pooled [B,O,yh/2,yw/2] channels-last and idx_i32 (same layout; MaxPool2D's argmax = flat offsets into the [B,O,yh,yw]
 * activation, max_pool2d.rs:21-88) are written, the full-size activation never reaches HBM.  Returns AGB_ERR_UNSUPPORTED — before
 * launching anything — when the layer is outside the fused kernel's envelope (TF32 mode, stride 1, output width >= 128, O <= 128,
 * channels-last x); the caller then runs agb_conv2d_fprop_fused_f32 + agb_maxpool2d_fwd. */
int  agb_conv2d_fprop_pool_f32(agb_ctx* ctx, const agb_tensor* x, const agb_tensor* w, const float* bias, int relu, agb_tensor* y_pooled,
                               int32_t* idx_i32, int pad, int stride, int dilation);
/* replaces Conv2DTranspose::compute (conv_ops/conv2d_transpose.rs:250-272): gy [B,O,yh,yw],
 * w [O,C,kh,kw] -> gx [B,C,xh,xw], xh = s(yh-1) - 2p + d(kh-1) + 1 (conv2d_transpose.rs:55-56). */
int  agb_conv2d_dgrad_f32(agb_ctx* ctx, const agb_tensor* gy, const agb_tensor* w, agb_tensor* gx,
                          int pad, int stride, int dilation);
/* fused form of Conv2DTranspose followed by the ReLU backward of the layer below, mul(greater(a, 0), gx)
 * (activation_ops.rs:161-166): gx = conv2d_transpose(gy, w) * (mask_src > 0).  mask_src (NULL = no mask) is [B,C,xh,xw]; when it
 * has the strides of gx the compare-and-zero runs on the accumulator registers of the tensor-core epilogue.
 * chan_sum (NULL = not wanted): C floats that receive sum_{b,h,w} gx[b,c,h,w] — the bias gradient of the layer below
 * (MaybeReduceSum of the broadcast AddOp operand, binary_ops.rs:39-105), accumulated in the same epilogue. */
int  agb_conv2d_dgrad_fused_f32(agb_ctx* ctx, const agb_tensor* gy, const agb_tensor* w, const agb_tensor* mask_src, float* chan_sum,
                                agb_tensor* gx, int pad, int stride, int dilation);
/* ReLU sign bits as a side channel of the two fused forms above (activation_ops.rs:161-166 needs only the SIGN of the activation): the forward
 * writes 1 bit per element (channels-last order: word (pixel * O + o) / 32, bit o % 32; numel(y) / 32 words) next to y when the kernel that runs
 * supports it (*bits_written = 1), and the masked dgrad reads those 4 bytes per pixel per 32 channels instead of 128 bytes of mask_src
 * (2.1 GB less HBM traffic per VGG training step).  mask_src is still passed and is what every non-fusing path reads. */
int  agb_conv2d_fprop_fused_bits_f32(agb_ctx* ctx, const agb_tensor* x, const agb_tensor* w, const float* bias, int relu, agb_tensor* y,
                                     uint32_t* relu_bits, int* bits_written, int pad, int stride, int dilation);
int  agb_conv2d_dgrad_fused_bits_f32(agb_ctx* ctx, const agb_tensor* gy, const agb_tensor* w, const agb_tensor* mask_src, const uint32_t* mask_bits,
                                     float* chan_sum, agb_tensor* gx, int pad, int stride, int dilation);
/* replaces Conv2DFilterGrad::compute (conv2d.rs:737-744) and Conv2DTransposeFilterGrad::compute
 * (conv2d_transpose.rs:433-451): gw[O,C,kh,kw] = sum_b g[b] (x) im2col(img[b]).
 * img [B,C,H,W] is the tensor that gets im2col'd, g [B,O,yh,yw] the one that multiplies it. */
int  agb_conv2d_wgrad_f32(agb_ctx* ctx, const agb_tensor* img, const agb_tensor* g, agb_tensor* gw,
                          int pad, int stride, int dilation);
/* Activation layouts.  Logical [B,C,H,W] tensors are accepted by the conv and pooling entry points in two dense memory orders:
 * NCHW (C-contiguous, the reference's layout) and channels-last (strides {H*W*C, 1, W*C, C}).  The tcgen05 conv kernels are
 * native channels-last (TMA needs 16-byte aligned innermost coordinates; the 3x3 taps shift W by one element), operands in
 * the other order are converted through a tiled transpose.  Returns 1 when a conv of this geometry runs on the tensor cores,
 * i.e. when the caller should keep its activations channels-last. */
int  agb_conv_prefers_channels_last(int in_channels, int out_channels, int kh, int kw, int stride, int out_w);
/* materialise im2col(x) = Conv2D output #1, [B,C,kh,kw,yh,yw] (only when user code evaluates it) */
int  agb_im2col_f32(agb_ctx* ctx, const agb_tensor* x, agb_tensor* cols, int kh, int kw,
                    int pad, int stride, int dilation);

/* ======================= pooling ======================= */
/* MaxPool2D::compute (conv_ops/max_pool2d.rs:166-227): strict '>' scan, index = flat offset into the
 * whole input, stored as float in idx_f32 (may be NULL) and as int32 in idx_i32 (may be NULL). */
int  agb_maxpool2d_fwd(agb_ctx* ctx, const agb_tensor* x, agb_tensor* y, float* idx_f32, int32_t* idx_i32,
                       int size, int pad, int stride);
/* MaxPool2DGrad::compute (max_pool2d.rs:245-279): gx = 0; gx[idx[i]] += gy[i]. exactly one of idx_* non-NULL */
int  agb_maxpool2d_bwd(agb_ctx* ctx, const agb_tensor* gy, const float* idx_f32, const int32_t* idx_i32,
                       agb_tensor* gx);
/* same, given the forward window (size, stride; 0 = unknown) and an optional gate laid out like gy: gx = scatter(gy * (gate > 0)).
 * Windows that tile gx exactly (size == stride) are written gather-form: no zero-fill, no atomics.  With gate = the pooled forward
 * output of a ReLU activation this also performs the ReLU backward that follows in conv -> relu -> pool stacks
 * (relu'(x[argmax]) == (max > 0)). */
int  agb_maxpool2d_bwd_fused(agb_ctx* ctx, const agb_tensor* gy, const float* idx_f32, const int32_t* idx_i32, const float* gate,
                             float* chan_sum /* NULL or C floats: sum_{b,h,w} gx, see agb_conv2d_dgrad_fused_f32 */,
                             agb_tensor* gx, int size, int stride);
/* MaxPool2DGradGrad::compute (max_pool2d.rs:297-331): ggy[i] = ggx[idx[i]] */
int  agb_maxpool2d_gradgrad(agb_ctx* ctx, const agb_tensor* ggx, const float* idx_f32, const int32_t* idx_i32,
                            agb_tensor* ggy);

/* index buffers are kept as int32 on the device (exact for any size); API-visible float copies are made on demand */
int  agb_convert_i32_f32(agb_ctx* ctx, const int32_t* src, float* dst, int64_t n);

/* ======================= elementwise ======================= */
enum { /* unary ops: math_ops.rs:277-1019, activation_ops.rs:113-226, array_ops.rs:537-574 */
  AGB_U_COPY = 0, AGB_U_ABS, AGB_U_NEG, AGB_U_SQUARE, AGB_U_INV, AGB_U_INVSQRT, AGB_U_SIGN, AGB_U_FLOOR,
  AGB_U_CEIL, AGB_U_SQRT, AGB_U_POW /*p0*/, AGB_U_LN, AGB_U_LOG2, AGB_U_LOG10, AGB_U_EXP, AGB_U_EXP2,
  AGB_U_EXP10, AGB_U_SIN, AGB_U_COS, AGB_U_TAN, AGB_U_ASIN, AGB_U_ACOS, AGB_U_ATAN, AGB_U_SINH, AGB_U_COSH,
  AGB_U_TANH, AGB_U_ASINH, AGB_U_ACOSH, AGB_U_ATANH, AGB_U_SIGMOID, AGB_U_RELU, AGB_U_SOFTPLUS,
  AGB_U_ELU /*p0=alpha*/, AGB_U_CLIP /*p0=min,p1=max*/, AGB_U_SCALE /* x*p0 */, AGB_U_ADD_SCALAR /* x+p0 */,
  AGB_U_RSUB_SCALAR /* p0-x */, AGB_U_RDIV_SCALAR /* p0/x */,
  AGB_U_LGAMMA /* ln|Gamma(x)| */, AGB_U_DIGAMMA /* d/dx ln Gamma(x) */ /* math_ops.rs:1021-1060, `special` 0.10 Gamma trait */, AGB_U_COUNT
};
enum { /* binary ops: binary_ops.rs:147-290, math_ops.rs:86-184, activation_ops.rs:204-226, array_ops.rs:556-574 */
  AGB_B_ADD = 0, AGB_B_SUB, AGB_B_MUL, AGB_B_DIV, AGB_B_EQ, AGB_B_NE, AGB_B_GT, AGB_B_LT, AGB_B_GE, AGB_B_LE,
  AGB_B_MAX, AGB_B_MIN, AGB_B_ELU_GRAD /* (x, gy), p0=alpha */, AGB_B_CLIP_GRAD /* (x, gy), p0,p1 */,
  AGB_B_SIGMOID_XENT /* (x, t) xent_ops.rs:33-46 */, AGB_B_RELU_GRAD /* (x, gy): (x>0)*gy */,
  AGB_B_COUNT
};
/* y = f(x); x may be any strided view, y is C-contiguous with x's shape */
int  agb_unary(agb_ctx* ctx, int op, float p0, float p1, const agb_tensor* x, agb_tensor* y);
/* y = f(a, b) with numpy-style broadcasting expressed by the caller as zero strides: a and b must
 * already have y's rank and shape (stride 0 on broadcast axes). */
int  agb_binary(agb_ctx* ctx, int op, float p0, float p1, const agb_tensor* a, const agb_tensor* b, agb_tensor* y);
/* AddN::compute (array_ops.rs:503-528): y = xs[0] + ... + xs[n-1], all same shape, left fold */
int  agb_add_n(agb_ctx* ctx, int n, const agb_tensor* const* xs, agb_tensor* y);
int  agb_fill(agb_ctx* ctx, agb_tensor* y, float value);
/* Fused elementwise program: one launch evaluates a small DAG of the unary / binary functors above over a
 * [rows, cols] index space and writes up to AGB_FUSE_MAX_OUT of its values.  It replaces a chain of ndarray `map` / `Zip` passes
 * such as the backward compositions gy * (y - square(y)) (activation_ops.rs:150) or gy * (1 - square(y)) (math_ops.rs:854-858) and
 * the LSTM cell sigmoid(f) * c + sigmoid(i) * tanh(g) (examples/lstm_lm.rs:36-45); every instruction applies the SAME functor as
 * agb_unary / agb_binary, so the values are bit-identical to the unfused sequence.
 * Leaf l is read at ptr + r * pitch + c * cstride (pitch / cstride 0 = broadcast, pitch != cols = a sliced view) into register
 * `reg`; instruction i computes regs[dst] = f(regs[a], regs[b] or imm); output o stores regs[reg] at ptr + r * pitch + c. */
#define AGB_FUSE_MAX_LEAVES 16
#define AGB_FUSE_MAX_INSTR 48
#define AGB_FUSE_MAX_OUT 8
#define AGB_FUSE_REGS 32
enum agb_fuse_kind { AGB_F_UNARY = 0 /* f(a; p0) */, AGB_F_BINARY /* f(a, b) */, AGB_F_BINARY_IMM_B /* f(a, p0) */, AGB_F_BINARY_IMM_A /* f(p0, b) */ };
typedef struct agb_fuse_leaf { const float* ptr; int64_t pitch, cstride; int32_t reg; } agb_fuse_leaf;
typedef struct agb_fuse_instr { int32_t kind, op, dst, a, b; float p0; } agb_fuse_instr;
typedef struct agb_fuse_out { float* ptr; int64_t pitch; int32_t reg; } agb_fuse_out;
int  agb_fused_ewise(agb_ctx* ctx, int64_t rows, int64_t cols, int n_leaves, const agb_fuse_leaf* leaves,
                     int n_instr, const agb_fuse_instr* instr, int n_out, const agb_fuse_out* outs);
/* dst (any strides, e.g. a sliced region of a larger buffer) <- src (any strides), same shape.
 * Serves deep_copy, Concat/Tile, SliceGrad/SplitGrad (array_ops.rs:576-825), MaybeBroadcast. */
int  agb_copy_strided(agb_ctx* ctx, const agb_tensor* src, agb_tensor* dst);
/* dst[s * rows + r, :] = srcs[s][r, :]: n equally-shaped row blocks (unit column stride, row pitch src_pitch[s]) stacked into one
 * contiguous [n * rows, cols] matrix, one launch per 64 blocks (Concat along axis 0, array_ops.rs:576-620).  The engine uses it to turn
 * the per-time-step weight-gradient GEMMs of an unrolled RNN, sum_t A_t^T * G_t (gradient accumulation by AddN, gradient.rs:168-173), into
 * ONE long-K GEMM.  cols and every pitch must be multiples of 4, pointers 16-byte aligned. */
int  agb_concat_rows(agb_ctx* ctx, int n, const float* const* srcs, const int64_t* src_pitch, int64_t rows, int64_t cols, float* dst);
/* Dropout::compute (random_ops.rs:218-237): train: y = x*mask (NOT rescaled); mask is supplied in
 * `mask` when seed==0, otherwise generated on device (Philox, (seed,offset)) and written to `mask`. */
int  agb_dropout(agb_ctx* ctx, const agb_tensor* x, agb_tensor* y, agb_tensor* mask,
                 float ratio, uint64_t seed, uint64_t offset);
/* The random_* generator ops (random_ops.rs:6-214; ArrayRng::{random_uniform, random_normal, bernoulli, exponential, log_normal, gamma},
 * ndarray_ext.rs:276-388) on the device: y is filled from the counter stream Philox-4x32-10(key = seed, counter = (element, offset)),
 * so a call is reproducible from (seed, offset) and independent of the launch geometry.  The reference's XorShift stream is
 * parity-unpinned; distributions and parameter meaning follow rand_distr 0.4: Uniform [p0, p1), Normal(mean p0, std p1),
 * Bernoulli = (U[0,1) < p0) as 0/1, Exp(rate p0), LogNormal(mu p0, sigma p1), Gamma(shape p0, scale p1) by Marsaglia-Tsang. */
enum agb_rand_kind { AGB_RAND_UNIFORM = 0, AGB_RAND_NORMAL, AGB_RAND_BERNOULLI, AGB_RAND_EXP, AGB_RAND_LOGNORMAL, AGB_RAND_GAMMA, AGB_RAND_COUNT };
int  agb_random(agb_ctx* ctx, int kind, float p0, float p1, uint64_t seed, uint64_t offset, agb_tensor* y);
/* The same generators with the stream position kept in device memory: `cell` points to agb_stream_cell_bytes() zero-initialised bytes
 * owned by the caller (one cell per op instance).  Every call draws from position *cell and advances it ON THE DEVICE, so an op that is
 * evaluated repeatedly continues its stream like the reference's ArrayRng (RefCell<R>, ndarray_ext.rs:250-264; Dropout's rng,
 * random_ops.rs:218-245) — including under CUDA-graph replay, where host-side kernel arguments are frozen at capture.  cell == NULL
 * behaves like the plain entry points. */
int  agb_stream_cell_bytes(void);
int  agb_dropout_stream(agb_ctx* ctx, const agb_tensor* x, agb_tensor* y, agb_tensor* mask, float ratio, uint64_t seed, uint64_t offset, uint32_t* cell);
int  agb_random_stream(agb_ctx* ctx, int kind, float p0, float p1, uint64_t seed, uint64_t offset, uint32_t* cell, agb_tensor* y);

/* ======================= reductions ======================= */
enum { AGB_R_SUM = 0, AGB_R_MEAN, AGB_R_PROD, AGB_R_MIN, AGB_R_MAX };
/* impl_reduce_forward!/ReduceMean (reduction_ops.rs:54-108,187-215): x is viewed as [outer, r, inner]
 * (C-contiguous), reduced over r into y [outer, inner].  Multi-axis reductions are expressed by the
 * caller as one call per contiguous axis group, highest axis first, like the reference's fold order. */
int  agb_reduce(agb_ctx* ctx, int op, const float* x, float* y, int64_t outer, int64_t r, int64_t inner);
/* ArgMax/ArgMin (reduction_ops.rs:365-457): first occurrence, result as float */
int  agb_argreduce(agb_ctx* ctx, int is_max, const float* x, float* y, int64_t outer, int64_t r, int64_t inner);

/* ======================= softmax family ======================= */
/* softmax_impl (activation_ops.rs:61-96), LogSoftmax (xent_ops.rs:17-22), logsumexp_forward
 * (math_ops.rs:540-593), all over [outer, r, inner] reducing r */
int  agb_softmax(agb_ctx* ctx, const float* x, float* y, int64_t outer, int64_t r, int64_t inner);
int  agb_log_softmax(agb_ctx* ctx, const float* x, float* y, int64_t outer, int64_t r, int64_t inner);
int  agb_logsumexp(agb_ctx* ctx, const float* x, float* y, int64_t outer, int64_t r, int64_t inner);
/* SparseSoftmaxCrossEntropy::compute (xent_ops.rs:63-113): logits [B,C], labels [B] as float ->
 * loss [B] (= [B,1]), log_x [B,C] */
int  agb_sparse_xent_fwd(agb_ctx* ctx, const float* logits, const float* labels, float* loss, float* log_x,
                         int64_t batch, int64_t classes);
/* SparseSoftmaxCrossEntropyGrad::compute (xent_ops.rs:139-152): gx = (exp(log_x) - onehot(t)) * gy,
 * gy has gy_len elements (B or 1) and is broadcast over classes */
int  agb_sparse_xent_bwd(agb_ctx* ctx, const float* log_x, const float* labels, const float* gy, int64_t gy_len,
                         float* gx, int64_t batch, int64_t classes);
/* SoftmaxCrossEntropy::compute (xent_ops.rs:160-177): dense labels t [B,C] -> loss [B], log_x [B,C] */
int  agb_softmax_xent_fwd(agb_ctx* ctx, const float* logits, const float* t, float* loss, float* log_x,
                          int64_t batch, int64_t classes);

/* ======================= gather / scatter ======================= */
/* Gather::compute (array_ops.rs:353-384): param viewed as [pre, axis_len, post] (C-contiguous),
 * out [pre, n_idx, post]; negative indices are wrapped when normalize_negative != 0 */
int  agb_gather(agb_ctx* ctx, const float* param, const float* indices, float* out,
                int64_t pre, int64_t axis_len, int64_t post, int64_t n_idx, int normalize_negative);
/* GatherGrad::compute (array_ops.rs:401-466): gx = 0; gx[:, idx[j], :] += gy[:, j, :] */
int  agb_gather_grad(agb_ctx* ctx, const float* gy, const float* indices, float* gx,
                     int64_t pre, int64_t axis_len, int64_t post, int64_t n_idx);
/* the accumulation half of GatherGrad alone: gx[:, idx[j], :] += gy[:, j, :] with NO zero fill, so that a sum of GatherGrads of one
 * table (an embedding looked up at every time step, gradient.rs:168-173) scatters into a single buffer */
int  agb_scatter_add(agb_ctx* ctx, const float* gy, const float* indices, float* gx,
                     int64_t pre, int64_t axis_len, int64_t post, int64_t n_idx);

/* ======================= optimizers ======================= */
/* One fused multi-tensor launch replaces n AdamOp::compute calls (gradient_descent_ops/adam.rs:11-58):
 * m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; p -= alpha * (m/(1-b1^t)) / (sqrt(v/(1-b2^t)) + eps); t += 1.
 * t[i] points at the per-variable 0-d step counter living in HBM (optimizers/adam.rs:97).
 * grad_scale multiplies g on read (1/world after the NCCL sum; 1.0 on one GPU). */
int  agb_multi_tensor_adam(agb_ctx* ctx, int n, float* const* p, const float* const* g, float* const* m,
                           float* const* v, float* const* t, const int64_t* sizes,
                           float alpha, float eps, float b1, float b2, float grad_scale);
/* SGDOp (sgd.rs:14-26): p -= alpha*g */
int  agb_multi_tensor_sgd(agb_ctx* ctx, int n, float* const* p, const float* const* g, const int64_t* sizes,
                          float alpha, float grad_scale);
/* MomentumSGDOp (sgd.rs:28-40): v = momentum*v - lr*g; p += v */
int  agb_multi_tensor_momentum(agb_ctx* ctx, int n, float* const* p, const float* const* g, float* const* v,
                               const int64_t* sizes, float lr, float momentum, float grad_scale);
/* AdaGradOp (adagrad.rs:8-21): h += g^2; p -= lr*g/(sqrt(h)+1e-7) */
int  agb_multi_tensor_adagrad(agb_ctx* ctx, int n, float* const* p, const float* const* g, float* const* h,
                              const int64_t* sizes, float lr, float grad_scale);

/* ======================= data parallel ======================= */
int  agb_nccl_unique_id(void* id128);                           /* rank 0: fills 128 bytes */
int  agb_nccl_init(agb_ctx* ctx, int rank, int world, const void* id128);
int  agb_allreduce_sum(agb_ctx* ctx, float* buf, int64_t n);    /* in place, on ctx stream */
/* overlapped form: the sum runs on the context's communication stream after the work enqueued so far; the compute stream continues and
 * agb_allreduce_wait() orders it behind every bucket issued so far (call before the optimizer reads the sums). */
int  agb_allreduce_sum_async(agb_ctx* ctx, float* buf, int64_t n);
int  agb_allreduce_wait(agb_ctx* ctx);
int  agb_nccl_destroy(agb_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* AGB200_H */
