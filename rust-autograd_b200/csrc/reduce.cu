// reduce.cu — reductions over a [outer, r, inner] view: Sum/Mean/Prod/Min/Max and ArgMax/ArgMin.
// HBM roofline: 4 B per input element (reduce_sum algorithmic bytes = 4n).
//
// Reference semantics followed:
//   impl_reduce_forward! (fold_axis, default = zero/one/max_value/min_value)  src/tensor_ops/reduction_ops.rs:54-108
//   ReduceMean: multiply by 1/len with len accumulated in f32                 src/tensor_ops/reduction_ops.rs:187-215
//   ReduceSumToScalar                                                          src/tensor_ops/reduction_ops.rs:123-128
//   MaybeReduceSum (un-broadcast = the same kernel)                            src/tensor_ops/binary_ops.rs:39-94
//   ArgMax/ArgMin: first occurrence of fold(max/min) along the axis            src/tensor_ops/reduction_ops.rs:365-457
// Float::max/min ignore NaN (fmaxf/fminf do too).  Summation order differs from ndarray's sequential
// fold (pairwise/tree here): results agree to fp32 rounding, the oracle accumulates in f64.
#include "common.cuh"
#include <float.h>

template <int OP> __device__ __forceinline__ float r_init() {
  return OP == AGB_R_PROD ? 1.0f : OP == AGB_R_MIN ? FLT_MAX : OP == AGB_R_MAX ? -FLT_MAX : 0.0f;
}
template <int OP> __device__ __forceinline__ float r_comb(float a, float b) {
  return OP == AGB_R_PROD ? a * b : OP == AGB_R_MIN ? fminf(a, b) : OP == AGB_R_MAX ? fmaxf(a, b) : a + b;
}
template <int OP> __device__ __forceinline__ float warp_red(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = r_comb<OP>(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
template <int OP> __device__ __forceinline__ float block_red(float v, float* sm) {
  v = warp_red<OP>(v);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (l == 0) sm[w] = v;
  __syncthreads();
  v = (threadIdx.x < nw) ? sm[threadIdx.x] : r_init<OP>();
  if (w == 0) v = warp_red<OP>(v);
  __syncthreads();
  return v;   // valid in warp 0
}

// ---- rows (inner == 1): x [rows, r].  grid = (chunks, rows): block (c, row) reduces a chunk of the row
//      into part[row*chunks + c] (chunks == 1: writes the result directly, scaled).
template <int OP>
__global__ void __launch_bounds__(256) reduce_rows_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                          int64_t r, int64_t chunk, int chunks, float scale) {
  __shared__ float sm[32];
  int64_t row = blockIdx.y;
  int64_t beg = (int64_t)blockIdx.x * chunk, end = beg + chunk; if (end > r) end = r;
  const float* p = x + row * r;
  float a0 = r_init<OP>(), a1 = r_init<OP>(), a2 = r_init<OP>(), a3 = r_init<OP>();
  // align to 16 B
  int64_t i = beg + threadIdx.x;
  int64_t head = beg; while (head < end && (((uintptr_t)(p + head)) & 15)) head++;
  if (i < head) a0 = r_comb<OP>(a0, p[i]);
  int64_t nv = (end - head) >> 2;
  const float* pv = p + head;
  int64_t j = threadIdx.x;
  for (; j + 3 * 256 < nv; j += 4 * 256) {      // 4 independent 128-bit loads in flight
    float4 v0 = ldg_stream4(pv + 4 * j), v1 = ldg_stream4(pv + 4 * (j + 256));
    float4 v2 = ldg_stream4(pv + 4 * (j + 512)), v3 = ldg_stream4(pv + 4 * (j + 768));
    a0 = r_comb<OP>(a0, r_comb<OP>(r_comb<OP>(v0.x, v0.y), r_comb<OP>(v0.z, v0.w)));
    a1 = r_comb<OP>(a1, r_comb<OP>(r_comb<OP>(v1.x, v1.y), r_comb<OP>(v1.z, v1.w)));
    a2 = r_comb<OP>(a2, r_comb<OP>(r_comb<OP>(v2.x, v2.y), r_comb<OP>(v2.z, v2.w)));
    a3 = r_comb<OP>(a3, r_comb<OP>(r_comb<OP>(v3.x, v3.y), r_comb<OP>(v3.z, v3.w)));
  }
  for (; j < nv; j += 256) {
    float4 v0 = ldg_stream4(pv + 4 * j);
    a0 = r_comb<OP>(a0, r_comb<OP>(r_comb<OP>(v0.x, v0.y), r_comb<OP>(v0.z, v0.w)));
  }
  for (int64_t k = head + (nv << 2) + threadIdx.x; k < end; k += 256) a1 = r_comb<OP>(a1, p[k]);
  float v = r_comb<OP>(r_comb<OP>(a0, a1), r_comb<OP>(a2, a3));
  v = block_red<OP>(v, sm);
  if (threadIdx.x == 0) out[row * chunks + blockIdx.x] = (chunks == 1) ? v * scale : v;
}

// ---- many short rows: one warp per row
template <int OP>
__global__ void __launch_bounds__(256) reduce_rows_warp_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                               int64_t rows, int64_t r, float scale) {
  int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  int l = threadIdx.x & 31;
  const float* p = x + row * r;
  float a = r_init<OP>();
  for (int64_t i = l; i < r; i += 32) a = r_comb<OP>(a, __ldg(p + i));
  a = warp_red<OP>(a);
  if (l == 0) out[row] = a * scale;
}

// ---- columns (inner > 1): x [outer, r, inner] -> out [outer, (splits), inner]; thread per inner element
template <int OP>
__global__ void __launch_bounds__(256) reduce_cols_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                          int64_t r, int64_t inner, int64_t rchunk, int splits, float scale) {
  int64_t col = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (col >= inner) return;
  int64_t o = blockIdx.z, s = blockIdx.y;
  int64_t beg = s * rchunk, end = beg + rchunk; if (end > r) end = r;
  const float* p = x + (o * r) * inner + col;
  float a0 = r_init<OP>(), a1 = r_init<OP>(), a2 = r_init<OP>(), a3 = r_init<OP>();
  int64_t k = beg;
  for (; k + 3 < end; k += 4) {
    a0 = r_comb<OP>(a0, __ldg(p + k * inner));
    a1 = r_comb<OP>(a1, __ldg(p + (k + 1) * inner));
    a2 = r_comb<OP>(a2, __ldg(p + (k + 2) * inner));
    a3 = r_comb<OP>(a3, __ldg(p + (k + 3) * inner));
  }
  for (; k < end; k++) a0 = r_comb<OP>(a0, __ldg(p + k * inner));
  float v = r_comb<OP>(r_comb<OP>(a0, a1), r_comb<OP>(a2, a3));
  out[(o * splits + s) * inner + col] = (splits == 1) ? v * scale : v;
}

// ---- columns with a SHORT inner extent (inner <= 1024, multiple of 4; e.g. bias gradients of channels-last activations,
//      [B*H*W, C] -> [C]): a thread-per-column mapping would leave most of the block idle, so 256 threads tile
//      (1024 / inner) rows x inner columns with 128-bit loads, then fold the row groups through shared memory.
template <int OP>
__global__ void __launch_bounds__(256) reduce_cols_small_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                                int64_t r, int inner, int64_t rchunk, int splits, float scale) {
  __shared__ float4 sm[256];
  const int vec = inner >> 2;                 // float4 per row
  const int rows_per_it = 256 / vec;          // vec divides 256 (inner in {4..1024} power-of-two multiples) or leaves idle threads
  const int tr = threadIdx.x / vec, tc = threadIdx.x - tr * vec;
  const int64_t o = blockIdx.z, s = blockIdx.y;
  int64_t beg = s * rchunk, end = beg + rchunk; if (end > r) end = r;
  const float* p = x + (o * r) * inner;
  float4 a = make_float4(r_init<OP>(), r_init<OP>(), r_init<OP>(), r_init<OP>()), b = a;
  if (tr < rows_per_it) {
    int64_t k = beg + tr;
    for (; k + rows_per_it < end; k += 2 * rows_per_it) {
      float4 v0 = ldg_stream4(p + k * inner + 4 * tc), v1 = ldg_stream4(p + (k + rows_per_it) * inner + 4 * tc);
      a.x = r_comb<OP>(a.x, v0.x); a.y = r_comb<OP>(a.y, v0.y); a.z = r_comb<OP>(a.z, v0.z); a.w = r_comb<OP>(a.w, v0.w);
      b.x = r_comb<OP>(b.x, v1.x); b.y = r_comb<OP>(b.y, v1.y); b.z = r_comb<OP>(b.z, v1.z); b.w = r_comb<OP>(b.w, v1.w);
    }
    for (; k < end; k += rows_per_it) {
      float4 v0 = ldg_stream4(p + k * inner + 4 * tc);
      a.x = r_comb<OP>(a.x, v0.x); a.y = r_comb<OP>(a.y, v0.y); a.z = r_comb<OP>(a.z, v0.z); a.w = r_comb<OP>(a.w, v0.w);
    }
  }
  a.x = r_comb<OP>(a.x, b.x); a.y = r_comb<OP>(a.y, b.y); a.z = r_comb<OP>(a.z, b.z); a.w = r_comb<OP>(a.w, b.w);
  sm[threadIdx.x] = a;
  __syncthreads();
  if (tr == 0 && tc < vec) {
    for (int g = 1; g < rows_per_it; g++) { float4 v = sm[g * vec + tc]; a.x = r_comb<OP>(a.x, v.x); a.y = r_comb<OP>(a.y, v.y); a.z = r_comb<OP>(a.z, v.z); a.w = r_comb<OP>(a.w, v.w); }
    if (splits == 1) { a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale; }
    *(float4*)(out + (o * splits + s) * inner + 4 * tc) = a;
  }
}

template <int OP>
static int reduce_impl(agb_ctx* ctx, const float* x, float* y, int64_t outer, int64_t r, int64_t inner, float scale) {
  if (outer * inner == 0) return AGB_OK;
  if (r == 0) {   // fold over an empty axis yields the identity
    agb_tensor t; t.ptr = y; t.rank = 1; t.shape[0] = outer * inner; t.stride[0] = 1;
    float init = OP == AGB_R_PROD ? 1.0f : OP == AGB_R_MIN ? FLT_MAX : OP == AGB_R_MAX ? -FLT_MAX : 0.0f;
    return agb_fill(ctx, &t, init * scale);
  }
  const int sms = ctx->sm_count;
  if (inner == 1) {
    int64_t rows = outer;
    if (r <= 1024 && rows >= 64) {
      int wpb = 8; int64_t blocks = (rows + wpb - 1) / wpb;
      reduce_rows_warp_kernel<OP><<<(unsigned)blocks, 256, 0, ctx->stream>>>(x, y, rows, r, scale);
      AGB_LAUNCHED(ctx); return AGB_OK;
    }
    if (rows > 65535) {   // grid.y limit: more rows than that already fill the machine with one block per row, launch in slabs
      for (int64_t r0 = 0; r0 < rows; r0 += 65535) {
        int64_t nr = rows - r0 < 65535 ? rows - r0 : 65535;
        reduce_rows_kernel<OP><<<dim3(1, (unsigned)nr), 256, 0, ctx->stream>>>(x + r0 * r, y + r0, r, (r + 3) & ~(int64_t)3, 1, scale);
        AGB_LAUNCHED(ctx);
      }
      return AGB_OK;
    }
    // chunks per row so that the grid covers >= 4 blocks per SM; each chunk >= 4096 elements
    int64_t want = ((int64_t)sms * 4 + rows - 1) / rows;
    int64_t maxc = (r + 4095) / 4096;
    int chunks = (int)(want < maxc ? want : maxc); if (chunks < 1) chunks = 1;
    int64_t chunk = (r + chunks - 1) / chunks; chunk = (chunk + 3) & ~(int64_t)3;
    chunks = (int)((r + chunk - 1) / chunk);
    if (chunks == 1) {
      reduce_rows_kernel<OP><<<dim3(1, (unsigned)rows), 256, 0, ctx->stream>>>(x, y, r, chunk, 1, scale);
      AGB_LAUNCHED(ctx); return AGB_OK;
    }
    float* part; AGB_TRY(agb_scratch(ctx, sizeof(float) * rows * chunks, (void**)&part));
    reduce_rows_kernel<OP><<<dim3(chunks, (unsigned)rows), 256, 0, ctx->stream>>>(x, part, r, chunk, chunks, 1.0f);
    AGB_LAUNCHED(ctx);
    reduce_rows_warp_kernel<OP><<<(unsigned)((rows + 7) / 8), 256, 0, ctx->stream>>>(part, y, rows, chunks, scale);
    AGB_LAUNCHED(ctx); return AGB_OK;
  }
  AGB_CHECK(outer <= 65535, AGB_ERR_UNSUPPORTED, "agb_reduce: outer too large for column reduction (%lld)", (long long)outer);
  int64_t cblocks = (inner + 255) / 256;
  int64_t want = ((int64_t)sms * 4) / (cblocks * outer);
  int64_t maxs = (r + 63) / 64;
  int splits = (int)(want < maxs ? want : maxs); if (splits < 1) splits = 1; if (splits > 1024) splits = 1024;
  int64_t rchunk = (r + splits - 1) / splits; splits = (int)((r + rchunk - 1) / rchunk);
  if (inner <= 1024 && inner % 4 == 0 && r >= 64 && (((uintptr_t)x | (uintptr_t)y) & 15) == 0 && outer <= 65535) {
    int64_t want = (4ll * sms + outer - 1) / outer; int64_t maxs = (r + 255) / 256;
    int sp = (int)(want < maxs ? want : maxs); if (sp < 1) sp = 1; if (sp > 65535) sp = 65535;
    int64_t rc = (r + sp - 1) / sp; sp = (int)((r + rc - 1) / rc);
    if (sp == 1) {
      reduce_cols_small_kernel<OP><<<dim3(1, 1, (unsigned)outer), 256, 0, ctx->stream>>>(x, y, r, (int)inner, rc, 1, scale);
      AGB_LAUNCHED(ctx); return AGB_OK;
    }
    float* part; AGB_TRY(agb_scratch(ctx, sizeof(float) * outer * sp * inner, (void**)&part));
    reduce_cols_small_kernel<OP><<<dim3(1, sp, (unsigned)outer), 256, 0, ctx->stream>>>(x, part, r, (int)inner, rc, sp, 1.0f);
    AGB_LAUNCHED(ctx);
    reduce_cols_kernel<OP><<<dim3((unsigned)((inner + 255) / 256), 1, (unsigned)outer), 256, 0, ctx->stream>>>(part, y, sp, inner, sp, 1, scale);
    AGB_LAUNCHED(ctx);
    return AGB_OK;
  }
  if (splits == 1) {
    reduce_cols_kernel<OP><<<dim3((unsigned)cblocks, 1, (unsigned)outer), 256, 0, ctx->stream>>>(x, y, r, inner, rchunk, 1, scale);
    AGB_LAUNCHED(ctx); return AGB_OK;
  }
  float* part; AGB_TRY(agb_scratch(ctx, sizeof(float) * outer * splits * inner, (void**)&part));
  reduce_cols_kernel<OP><<<dim3((unsigned)cblocks, splits, (unsigned)outer), 256, 0, ctx->stream>>>(x, part, r, inner, rchunk, splits, 1.0f);
  AGB_LAUNCHED(ctx);
  reduce_cols_kernel<OP><<<dim3((unsigned)cblocks, 1, (unsigned)outer), 256, 0, ctx->stream>>>(part, y, splits, inner, splits, 1, scale);
  AGB_LAUNCHED(ctx);
  return AGB_OK;
}

extern "C" int agb_reduce(agb_ctx* ctx, int op, const float* x, float* y, int64_t outer, int64_t r, int64_t inner) {
  switch (op) {
    case AGB_R_SUM: return reduce_impl<AGB_R_SUM>(ctx, x, y, outer, r, inner, 1.0f);
    case AGB_R_MEAN: {
      // reduction_ops.rs:198-209: len is accumulated in f32 and applied as multiply-by-reciprocal
      float len = (float)r;
      return reduce_impl<AGB_R_SUM>(ctx, x, y, outer, r, inner, 1.0f / len);
    }
    case AGB_R_PROD: return reduce_impl<AGB_R_PROD>(ctx, x, y, outer, r, inner, 1.0f);
    case AGB_R_MIN: return reduce_impl<AGB_R_MIN>(ctx, x, y, outer, r, inner, 1.0f);
    case AGB_R_MAX: return reduce_impl<AGB_R_MAX>(ctx, x, y, outer, r, inner, 1.0f);
  }
  agb_set_error("agb_reduce: bad op %d", op); return AGB_ERR_INVALID_DIMS;
}

// ----------------------------------------------------------------------------------------------
// argmax / argmin
// ----------------------------------------------------------------------------------------------
struct ArgPair { float v; int64_t i; };
template <bool MAX> __device__ __forceinline__ ArgPair arg_comb(ArgPair a, ArgPair b) {
  // the extreme wins; on ties the smaller index (first occurrence) wins; unset index = INT64_MAX
  bool take_b = MAX ? (b.v > a.v) : (b.v < a.v);
  if (b.v == a.v && b.i < a.i) take_b = true;
  return take_b ? b : a;
}
template <bool MAX> __device__ __forceinline__ void arg_step(ArgPair& a, float v, int64_t i) {
  // reduction_ops.rs:372-395: m = fold(default, max); mask = first position with x == m
  if (MAX ? (v > a.v) : (v < a.v)) { a.v = v; a.i = i; }
  else if (v == a.v && a.i == INT64_MAX) a.i = i;
}

template <bool MAX>
__global__ void __launch_bounds__(256) arg_rows_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t rows, int64_t r) {
  int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  int l = threadIdx.x & 31;
  const float* p = x + row * r;
  ArgPair a; a.v = MAX ? -FLT_MAX : FLT_MAX; a.i = INT64_MAX;
  for (int64_t i = l; i < r; i += 32) arg_step<MAX>(a, __ldg(p + i), i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ArgPair b; b.v = __shfl_xor_sync(0xffffffffu, a.v, o); b.i = __shfl_xor_sync(0xffffffffu, a.i, o);
    a = arg_comb<MAX>(a, b);
  }
  if (l == 0) y[row] = (a.i == INT64_MAX) ? 0.0f : (float)a.i;
}
template <bool MAX>
__global__ void __launch_bounds__(256) arg_cols_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t r, int64_t inner) {
  int64_t col = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (col >= inner) return;
  int64_t o = blockIdx.y;
  const float* p = x + o * r * inner + col;
  ArgPair a; a.v = MAX ? -FLT_MAX : FLT_MAX; a.i = INT64_MAX;
  for (int64_t k = 0; k < r; k++) arg_step<MAX>(a, __ldg(p + k * inner), k);
  y[o * inner + col] = (a.i == INT64_MAX) ? 0.0f : (float)a.i;
}

extern "C" int agb_argreduce(agb_ctx* ctx, int is_max, const float* x, float* y, int64_t outer, int64_t r, int64_t inner) {
  if (outer * inner == 0) return AGB_OK;
  if (inner == 1) {
    unsigned blocks = (unsigned)((outer + 7) / 8);
    if (is_max) arg_rows_kernel<true><<<blocks, 256, 0, ctx->stream>>>(x, y, outer, r);
    else arg_rows_kernel<false><<<blocks, 256, 0, ctx->stream>>>(x, y, outer, r);
  } else {
    AGB_CHECK(outer <= 65535, AGB_ERR_UNSUPPORTED, "agb_argreduce: outer too large");
    dim3 grid((unsigned)((inner + 255) / 256), (unsigned)outer);
    if (is_max) arg_cols_kernel<true><<<grid, 256, 0, ctx->stream>>>(x, y, r, inner);
    else arg_cols_kernel<false><<<grid, 256, 0, ctx->stream>>>(x, y, r, inner);
  }
  AGB_LAUNCHED(ctx);
  return AGB_OK;
}
