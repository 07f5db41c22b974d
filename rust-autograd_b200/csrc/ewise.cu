// ewise.cu — bandwidth-bound elementwise kernels: unary / binary-with-broadcast / add_n / fill /
// strided copy / dropout.  HBM roofline: unary 8 B/elem, binary 12 B/elem.
// 128-bit vectorised streaming loads/stores on the contiguous fast paths, grid = SMs x 8 grid-stride.
//
// Reference semantics followed:
//   binary arithmetic + broadcasting     src/tensor_ops/binary_ops.rs:147-290,304-347
//   compare/select (0/1 floats)          src/tensor_ops/math_ops.rs:86-184
//   unary math                           src/tensor_ops/math_ops.rs:277-1019
//   Sigmoid/ReLU/Softplus/ELU(+Grad)     src/tensor_ops/activation_ops.rs:113-226
//   Clip/ClipGrad, AddN                  src/tensor_ops/array_ops.rs:503-574
//   SigmoidCrossEntropy                  src/tensor_ops/xent_ops.rs:33-46
//   Dropout (non-inverted)               src/tensor_ops/random_ops.rs:218-237
#include "common.cuh"

// ----------------------------------------------------------------------------------------------
// functors
// ----------------------------------------------------------------------------------------------
// digamma (the `special` crate's Gamma::digamma, math_ops.rs:1029): reflection for x <= 0, recurrence up to x >= 6, then the
// asymptotic series; evaluated in double (the op is bandwidth-bound) and rounded once
__device__ __noinline__ float digamma_apply(float xf) {
  double x = (double)xf, r = 0.0;
  if (x <= 0.0) {
    if (x == floor(x)) return __int_as_float(0x7fc00000);      // poles
    r = -3.14159265358979323846 / tan(3.14159265358979323846 * x); x = 1.0 - x;
  }
  while (x < 6.0) { r -= 1.0 / x; x += 1.0; }
  const double f = 1.0 / (x * x);
  r += log(x) - 0.5 / x - f * (1.0 / 12.0 - f * (1.0 / 120.0 - f * (1.0 / 252.0 - f * (1.0 / 240.0 - f * (1.0 / 132.0)))));
  return (float)r;
}

__device__ __forceinline__ float unary_apply(int op, float x, float p0, float p1) {
  switch (op) {
    case AGB_U_COPY: return x;
    case AGB_U_ABS: return fabsf(x);
    case AGB_U_NEG: return -x;
    case AGB_U_SQUARE: return x * x;
    case AGB_U_INV: return 1.0f / x;
    case AGB_U_INVSQRT: return 1.0f / sqrtf(x);           // math_ops.rs: a.sqrt().recip()
    case AGB_U_SIGN: return (x == 0.0f) ? 0.0f : (x != x ? x : copysignf(1.0f, x));   // math_ops.rs:370-381: 0 -> 0, else signum
    case AGB_U_FLOOR: return floorf(x);
    case AGB_U_CEIL: return ceilf(x);
    case AGB_U_SQRT: return sqrtf(x);
    case AGB_U_POW: return powf(x, p0);
    case AGB_U_LN: return logf(x);
    case AGB_U_LOG2: return log2f(x);
    case AGB_U_LOG10: return log10f(x);
    case AGB_U_EXP: return expf(x);
    case AGB_U_EXP2: return exp2f(x);
    case AGB_U_EXP10: return exp10f(x);
    case AGB_U_SIN: return sinf(x);
    case AGB_U_COS: return cosf(x);
    case AGB_U_TAN: return tanf(x);
    case AGB_U_ASIN: return asinf(x);
    case AGB_U_ACOS: return acosf(x);
    case AGB_U_ATAN: return atanf(x);
    case AGB_U_SINH: return sinhf(x);
    case AGB_U_COSH: return coshf(x);
    case AGB_U_TANH: return tanhf(x);
    case AGB_U_ASINH: return asinhf(x);
    case AGB_U_ACOSH: return acoshf(x);
    case AGB_U_ATANH: return atanhf(x);
    case AGB_U_SIGMOID: return tanhf(x * 0.5f) * 0.5f + 0.5f;        // activation_ops.rs:138-141
    case AGB_U_RELU: return fmaxf(x, 0.0f);                           // Float::max: NaN -> 0
    case AGB_U_SOFTPLUS: return logf(expf(x) + 1.0f);                 // unguarded, as the reference
    case AGB_U_ELU: return x > 0.0f ? x : p0 * (expf(x) - 1.0f);
    case AGB_U_CLIP: return fmaxf(fminf(x, p1), p0);                  // a.min(max).max(min)
    case AGB_U_SCALE: return x * p0;
    case AGB_U_ADD_SCALAR: return x + p0;
    case AGB_U_RSUB_SCALAR: return p0 - x;
    case AGB_U_RDIV_SCALAR: return p0 / x;
    case AGB_U_LGAMMA: return lgammaf(x);                              // ln_gamma().0 = ln|Gamma(x)|
    case AGB_U_DIGAMMA: return digamma_apply(x);
  }
  return x;
}

__device__ __forceinline__ float binary_apply(int op, float a, float b, float p0, float p1) {
  switch (op) {
    case AGB_B_ADD: return a + b;
    case AGB_B_SUB: return a - b;
    case AGB_B_MUL: return a * b;
    case AGB_B_DIV: return a / b;
    case AGB_B_EQ: return a == b ? 1.0f : 0.0f;
    case AGB_B_NE: return a != b ? 1.0f : 0.0f;
    case AGB_B_GT: return a > b ? 1.0f : 0.0f;
    case AGB_B_LT: return a < b ? 1.0f : 0.0f;
    case AGB_B_GE: return a >= b ? 1.0f : 0.0f;
    case AGB_B_LE: return a <= b ? 1.0f : 0.0f;
    case AGB_B_MAX: return a > b ? a : b;                 // math_ops.rs maximum_fn
    case AGB_B_MIN: return a < b ? a : b;
    case AGB_B_ELU_GRAD: return (a > 0.0f ? 1.0f : p0 * (expf(a) - 1.0f) + p0) * b;
    case AGB_B_CLIP_GRAD: return ((a > p0) ? 1.0f : 0.0f) * ((a < p1) ? 1.0f : 0.0f) * b;
    case AGB_B_SIGMOID_XENT: return logf(expf(-fabsf(a)) + 1.0f) + fmaxf(0.0f, a) - b * a;
    case AGB_B_RELU_GRAD: return a > 0.0f ? b : 0.0f * b;
  }
  return 0.0f;
}

// ----------------------------------------------------------------------------------------------
// contiguous unary
// ----------------------------------------------------------------------------------------------
template <int OP>
__global__ void __launch_bounds__(256) unary_contig_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                           int64_t n, float p0, float p1) {
  int64_t n4 = n >> 2;
  int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // 2 independent 128-bit loads in flight per thread per iteration
  int64_t i = tid;
  for (; i + stride < n4; i += 2 * stride) {
    float4 a = ldg_stream4(x + 4 * i);
    float4 b = ldg_stream4(x + 4 * (i + stride));
    a.x = unary_apply(OP, a.x, p0, p1); a.y = unary_apply(OP, a.y, p0, p1);
    a.z = unary_apply(OP, a.z, p0, p1); a.w = unary_apply(OP, a.w, p0, p1);
    b.x = unary_apply(OP, b.x, p0, p1); b.y = unary_apply(OP, b.y, p0, p1);
    b.z = unary_apply(OP, b.z, p0, p1); b.w = unary_apply(OP, b.w, p0, p1);
    stg_stream4(y + 4 * i, a);
    stg_stream4(y + 4 * (i + stride), b);
  }
  for (; i < n4; i += stride) {
    float4 a = ldg_stream4(x + 4 * i);
    a.x = unary_apply(OP, a.x, p0, p1); a.y = unary_apply(OP, a.y, p0, p1);
    a.z = unary_apply(OP, a.z, p0, p1); a.w = unary_apply(OP, a.w, p0, p1);
    stg_stream4(y + 4 * i, a);
  }
  for (int64_t j = (n4 << 2) + tid; j < n; j += stride) y[j] = unary_apply(OP, x[j], p0, p1);
}

template <int OP>
__global__ void __launch_bounds__(256) unary_scalar_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                           int64_t n, float p0, float p1) {
  int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = tid; j < n; j += stride) y[j] = unary_apply(OP, x[j], p0, p1);
}

typedef void (*unary_fn)(const float*, float*, int64_t, float, float);
template <int... I> struct useq {};
template <int N, int... I> struct make_useq : make_useq<N - 1, N - 1, I...> {};
template <int... I> struct make_useq<0, I...> { typedef useq<I...> type; };
template <int... I> static const unary_fn* unary_table_v(useq<I...>) { static const unary_fn t[] = {unary_contig_kernel<I>...}; return t; }
template <int... I> static const unary_fn* unary_table_s(useq<I...>) { static const unary_fn t[] = {unary_scalar_kernel<I>...}; return t; }

// ----------------------------------------------------------------------------------------------
// generic strided binary (broadcast by zero strides); output contiguous (or strided for copy)
// ----------------------------------------------------------------------------------------------
#define AGB_EW_MAXD 6
struct StridedParams {
  int rank;                      // collapsed rank (>=1)
  int64_t shape[AGB_EW_MAXD];    // collapsed shape; innermost is dim rank-1 (already divided by 4 for vec)
  int64_t sa[AGB_EW_MAXD], sb[AGB_EW_MAXD], sy[AGB_EW_MAXD];
};

template <int OP, bool VEC, typename IDX>
__global__ void __launch_bounds__(256) binary_strided_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                             float* __restrict__ y, StridedParams P, int64_t total,
                                                             float p0, float p1) {
  IDX tid = (IDX)blockIdx.x * blockDim.x + threadIdx.x;
  IDX stride = (IDX)gridDim.x * blockDim.x;
  for (IDX i = tid; i < (IDX)total; i += stride) {
    IDX rem = i; int64_t oa = 0, ob = 0, oy = 0;
#pragma unroll
    for (int d = AGB_EW_MAXD - 1; d >= 0; d--) {
      if (d < P.rank) {
        IDX q = (d == 0) ? 0 : rem / (IDX)P.shape[d];
        IDX c = (d == 0) ? rem : rem - q * (IDX)P.shape[d];
        rem = q;
        if (VEC && d == P.rank - 1) { oa += (int64_t)c * 4 * P.sa[d]; ob += (int64_t)c * 4 * P.sb[d]; oy += (int64_t)c * 4 * P.sy[d]; }
        else { oa += (int64_t)c * P.sa[d]; ob += (int64_t)c * P.sb[d]; oy += (int64_t)c * P.sy[d]; }
      }
    }
    if (VEC) {
      const int last = P.rank - 1;
      float4 va, vb, r;
      if (P.sa[last] == 1) va = ldg_stream4(a + oa); else { float s = __ldg(a + oa); va = make_float4(s, s, s, s); }
      if (OP == -1) { vb = va; }
      else if (P.sb[last] == 1) vb = ldg_stream4(b + ob); else { float s = __ldg(b + ob); vb = make_float4(s, s, s, s); }
      if (OP == -1) r = va;
      else {
        r.x = binary_apply(OP, va.x, vb.x, p0, p1); r.y = binary_apply(OP, va.y, vb.y, p0, p1);
        r.z = binary_apply(OP, va.z, vb.z, p0, p1); r.w = binary_apply(OP, va.w, vb.w, p0, p1);
      }
      stg_stream4(y + oy, r);
    } else {
      float va = __ldg(a + oa);
      if (OP == -1) y[oy] = va;
      else y[oy] = binary_apply(OP, va, __ldg(b + ob), p0, p1);
    }
  }
}

// ---- row / channel broadcast fast paths (bias add and friends): y[r, c] = f(a[r, c], b[c]) on contiguous a, y.
//      CSTRIDE == 1: b varies with the innermost index (channels-last bias, [rows, C] + [C]);
//      CSTRIDE  > 1: b varies with the middle index of [outer, C, inner] (NCHW bias [1,C,1,1]); inner % 4 == 0.
template <int OP, bool SWAP>
__global__ void __launch_bounds__(256) binary_bias_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y,
                                                          int64_t n4, int C, int inner4, float p0, float p1) {
  int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = tid; i < n4; i += stride) {
    float4 va = ldg_stream4(a + 4 * i), vb, r;
    if (inner4 == 0) { int c = (int)((4 * i) % C); vb = __ldg((const float4*)(b + c)); }
    else { float s = __ldg(b + (int)((i / inner4) % C)); vb = make_float4(s, s, s, s); }
    if (!SWAP) { r.x = binary_apply(OP, va.x, vb.x, p0, p1); r.y = binary_apply(OP, va.y, vb.y, p0, p1); r.z = binary_apply(OP, va.z, vb.z, p0, p1); r.w = binary_apply(OP, va.w, vb.w, p0, p1); }
    else { r.x = binary_apply(OP, vb.x, va.x, p0, p1); r.y = binary_apply(OP, vb.y, va.y, p0, p1); r.z = binary_apply(OP, vb.z, va.z, p0, p1); r.w = binary_apply(OP, vb.w, va.w, p0, p1); }
    stg_stream4(y + 4 * i, r);
  }
}

// collapse adjacent dims whenever all three stride sets allow it; drop size-1 dims
static void collapse(int rank, const int64_t* shape, const int64_t* sa, const int64_t* sb, const int64_t* sy, StridedParams& P) {
  int64_t sh[AGB_MAX_RANK], a[AGB_MAX_RANK], b[AGB_MAX_RANK], y[AGB_MAX_RANK]; int r = 0;
  for (int i = 0; i < rank; i++) {
    if (shape[i] == 1) continue;
    if (r > 0 && a[r - 1] == shape[i] * sa[i] && b[r - 1] == shape[i] * sb[i] && y[r - 1] == shape[i] * sy[i]) {
      sh[r - 1] *= shape[i]; a[r - 1] = sa[i]; b[r - 1] = sb[i]; y[r - 1] = sy[i];
    } else { sh[r] = shape[i]; a[r] = sa[i]; b[r] = sb[i]; y[r] = sy[i]; r++; }
  }
  if (r == 0) { sh[0] = 1; a[0] = b[0] = y[0] = 1; r = 1; }
  P.rank = r;
  for (int i = 0; i < r; i++) { P.shape[i] = sh[i]; P.sa[i] = a[i]; P.sb[i] = b[i]; P.sy[i] = y[i]; }
}

template <int OP>
static int launch_strided(agb_ctx* ctx, const float* a, const float* b, float* y, int rank, const int64_t* shape,
                          const int64_t* sa, const int64_t* sb, const int64_t* sy, float p0, float p1) {
  StridedParams P; collapse(rank, shape, sa, sb, sy, P);
  AGB_CHECK(P.rank <= AGB_EW_MAXD, AGB_ERR_UNSUPPORTED, "elementwise: more than %d non-collapsible dims", AGB_EW_MAXD);
  int64_t total = 1; for (int i = 0; i < P.rank; i++) total *= P.shape[i];
  if (total == 0) return AGB_OK;
  const int last = P.rank - 1;
  auto al16 = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
  if (OP >= 0 && total < (1ll << 40) && al16(a) && al16(b) && al16(y)) {
    // bias patterns: one operand full and contiguous, the other a [C] vector broadcast along rows (rank 2) or along
    // [outer, C, inner] (rank 3); output contiguous
    for (int swap = 0; swap < 2; swap++) {
      const int64_t* sf = swap ? P.sb : P.sa; const int64_t* sv = swap ? P.sa : P.sb;
      const float* full = swap ? b : a; const float* vecp = swap ? a : b;
      if (P.rank == 2 && sf[0] == P.shape[1] && sf[1] == 1 && P.sy[0] == P.shape[1] && P.sy[1] == 1 && sv[0] == 0 && sv[1] == 1 && P.shape[1] % 4 == 0 && P.shape[1] < (1ll << 30)) {
        int64_t n4 = total / 4; int grid = agb_grid_for(n4, 256, ctx->sm_count, 8);
        if (swap) binary_bias_kernel<OP, true><<<grid, 256, 0, ctx->stream>>>(full, vecp, y, n4, (int)P.shape[1], 0, p0, p1);
        else binary_bias_kernel<OP, false><<<grid, 256, 0, ctx->stream>>>(full, vecp, y, n4, (int)P.shape[1], 0, p0, p1);
        AGB_LAUNCHED(ctx); return AGB_OK;
      }
      if (P.rank == 3 && sf[2] == 1 && sf[1] == P.shape[2] && sf[0] == P.shape[1] * P.shape[2] && P.sy[2] == 1 && P.sy[1] == P.shape[2] && P.sy[0] == P.shape[1] * P.shape[2] &&
          sv[0] == 0 && sv[1] == 1 && sv[2] == 0 && P.shape[2] % 4 == 0 && P.shape[1] < (1ll << 30) && P.shape[2] < (1ll << 32)) {
        int64_t n4 = total / 4; int grid = agb_grid_for(n4, 256, ctx->sm_count, 8);
        if (swap) binary_bias_kernel<OP, true><<<grid, 256, 0, ctx->stream>>>(full, vecp, y, n4, (int)P.shape[1], (int)(P.shape[2] / 4), p0, p1);
        else binary_bias_kernel<OP, false><<<grid, 256, 0, ctx->stream>>>(full, vecp, y, n4, (int)P.shape[1], (int)(P.shape[2] / 4), p0, p1);
        AGB_LAUNCHED(ctx); return AGB_OK;
      }
    }
  }
  bool vec = (P.shape[last] % 4 == 0) && (P.sa[last] == 1 || P.sa[last] == 0) && (P.sb[last] == 1 || P.sb[last] == 0) &&
             P.sy[last] == 1 && al16(y) && (P.sa[last] == 0 || al16(a)) && (P.sb[last] == 0 || al16(b));
  for (int i = 0; i < last && vec; i++) {
    if (P.sa[last] == 1 && P.sa[i] % 4) vec = false;
    if (P.sb[last] == 1 && P.sb[i] % 4) vec = false;
    if (P.sy[i] % 4) vec = false;
  }
  if (vec) { P.shape[last] /= 4; total /= 4; }
  int grid = agb_grid_for(total, 256, ctx->sm_count, 8);
  bool small = total * (vec ? 4 : 1) < (1ll << 31) ;
  for (int i = 0; i < P.rank && small; i++) {
    if (llabs(P.sa[i]) * P.shape[i] >= (1ll << 31) || llabs(P.sb[i]) * P.shape[i] >= (1ll << 31)) small = false;
  }
  if (vec) {
    if (small) binary_strided_kernel<OP, true, uint32_t><<<grid, 256, 0, ctx->stream>>>(a, b, y, P, total, p0, p1);
    else binary_strided_kernel<OP, true, int64_t><<<grid, 256, 0, ctx->stream>>>(a, b, y, P, total, p0, p1);
  } else {
    if (small) binary_strided_kernel<OP, false, uint32_t><<<grid, 256, 0, ctx->stream>>>(a, b, y, P, total, p0, p1);
    else binary_strided_kernel<OP, false, int64_t><<<grid, 256, 0, ctx->stream>>>(a, b, y, P, total, p0, p1);
  }
  AGB_LAUNCHED(ctx);
  return AGB_OK;
}

typedef int (*strided_fn)(agb_ctx*, const float*, const float*, float*, int, const int64_t*, const int64_t*, const int64_t*, const int64_t*, float, float);
template <int... I> static const strided_fn* binary_table(useq<I...>) { static const strided_fn t[] = {launch_strided<I>...}; return t; }

static void contig_strides(int rank, const int64_t* shape, int64_t* st) {
  int64_t s = 1; for (int i = rank - 1; i >= 0; i--) { st[i] = s; s *= shape[i]; }
}

extern "C" int agb_unary(agb_ctx* ctx, int op, float p0, float p1, const agb_tensor* x, agb_tensor* y) {
  AGB_CHECK(op >= 0 && op < AGB_U_COUNT, AGB_ERR_INVALID_DIMS, "agb_unary: bad op %d", op);
  AGB_CHECK(x->rank == y->rank, AGB_ERR_INCOMPATIBLE_SHAPE, "agb_unary: rank mismatch");
  for (int i = 0; i < x->rank; i++) AGB_CHECK(x->shape[i] == y->shape[i], AGB_ERR_INCOMPATIBLE_SHAPE, "agb_unary: shape mismatch on axis %d", i);
  AGB_CHECK(agb_is_contig(y), AGB_ERR_UNSUPPORTED, "agb_unary: output must be C-contiguous");
  int64_t n = agb_numel(x);
  if (n == 0) return AGB_OK;
  AgbProfScope prof(ctx, AGB_PROF_EWISE, 8.0 * (double)n);
  const float* xp = x->ptr; float* tmp = nullptr;
  if (!agb_is_contig(x)) {       // materialise the view first (reference: ndarray map over a strided view)
    if (op == AGB_U_COPY) return agb_copy_strided(ctx, x, y);
    agb_tensor t = *y; AGB_TRY(agb_alloc(ctx, n * sizeof(float), (void**)&tmp)); t.ptr = tmp;
    AGB_TRY(agb_copy_strided(ctx, x, &t)); xp = tmp;
  }
  bool aligned = (((uintptr_t)xp | (uintptr_t)y->ptr) & 15) == 0;
  const unary_fn* tv = unary_table_v(make_useq<AGB_U_COUNT>::type());
  const unary_fn* ts = unary_table_s(make_useq<AGB_U_COUNT>::type());
  if (aligned) tv[op]<<<agb_grid_occ(ctx, tv[op], (n + 3) / 4, 256), 256, 0, ctx->stream>>>(xp, y->ptr, n, p0, p1);
  else ts[op]<<<agb_grid_for(n, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(xp, y->ptr, n, p0, p1);
  AGB_LAUNCHED(ctx);
  if (tmp) AGB_TRY(agb_free(ctx, tmp));
  return AGB_OK;
}

extern "C" int agb_binary(agb_ctx* ctx, int op, float p0, float p1, const agb_tensor* a, const agb_tensor* b, agb_tensor* y) {
  AGB_CHECK(op >= 0 && op < AGB_B_COUNT, AGB_ERR_INVALID_DIMS, "agb_binary: bad op %d", op);
  AGB_CHECK(a->rank == y->rank && b->rank == y->rank, AGB_ERR_INCOMPATIBLE_SHAPE, "agb_binary: operands must be pre-broadcast to the output rank");
  for (int i = 0; i < y->rank; i++)
    AGB_CHECK(a->shape[i] == y->shape[i] && b->shape[i] == y->shape[i], AGB_ERR_INCOMPATIBLE_SHAPE,
              "agb_binary: operands must be pre-broadcast to the output shape (axis %d)", i);
  AGB_CHECK(agb_is_contig(y), AGB_ERR_UNSUPPORTED, "agb_binary: output must be C-contiguous");
  int64_t sy[AGB_MAX_RANK]; contig_strides(y->rank, y->shape, sy);
  AgbProfScope prof(ctx, AGB_PROF_EWISE, 12.0 * (double)agb_numel(y));
  const strided_fn* t = binary_table(make_useq<AGB_B_COUNT>::type());
  return t[op](ctx, a->ptr, b->ptr, y->ptr, y->rank, y->shape, a->stride, b->stride, sy, p0, p1);
}

// batched 2-D transpose through shared memory: src [B][R][S] (S contiguous) -> dst [B][S][R] (R contiguous).  Both sides are
// read / written in 128-byte rows; this is the NCHW <-> channels-last layout change (R = C, S = H*W or the reverse).
__global__ void __launch_bounds__(256) transpose_batched_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t R, int64_t S) {
  __shared__ float tile[32][33];
  const int64_t b = blockIdx.z; const int64_t s0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
  const float* sp = src + b * R * S; float* dp = dst + b * R * S;
#pragma unroll
  for (int k = 0; k < 32; k += 8) { int64_t r = r0 + ty + k, s = s0 + tx; if (r < R && s < S) tile[ty + k][tx] = __ldg(sp + r * S + s); }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 32; k += 8) { int64_t s = s0 + ty + k, r = r0 + tx; if (r < R && s < S) dp[s * R + r] = tile[tx][ty + k]; }
}
static int launch_transpose(agb_ctx* ctx, const float* src, float* dst, int64_t B, int64_t R, int64_t S) {
  for (int64_t b0 = 0; b0 < B; b0 += 65535) {
    int64_t nb = B - b0 < 65535 ? B - b0 : 65535;
    dim3 grid((unsigned)((S + 31) / 32), (unsigned)((R + 31) / 32), (unsigned)nb);
    AGB_CHECK(grid.y <= 65535, AGB_ERR_UNSUPPORTED, "transpose: too many row tiles");
    transpose_batched_kernel<<<grid, 256, 0, ctx->stream>>>(src + b0 * R * S, dst + b0 * R * S, R, S);
    AGB_LAUNCHED(ctx);
  }
  return AGB_OK;
}

extern "C" int agb_copy_strided(agb_ctx* ctx, const agb_tensor* src, agb_tensor* dst) {
  AGB_CHECK(src->rank == dst->rank, AGB_ERR_INCOMPATIBLE_SHAPE, "agb_copy_strided: rank mismatch %d vs %d", src->rank, dst->rank);
  for (int i = 0; i < src->rank; i++) AGB_CHECK(src->shape[i] == dst->shape[i], AGB_ERR_INCOMPATIBLE_SHAPE, "agb_copy_strided: shape mismatch on axis %d", i);
  if (agb_is_contig(src) && agb_is_contig(dst)) return agb_d2d(ctx, dst->ptr, src->ptr, agb_numel(src) * sizeof(float));
  {   // [B][R][S] <-> [B][S][R] pattern (layout changes): tiled transpose instead of the generic gather
    StridedParams P; collapse(src->rank, src->shape, src->stride, src->stride, dst->stride, P);
    if (P.rank == 2 || P.rank == 3) {
      const int o = P.rank - 2; const int64_t B = o ? P.shape[0] : 1, d1 = P.shape[o], d2 = P.shape[o + 1];
      const bool bs_ok = !o || (P.sa[0] == d1 * d2 && P.sy[0] == d1 * d2);
      if (bs_ok && d1 >= 8 && d2 >= 8) {
        if (P.sa[o] == d2 && P.sa[o + 1] == 1 && P.sy[o] == 1 && P.sy[o + 1] == d1) return launch_transpose(ctx, src->ptr, dst->ptr, B, d1, d2);      // src row-major, dst transposed
        if (P.sa[o] == 1 && P.sa[o + 1] == d1 && P.sy[o] == d2 && P.sy[o + 1] == 1) return launch_transpose(ctx, src->ptr, dst->ptr, B, d2, d1);      // src transposed, dst row-major
      }
    }
  }
  return launch_strided<-1>(ctx, src->ptr, src->ptr, dst->ptr, src->rank, src->shape, src->stride, src->stride, dst->stride, 0.f, 0.f);
}

// ----------------------------------------------------------------------------------------------
// add_n, fill
// ----------------------------------------------------------------------------------------------
#define AGB_ADDN_MAX 8
struct AddNPtrs { const float* p[AGB_ADDN_MAX]; };
__global__ void __launch_bounds__(256) add_n_kernel(AddNPtrs P, int n, float* __restrict__ y, int64_t numel, int accumulate) {
  int64_t n4 = numel >> 2;
  int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = tid; i < n4; i += stride) {
    // all operand loads are issued before the first add (a runtime-length loop would serialise n memory round trips)
    float4 v[AGB_ADDN_MAX];
#pragma unroll
    for (int k = 0; k < AGB_ADDN_MAX; k++) if (k < n) v[k] = ldg_stream4(P.p[k] + 4 * i);
    float4 acc = accumulate ? *(const float4*)(y + 4 * i) : v[0];
#pragma unroll
    for (int k = 0; k < AGB_ADDN_MAX; k++) {
      if (k < n && (accumulate || k > 0)) { acc.x += v[k].x; acc.y += v[k].y; acc.z += v[k].z; acc.w += v[k].w; }   // left fold, like `base += &ctx.input(i)`
    }
    *(float4*)(y + 4 * i) = acc;
  }
  for (int64_t j = (n4 << 2) + tid; j < numel; j += stride) {
    float acc = accumulate ? y[j] : P.p[0][j];
    for (int k = accumulate ? 0 : 1; k < n; k++) acc += P.p[k][j];
    y[j] = acc;
  }
}

extern "C" int agb_add_n(agb_ctx* ctx, int n, const agb_tensor* const* xs, agb_tensor* y) {
  AGB_CHECK(n >= 1, AGB_ERR_INVALID_DIMS, "agb_add_n: n must be >= 1");
  int64_t numel = agb_numel(y);
  AGB_CHECK(agb_is_contig(y), AGB_ERR_UNSUPPORTED, "agb_add_n: output must be contiguous");
  for (int i = 0; i < n; i++) {
    AGB_CHECK(agb_numel(xs[i]) == numel && agb_is_contig(xs[i]), AGB_ERR_INCOMPATIBLE_SHAPE, "agb_add_n: input %d must be contiguous with the output's size", i);
    AGB_CHECK((((uintptr_t)xs[i]->ptr) & 15) == 0, AGB_ERR_UNSUPPORTED, "agb_add_n: input %d not 16B aligned", i);
  }
  if (numel == 0) return AGB_OK;
  int grid = agb_grid_for((numel + 3) / 4, 256, ctx->sm_count, 8);
  for (int base = 0; base < n; base += AGB_ADDN_MAX) {
    AddNPtrs P; int m = n - base < AGB_ADDN_MAX ? n - base : AGB_ADDN_MAX;
    for (int i = 0; i < m; i++) P.p[i] = xs[base + i]->ptr;
    add_n_kernel<<<grid, 256, 0, ctx->stream>>>(P, m, y->ptr, numel, base > 0);
    AGB_LAUNCHED(ctx);
  }
  return AGB_OK;
}

// ----------------------------------------------------------------------------------------------
// stack n row blocks into one matrix (pointer table in the kernel parameters, blockIdx.y = block)
// ----------------------------------------------------------------------------------------------
#define AGB_CONCAT_MAX 64
struct ConcatParams { const float* src[AGB_CONCAT_MAX]; int64_t pitch[AGB_CONCAT_MAX]; };
__global__ void __launch_bounds__(256) concat_rows_kernel(const __grid_constant__ ConcatParams P, int64_t rows, int64_t cols4, float* __restrict__ dst) {
  const int s = blockIdx.y;
  const float* __restrict__ src = P.src[s]; const int64_t pitch = P.pitch[s];
  float* d = dst + (int64_t)s * rows * cols4 * 4;
  const int64_t total = rows * cols4, stride = (int64_t)gridDim.x * 256;
  for (int64_t i = blockIdx.x * (int64_t)256 + threadIdx.x; i < total; i += stride) {
    const int64_t r = i / cols4, c = i - r * cols4;
    *reinterpret_cast<float4*>(d + 4 * i) = ldg_stream4(src + r * pitch + 4 * c);      // plain store: the GEMM that follows reads it from L2
  }
}
extern "C" int agb_concat_rows(agb_ctx* ctx, int n, const float* const* srcs, const int64_t* src_pitch, int64_t rows, int64_t cols, float* dst) {
  AGB_CHECK(n >= 1 && rows >= 0 && cols >= 0, AGB_ERR_INVALID_DIMS, "agb_concat_rows: bad extents");
  AGB_CHECK(cols % 4 == 0 && (((uintptr_t)dst) & 15) == 0, AGB_ERR_UNSUPPORTED, "agb_concat_rows: cols must be a multiple of 4 and dst 16-byte aligned");
  for (int i = 0; i < n; i++)
    AGB_CHECK(srcs[i] != nullptr && (((uintptr_t)srcs[i]) & 15) == 0 && src_pitch[i] % 4 == 0, AGB_ERR_UNSUPPORTED, "agb_concat_rows: block %d is not 16-byte aligned / pitch %% 4 != 0", i);
  if (rows * cols == 0) return AGB_OK;
  AgbProfScope prof(ctx, AGB_PROF_EWISE, 8.0 * (double)n * rows * cols);
  for (int base = 0; base < n; base += AGB_CONCAT_MAX) {
    const int m = n - base < AGB_CONCAT_MAX ? n - base : AGB_CONCAT_MAX;
    ConcatParams P; memset(&P, 0, sizeof(P));
    for (int i = 0; i < m; i++) { P.src[i] = srcs[base + i]; P.pitch[i] = src_pitch[base + i]; }
    int gx = agb_grid_for(rows * cols / 4, 256, ctx->sm_count, 8); gx = (gx + m - 1) / m; if (gx < 1) gx = 1;
    concat_rows_kernel<<<dim3(gx, m), 256, 0, ctx->stream>>>(P, rows, cols / 4, dst + (int64_t)base * rows * cols);
    AGB_LAUNCHED(ctx);
  }
  return AGB_OK;
}

// ----------------------------------------------------------------------------------------------
// fused elementwise program: a small register machine per element.  The host (engine/fuse.cc) compiles a
// DAG of deferred unary / binary ops into <= 48 instructions over <= 32 registers; every thread keeps its register file in its
// own shared-memory column (conflict-free, no synchronisation), loads all leaves up front (independent loads in flight), runs the
// program with the same functors as the single-op kernels and stores the requested registers.  One launch replaces the whole
// chain, and sliced / broadcast operands are read in place (no deep copy first).
// ----------------------------------------------------------------------------------------------
struct FuseParams {
  int n_leaves, n_instr, n_out;
  int64_t cols, total;
  const float* lptr[AGB_FUSE_MAX_LEAVES]; int64_t lpitch[AGB_FUSE_MAX_LEAVES], lcs[AGB_FUSE_MAX_LEAVES];
  float* optr[AGB_FUSE_MAX_OUT]; int64_t opitch[AGB_FUSE_MAX_OUT];
  uint32_t code[AGB_FUSE_MAX_INSTR];      // kind (2 bits) | op (6) | dst (5) | a (5) | b (5)
  float imm[AGB_FUSE_MAX_INSTR];
  uint8_t lreg[AGB_FUSE_MAX_LEAVES], oreg[AGB_FUSE_MAX_OUT];
};

__global__ void __launch_bounds__(256, 4) fused_ewise_kernel(const __grid_constant__ FuseParams P) {
  __shared__ float R[AGB_FUSE_REGS][256];
  const int t = threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * 256;
  for (int64_t i = blockIdx.x * (int64_t)256 + t; i < P.total; i += stride) {
    int64_t r, c;
    if (P.cols == P.total) { r = 0; c = i; }
    else if (P.total < (int64_t)0x7fffffff) { r = (uint32_t)i / (uint32_t)P.cols; c = i - r * P.cols; }
    else { r = i / P.cols; c = i - r * P.cols; }
    float v[AGB_FUSE_MAX_LEAVES];
#pragma unroll
    for (int l = 0; l < AGB_FUSE_MAX_LEAVES; l++) if (l < P.n_leaves) v[l] = __ldg(P.lptr[l] + r * P.lpitch[l] + c * P.lcs[l]);
#pragma unroll
    for (int l = 0; l < AGB_FUSE_MAX_LEAVES; l++) if (l < P.n_leaves) R[P.lreg[l]][t] = v[l];
    float prev = 0.0f; int prev_dst = -1;          // the previous result is forwarded in a register: a dependent chain does not wait for its own store
    for (int k = 0; k < P.n_instr; k++) {
      const uint32_t w = P.code[k];
      const int kind = w & 3, op = (w >> 2) & 63, ia = (w >> 13) & 31, ib = (w >> 18) & 31;
      const float p0 = P.imm[k];
      float a, b, y;
      if (ia == prev_dst) a = prev; else a = R[ia][t];                 // (warp-uniform branches)
      if (kind != AGB_F_BINARY && kind != AGB_F_BINARY_IMM_A) b = 0.0f; else if (ib == prev_dst) b = prev; else b = R[ib][t];
      if (kind == AGB_F_UNARY) y = unary_apply(op, a, p0, 0.0f);
      else {
        if (kind == AGB_F_BINARY_IMM_B) b = p0; else if (kind == AGB_F_BINARY_IMM_A) a = p0;
        y = binary_apply(op, a, b, 0.0f, 0.0f);
      }
      prev_dst = (w >> 8) & 31; prev = y;
      R[prev_dst][t] = y;
    }
    for (int o = 0; o < P.n_out; o++) P.optr[o][r * P.opitch[o] + c] = R[P.oreg[o]][t];
  }
}

// The same machine with E elements per thread: the program is decoded and dispatched ONCE per instruction for E values, which is what
// bounds the one-element kernel on mid-sized tensors (issue-bound: ~40 SASS instructions of decode + dispatch per interpreted instruction
// against 1 for an add).  Register file = dynamic shared memory [nreg][E][256]; leaves are fetched four at a time (4 * E loads in flight).
#define AGB_FUSE_UNARY_OPS(X) X(AGB_U_COPY) X(AGB_U_ABS) X(AGB_U_NEG) X(AGB_U_SQUARE) X(AGB_U_INV) X(AGB_U_INVSQRT) X(AGB_U_SIGN) X(AGB_U_FLOOR) \
  X(AGB_U_CEIL) X(AGB_U_SQRT) X(AGB_U_POW) X(AGB_U_LN) X(AGB_U_LOG2) X(AGB_U_LOG10) X(AGB_U_EXP) X(AGB_U_EXP2) X(AGB_U_EXP10) X(AGB_U_SIN) \
  X(AGB_U_COS) X(AGB_U_TAN) X(AGB_U_ASIN) X(AGB_U_ACOS) X(AGB_U_ATAN) X(AGB_U_SINH) X(AGB_U_COSH) X(AGB_U_TANH) X(AGB_U_ASINH) X(AGB_U_ACOSH) \
  X(AGB_U_ATANH) X(AGB_U_SIGMOID) X(AGB_U_RELU) X(AGB_U_SOFTPLUS) X(AGB_U_ELU) X(AGB_U_SCALE) X(AGB_U_ADD_SCALAR) X(AGB_U_RSUB_SCALAR) \
  X(AGB_U_RDIV_SCALAR) X(AGB_U_LGAMMA) X(AGB_U_DIGAMMA)
#define AGB_FUSE_BINARY_OPS(X) X(AGB_B_ADD) X(AGB_B_SUB) X(AGB_B_MUL) X(AGB_B_DIV) X(AGB_B_EQ) X(AGB_B_NE) X(AGB_B_GT) X(AGB_B_LT) X(AGB_B_GE) \
  X(AGB_B_LE) X(AGB_B_MAX) X(AGB_B_MIN)

template <int E>
__global__ void __launch_bounds__(256) fused_ewise_kernel_e(const __grid_constant__ FuseParams P) {
  extern __shared__ float Rf[];                        // [nreg][E][256]
  const int t = threadIdx.x;
  const int64_t chunk = 256 * E;
#define RF(reg, j) Rf[((reg) * E + (j)) * 256 + t]
  for (int64_t base = blockIdx.x * chunk; base < P.total; base += (int64_t)gridDim.x * chunk) {
    uint32_t r[E], c[E]; bool ok[E];
#pragma unroll
    for (int j = 0; j < E; j++) {
      const int64_t i = base + j * 256 + t;
      ok[j] = i < P.total;
      const uint32_t ii = ok[j] ? (uint32_t)i : 0u;                     // out-of-range slots read element 0 and store nothing (total < 2^31 on this path)
      if (P.cols == P.total) { r[j] = 0; c[j] = ii; } else { r[j] = ii / (uint32_t)P.cols; c[j] = ii - r[j] * (uint32_t)P.cols; }
    }
    for (int l0 = 0; l0 < P.n_leaves; l0 += 4) {
      float v[4][E];
#pragma unroll
      for (int q = 0; q < 4; q++) if (l0 + q < P.n_leaves) {
        const float* p = P.lptr[l0 + q]; const int64_t pitch = P.lpitch[l0 + q], cs = P.lcs[l0 + q];
#pragma unroll
        for (int j = 0; j < E; j++) v[q][j] = __ldg(p + (int64_t)r[j] * pitch + (int64_t)c[j] * cs);
      }
#pragma unroll
      for (int q = 0; q < 4; q++) if (l0 + q < P.n_leaves) {
        const int reg = P.lreg[l0 + q];
#pragma unroll
        for (int j = 0; j < E; j++) RF(reg, j) = v[q][j];
      }
    }
    float prev[E]; int prev_dst = -1;
#pragma unroll
    for (int j = 0; j < E; j++) prev[j] = 0.0f;
    uint32_t w_next = P.code[0]; float p_next = P.imm[0];
    for (int k = 0; k < P.n_instr; k++) {
      const uint32_t w = w_next; const float p0 = p_next;
      if (k + 1 < P.n_instr) { w_next = P.code[k + 1]; p_next = P.imm[k + 1]; }        // the next instruction word is fetched under this instruction
      const int kind = w & 3, op = (w >> 2) & 63, ia = (w >> 13) & 31, ib = (w >> 18) & 31, dst = (w >> 8) & 31;
      float a[E], b[E], y[E];
      if (kind == AGB_F_BINARY_IMM_A) {
#pragma unroll
        for (int j = 0; j < E; j++) a[j] = p0;
      } else if (ia == prev_dst) {
#pragma unroll
        for (int j = 0; j < E; j++) a[j] = prev[j];
      } else {
#pragma unroll
        for (int j = 0; j < E; j++) a[j] = RF(ia, j);
      }
      if (kind == AGB_F_BINARY || kind == AGB_F_BINARY_IMM_A) {
        if (ib == prev_dst) {
#pragma unroll
          for (int j = 0; j < E; j++) b[j] = prev[j];
        } else {
#pragma unroll
          for (int j = 0; j < E; j++) b[j] = RF(ib, j);
        }
      } else {
#pragma unroll
        for (int j = 0; j < E; j++) b[j] = p0;                          // AGB_F_BINARY_IMM_B (unused by unary ops)
      }
      if (kind == AGB_F_UNARY) {
        switch (op) {
#define X(OP) case OP: _Pragma("unroll") for (int j = 0; j < E; j++) y[j] = unary_apply(OP, a[j], p0, 0.0f); break;
          AGB_FUSE_UNARY_OPS(X)
#undef X
          default: _Pragma("unroll") for (int j = 0; j < E; j++) y[j] = a[j]; break;
        }
      } else {
        switch (op) {
#define X(OP) case OP: _Pragma("unroll") for (int j = 0; j < E; j++) y[j] = binary_apply(OP, a[j], b[j], 0.0f, 0.0f); break;
          AGB_FUSE_BINARY_OPS(X)
#undef X
          default: _Pragma("unroll") for (int j = 0; j < E; j++) y[j] = 0.0f; break;
        }
      }
#pragma unroll
      for (int j = 0; j < E; j++) { RF(dst, j) = y[j]; prev[j] = y[j]; }
      prev_dst = dst;
    }
    for (int o = 0; o < P.n_out; o++) {
      float* q = P.optr[o]; const int64_t pitch = P.opitch[o]; const int reg = P.oreg[o];
#pragma unroll
      for (int j = 0; j < E; j++) if (ok[j]) q[(int64_t)r[j] * pitch + c[j]] = RF(reg, j);
    }
  }
#undef RF
}

template <int E>
static int launch_fused_e(agb_ctx* ctx, const FuseParams& P, int n_leaves, const agb_fuse_leaf* leaves, int n_instr, const agb_fuse_instr* instr) {
  int nreg = 0;
  for (int l = 0; l < n_leaves; l++) nreg = leaves[l].reg + 1 > nreg ? leaves[l].reg + 1 : nreg;
  for (int k = 0; k < n_instr; k++) nreg = instr[k].dst + 1 > nreg ? instr[k].dst + 1 : nreg;
  const size_t smem = (size_t)nreg * E * 256 * sizeof(float);
  static bool attr = false;
  if (!attr) { AGB_CUDA(cudaFuncSetAttribute(fused_ewise_kernel_e<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, AGB_FUSE_REGS * E * 256 * (int)sizeof(float))); attr = true; }
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fused_ewise_kernel_e<E>, 256, smem) != cudaSuccess || nb < 1) nb = 1;
  fused_ewise_kernel_e<E><<<agb_grid_for((P.total + E - 1) / E, 256, ctx->sm_count, nb), 256, smem, ctx->stream>>>(P);
  AGB_LAUNCHED(ctx);
  return AGB_OK;
}

extern "C" int agb_fused_ewise(agb_ctx* ctx, int64_t rows, int64_t cols, int n_leaves, const agb_fuse_leaf* leaves,
                               int n_instr, const agb_fuse_instr* instr, int n_out, const agb_fuse_out* outs) {
  AGB_CHECK(rows >= 0 && cols >= 0, AGB_ERR_INVALID_DIMS, "agb_fused_ewise: negative extent");
  AGB_CHECK(n_leaves >= 0 && n_leaves <= AGB_FUSE_MAX_LEAVES && n_instr >= 1 && n_instr <= AGB_FUSE_MAX_INSTR && n_out >= 1 && n_out <= AGB_FUSE_MAX_OUT,
            AGB_ERR_INVALID_DIMS, "agb_fused_ewise: program too large (%d leaves, %d instructions, %d outputs)", n_leaves, n_instr, n_out);
  FuseParams P; memset(&P, 0, sizeof(P));
  P.n_leaves = n_leaves; P.n_instr = n_instr; P.n_out = n_out; P.cols = cols; P.total = rows * cols;
  bool written[AGB_FUSE_REGS] = {false};
  for (int l = 0; l < n_leaves; l++) {
    AGB_CHECK(leaves[l].ptr != nullptr && leaves[l].reg >= 0 && leaves[l].reg < AGB_FUSE_REGS, AGB_ERR_INVALID_DIMS, "agb_fused_ewise: bad leaf %d", l);
    P.lptr[l] = leaves[l].ptr; P.lpitch[l] = leaves[l].pitch; P.lcs[l] = leaves[l].cstride; P.lreg[l] = (uint8_t)leaves[l].reg; written[leaves[l].reg] = true;
  }
  for (int k = 0; k < n_instr; k++) {
    const agb_fuse_instr& I = instr[k];
    AGB_CHECK(I.kind >= AGB_F_UNARY && I.kind <= AGB_F_BINARY_IMM_A, AGB_ERR_UNSUPPORTED, "agb_fused_ewise: instruction %d: bad kind %d", k, I.kind);
    AGB_CHECK(I.op >= 0 && I.op < (I.kind == AGB_F_UNARY ? (int)AGB_U_COUNT : (int)AGB_B_MIN + 1) && !(I.kind == AGB_F_UNARY && I.op == AGB_U_CLIP), AGB_ERR_UNSUPPORTED,
              "agb_fused_ewise: instruction %d: op %d is not fusable", k, I.op);
    AGB_CHECK(I.dst >= 0 && I.dst < AGB_FUSE_REGS && I.a >= 0 && I.a < AGB_FUSE_REGS && I.b >= 0 && I.b < AGB_FUSE_REGS, AGB_ERR_INVALID_DIMS, "agb_fused_ewise: instruction %d: register out of range", k);
    const bool need_a = I.kind != AGB_F_BINARY_IMM_A, need_b = I.kind == AGB_F_BINARY || I.kind == AGB_F_BINARY_IMM_A;
    AGB_CHECK((!need_a || written[I.a]) && (!need_b || written[I.b]), AGB_ERR_INVALID_DIMS, "agb_fused_ewise: instruction %d reads a register nothing wrote", k);
    written[I.dst] = true;
    P.code[k] = (uint32_t)I.kind | ((uint32_t)I.op << 2) | ((uint32_t)I.dst << 8) | ((uint32_t)I.a << 13) | ((uint32_t)I.b << 18);
    P.imm[k] = I.p0;
  }
  for (int o = 0; o < n_out; o++) {
    AGB_CHECK(outs[o].ptr != nullptr && outs[o].reg >= 0 && outs[o].reg < AGB_FUSE_REGS && written[outs[o].reg], AGB_ERR_INVALID_DIMS, "agb_fused_ewise: bad output %d", o);
    P.optr[o] = outs[o].ptr; P.opitch[o] = outs[o].pitch; P.oreg[o] = (uint8_t)outs[o].reg;
  }
  if (P.total == 0) return AGB_OK;
  AgbProfScope prof(ctx, AGB_PROF_EWISE, 4.0 * (double)P.total * (n_leaves + n_out));
  if (P.total >= (1 << 15) && P.total < (int64_t)0x7fffffff) {        // enough work to amortise the dispatch over several elements per thread
    static const int e_sel = [] { const char* e = getenv("AGB_FUSE_E"); return e ? atoi(e) : 2; }();       // tuning knob: elements per thread (2: 3.5 warps per scheduler on a [128, 1024] cell; 4 leaves 1.7)
    if (e_sel == 4) return launch_fused_e<4>(ctx, P, n_leaves, leaves, n_instr, instr);
    if (e_sel != 1) return launch_fused_e<2>(ctx, P, n_leaves, leaves, n_instr, instr);
  }
  fused_ewise_kernel<<<agb_grid_occ(ctx, fused_ewise_kernel, P.total, 256), 256, 0, ctx->stream>>>(P);
  AGB_LAUNCHED(ctx);
  return AGB_OK;
}

__global__ void __launch_bounds__(256) fill_kernel(float* __restrict__ y, int64_t n, float v) {
  int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t n4 = n >> 2;
  float4 v4 = make_float4(v, v, v, v);
  if ((((uintptr_t)y) & 15) == 0) {
    for (int64_t i = tid; i < n4; i += stride) stg_stream4(y + 4 * i, v4);
    for (int64_t j = (n4 << 2) + tid; j < n; j += stride) y[j] = v;
  } else {
    for (int64_t j = tid; j < n; j += stride) y[j] = v;
  }
}
extern "C" int agb_fill(agb_ctx* ctx, agb_tensor* y, float value) {
  AGB_CHECK(agb_is_contig(y), AGB_ERR_UNSUPPORTED, "agb_fill: output must be contiguous");
  int64_t n = agb_numel(y); if (n == 0) return AGB_OK;
  if (value == 0.0f) return agb_memset0(ctx, y->ptr, n * sizeof(float));
  fill_kernel<<<agb_grid_for((n + 3) / 4, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(y->ptr, n, value);
  AGB_LAUNCHED(ctx);
  return AGB_OK;
}

// ----------------------------------------------------------------------------------------------
// dropout: y = x * mask, mask = (u < 1 - ratio) as float.  Not rescaled (random_ops.rs:224-237).
// seed == 0: the caller supplies `mask` (parity runs: same mask as the oracle).
// seed != 0: Philox-4x32-10 counter RNG; the reference's XorShift stream is parity-unpinned.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr;
}
// Stream position kept ON THE DEVICE (cell[0] = evaluations so far, cell[1] = block ticket): an op instance that is evaluated repeatedly
// continues its stream like the reference's ArrayRng (RefCell<R>, ndarray_ext.rs:250-264) — also when the evaluation is a replayed CUDA
// graph, whose kernel arguments are frozen at capture.  Every block reads cell[0] when it starts; the block that finishes last advances it.
__device__ __forceinline__ uint32_t stream_pos_begin(const uint32_t* cell) { return cell ? *(const volatile uint32_t*)cell : 0u; }
__device__ __forceinline__ void stream_pos_end(uint32_t* cell, uint32_t pos) {
  if (cell == nullptr) return;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicInc(cell + 1, gridDim.x - 1) == gridDim.x - 1) { cell[0] = pos + 1; __threadfence(); }
  }
}
__global__ void __launch_bounds__(256) dropout_gen_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                          float* __restrict__ mask, int64_t n, float keep,
                                                          uint64_t seed, uint64_t offset, uint32_t* cell) {
  int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t n4 = (n + 3) >> 2;
  const uint32_t pos = stream_pos_begin(cell);
  for (int64_t i = tid; i < n4; i += stride) {
    uint64_t c = offset + (uint64_t)i;
    uint4 r = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), pos, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      int64_t j = 4 * i + k;
      if (j < n) {
        float u = (rr[k] >> 8) * (1.0f / 16777216.0f);   // U[0,1)
        float m = u < keep ? 1.0f : 0.0f;
        mask[j] = m; y[j] = x[j] * m;
      }
    }
  }
  stream_pos_end(cell, pos);
}
extern "C" int agb_stream_cell_bytes(void) { return 8; }
extern "C" int agb_dropout(agb_ctx* ctx, const agb_tensor* x, agb_tensor* y, agb_tensor* mask, float ratio, uint64_t seed, uint64_t offset) {
  return agb_dropout_stream(ctx, x, y, mask, ratio, seed, offset, nullptr);
}
extern "C" int agb_dropout_stream(agb_ctx* ctx, const agb_tensor* x, agb_tensor* y, agb_tensor* mask, float ratio, uint64_t seed, uint64_t offset, uint32_t* cell) {
  AGB_CHECK(agb_is_contig(x) && agb_is_contig(y) && agb_is_contig(mask), AGB_ERR_UNSUPPORTED, "agb_dropout: tensors must be contiguous");
  int64_t n = agb_numel(x);
  AGB_CHECK(agb_numel(y) == n && agb_numel(mask) == n, AGB_ERR_INCOMPATIBLE_SHAPE, "agb_dropout: size mismatch");
  if (n == 0) return AGB_OK;
  if (seed == 0) {
    agb_tensor xm = *x, mm = *mask, ym = *y;
    return agb_binary(ctx, AGB_B_MUL, 0.f, 0.f, &mm, &xm, &ym);    // mask * x
  }
  dropout_gen_kernel<<<agb_grid_for((n + 3) / 4, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(x->ptr, y->ptr, mask->ptr, n, 1.0f - ratio, seed, offset, cell);
  AGB_LAUNCHED(ctx);
  return AGB_OK;
}

// ----------------------------------------------------------------------------------------------
// random_* generator ops (random_ops.rs:6-214): counter-based, 4 elements per Philox call; Gamma is a per-element rejection loop on its
// own counter sub-space (word 3 = 0x80000000 | iteration)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float u01_co(uint32_t x) { return (x >> 8) * (1.0f / 16777216.0f); }             // [0, 1)
__device__ __forceinline__ float u01_oc(uint32_t x) { return ((x >> 8) + 1u) * (1.0f / 16777216.0f); }      // (0, 1]
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
  float r = sqrtf(-2.0f * logf(u01_oc(a))), s, c;
  sincospif(2.0f * u01_co(b), &s, &c);
  z0 = r * c; z1 = r * s;
}
__device__ float gamma_sample(int64_t elem, float k, float scale, uint64_t seed, uint32_t off) {
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  float boost = 1.0f;
  const bool small = k < 1.0f;
  const float kk = small ? k + 1.0f : k;
  const float d = kk - 1.0f / 3.0f, c = rsqrtf(9.0f * d);
  for (uint32_t it = 0; it < 64; it++) {
    uint4 r = philox4x32_10(make_uint4((uint32_t)elem, (uint32_t)((uint64_t)elem >> 32), off, 0x80000000u | it), key);
    if (it == 0 && small) boost = powf(u01_oc(r.w), 1.0f / k);          // Gamma(k) = Gamma(k + 1) * U^(1/k)
    float z, unused; box_muller(r.x, r.y, z, unused);
    float v = 1.0f + c * z;
    if (v <= 0.0f) continue;
    v = v * v * v;
    if (logf(u01_oc(r.z)) < 0.5f * z * z + d - d * v + d * logf(v)) return d * v * scale * boost;
  }
  return d * scale * boost;      // (probability ~ 1e-30) the mode
}
__global__ void __launch_bounds__(256) random_kernel(float* __restrict__ y, int64_t n, int kind, float p0, float p1, uint64_t seed, uint64_t offset, uint32_t* cell) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, n4 = (n + 3) >> 2;
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  const uint32_t pos = stream_pos_begin(cell);
  offset += pos;                                           // the op's evaluations so far (device-resident stream position)
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    float v[4];
    if (kind == AGB_RAND_GAMMA) {
#pragma unroll
      for (int k = 0; k < 4; k++) v[k] = 4 * i + k < n ? gamma_sample(4 * i + k, p0, p1, seed, (uint32_t)offset) : 0.0f;
    } else {
      uint4 r = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)((uint64_t)i >> 32), (uint32_t)offset, (uint32_t)(offset >> 32) & 0x7fffffffu), key);
      const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
      if (kind == AGB_RAND_NORMAL || kind == AGB_RAND_LOGNORMAL) {
        box_muller(r.x, r.y, v[0], v[1]); box_muller(r.z, r.w, v[2], v[3]);
#pragma unroll
        for (int k = 0; k < 4; k++) { v[k] = p0 + p1 * v[k]; if (kind == AGB_RAND_LOGNORMAL) v[k] = expf(v[k]); }
      } else {
#pragma unroll
        for (int k = 0; k < 4; k++) {
          if (kind == AGB_RAND_UNIFORM) { v[k] = p0 + (p1 - p0) * u01_co(rr[k]); if (v[k] >= p1 && p1 > p0) v[k] = p0; }     // rounding must not reach the open end
          else if (kind == AGB_RAND_BERNOULLI) v[k] = u01_co(rr[k]) < p0 ? 1.0f : 0.0f;
          else v[k] = -logf(u01_oc(rr[k])) / p0;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) if (4 * i + k < n) y[4 * i + k] = v[k];
  }
  stream_pos_end(cell, pos);
}
extern "C" int agb_random(agb_ctx* ctx, int kind, float p0, float p1, uint64_t seed, uint64_t offset, agb_tensor* y) {
  return agb_random_stream(ctx, kind, p0, p1, seed, offset, nullptr, y);
}
extern "C" int agb_random_stream(agb_ctx* ctx, int kind, float p0, float p1, uint64_t seed, uint64_t offset, uint32_t* cell, agb_tensor* y) {
  AGB_CHECK(kind >= 0 && kind < AGB_RAND_COUNT, AGB_ERR_UNSUPPORTED, "agb_random: unknown distribution %d", kind);
  AGB_CHECK(agb_is_contig(y), AGB_ERR_UNSUPPORTED, "agb_random: output must be contiguous");
  if (kind == AGB_RAND_NORMAL || kind == AGB_RAND_LOGNORMAL) AGB_CHECK(p1 >= 0.0f, AGB_ERR_INVALID_DIMS, "agb_random: standard deviation must be >= 0");     // Normal::new(..).unwrap() panics
  if (kind == AGB_RAND_EXP) AGB_CHECK(p0 > 0.0f, AGB_ERR_INVALID_DIMS, "agb_random: exponential rate must be > 0");
  if (kind == AGB_RAND_GAMMA) AGB_CHECK(p0 > 0.0f && p1 > 0.0f, AGB_ERR_INVALID_DIMS, "agb_random: gamma shape and scale must be > 0");
  if (kind == AGB_RAND_UNIFORM) AGB_CHECK(p0 < p1, AGB_ERR_INVALID_DIMS, "agb_random: uniform range must satisfy low < high");                              // Uniform::new panics otherwise
  const int64_t n = agb_numel(y); if (n == 0) return AGB_OK;
  random_kernel<<<agb_grid_for((n + 3) / 4, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(y->ptr, n, kind, p0, p1, seed, offset, cell);
  AGB_LAUNCHED(ctx);
  return AGB_OK;
}
