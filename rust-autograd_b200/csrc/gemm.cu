// gemm.cu — agb_gemm_f32: MatMul / BatchMatMul entry point and dispatch.
//
// Reference semantics followed (src/tensor_ops/dot_ops.rs):
//   MatMul::compute       :565-606  2-D only, transposes = stride swap (:574-579), k mismatch -> IncompatibleShape
//   BatchMatMul::compute  :632-695  rank >= 2, identical leading dims (no broadcast, :661), transposes on last two axes
// Output layout: always C-contiguous (the reference may emit Fortran order, dot_ops.rs:585-594; SURVEY §9.17 lets
// the device path pick one layout and carry strides in the descriptor).
//
// Dispatch: tcgen05 tensor-core kernel (tc_gemm.cu; TF32 or 3xTF32) when the operands satisfy TMA's
// alignment rules and the problem is big enough to fill a 128-lane tile; otherwise the CUDA-core fp32 kernel.
#include "simt_gemm.cuh"

int agb_tc_gemm(agb_ctx* ctx, int mode, const float* A, const float* B, float* C,
                int64_t M, int64_t N, int64_t K, int64_t batch,
                int64_t rsa, int64_t csa, int64_t bsa, int64_t rsb, int64_t csb, int64_t bsb, int64_t bsc, float beta);

// loaders: element (z, m, k) of op(A) / (z, k, n) of op(B).  Split-K (kchunk > 0): z indexes a K slab instead of a batch entry.
#define LOADER_FIELDS const float* p; int64_t rs, cs, bs; int64_t kchunk = 0, ktot = 0;
struct StridedA { LOADER_FIELDS static const bool K_CONTIG = true;
  __device__ __forceinline__ float load(int z, int64_t m, int64_t k) const { if (kchunk && z * kchunk + k >= ktot) return 0.0f; return __ldg(p + z * bs + m * rs + k * cs); } };
struct StridedA_M { LOADER_FIELDS static const bool K_CONTIG = false;
  __device__ __forceinline__ float load(int z, int64_t m, int64_t k) const { if (kchunk && z * kchunk + k >= ktot) return 0.0f; return __ldg(p + z * bs + m * rs + k * cs); } };
struct StridedB { LOADER_FIELDS static const bool K_CONTIG = false;   // n contiguous
  __device__ __forceinline__ float load(int z, int64_t k, int64_t n) const { if (kchunk && z * kchunk + k >= ktot) return 0.0f; return __ldg(p + z * bs + k * rs + n * cs); } };
struct StridedB_K { LOADER_FIELDS static const bool K_CONTIG = true;
  __device__ __forceinline__ float load(int z, int64_t k, int64_t n) const { if (kchunk && z * kchunk + k >= ktot) return 0.0f; return __ldg(p + z * bs + k * rs + n * cs); } };
struct StoreC { float* p; int64_t ld, bs; int accumulate; int atomic = 0;
  __device__ __forceinline__ void store(int z, int64_t m, int64_t n, float v) const {
    if (atomic) { atomicAdd(p + m * ld + n, v); return; }
    float* q = p + z * bs + m * ld + n; *q = accumulate ? *q + v : v; } };

static int simt_gemm(agb_ctx* ctx, const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, int64_t batch,
                     int64_t rsa, int64_t csa, int64_t bsa, int64_t rsb, int64_t csb, int64_t bsb, int64_t bsc, float beta) {
  StoreC cs{C, N, bsc, beta != 0.0f};
  bool a_kc = (csa == 1) || (rsa != 1), b_kc = (rsb == 1) && (csb != 1);
  // skinny outputs with a long reduction (e.g. the classifier of a CNN: [256 x 65536] . [65536 x 10]): split K across the
  // SMs, fp32 atomics into the (zeroed) result
  int64_t tile = (M >= 96 && N >= 96) ? 128 : 64;
  int64_t tiles = ((M + tile - 1) / tile) * ((N + tile - 1) / tile);
  if (batch == 1 && K >= 512 && tiles * 2 <= ctx->sm_count) {       // (K >= 512: the 200 x 784 x 10 softmax-regression GEMM ran 73 us on 4 CTAs)
    int64_t splits = (2 * (int64_t)ctx->sm_count + tiles - 1) / tiles;
    int64_t kchunk = (K + splits - 1) / splits; kchunk = (kchunk + 15) / 16 * 16; if (kchunk < 128) kchunk = 128;
    splits = (K + kchunk - 1) / kchunk;
    if (splits > 1 && splits <= 65535) {
      // small outputs: every split stores its partial product, one reduction adds them in a fixed order (deterministic, and the same
      // number of graph nodes as zero-fill + atomics); large outputs keep the atomic accumulation
      float* part = nullptr;
      if (beta == 0.0f && splits * M * N <= (1ll << 22)) { void* pv = nullptr; AGB_TRY(agb_alloc(ctx, (size_t)(splits * M * N) * sizeof(float), &pv)); part = (float*)pv; }
      if (!part && beta == 0.0f) AGB_TRY(agb_memset0(ctx, C, (size_t)M * N * sizeof(float)));
      StoreC ca{C, N, 0, 1, 1};
      if (part) ca = StoreC{part, N, M * N, 0, 0};
      struct Finish { agb_ctx* ctx; float* part; float* C; int64_t splits, mn;
        int operator()(int r) const { if (!part) return r; if (r == AGB_OK) r = agb_reduce(ctx, AGB_R_SUM, part, C, 1, splits, mn); agb_free(ctx, part); return r; } } finish{ctx, part, C, splits, M * N};
      int r;
      if (a_kc && !b_kc) r = simt_gemm_launch(ctx, StridedA{A, rsa, csa, kchunk * csa, kchunk, K}, StridedB{B, rsb, csb, kchunk * rsb, kchunk, K}, ca, M, N, kchunk, splits);
      else if (a_kc && b_kc) r = simt_gemm_launch(ctx, StridedA{A, rsa, csa, kchunk * csa, kchunk, K}, StridedB_K{B, rsb, csb, kchunk * rsb, kchunk, K}, ca, M, N, kchunk, splits);
      else if (!a_kc && !b_kc) r = simt_gemm_launch(ctx, StridedA_M{A, rsa, csa, kchunk * csa, kchunk, K}, StridedB{B, rsb, csb, kchunk * rsb, kchunk, K}, ca, M, N, kchunk, splits);
      else r = simt_gemm_launch(ctx, StridedA_M{A, rsa, csa, kchunk * csa, kchunk, K}, StridedB_K{B, rsb, csb, kchunk * rsb, kchunk, K}, ca, M, N, kchunk, splits);
      return finish(r);
    }
  }
  for (int64_t z0 = 0; z0 < batch; z0 += 65535) {
    int64_t zc = batch - z0 < 65535 ? batch - z0 : 65535;
    const float* a = A + z0 * bsa; const float* b = B + z0 * bsb; StoreC c = cs; c.p = C + z0 * bsc;
    int r;
    if (a_kc && !b_kc) r = simt_gemm_launch(ctx, StridedA{a, rsa, csa, bsa}, StridedB{b, rsb, csb, bsb}, c, M, N, K, zc);
    else if (a_kc && b_kc) r = simt_gemm_launch(ctx, StridedA{a, rsa, csa, bsa}, StridedB_K{b, rsb, csb, bsb}, c, M, N, K, zc);
    else if (!a_kc && !b_kc) r = simt_gemm_launch(ctx, StridedA_M{a, rsa, csa, bsa}, StridedB{b, rsb, csb, bsb}, c, M, N, K, zc);
    else r = simt_gemm_launch(ctx, StridedA_M{a, rsa, csa, bsa}, StridedB_K{b, rsb, csb, bsb}, c, M, N, K, zc);
    AGB_TRY(r);
  }
  return AGB_OK;
}

// ----------------------------------------------------------------------------------------------
// Skinny GEMMs: one extent <= 16 next to a large operand — the classifier of the VGG stack (FC 65536 -> 10 at batch 256): forward
// [256, 65536] x [65536, 10], weight gradient [65536, 256] x [256, 10], input gradient [256, 10] x [10, 65536].  They are one streaming pass over
// the large operand (67 MB: ~10 us at the HBM rate); on the 128-lane tensor-core tiles they used 10 of 128 lanes or 10 of 32 k and took 56 / 44 /
// 53 us behind padded copies.  Three CUDA-core kernels in exact fp32 FMA (dot_ops.rs:383-422 is an f32 sgemm):
//   thin N, k contiguous in A   : CTA = 64 rows x 1024-k slice, B^T slice in shared memory, a warp walks 4 rows at a time (every 128-bit
//                                 shared-memory read of B^T feeds 4 rows), per-slice partial sums added in slice order (deterministic)
//   thin N, m contiguous in A   : thread per output row m, B in shared memory (broadcast reads), k sequential
//   thin K                      : thread per output column n with its B column in registers, A in shared memory, coalesced stores of C
// ----------------------------------------------------------------------------------------------
#define SK_MAX 16
#define SK_KSLICE 1024
template <int NN>
__global__ void __launch_bounds__(256) skinny_n_kcontig_kernel(const float* __restrict__ A, int64_t rsa, const float* __restrict__ B, int64_t rsb, int64_t csb,
                                                               float* __restrict__ part, int M, int N, int64_t K) {
  extern __shared__ __align__(16) float Wt[];                   // [NN][SK_KSLICE + 4]
  constexpr int PITCH = SK_KSLICE + 4;
  const int64_t k0 = (int64_t)blockIdx.x * SK_KSLICE;
  const int kn = (int)min((int64_t)SK_KSLICE, K - k0);
  for (int i = threadIdx.x; i < NN * SK_KSLICE; i += blockDim.x) {
    int kk, n;
    if (csb == 1) { kk = i / NN; n = i - kk * NN; } else { n = i / SK_KSLICE; kk = i - n * SK_KSLICE; }      // walk B in its contiguous direction
    Wt[n * PITCH + kk] = (kk < kn && n < N) ? __ldg(B + (k0 + kk) * rsb + n * csb) : 0.0f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_base = blockIdx.y * 64 + warp * 8;
#pragma unroll 1
  for (int g = 0; g < 2; g++) {
    const int r0 = row_base + g * 4;
    if (r0 >= M) break;
    float acc[4][NN];
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
      for (int n = 0; n < NN; n++) acc[j][n] = 0.0f;
    const float* ap[4];
#pragma unroll
    for (int j = 0; j < 4; j++) ap[j] = A + (int64_t)min(r0 + j, M - 1) * rsa + k0;
#pragma unroll 2
    for (int kk = lane * 4; kk < kn; kk += 128) {
      float4 a[4];
#pragma unroll
      for (int j = 0; j < 4; j++) a[j] = ldg_stream4(ap[j] + kk);
#pragma unroll
      for (int n = 0; n < NN; n++) {
        const float4 w = *(const float4*)(Wt + n * PITCH + kk);
#pragma unroll
        for (int j = 0; j < 4; j++) acc[j][n] = fmaf(a[j].x, w.x, fmaf(a[j].y, w.y, fmaf(a[j].z, w.z, fmaf(a[j].w, w.w, acc[j][n]))));
      }
    }
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
      for (int n = 0; n < NN; n++) acc[j][n] = warp_sum(acc[j][n]);
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (r0 + j < M) {
#pragma unroll
          for (int n = 0; n < NN; n++) if (n < N) part[((int64_t)blockIdx.x * M + r0 + j) * N + n] = acc[j][n];
        }
    }
  }
}
template <int NN>
__global__ void __launch_bounds__(256) skinny_n_mcontig_kernel(const float* __restrict__ A, int64_t csa, const float* __restrict__ B, int64_t rsb, int64_t csb,
                                                               float* __restrict__ C, int64_t M, int N, int K, int accumulate) {
  __shared__ __align__(16) float Bs[256 * NN];                  // a chunk of up to 256 k
  const int64_t m = blockIdx.x * (int64_t)256 + threadIdx.x;
  float acc[NN];
#pragma unroll
  for (int n = 0; n < NN; n++) acc[n] = 0.0f;
  for (int kc = 0; kc < K; kc += 256) {
    const int kn = min(256, K - kc);
    __syncthreads();
    for (int i = threadIdx.x; i < kn * NN; i += blockDim.x) { const int kk = i / NN, n = i - kk * NN; Bs[i] = n < N ? __ldg(B + (int64_t)(kc + kk) * rsb + n * csb) : 0.0f; }
    __syncthreads();
    if (m < M) {
      const float* ap = A + m + (int64_t)kc * csa;
      int kk = 0;
      for (; kk + 8 <= kn; kk += 8) {                           // eight independent loads in flight per thread
        float a[8];
#pragma unroll
        for (int u = 0; u < 8; u++) a[u] = __ldg(ap + (int64_t)(kk + u) * csa);
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
          for (int n = 0; n < NN; n++) acc[n] = fmaf(a[u], Bs[(kk + u) * NN + n], acc[n]);
      }
      for (; kk < kn; kk++) {
        const float a = __ldg(ap + (int64_t)kk * csa);
#pragma unroll
        for (int n = 0; n < NN; n++) acc[n] = fmaf(a, Bs[kk * NN + n], acc[n]);
      }
    }
  }
  if (m < M) {
#pragma unroll
    for (int n = 0; n < NN; n++) if (n < N) { float* c = C + m * N + n; *c = accumulate ? *c + acc[n] : acc[n]; }
  }
}
__global__ void __launch_bounds__(256) skinny_k_kernel(const float* __restrict__ A, int64_t rsa, int64_t csa, const float* __restrict__ B, int64_t rsb, int64_t csb,
                                                       float* __restrict__ C, int M, int64_t N, int K, int accumulate) {
  __shared__ float As[32][SK_MAX];
  const int64_t n = blockIdx.x * (int64_t)256 + threadIdx.x;
  const int m0 = blockIdx.y * 32, mn = min(32, M - m0);
  for (int i = threadIdx.x; i < 32 * SK_MAX; i += blockDim.x) { const int mm = i / SK_MAX, k = i - mm * SK_MAX; As[mm][k] = (mm < mn && k < K) ? __ldg(A + (int64_t)(m0 + mm) * rsa + k * csa) : 0.0f; }
  __syncthreads();
  if (n >= N) return;
  float b[SK_MAX];
#pragma unroll
  for (int k = 0; k < SK_MAX; k++) b[k] = k < K ? __ldg(B + (int64_t)k * rsb + n * csb) : 0.0f;
  for (int mm = 0; mm < mn; mm++) {
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < SK_MAX; k++) s = fmaf(As[mm][k], b[k], s);
    float* c = C + (int64_t)(m0 + mm) * N + n;
    *c = accumulate ? *c + s : s;
  }
}
// returns AGB_ERR_UNSUPPORTED when no skinny form applies
static int skinny_gemm(agb_ctx* ctx, const float* A, const float* B, float* C, int64_t m, int64_t n, int64_t k, int64_t rsa, int64_t csa, int64_t rsb, int64_t csb, float beta) {
  static const int enabled = [] { const char* e = getenv("AGB_SKINNY_GEMM"); return (e && e[0] == '0') ? 0 : 1; }();
  if (!enabled || m >= (1ll << 31) || n >= (1ll << 31) || k >= (1ll << 31)) return AGB_ERR_UNSUPPORTED;
  const int acc = beta != 0.0f;
  if (n <= SK_MAX && m * k >= (1 << 18)) {
    const int NN = n <= 4 ? 4 : n <= 8 ? 8 : n <= 12 ? 12 : 16;
    if (csa == 1 && k >= 4 * SK_KSLICE && k % 4 == 0 && rsa % 4 == 0 && (((uintptr_t)A) & 15) == 0) {
      const int64_t slices = (k + SK_KSLICE - 1) / SK_KSLICE;
      float* part; AGB_TRY(agb_scratch2(ctx, (size_t)slices * m * n * sizeof(float), (void**)&part));
      const dim3 grid((unsigned)slices, (unsigned)((m + 63) / 64));
      const size_t smem = (size_t)NN * (SK_KSLICE + 4) * sizeof(float);
#define SKA(NN_) do { static bool attr = false; if (!attr) { AGB_CUDA(cudaFuncSetAttribute(skinny_n_kcontig_kernel<NN_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * (SK_KSLICE + 4) * 4)); attr = true; } \
                      skinny_n_kcontig_kernel<NN_><<<grid, 256, smem, ctx->stream>>>(A, rsa, B, rsb, csb, part, (int)m, (int)n, k); } while (0)
      if (NN == 4) SKA(4); else if (NN == 8) SKA(8); else if (NN == 12) SKA(12); else SKA(16);
#undef SKA
      AGB_LAUNCHED(ctx);
      return agb_reduce_partials(ctx, part, C, (int)slices, m * n, m * n, acc);
    }
    if (rsa == 1 && m >= 4096) {
      const unsigned grid = (unsigned)((m + 255) / 256);
      if (NN == 4) skinny_n_mcontig_kernel<4><<<grid, 256, 0, ctx->stream>>>(A, csa, B, rsb, csb, C, m, (int)n, (int)k, acc);
      else if (NN == 8) skinny_n_mcontig_kernel<8><<<grid, 256, 0, ctx->stream>>>(A, csa, B, rsb, csb, C, m, (int)n, (int)k, acc);
      else if (NN == 12) skinny_n_mcontig_kernel<12><<<grid, 256, 0, ctx->stream>>>(A, csa, B, rsb, csb, C, m, (int)n, (int)k, acc);
      else skinny_n_mcontig_kernel<16><<<grid, 256, 0, ctx->stream>>>(A, csa, B, rsb, csb, C, m, (int)n, (int)k, acc);
      AGB_LAUNCHED(ctx);
      return AGB_OK;
    }
  }
  if (k <= SK_MAX && m * n >= (1 << 18) && n >= 4096) {
    skinny_k_kernel<<<dim3((unsigned)((n + 255) / 256), (unsigned)((m + 31) / 32)), 256, 0, ctx->stream>>>(A, rsa, csa, B, rsb, csb, C, (int)m, n, (int)k, acc);
    AGB_LAUNCHED(ctx);
    return AGB_OK;
  }
  return AGB_ERR_UNSUPPORTED;
}

extern "C" int agb_gemm_f32(agb_ctx* ctx, int trans_a, int trans_b, const agb_tensor* a, const agb_tensor* b, agb_tensor* c, float beta) {
  AGB_CHECK(a->rank >= 2 && b->rank >= 2, AGB_ERR_INCOMPATIBLE_SHAPE, "matmul: inputs must have ndim >= 2 (got %d and %d)", a->rank, b->rank);
  AGB_CHECK(a->rank == b->rank && c->rank == a->rank, AGB_ERR_INCOMPATIBLE_SHAPE, "matmul: rank mismatch: %d vs %d (out %d)", a->rank, b->rank, c->rank);
  AGB_CHECK(beta == 0.0f || beta == 1.0f, AGB_ERR_INVALID_DIMS, "matmul: beta must be 0 or 1");
  const int R = a->rank;
  int64_t m = a->shape[R - 2], k = a->shape[R - 1], rsa = a->stride[R - 2], csa = a->stride[R - 1];
  if (trans_a) { int64_t t = m; m = k; k = t; t = rsa; rsa = csa; csa = t; }
  int64_t k2 = b->shape[R - 2], n = b->shape[R - 1], rsb = b->stride[R - 2], csb = b->stride[R - 1];
  if (trans_b) { int64_t t = k2; k2 = n; n = t; t = rsb; rsb = csb; csb = t; }
  AGB_CHECK(k == k2, AGB_ERR_INCOMPATIBLE_SHAPE, "inputs %lld x %lld and %lld x %lld are not compatible for matrix multiplication",
            (long long)m, (long long)k, (long long)k2, (long long)n);    // dot_shape_error text, dot_ops.rs
  int64_t batch = 1;
  for (int i = 0; i < R - 2; i++) {
    AGB_CHECK(a->shape[i] == b->shape[i], AGB_ERR_INCOMPATIBLE_SHAPE, "Input shapes mismatch on batch axis %d: %lld vs %lld", i, (long long)a->shape[i], (long long)b->shape[i]);
    AGB_CHECK(c->shape[i] == a->shape[i], AGB_ERR_INCOMPATIBLE_SHAPE, "matmul: output batch axis %d mismatch", i);
    batch *= a->shape[i];
  }
  AGB_CHECK(c->shape[R - 2] == m && c->shape[R - 1] == n, AGB_ERR_INCOMPATIBLE_SHAPE, "matmul: output must be [.., %lld, %lld]", (long long)m, (long long)n);
  AGB_CHECK(agb_is_contig(c), AGB_ERR_UNSUPPORTED, "matmul: output must be C-contiguous");
  // batch dims must collapse to a single stride (the reference deep-copies otherwise, dot_ops.rs:444-453,524-531;
  // the host evaluator does the same before calling here)
  int64_t bsa = 0, bsb = 0;
  if (R > 2) {
    bsa = a->stride[R - 3]; bsb = b->stride[R - 3];
    for (int i = R - 4; i >= 0; i--) {
      AGB_CHECK(a->shape[i] == 1 || a->stride[i] == a->stride[i + 1] * a->shape[i + 1], AGB_ERR_UNSUPPORTED, "batch_matmul: lhs batch dims are not collapsible; copy first");
      AGB_CHECK(b->shape[i] == 1 || b->stride[i] == b->stride[i + 1] * b->shape[i + 1], AGB_ERR_UNSUPPORTED, "batch_matmul: rhs batch dims are not collapsible; copy first");
    }
  }
  if (m == 0 || n == 0 || batch == 0) return AGB_OK;
  if (k == 0) { if (beta == 0.0f) return agb_memset0(ctx, c->ptr, agb_numel(c) * sizeof(float)); return AGB_OK; }
  AgbProfScope prof(ctx, AGB_PROF_GEMM, 2.0 * (double)m * (double)n * (double)k * (double)batch);
  int mode = ctx->math_mode;
  if (batch == 1) {
    int r = skinny_gemm(ctx, a->ptr, b->ptr, c->ptr, m, n, k, rsa, csa, rsb, csb, beta);
    if (r != AGB_ERR_UNSUPPORTED) return r;
  }
  if (mode != AGB_MATH_FP32) {
    int r = agb_tc_gemm(ctx, mode, a->ptr, b->ptr, c->ptr, m, n, k, batch, rsa, csa, bsa, rsb, csb, bsb, m * n, beta);
    if (r != AGB_ERR_UNSUPPORTED) return r;
  }
  return simt_gemm(ctx, a->ptr, b->ptr, c->ptr, m, n, k, batch, rsa, csa, bsa, rsb, csb, bsb, m * n, beta);
}
