// softmax.cu — softmax / log_softmax / logsumexp over a [outer, r, inner] view and the fused
// sparse-softmax-cross-entropy forward/backward.  HBM roofline: 8 B/elem (read once + write once,
// SURVEY §8d) — rows are staged once in shared memory / registers so x is read from HBM exactly once
// whenever a row fits on chip (<= 48 K floats); longer rows take a second (L2-served) read.
//
// Reference semantics followed:
//   softmax_impl (max-subtract, exp, sum, divide)      src/tensor_ops/activation_ops.rs:61-96
//   logsumexp_forward (max, exp, sum, ln, + max)       src/tensor_ops/math_ops.rs:540-593
//   LogSoftmax = x - logsumexp(x)                      src/tensor_ops/xent_ops.rs:17-22
//   SparseSoftmaxCrossEntropy (loss [B,1], log_x)      src/tensor_ops/xent_ops.rs:63-113
//   SparseSoftmaxCrossEntropyGrad                      src/tensor_ops/xent_ops.rs:139-152
//   SoftmaxCrossEntropy                                src/tensor_ops/xent_ops.rs:160-177
// The reference max-fold starts from T::min_value() (= -FLT_MAX); so does this one.
#include "common.cuh"
#include <float.h>

enum { SM_SOFTMAX = 0, SM_LOGSOFTMAX = 1, SM_LSE = 2, SM_SPARSE_XENT = 3, SM_DENSE_XENT = 4 };

__device__ __forceinline__ float block_max(float v, float* sm) {
  v = warp_max(v);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (l == 0) sm[w] = v;
  __syncthreads();
  float r = (l < nw) ? sm[l] : -FLT_MAX;
  r = warp_max(r);
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_sum(float v, float* sm) {
  v = warp_sum(v);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (l == 0) sm[w] = v;
  __syncthreads();
  float r = (l < nw) ? sm[l] : 0.0f;
  r = warp_sum(r);
  __syncthreads();
  return r;
}

// finalise one row element
template <int MODE> __device__ __forceinline__ float sm_out(float x, float mx, float sum, float lse) {
  if (MODE == SM_SOFTMAX) return expf(x - mx) / sum;
  return x - lse;
}

// ---- short rows: one warp per row, row cached in registers (r <= 32*CAP) ----
template <int MODE, int CAP>
__global__ void __launch_bounds__(256) softmax_warp_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                           const float* __restrict__ aux, float* __restrict__ loss,
                                                           int64_t rows, int r, int* err) {
  int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  int l = threadIdx.x & 31;
  const float* p = x + row * (int64_t)r;
  float v[CAP];
  float mx = -FLT_MAX;
#pragma unroll
  for (int k = 0; k < CAP; k++) { int i = l + 32 * k; v[k] = (i < r) ? __ldg(p + i) : -FLT_MAX; if (i < r) mx = fmaxf(mx, v[k]); }
  mx = warp_max(mx);
  float s = 0.0f;
#pragma unroll
  for (int k = 0; k < CAP; k++) { int i = l + 32 * k; if (i < r) s += expf(v[k] - mx); }
  s = warp_sum(s);
  float lse = logf(s) + mx;
  if (MODE == SM_LSE) { if (l == 0) y[row] = lse; return; }
  float* q = y + row * (int64_t)r;
  float dense = 0.0f;
#pragma unroll
  for (int k = 0; k < CAP; k++) {
    int i = l + 32 * k;
    if (i < r) {
      float o = sm_out<MODE>(v[k], mx, s, lse);
      q[i] = o;
      if (MODE == SM_DENSE_XENT) dense += __ldg(aux + row * (int64_t)r + i) * o;
    }
  }
  if (MODE == SM_SPARSE_XENT && l == 0) {
    float tf = __ldg(aux + row); int t = (int)tf;
    if (tf < 0.0f || t >= r || tf != tf) { atomicExch(err, 1); loss[row] = nanf(""); }
    else loss[row] = -(__ldg(p + t) - lse);
  }
  if (MODE == SM_DENSE_XENT) { dense = warp_sum(dense); if (l == 0) loss[row] = -dense; }
}

// ---- medium rows (32 < r <= 2048, r % 4 == 0, 16-byte aligned rows): one warp per row with 128-bit loads / stores — the scalar
//      version above issued four times as many memory instructions and ran at 0.37 of the HBM rate on 1024-column rows ----
template <int MODE, int CAP4>
__global__ void __launch_bounds__(256) softmax_warp4_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                            const float* __restrict__ aux, float* __restrict__ loss,
                                                            int64_t rows, int r, int* err) {
  int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int l = threadIdx.x & 31, n4 = r >> 2;
  const float* p = x + row * (int64_t)r;
  float4 v[CAP4];
  float mx = -FLT_MAX;
#pragma unroll
  for (int k = 0; k < CAP4; k++) {
    const int i = l + 32 * k;
    if (i < n4) { v[k] = ldg_stream4(p + 4 * i); mx = fmaxf(mx, fmaxf(fmaxf(v[k].x, v[k].y), fmaxf(v[k].z, v[k].w))); }
    else v[k] = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
  }
  mx = warp_max(mx);
  float s = 0.0f;
#pragma unroll
  for (int k = 0; k < CAP4; k++) if (l + 32 * k < n4) s += expf(v[k].x - mx) + expf(v[k].y - mx) + expf(v[k].z - mx) + expf(v[k].w - mx);
  s = warp_sum(s);
  const float lse = logf(s) + mx;
  if (MODE == SM_LSE) { if (l == 0) y[row] = lse; return; }
  float* q = y + row * (int64_t)r;
  float dense = 0.0f;
#pragma unroll
  for (int k = 0; k < CAP4; k++) {
    const int i = l + 32 * k;
    if (i < n4) {
      float4 o;
      o.x = sm_out<MODE>(v[k].x, mx, s, lse); o.y = sm_out<MODE>(v[k].y, mx, s, lse); o.z = sm_out<MODE>(v[k].z, mx, s, lse); o.w = sm_out<MODE>(v[k].w, mx, s, lse);
      stg_stream4(q + 4 * i, o);
      if (MODE == SM_DENSE_XENT) { const float4 a = __ldg((const float4*)(aux + row * (int64_t)r + 4 * i)); dense += a.x * o.x + a.y * o.y + a.z * o.z + a.w * o.w; }
    }
  }
  if (MODE == SM_SPARSE_XENT && l == 0) {
    float tf = __ldg(aux + row); int t = (int)tf;
    if (tf < 0.0f || t >= r || tf != tf) { atomicExch(err, 1); loss[row] = nanf(""); }
    else loss[row] = -(__ldg(p + t) - lse);
  }
  if (MODE == SM_DENSE_XENT) { dense = warp_sum(dense); if (l == 0) loss[row] = -dense; }
}

// ---- long rows: one block per row.  Online (single-pass) max + sum: one block reduction instead of two.
//      CACHED (row <= 48 KB, several CTAs per SM): the row is staged in shared memory, HBM sees it exactly once.
//      !CACHED (longer rows): the row is re-read for the output pass — it is at most a few MB and still L2-resident, so HBM
//      traffic stays at the algorithmic 8 B/elem while 4+ CTAs per SM keep the load / reduce / store phases of different rows
//      overlapped (a 128 KB shared-memory row would pin the SM to one row at a time).
struct MaxSum { float m, s; };
__device__ __forceinline__ MaxSum ms_comb(MaxSum a, MaxSum b) {
  MaxSum r; r.m = fmaxf(a.m, b.m);
  r.s = a.s * __expf(a.m - r.m) + b.s * __expf(b.m - r.m);
  return r;
}
__device__ __forceinline__ MaxSum block_maxsum(MaxSum v, MaxSum* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { MaxSum t; t.m = __shfl_xor_sync(0xffffffffu, v.m, o); t.s = __shfl_xor_sync(0xffffffffu, v.s, o); v = ms_comb(v, t); }
  int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (l == 0) sm[w] = v;
  __syncthreads();
  MaxSum r; r.m = -FLT_MAX; r.s = 0.0f;
  if (l < nw) r = sm[l];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { MaxSum t; t.m = __shfl_xor_sync(0xffffffffu, r.m, o); t.s = __shfl_xor_sync(0xffffffffu, r.s, o); r = ms_comb(r, t); }
  __syncthreads();
  return r;
}
__device__ __forceinline__ void ms_push4(MaxSum& a, float4 v) {
  float m4 = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
  float m = fmaxf(a.m, m4);
  a.s = a.s * __expf(a.m - m) + __expf(v.x - m) + __expf(v.y - m) + __expf(v.z - m) + __expf(v.w - m);
  a.m = m;
}
__device__ __forceinline__ void ms_push1(MaxSum& a, float v) {
  float m = fmaxf(a.m, v);
  a.s = a.s * __expf(a.m - m) + __expf(v - m);
  a.m = m;
}

template <int MODE, bool CACHED>
__global__ void __launch_bounds__(512) softmax_block_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                            const float* __restrict__ aux, float* __restrict__ loss,
                                                            int64_t r, int* err) {
  extern __shared__ __align__(16) float rowbuf[];
  __shared__ MaxSum red[32];
  int64_t row = blockIdx.x;
  const float* p = x + row * r;
  const int T = blockDim.x;
  bool vec = ((((uintptr_t)p) & 15) == 0) && (r % 4 == 0);
  MaxSum acc; acc.m = -FLT_MAX; acc.s = 0.0f;
  if (vec) {
    const int64_t n4 = r >> 2;
    int64_t i = threadIdx.x;
    for (; i + T < n4; i += 2 * T) {               // two independent 128-bit loads in flight (four cost occupancy: 4096-column rows 0.83 -> 0.68 of HBM)
      float4 v0 = ldg_stream4(p + 4 * i), v1 = ldg_stream4(p + 4 * (i + T));
      if (CACHED) { *(float4*)(rowbuf + 4 * i) = v0; *(float4*)(rowbuf + 4 * (i + T)) = v1; }
      ms_push4(acc, v0); ms_push4(acc, v1);
    }
    for (; i < n4; i += T) { float4 v = ldg_stream4(p + 4 * i); if (CACHED) *(float4*)(rowbuf + 4 * i) = v; ms_push4(acc, v); }
  } else {
    for (int64_t i = threadIdx.x; i < r; i += T) { float v = __ldg(p + i); if (CACHED) rowbuf[i] = v; ms_push1(acc, v); }
  }
  acc = block_maxsum(acc, red);                     // also orders the rowbuf writes (syncthreads inside)
  const float mx = acc.m, s = acc.s;
  const float lse = logf(s) + mx, inv = 1.0f / s;
  if (MODE == SM_LSE) { if (threadIdx.x == 0) y[row] = lse; return; }
  const float* src = CACHED ? rowbuf : p;
  float* q = y + row * r;
  float dense = 0.0f;
  bool vecq = vec && ((((uintptr_t)q) & 15) == 0);
  if (vecq && MODE != SM_DENSE_XENT) {
    for (int64_t i = threadIdx.x; i < (r >> 2); i += T) {
      float4 v = CACHED ? *(const float4*)(rowbuf + 4 * i) : __ldg((const float4*)(p + 4 * i));
      if (MODE == SM_SOFTMAX) { v.x = __expf(v.x - mx) * inv; v.y = __expf(v.y - mx) * inv; v.z = __expf(v.z - mx) * inv; v.w = __expf(v.w - mx) * inv; }
      else { v.x -= lse; v.y -= lse; v.z -= lse; v.w -= lse; }
      stg_stream4(q + 4 * i, v);
    }
  } else {
    for (int64_t i = threadIdx.x; i < r; i += T) {
      float o = (MODE == SM_SOFTMAX) ? __expf(src[i] - mx) * inv : src[i] - lse;
      q[i] = o;
      if (MODE == SM_DENSE_XENT) dense += __ldg(aux + row * r + i) * o;
    }
  }
  if (MODE == SM_SPARSE_XENT && threadIdx.x == 0) {
    float tf = __ldg(aux + row); int64_t t = (int64_t)tf;
    if (tf < 0.0f || t >= r || tf != tf) { atomicExch(err, 1); loss[row] = nanf(""); }
    else loss[row] = -(src[t] - lse);
  }
  if (MODE == SM_DENSE_XENT) {
    __shared__ float redf[32];
    dense = block_sum(dense, redf); if (threadIdx.x == 0) loss[row] = -dense;
  }
}

// ---- very long rows (64 KB < row <= 8 x 64 KB): one thread-block CLUSTER per row.  Each CTA stages its slice of the row in its own shared
//      memory (three 64 KB slices per SM, so the load / reduce / store phases of different rows overlap), the per-CTA (max, sum) pairs are
//      exchanged through distributed shared memory, and HBM sees the row exactly once — the L2 re-read form above ran at 0.63-0.66 of the
//      HBM rate on 32 k .. 128 k-column rows.
template <int MODE>
__global__ void __launch_bounds__(512) softmax_cluster_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                              const float* __restrict__ aux, float* __restrict__ loss,
                                                              int64_t r, int seg, int* err) {
  extern __shared__ __align__(16) float rowbuf[];
  __shared__ MaxSum red[32];
  __shared__ MaxSum mine;
  __shared__ float dense_mine;
  uint32_t cs, rank;
  asm volatile("mov.u32 %0, %%cluster_nctaid.x;" : "=r"(cs));
  asm volatile("mov.u32 %0, %%cluster_ctaid.x;" : "=r"(rank));
  const int64_t row = blockIdx.x / cs;
  const int64_t lo = (int64_t)rank * seg;
  int64_t n = r - lo; if (n > seg) n = seg; if (n < 0) n = 0;           // this CTA's slice: [lo, lo + n)
  const float* p = x + row * r + lo;
  const int T = blockDim.x;
  const bool vec = ((((uintptr_t)p) & 15) == 0) && (n % 4 == 0);
  MaxSum acc; acc.m = -FLT_MAX; acc.s = 0.0f;
  if (vec) {
    const int64_t n4 = n >> 2;
    int64_t i = threadIdx.x;
    for (; i + 3 * T < n4; i += 4 * T) {
      float4 v0 = ldg_stream4(p + 4 * i), v1 = ldg_stream4(p + 4 * (i + T)), v2 = ldg_stream4(p + 4 * (i + 2 * T)), v3 = ldg_stream4(p + 4 * (i + 3 * T));
      *(float4*)(rowbuf + 4 * i) = v0; *(float4*)(rowbuf + 4 * (i + T)) = v1; *(float4*)(rowbuf + 4 * (i + 2 * T)) = v2; *(float4*)(rowbuf + 4 * (i + 3 * T)) = v3;
      ms_push4(acc, v0); ms_push4(acc, v1); ms_push4(acc, v2); ms_push4(acc, v3);
    }
    for (; i < n4; i += T) { float4 v = ldg_stream4(p + 4 * i); *(float4*)(rowbuf + 4 * i) = v; ms_push4(acc, v); }
  } else {
    for (int64_t i = threadIdx.x; i < n; i += T) { float v = __ldg(p + i); rowbuf[i] = v; ms_push1(acc, v); }
  }
  acc = block_maxsum(acc, red);
  if (threadIdx.x == 0) mine = acc;
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  MaxSum all; all.m = -FLT_MAX; all.s = 0.0f;
  {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(&mine);
    for (uint32_t k = 0; k < cs; k++) {                                   // same order in every CTA: all of them get the same bits
      uint32_t ra; MaxSum t;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(k));
      asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(t.m) : "r"(ra) : "memory");
      asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(t.s) : "r"(ra + 4) : "memory");
      all = ms_comb(all, t);
    }
  }
  const float mx = all.m, s = all.s;
  const float lse = logf(s) + mx, inv = 1.0f / s;
  if (MODE == SM_LSE) {
    if (threadIdx.x == 0 && rank == 0) y[row] = lse;
  } else {
    float* q = y + row * r + lo;
    float dense = 0.0f;
    if (vec && ((((uintptr_t)q) & 15) == 0) && MODE != SM_DENSE_XENT) {
      for (int64_t i = threadIdx.x; i < (n >> 2); i += T) {
        float4 v = *(const float4*)(rowbuf + 4 * i);
        if (MODE == SM_SOFTMAX) { v.x = __expf(v.x - mx) * inv; v.y = __expf(v.y - mx) * inv; v.z = __expf(v.z - mx) * inv; v.w = __expf(v.w - mx) * inv; }
        else { v.x -= lse; v.y -= lse; v.z -= lse; v.w -= lse; }
        stg_stream4(q + 4 * i, v);
      }
    } else {
      for (int64_t i = threadIdx.x; i < n; i += T) {
        float o = (MODE == SM_SOFTMAX) ? __expf(rowbuf[i] - mx) * inv : rowbuf[i] - lse;
        q[i] = o;
        if (MODE == SM_DENSE_XENT) dense += __ldg(aux + row * r + lo + i) * o;
      }
    }
    if (MODE == SM_SPARSE_XENT && threadIdx.x == 0) {
      float tf = __ldg(aux + row); int64_t t = (int64_t)tf;
      if (tf < 0.0f || t >= r || tf != tf) { if (rank == 0) { atomicExch(err, 1); loss[row] = nanf(""); } }
      else if (t >= lo && t < lo + n) loss[row] = -(rowbuf[t - lo] - lse);
    }
    if (MODE == SM_DENSE_XENT) {
      __shared__ float redf[32];
      dense = block_sum(dense, redf);
      if (threadIdx.x == 0) dense_mine = dense;
      asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
      if (threadIdx.x == 0 && rank == 0) {
        const uint32_t a = (uint32_t)__cvta_generic_to_shared(&dense_mine);
        float tot = 0.0f;
        for (uint32_t k = 0; k < cs; k++) {
          uint32_t ra; float t;
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(k));
          asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(t) : "r"(ra) : "memory");
          tot += t;
        }
        loss[row] = -tot;
      }
    }
  }
  // no CTA may exit while a peer can still read its shared memory
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- inner > 1: one thread per (outer, inner) column, three strided passes (coalesced across inner) ----
template <int MODE>
__global__ void __launch_bounds__(256) softmax_cols_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t r, int64_t inner) {
  int64_t col = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (col >= inner) return;
  int64_t o = blockIdx.y;
  const float* p = x + o * r * inner + col;
  float mx = -FLT_MAX;
  for (int64_t k = 0; k < r; k++) mx = fmaxf(mx, __ldg(p + k * inner));
  float s = 0.0f;
  for (int64_t k = 0; k < r; k++) s += expf(__ldg(p + k * inner) - mx);
  float lse = logf(s) + mx;
  if (MODE == SM_LSE) { y[o * inner + col] = lse; return; }
  float* q = y + o * r * inner + col;
  for (int64_t k = 0; k < r; k++) q[k * inner] = sm_out<MODE>(__ldg(p + k * inner), mx, s, lse);
}

template <int MODE>
static int softmax_rows(agb_ctx* ctx, const float* x, float* y, const float* aux, float* loss, int64_t rows, int64_t r) {
  if (rows == 0) return AGB_OK;
  AGB_CHECK(r > 0, AGB_ERR_INCOMPATIBLE_SHAPE, "softmax: reduction axis has length 0");
  const bool vec = r % 4 == 0 && ((((uintptr_t)x | (uintptr_t)y) & 15) == 0) && (MODE != SM_DENSE_XENT || (((uintptr_t)aux) & 15) == 0);
  if (r <= 32 * 4 && !(vec && r > 32)) {
    unsigned blocks = (unsigned)((rows + 7) / 8);
    if (r <= 32) softmax_warp_kernel<MODE, 1><<<blocks, 256, 0, ctx->stream>>>(x, y, aux, loss, rows, (int)r, ctx->dev_err);
    else softmax_warp_kernel<MODE, 4><<<blocks, 256, 0, ctx->stream>>>(x, y, aux, loss, rows, (int)r, ctx->dev_err);
    AGB_LAUNCHED(ctx); return AGB_OK;
  }
  if (r <= 1024 || (vec && r <= 2048)) {
    unsigned blocks = (unsigned)((rows + 7) / 8);
    if (vec && r <= 128) softmax_warp4_kernel<MODE, 1><<<blocks, 256, 0, ctx->stream>>>(x, y, aux, loss, rows, (int)r, ctx->dev_err);
    else if (vec && r <= 256) softmax_warp4_kernel<MODE, 2><<<blocks, 256, 0, ctx->stream>>>(x, y, aux, loss, rows, (int)r, ctx->dev_err);
    else if (vec && r <= 512) softmax_warp4_kernel<MODE, 4><<<blocks, 256, 0, ctx->stream>>>(x, y, aux, loss, rows, (int)r, ctx->dev_err);
    else if (vec && r <= 1024) softmax_warp4_kernel<MODE, 8><<<blocks, 256, 0, ctx->stream>>>(x, y, aux, loss, rows, (int)r, ctx->dev_err);
    else if (vec) softmax_warp4_kernel<MODE, 16><<<blocks, 256, 0, ctx->stream>>>(x, y, aux, loss, rows, (int)r, ctx->dev_err);
    else softmax_warp_kernel<MODE, 32><<<blocks, 256, 0, ctx->stream>>>(x, y, aux, loss, rows, (int)r, ctx->dev_err);
    AGB_LAUNCHED(ctx); return AGB_OK;
  }
  AGB_CHECK(rows < (1ll << 31), AGB_ERR_UNSUPPORTED, "softmax: too many rows");
  int threads = r >= 8192 ? 512 : 256;
  size_t smem = (size_t)r * sizeof(float);
  if (smem <= 64 * 1024) {           // rows up to 16 k columns stay in shared memory (three 64 KB rows per SM): one trip through L2 instead of two
    static bool attr = false;
    if (!attr) { AGB_CUDA(cudaFuncSetAttribute(softmax_block_kernel<MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)); attr = true; }
    softmax_block_kernel<MODE, true><<<(unsigned)rows, threads, smem, ctx->stream>>>(x, y, aux, loss, r, ctx->dev_err);
  } else if (smem <= 8 * 64 * 1024 && rows * 8 < (1ll << 31)) {      // a cluster of 2 / 4 / 8 CTAs per row, 64 KB slices
    const int cs = smem <= 2 * 64 * 1024 ? 2 : smem <= 4 * 64 * 1024 ? 4 : 8;
    const int seg = (int)((((r + cs - 1) / cs) + 3) & ~(int64_t)3);
    static bool attr = false;
    if (!attr) { AGB_CUDA(cudaFuncSetAttribute(softmax_cluster_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)); attr = true; }
    cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(rows * cs)); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = (size_t)seg * sizeof(float); cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    AGB_CUDA(cudaLaunchKernelEx(&cfg, softmax_cluster_kernel<MODE>, x, y, aux, loss, r, seg, ctx->dev_err));
  } else {
    softmax_block_kernel<MODE, false><<<(unsigned)rows, 512, 0, ctx->stream>>>(x, y, aux, loss, r, ctx->dev_err);
  }
  AGB_LAUNCHED(ctx);
  return AGB_OK;
}

template <int MODE>
static int softmax_any(agb_ctx* ctx, const float* x, float* y, int64_t outer, int64_t r, int64_t inner) {
  if (outer * inner == 0) return AGB_OK;
  if (inner == 1) return softmax_rows<MODE>(ctx, x, y, nullptr, nullptr, outer, r);
  AGB_CHECK(outer <= 65535, AGB_ERR_UNSUPPORTED, "softmax: outer too large for the strided path");
  AGB_CHECK(r > 0, AGB_ERR_INCOMPATIBLE_SHAPE, "softmax: reduction axis has length 0");
  softmax_cols_kernel<MODE><<<dim3((unsigned)((inner + 255) / 256), (unsigned)outer), 256, 0, ctx->stream>>>(x, y, r, inner);
  AGB_LAUNCHED(ctx);
  return AGB_OK;
}

extern "C" int agb_softmax(agb_ctx* ctx, const float* x, float* y, int64_t outer, int64_t r, int64_t inner) {
  return softmax_any<SM_SOFTMAX>(ctx, x, y, outer, r, inner);
}
extern "C" int agb_log_softmax(agb_ctx* ctx, const float* x, float* y, int64_t outer, int64_t r, int64_t inner) {
  return softmax_any<SM_LOGSOFTMAX>(ctx, x, y, outer, r, inner);
}
extern "C" int agb_logsumexp(agb_ctx* ctx, const float* x, float* y, int64_t outer, int64_t r, int64_t inner) {
  return softmax_any<SM_LSE>(ctx, x, y, outer, r, inner);
}
extern "C" int agb_sparse_xent_fwd(agb_ctx* ctx, const float* logits, const float* labels, float* loss, float* log_x,
                                   int64_t batch, int64_t classes) {
  return softmax_rows<SM_SPARSE_XENT>(ctx, logits, log_x, labels, loss, batch, classes);
}
extern "C" int agb_softmax_xent_fwd(agb_ctx* ctx, const float* logits, const float* t, float* loss, float* log_x,
                                    int64_t batch, int64_t classes) {
  return softmax_rows<SM_DENSE_XENT>(ctx, logits, log_x, t, loss, batch, classes);
}

// gx[b,c] = (exp(log_x[b,c]) - (c == t[b])) * gy[b]
__global__ void __launch_bounds__(256) sparse_xent_bwd_kernel(const float* __restrict__ log_x, const float* __restrict__ labels,
                                                              const float* __restrict__ gy, int64_t gy_len, float* __restrict__ gx,
                                                              int64_t batch, int64_t classes) {
  int64_t n = batch * classes;
  int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = tid; i < n; i += stride) {
    int64_t b = i / classes, c = i - b * classes;
    float g = __ldg(gy + (gy_len == 1 ? 0 : b));
    int64_t t = (int64_t)__ldg(labels + b);
    float v = expf(__ldg(log_x + i));
    if (c == t) v -= 1.0f;
    gx[i] = v * g;
  }
}
extern "C" int agb_sparse_xent_bwd(agb_ctx* ctx, const float* log_x, const float* labels, const float* gy, int64_t gy_len,
                                   float* gx, int64_t batch, int64_t classes) {
  AGB_CHECK(gy_len == 1 || gy_len == batch, AGB_ERR_INCOMPATIBLE_SHAPE, "sparse_xent_bwd: gy must have 1 or batch elements (got %lld)", (long long)gy_len);
  int64_t n = batch * classes; if (n == 0) return AGB_OK;
  sparse_xent_bwd_kernel<<<agb_grid_for(n, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(log_x, labels, gy, gy_len, gx, batch, classes);
  AGB_LAUNCHED(ctx);
  return AGB_OK;
}
