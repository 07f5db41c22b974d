/* cpu_ref.c — f32 CPU restatement of the reference's DEFAULT-BUILD hot path, as a timing baseline.
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY: built into oracle/lib/libcpuref.so, loaded by oracle/cpu_ref.py, used by tests/ (checked
 * against the numpy oracle) and by bench.py's reference arm / cpu_baseline leg.  Nothing under rust-autograd_b200/ links or loads it.
 *
 * What it restates (reference file:line, crate autograd 2.0.0-rc3), with the reference's own parallel structure:
 *   conv2d forward        conv_ops/conv2d.rs:115-211   per sample: materialised im2col (conv_ops/mod.rs:73-124) + ONE sgemm
 *                                                       [O, C*kh*kw] x [C*kh*kw, yh*yw]; samples in parallel (rayon par_iter -> a pthread pool)
 *   conv2d_transpose      conv2d_transpose.rs:89-247    per sample: sgemm W^T x gy -> cols, then col2im (mod.rs:178-223); samples in parallel
 *   conv2d filter grad    conv2d.rs:631-734             SEQUENTIAL loop over the batch, sgemm gy_b x cols_b^T with beta = (b == 0 ? 0 : 1)
 *   matmul                dot_ops.rs:383-422            one single-threaded sgemm (Cargo.toml:21: matrixmultiply without `threading`)
 *   max_pool2d (+grad)    max_pool2d.rs:21-135          scalar loops, strict `>`, flat argmax offsets, scatter-add backward; single thread
 *   elementwise / reduce  binary_ops.rs:304-347, activation_ops.rs:154-166, reduction_ops.rs:54-108   one pass per op, single thread
 *   sparse softmax xent   xent_ops.rs:63-158            max / exp-sum / log passes per row, single thread
 *   Adam                  gradient_descent_ops/adam.rs:11-58   five passes over each variable, single thread
 * The sgemm microkernel: the crate's `matrixmultiply::sgemm` is not available here (no Rust toolchain); a single-threaded
 * `cblas_sgemm` of the OpenBLAS that numpy ships stands in for it when the caller provides it (cr_set_sgemm), otherwise the
 * packed AVX2 kernel below.  Either is at least as fast as matrixmultiply's, so the baseline errs on the fast side. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <pthread.h>
#include <unistd.h>

typedef void (*cblas_sgemm64_fn)(int order, int ta, int tb, int64_t m, int64_t n, int64_t k, float alpha, const float* a, int64_t lda,
                                 const float* b, int64_t ldb, float beta, float* c, int64_t ldc);
static cblas_sgemm64_fn g_sgemm = 0;
void cr_set_sgemm(void* fn) { g_sgemm = (cblas_sgemm64_fn)fn; }
/* rayon-style parallel-for over samples: a pool of `cr_threads()` workers pulls sample indices from an atomic counter */
static int g_threads = 0;
void cr_set_threads(int n) { g_threads = n; }
int cr_threads(void) {
  if (g_threads > 0) return g_threads;
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}
typedef void (*sample_fn)(int b, void* arg, int worker);
typedef struct { sample_fn fn; void* arg; int n; int* next; int worker; } par_job;
static void* par_worker(void* p) {
  par_job* j = (par_job*)p;
  for (;;) { int b = __atomic_fetch_add(j->next, 1, __ATOMIC_RELAXED); if (b >= j->n) break; j->fn(b, j->arg, j->worker); }
  return 0;
}
static void par_for(int n, sample_fn fn, void* arg) {
  int nt = cr_threads(); if (nt > n) nt = n; if (nt < 1) nt = 1;
  int next = 0; pthread_t th[256]; par_job jobs[256]; if (nt > 256) nt = 256;
  for (int t = 0; t < nt; t++) { jobs[t] = (par_job){fn, arg, n, &next, t}; if (t > 0) pthread_create(&th[t], 0, par_worker, &jobs[t]); }
  par_worker(&jobs[0]);
  for (int t = 1; t < nt; t++) pthread_join(th[t], 0);
}

/* ---- built-in row-major sgemm: C[m,n] = alpha * op(A) op(B) + beta * C; strides given per element (rs, cs) like matrixmultiply ---- */
#define MR 6
#define NR 16
#define KC 256
#define MC 96
#define NC 2048
typedef float v8 __attribute__((vector_size(32), aligned(4)));
static void micro_6x16(int kc, const float* pa, const float* pb, float* c, int64_t rsc, int64_t csc, int mr, int nr, float alpha, float beta) {
  v8 acc[MR][2];
  for (int i = 0; i < MR; i++) { acc[i][0] = (v8){0}; acc[i][1] = (v8){0}; }
  for (int p = 0; p < kc; p++) {
    v8 b0 = *(const v8*)(pb + p * NR), b1 = *(const v8*)(pb + p * NR + 8);
    for (int i = 0; i < MR; i++) { float a = pa[p * MR + i]; v8 av = {a, a, a, a, a, a, a, a}; acc[i][0] += av * b0; acc[i][1] += av * b1; }
  }
  for (int i = 0; i < mr; i++)
    for (int j = 0; j < nr; j++) {
      float v = alpha * acc[i][j >> 3][j & 7];
      float* d = c + i * rsc + j * csc;
      *d = beta == 0.0f ? v : beta * *d + v;
    }
}
static void builtin_sgemm(int64_t m, int64_t k, int64_t n, float alpha, const float* a, int64_t rsa, int64_t csa, const float* b, int64_t rsb, int64_t csb,
                          float beta, float* c, int64_t rsc, int64_t csc) {
  float* pa = (float*)aligned_alloc(64, sizeof(float) * (MC + MR) * KC);
  float* pb = (float*)aligned_alloc(64, sizeof(float) * (NC + NR) * KC);
  for (int64_t jc = 0; jc < n; jc += NC) {
    int64_t nc = n - jc < NC ? n - jc : NC;
    for (int64_t pc = 0; pc < k; pc += KC) {
      int kc = (int)(k - pc < KC ? k - pc : KC);
      float bet = pc == 0 ? beta : 1.0f;
      for (int64_t j = 0; j < nc; j += NR)                      /* pack B panel: [kc][NR] */
        for (int p = 0; p < kc; p++)
          for (int jj = 0; jj < NR; jj++) pb[(j / NR) * (int64_t)KC * NR + p * NR + jj] = j + jj < nc ? b[(pc + p) * rsb + (jc + j + jj) * csb] : 0.0f;
      for (int64_t ic = 0; ic < m; ic += MC) {
        int64_t mc = m - ic < MC ? m - ic : MC;
        for (int64_t i = 0; i < mc; i += MR)                    /* pack A panel: [kc][MR] */
          for (int p = 0; p < kc; p++)
            for (int ii = 0; ii < MR; ii++) pa[(i / MR) * (int64_t)KC * MR + p * MR + ii] = i + ii < mc ? a[(ic + i + ii) * rsa + (pc + p) * csa] : 0.0f;
        for (int64_t j = 0; j < nc; j += NR)
          for (int64_t i = 0; i < mc; i += MR)
            micro_6x16(kc, pa + (i / MR) * (int64_t)KC * MR, pb + (j / NR) * (int64_t)KC * NR, c + (ic + i) * rsc + (jc + j) * csc, rsc, csc,
                       (int)(mc - i < MR ? mc - i : MR), (int)(nc - j < NR ? nc - j : NR), alpha, bet);
      }
    }
  }
  free(pa); free(pb);
}
/* row-major C[m,n] = alpha * A' B' + beta C with A' = A (lda) or A^T, B' = B or B^T */
static void sgemm_rm(int ta, int tb, int64_t m, int64_t n, int64_t k, float alpha, const float* a, int64_t lda, const float* b, int64_t ldb, float beta, float* c, int64_t ldc) {
  if (g_sgemm) { g_sgemm(101 /*RowMajor*/, ta ? 112 : 111, tb ? 112 : 111, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc); return; }
  builtin_sgemm(m, k, n, alpha, a, ta ? 1 : lda, ta ? lda : 1, b, tb ? 1 : ldb, tb ? ldb : 1, beta, c, ldc, 1);
}
void cr_matmul(const float* a, const float* b, float* c, int64_t m, int64_t n, int64_t k, int ta, int tb) {       /* dot_ops.rs:383-422 */
  sgemm_rm(ta, tb, m, n, k, 1.0f, a, ta ? m : k, b, tb ? k : n, 0.0f, c, n);
}

/* ---- im2col / col2im of ONE sample, the loop nest of conv_ops/mod.rs:73-124,178-223 ---- */
static void im2col1(const float* x, float* cols, int C, int H, int W, int kh, int kw, int pad, int s, int d, int yh, int yw) {
  for (int c = 0; c < C; c++, x += (size_t)H * W)
    for (int i = 0; i < kh; i++)
      for (int j = 0; j < kw; j++)
        for (int oy = 0; oy < yh; oy++, cols += yw) {
          const int iy = oy * s + i * d - pad;
          if ((unsigned)iy >= (unsigned)H) { memset(cols, 0, sizeof(float) * yw); continue; }
          const float* row = x + (size_t)iy * W;
          int ix = j * d - pad;
          for (int ox = 0; ox < yw; ox++, ix += s) cols[ox] = (unsigned)ix < (unsigned)W ? row[ix] : 0.0f;
        }
}
static void col2im1(const float* cols, float* x, int C, int H, int W, int kh, int kw, int pad, int s, int d, int yh, int yw) {
  for (int c = 0; c < C; c++, x += (size_t)H * W)
    for (int i = 0; i < kh; i++)
      for (int j = 0; j < kw; j++)
        for (int oy = 0; oy < yh; oy++, cols += yw) {
          const int iy = oy * s + i * d - pad;
          if ((unsigned)iy >= (unsigned)H) continue;
          float* row = x + (size_t)iy * W;
          int ix = j * d - pad;
          for (int ox = 0; ox < yw; ox++, ix += s) if ((unsigned)ix < (unsigned)W) row[ix] += cols[ox];
        }
}
/* Conv2D::compute: y [B,O,yh,yw], cols [B, C*kh*kw, yh*yw] (kept for the filter gradient, conv2d.rs:484) */
typedef struct { const float* x; const float* w; float* y; float* cols; float** scratch; int C, H, W, O, kh, kw, pad, s, d, yh, yw; } conv_args;
static void conv_fwd_sample(int b, void* p, int worker) {
  conv_args* a = (conv_args*)p; (void)worker;
  const int64_t K = (int64_t)a->C * a->kh * a->kw, N = (int64_t)a->yh * a->yw;
  float* cb = a->cols + (size_t)b * K * N;
  im2col1(a->x + (size_t)b * a->C * a->H * a->W, cb, a->C, a->H, a->W, a->kh, a->kw, a->pad, a->s, a->d, a->yh, a->yw);
  sgemm_rm(0, 0, a->O, N, K, 1.0f, a->w, K, cb, N, 0.0f, a->y + (size_t)b * a->O * N, N);
}
void cr_conv2d(const float* x, const float* w, float* y, float* cols, int B, int C, int H, int W, int O, int kh, int kw, int pad, int s, int d) {
  const int yh = (H + 2 * pad - (d * (kh - 1) + 1)) / s + 1, yw = (W + 2 * pad - (d * (kw - 1) + 1)) / s + 1;
  conv_args a = {x, w, y, cols, 0, C, H, W, O, kh, kw, pad, s, d, yh, yw};
  par_for(B, conv_fwd_sample, &a);
}
static void conv_dgrad_sample(int b, void* p, int worker) {
  conv_args* a = (conv_args*)p;
  const int64_t K = (int64_t)a->C * a->kh * a->kw, N = (int64_t)a->yh * a->yw;
  if (!a->scratch[worker]) a->scratch[worker] = (float*)malloc(sizeof(float) * (size_t)K * N);
  float* cols = a->scratch[worker];
  sgemm_rm(1, 0, K, N, a->O, 1.0f, a->w, K, a->x + (size_t)b * a->O * N, N, 0.0f, cols, N);      /* a->x = gy, a->y = gx */
  col2im1(cols, a->y + (size_t)b * a->C * a->H * a->W, a->C, a->H, a->W, a->kh, a->kw, a->pad, a->s, a->d, a->yh, a->yw);
}
/* Conv2DTranspose::compute (the input gradient of conv2d): gy [B,O,yh,yw], w [O,C,kh,kw] -> gx [B,C,H,W] */
void cr_conv2d_transpose(const float* gy, const float* w, float* gx, int B, int C, int H, int W, int O, int kh, int kw, int pad, int s, int d) {
  const int yh = (H + 2 * pad - (d * (kh - 1) + 1)) / s + 1, yw = (W + 2 * pad - (d * (kw - 1) + 1)) / s + 1;
  const int64_t K = (int64_t)C * kh * kw, N = (int64_t)yh * yw;
  memset(gx, 0, sizeof(float) * (size_t)B * C * H * W);
  float* scratch[256] = {0};
  conv_args a = {gy, w, gx, 0, scratch, C, H, W, O, kh, kw, pad, s, d, yh, yw};
  (void)K; (void)N;
  par_for(B, conv_dgrad_sample, &a);
  for (int t = 0; t < 256; t++) free(scratch[t]);
}
/* Conv2DFilterGrad::compute: sequential over the batch with beta = 1 after the first sample (conv2d.rs:703-722) */
void cr_conv2d_filter_grad(const float* cols, const float* gy, float* gw, int B, int C, int O, int kh, int kw, int yh, int yw) {
  const int64_t K = (int64_t)C * kh * kw, N = (int64_t)yh * yw;
  for (int b = 0; b < B; b++)
    sgemm_rm(0, 1, O, K, N, 1.0f, gy + (size_t)b * O * N, N, cols + (size_t)b * K * N, N, b == 0 ? 0.0f : 1.0f, gw, K);
}

/* ---- max pooling (max_pool2d.rs:21-88,111-135), pad = 0 ---- */
void cr_max_pool2d(const float* x, float* y, float* idx, int B, int C, int H, int W, int size, int s) {
  const int yh = (H - size) / s + 1, yw = (W - size) / s + 1;
  size_t o = 0;
  for (int b = 0; b < B; b++)
    for (int c = 0; c < C; c++)
      for (int oy = 0; oy < yh; oy++)
        for (int ox = 0; ox < yw; ox++, o++) {
          float best = -FLT_MAX; size_t bi = 0;
          const int h1 = oy * s + size < H ? oy * s + size : H, w1 = ox * s + size < W ? ox * s + size : W;
          for (int h = oy * s; h < h1; h++)
            for (int w_ = ox * s; w_ < w1; w_++) {
              const size_t i = (size_t)w_ + (size_t)W * (h + (size_t)H * (c + (size_t)b * C));
              if (x[i] > best) { best = x[i]; bi = i; }
            }
          y[o] = best; idx[o] = (float)bi;
        }
}
void cr_max_pool2d_grad(const float* gy, const float* idx, float* gx, int64_t n_out, int64_t n_in) {
  memset(gx, 0, sizeof(float) * (size_t)n_in);
  for (int64_t i = 0; i < n_out; i++) gx[(size_t)idx[i]] += gy[i];
}

/* ---- elementwise / reductions: one pass per reference op ---- */
void cr_add_bias_nchw(const float* x, const float* bias, float* y, int B, int C, int64_t hw) {       /* AddOp with a [1,C,1,1] operand (broadcast) */
  for (int b = 0; b < B; b++) for (int c = 0; c < C; c++) { const float v = bias[c]; const float* s = x + ((size_t)b * C + c) * hw; float* t = y + ((size_t)b * C + c) * hw; for (int64_t i = 0; i < hw; i++) t[i] = s[i] + v; }
}
void cr_add_rowvec(const float* x, const float* bias, float* y, int64_t rows, int64_t cols) { for (int64_t r = 0; r < rows; r++) for (int64_t c = 0; c < cols; c++) y[r * cols + c] = x[r * cols + c] + bias[c]; }
void cr_relu(const float* x, float* y, int64_t n) { for (int64_t i = 0; i < n; i++) y[i] = x[i] > 0.0f ? x[i] : 0.0f; }                    /* activation_ops.rs:154-160 */
void cr_greater0(const float* x, float* y, int64_t n) { for (int64_t i = 0; i < n; i++) y[i] = x[i] > 0.0f ? 1.0f : 0.0f; }                 /* ReLU::grad = gy * greater(x, 0) (:162-166): two passes */
void cr_mul(const float* a, const float* b, float* y, int64_t n) { for (int64_t i = 0; i < n; i++) y[i] = a[i] * b[i]; }
void cr_scale(const float* a, float s, float* y, int64_t n) { for (int64_t i = 0; i < n; i++) y[i] = a[i] * s; }
void cr_sum_to_channels(const float* g, float* out, int B, int C, int64_t hw) {       /* MaybeReduceSum to [1,C,1,1] (binary_ops.rs:37-95): fold over axes 0, 2, 3 */
  for (int c = 0; c < C; c++) out[c] = 0.0f;
  for (int b = 0; b < B; b++) for (int c = 0; c < C; c++) { const float* s = g + ((size_t)b * C + c) * hw; float acc = 0.0f; for (int64_t i = 0; i < hw; i++) acc += s[i]; out[c] += acc; }
}
void cr_sum_rows(const float* g, float* out, int64_t rows, int64_t cols) { for (int64_t c = 0; c < cols; c++) out[c] = 0.0f; for (int64_t r = 0; r < rows; r++) for (int64_t c = 0; c < cols; c++) out[c] += g[r * cols + c]; }
float cr_mean(const float* x, int64_t n) { float acc = 0.0f; for (int64_t i = 0; i < n; i++) acc += x[i]; return acc / (float)n; }

/* ---- sparse softmax cross-entropy (xent_ops.rs:63-158): loss [B,1], log_x [B,C]; grad = (exp(log_x) - onehot) * gy ---- */
void cr_sparse_xent(const float* x, const float* t, float* loss, float* log_x, int64_t B, int64_t C) {
  for (int64_t b = 0; b < B; b++) {
    const float* r = x + b * C; float m = -FLT_MAX, se = 0.0f;
    for (int64_t c = 0; c < C; c++) m = r[c] > m ? r[c] : m;
    for (int64_t c = 0; c < C; c++) se += expf(r[c] - m);
    const float lse = logf(se) + m;
    for (int64_t c = 0; c < C; c++) log_x[b * C + c] = r[c] - lse;
    loss[b] = -log_x[b * C + (int64_t)t[b]];
  }
}
void cr_sparse_xent_grad(const float* log_x, const float* t, const float* gy, float* gx, int64_t B, int64_t C) {
  for (int64_t b = 0; b < B; b++) {
    for (int64_t c = 0; c < C; c++) gx[b * C + c] = expf(log_x[b * C + c]);
    gx[b * C + (int64_t)t[b]] -= 1.0f;
    for (int64_t c = 0; c < C; c++) gx[b * C + c] *= gy[b];
  }
}

/* ---- Adam (gradient_descent_ops/adam.rs:11-58): m, v updates, two bias-corrected temporaries, the parameter update: five passes ---- */
void cr_adam(float* p, const float* g, float* m, float* v, float* t, int64_t n, float alpha, float eps, float b1, float b2, float* tmp) {
  for (int64_t i = 0; i < n; i++) m[i] = m[i] * b1 + (1.0f - b1) * g[i];
  for (int64_t i = 0; i < n; i++) v[i] = v[i] * b2 + (1.0f - b2) * g[i] * g[i];
  const float rm = 1.0f / (1.0f - powf(b1, *t)), rv = 1.0f / (1.0f - powf(b2, *t));
  for (int64_t i = 0; i < n; i++) tmp[i] = m[i] * rm;                                   /* m_hat */
  for (int64_t i = 0; i < n; i++) tmp[i] = tmp[i] / (sqrtf(v[i] * rv) + eps);           /* m_hat / (sqrt(v_hat) + eps) */
  for (int64_t i = 0; i < n; i++) p[i] -= alpha * tmp[i];
  *t += 1.0f;
}
