// optim.cu — fused multi-tensor optimizer updates: one launch updates every parameter tensor
// (the reference evaluates one AdamOp node per variable, 5 passes over memory each).
// HBM roofline: Adam 28 B/param (read p,g,m,v; write p,m,v), SGD 12, Momentum 20, AdaGrad 20.
//
// Reference semantics followed:
//   AdamOp::compute      src/tensor_ops/gradient_descent_ops/adam.rs:11-58
//     m = m*b1 + (1-b1)*g ; v = v*b2 + (1-b2)*g*g
//     m_hat = m * (1/(1-b1^t)) ; v_hat = v * (1/(1-b2^t)) ; p -= alpha * (m_hat / (sqrt(v_hat) + eps)) ; t += 1
//     (t is a per-variable 0-d tensor starting at 1.0, src/optimizers/adam.rs:97)
//   SGDOp / MomentumSGDOp  src/tensor_ops/gradient_descent_ops/sgd.rs:14-40
//   AdaGradOp              src/tensor_ops/gradient_descent_ops/adagrad.rs:8-21
// `grad_scale` multiplies g on read: 1/world after the NCCL gradient sum, 1.0 otherwise.
#include "common.cuh"

#define OPT_MAX_T 40
#define OPT_CHUNK 4096   // elements per block: 256 threads x 4 float4

struct OptArgs {
  float* p[OPT_MAX_T]; const float* g[OPT_MAX_T]; float* s0[OPT_MAX_T]; float* s1[OPT_MAX_T]; float* t[OPT_MAX_T];
  int64_t n[OPT_MAX_T];
  int chunk_begin[OPT_MAX_T + 1];
  int count;
};
enum { OPT_ADAM = 0, OPT_SGD, OPT_MOMENTUM, OPT_ADAGRAD };

template <int KIND>
__device__ __forceinline__ void opt_elem(float& p, float g, float& s0, float& s1, float h0, float h1, float h2, float h3, float c1, float c2) {
  if (KIND == OPT_ADAM) {          // h0=alpha h1=eps h2=b1 h3=b2 ; c1 = 1/(1-b1^t), c2 = 1/(1-b2^t)
    s0 = s0 * h2 + (1.0f - h2) * g;
    s1 = s1 * h3 + (1.0f - h3) * g * g;
    float m_hat = s0 * c1, v_hat = s1 * c2;
    p -= h0 * (m_hat / (sqrtf(v_hat) + h1));
  } else if (KIND == OPT_SGD) {    // h0=alpha
    p -= h0 * g;
  } else if (KIND == OPT_MOMENTUM) { // h0=lr h1=momentum ; s0 = v
    s0 = s0 * h1 - h0 * g;
    p += s0;
  } else {                         // AdaGrad: h0=lr ; s0 = h
    s0 += g * g;
    p -= h0 * g / (sqrtf(s0) + 1e-7f);
  }
}

template <int KIND>
__global__ void __launch_bounds__(256) multi_tensor_kernel(OptArgs A, float h0, float h1, float h2, float h3, float gscale) {
  // locate this block's tensor
  int ti = 0;
  while (ti + 1 < A.count && (int)blockIdx.x >= A.chunk_begin[ti + 1]) ti++;
  int64_t start = (int64_t)(blockIdx.x - A.chunk_begin[ti]) * OPT_CHUNK;
  int64_t n = A.n[ti];
  int64_t end = start + OPT_CHUNK; if (end > n) end = n;
  float* p = A.p[ti]; const float* g = A.g[ti]; float* s0 = A.s0[ti]; float* s1 = A.s1[ti];
  float c1 = 1.0f, c2 = 1.0f;
  if (KIND == OPT_ADAM) {
    float t = __ldg(A.t[ti]);
    c2 = 1.0f / (1.0f - powf(h3, t));
    c1 = 1.0f / (1.0f - powf(h2, t));
  }
  bool vec = ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)(s0 ? s0 : p) | (uintptr_t)(s1 ? s1 : p)) & 15) == 0);
  if (vec) {
    int64_t i = start + 4 * threadIdx.x;
    for (; i + 3 < end; i += 4 * 256) {
      float4 pv = *(float4*)(p + i); float4 gv = ldg_stream4(g + i);
      float4 av = make_float4(0, 0, 0, 0), bv = make_float4(0, 0, 0, 0);
      if (KIND != OPT_SGD) av = *(float4*)(s0 + i);
      if (KIND == OPT_ADAM) bv = *(float4*)(s1 + i);
      opt_elem<KIND>(pv.x, gv.x * gscale, av.x, bv.x, h0, h1, h2, h3, c1, c2);
      opt_elem<KIND>(pv.y, gv.y * gscale, av.y, bv.y, h0, h1, h2, h3, c1, c2);
      opt_elem<KIND>(pv.z, gv.z * gscale, av.z, bv.z, h0, h1, h2, h3, c1, c2);
      opt_elem<KIND>(pv.w, gv.w * gscale, av.w, bv.w, h0, h1, h2, h3, c1, c2);
      *(float4*)(p + i) = pv;
      if (KIND != OPT_SGD) *(float4*)(s0 + i) = av;
      if (KIND == OPT_ADAM) *(float4*)(s1 + i) = bv;
    }
    // tail (< 4 elements of this chunk), handled by the thread that would own it
    if (i < end) {
      for (int64_t k = i; k < end; k++) {
        float pv = p[k], av = (KIND != OPT_SGD) ? s0[k] : 0.f, bv = (KIND == OPT_ADAM) ? s1[k] : 0.f;
        opt_elem<KIND>(pv, g[k] * gscale, av, bv, h0, h1, h2, h3, c1, c2);
        p[k] = pv; if (KIND != OPT_SGD) s0[k] = av; if (KIND == OPT_ADAM) s1[k] = bv;
      }
    }
  } else {
    for (int64_t k = start + threadIdx.x; k < end; k += 256) {
      float pv = p[k], av = (KIND != OPT_SGD) ? s0[k] : 0.f, bv = (KIND == OPT_ADAM) ? s1[k] : 0.f;
      opt_elem<KIND>(pv, g[k] * gscale, av, bv, h0, h1, h2, h3, c1, c2);
      p[k] = pv; if (KIND != OPT_SGD) s0[k] = av; if (KIND == OPT_ADAM) s1[k] = bv;
    }
  }
}

struct BumpArgs { float* t[OPT_MAX_T]; int count; };
__global__ void adam_bump_t_kernel(BumpArgs B) {
  int i = threadIdx.x;
  if (i < B.count) *B.t[i] += 1.0f;       // adam.rs:52-54
}

template <int KIND>
static int launch_multi(agb_ctx* ctx, int n, float* const* p, const float* const* g, float* const* s0, float* const* s1,
                        float* const* t, const int64_t* sizes, float h0, float h1, float h2, float h3, float gscale) {
  AGB_CHECK(n >= 0, AGB_ERR_INVALID_DIMS, "multi_tensor: negative tensor count");
  for (int base = 0; base < n; base += OPT_MAX_T) {
    OptArgs A; memset(&A, 0, sizeof(A));
    int m = n - base < OPT_MAX_T ? n - base : OPT_MAX_T;
    int chunks = 0;
    for (int i = 0; i < m; i++) {
      A.p[i] = p[base + i]; A.g[i] = g[base + i];
      A.s0[i] = s0 ? s0[base + i] : nullptr; A.s1[i] = s1 ? s1[base + i] : nullptr; A.t[i] = t ? t[base + i] : nullptr;
      A.n[i] = sizes[base + i];
      A.chunk_begin[i] = chunks;
      chunks += (int)((sizes[base + i] + OPT_CHUNK - 1) / OPT_CHUNK);
      // every tensor gets at least one block so that the lookup loop stays simple
      if (sizes[base + i] == 0) chunks += 1;
    }
    A.chunk_begin[m] = chunks; A.count = m;
    if (chunks == 0) continue;
    multi_tensor_kernel<KIND><<<chunks, 256, 0, ctx->stream>>>(A, h0, h1, h2, h3, gscale);
    AGB_LAUNCHED(ctx);
    if (KIND == OPT_ADAM) {
      BumpArgs B; B.count = m; for (int i = 0; i < m; i++) B.t[i] = t[base + i];
      adam_bump_t_kernel<<<1, OPT_MAX_T, 0, ctx->stream>>>(B);
      AGB_LAUNCHED(ctx);
    }
  }
  return AGB_OK;
}

extern "C" int agb_multi_tensor_adam(agb_ctx* ctx, int n, float* const* p, const float* const* g, float* const* m,
                                     float* const* v, float* const* t, const int64_t* sizes,
                                     float alpha, float eps, float b1, float b2, float grad_scale) {
  return launch_multi<OPT_ADAM>(ctx, n, p, g, m, v, t, sizes, alpha, eps, b1, b2, grad_scale);
}
extern "C" int agb_multi_tensor_sgd(agb_ctx* ctx, int n, float* const* p, const float* const* g, const int64_t* sizes,
                                    float alpha, float grad_scale) {
  return launch_multi<OPT_SGD>(ctx, n, p, g, nullptr, nullptr, nullptr, sizes, alpha, 0.f, 0.f, 0.f, grad_scale);
}
extern "C" int agb_multi_tensor_momentum(agb_ctx* ctx, int n, float* const* p, const float* const* g, float* const* v,
                                         const int64_t* sizes, float lr, float momentum, float grad_scale) {
  return launch_multi<OPT_MOMENTUM>(ctx, n, p, g, v, nullptr, nullptr, sizes, lr, momentum, 0.f, 0.f, grad_scale);
}
extern "C" int agb_multi_tensor_adagrad(agb_ctx* ctx, int n, float* const* p, const float* const* g, float* const* h,
                                        const int64_t* sizes, float lr, float grad_scale) {
  return launch_multi<OPT_ADAGRAD>(ctx, n, p, g, h, nullptr, nullptr, sizes, lr, 0.f, 0.f, 0.f, grad_scale);
}
