// simt_gemm.cuh — CUDA-core fp32 tiled GEMM core shared by the exact-fp32 GEMM path and the general
// (any pad/stride/dilation, any channel count) implicit-GEMM convolutions.  Operands are fetched through
// loader functors so that im2col is never materialised.  fp32 FMA accumulation, like the reference's
// matrixmultiply::sgemm microkernel (src/tensor_ops/dot_ops.rs:383-422).
#pragma once
#include "common.cuh"

// AL: float load(int z, int64 m, int64 k) const;  static const bool K_CONTIG (consecutive k are adjacent in memory)
// BL: float load(int z, int64 k, int64 n) const;  static const bool K_CONTIG
// CS: void  store(int z, int64 m, int64 n, float v) const
template <int TM, class AL, class BL, class CS>
__global__ void __launch_bounds__(256) simt_gemm_kernel(AL A, BL B, CS C, int64_t M, int64_t N, int64_t K) {
  constexpr int BM = 16 * TM, BN = 16 * TM, BK = 16;
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int z = blockIdx.z;
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  constexpr int PER = BM * BK / 256;   // elements per thread per operand tile
  float ra[PER], rb[PER];
  float acc[TM][TM];
#pragma unroll
  for (int i = 0; i < TM; i++)
#pragma unroll
    for (int j = 0; j < TM; j++) acc[i][j] = 0.0f;

  auto fetch = [&](int64_t k0) {
#pragma unroll
    for (int e = 0; e < PER; e++) {
      int lin = tid + e * 256, mm, kk;
      if (AL::K_CONTIG) { kk = lin % BK; mm = lin / BK; } else { mm = lin % BM; kk = lin / BM; }
      int64_t gm = m0 + mm, gk = k0 + kk;
      ra[e] = (gm < M && gk < K) ? A.load(z, gm, gk) : 0.0f;
      int nn;
      if (BL::K_CONTIG) { kk = lin % BK; nn = lin / BK; } else { nn = lin % BN; kk = lin / BN; }
      int64_t gn = n0 + nn; gk = k0 + kk;
      rb[e] = (gn < N && gk < K) ? B.load(z, gk, gn) : 0.0f;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int e = 0; e < PER; e++) {
      int lin = tid + e * 256, mm, kk, nn;
      if (AL::K_CONTIG) { kk = lin % BK; mm = lin / BK; } else { mm = lin % BM; kk = lin / BM; }
      As[buf][kk][mm] = ra[e];
      if (BL::K_CONTIG) { kk = lin % BK; nn = lin / BK; } else { nn = lin % BN; kk = lin / BN; }
      Bs[buf][kk][nn] = rb[e];
    }
  };

  int64_t nk = (K + BK - 1) / BK;
  if (nk > 0) { fetch(0); stash(0); }
  __syncthreads();
  for (int64_t kb = 0; kb < nk; kb++) {
    int buf = (int)(kb & 1);
    if (kb + 1 < nk) fetch((kb + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; k++) {
      float a[TM], b[TM];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        float4 v = *(const float4*)&As[buf][k][(i / 4) * 64 + ty * 4];
        a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
        float4 w = *(const float4*)&Bs[buf][k][(i / 4) * 64 + tx * 4];
        b[i] = w.x; b[i + 1] = w.y; b[i + 2] = w.z; b[i + 3] = w.w;
      }
#pragma unroll
      for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TM; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kb + 1 < nk) stash(buf ^ 1);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; i++) {
    int64_t gm = m0 + (i / 4) * 64 + ty * 4 + (i & 3);
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < TM; j++) {
      int64_t gn = n0 + (j / 4) * 64 + tx * 4 + (j & 3);
      if (gn < N) C.store(z, gm, gn, acc[i][j]);
    }
  }
}

template <class AL, class BL, class CS>
static int simt_gemm_launch(agb_ctx* ctx, AL A, BL B, CS C, int64_t M, int64_t N, int64_t K, int64_t Z) {
  if (M <= 0 || N <= 0 || Z <= 0) return AGB_OK;
  AGB_CHECK(Z <= 65535, AGB_ERR_UNSUPPORTED, "simt gemm: batch/split dimension too large (%lld)", (long long)Z);
  bool big = (M >= 96 && N >= 96);
  if (big) {
    dim3 grid((unsigned)((N + 127) / 128), (unsigned)((M + 127) / 128), (unsigned)Z);
    AGB_CHECK(grid.y <= 65535, AGB_ERR_UNSUPPORTED, "simt gemm: M too large");
    simt_gemm_kernel<8, AL, BL, CS><<<grid, 256, 0, ctx->stream>>>(A, B, C, M, N, K);
  } else {
    dim3 grid((unsigned)((N + 63) / 64), (unsigned)((M + 63) / 64), (unsigned)Z);
    AGB_CHECK(grid.y <= 65535, AGB_ERR_UNSUPPORTED, "simt gemm: M too large");
    simt_gemm_kernel<4, AL, BL, CS><<<grid, 256, 0, ctx->stream>>>(A, B, C, M, N, K);
  }
  AGB_LAUNCHED(ctx);
  return AGB_OK;
}
