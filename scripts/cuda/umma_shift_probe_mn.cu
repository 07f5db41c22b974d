// Stand-alone probe (companion of umma_shift_probe.cu): MN-major TF32 operands (UMMA layout 128B_BASE32B, TMA SWIZZLE_128B_ATOM_32B).
// Can the A operand start at a K-row (pixel) that is not a multiple of the 4-row swizzle atom / 8-row k-step, and may the M-dimension
// box stride (LBO) be something other than the dense 4096 B?  This is what an all-taps-per-CTA wgrad needs: one haloed window of x in
// shared memory, each tap a descriptor shifted by dx pixel rows, boxes of (32 + halo) rows.
//   D[m, n] = sum_{k<32} A[k + shift][m] * Q[k][n],  A: [KR rows][128 m] as 4 boxes [KR][32] at stride BOXS,  Q: [32][64] dense MN-major
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I../../rust-autograd_b200/csrc -o umma_shift_probe_mn umma_shift_probe_mn.cu -lcuda
#include "tc_common.cuh"
#include <stdio.h>
#include <vector>
#include <math.h>
void agb_set_error(const char*, ...) {}
#define KR 48
#define BOXS (KR * 128)          // 6144 B, 1024-aligned
__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmQ, float* out, int shift) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem; uint8_t* sQ = smem + 4 * BOXS;
  __shared__ uint64_t full, done; __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&full, 1); mbar_init(&done, 1); fence_barrier_init(); }
  if (warp == 1) { tmem_alloc(&tslot, 64); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tslot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&full, 4 * BOXS + 2 * 4096);
    for (int g = 0; g < 4; g++) tma_load_3d(sA + g * BOXS, &tmA, &full, 32 * g, 0, 0);
    for (int g = 0; g < 2; g++) tma_load_3d(sQ + g * 4096, &tmQ, &full, 32 * g, 0, 0);
    mbar_wait(&full, 0);
    tc_fence_after();
    constexpr uint32_t idesc = umma_idesc_tf32(128, 64, 1, 1);
    for (int k = 0; k < 4; k++) {
      const uint64_t dA = umma_smem_desc(smem_u32(sA) + shift * 128 + k * 1024, BOXS, 512, 1);
      const uint64_t dQ = umma_smem_desc(smem_u32(sQ) + k * 1024, 4096, 512, 1);
      umma_tf32(tmem, dA, dQ, idesc, k != 0);
    }
    umma_commit(&done);
  }
  mbar_wait(&done, 0);
  tc_fence_after();
  const uint32_t tl = tmem + ((uint32_t)(32 * warp) << 16);
  for (int c0 = 0; c0 < 64; c0 += 32) {
    float v[32]; tmem_ld32(tl + c0, v); tmem_ld_wait();
    for (int j = 0; j < 32; j++) out[(32 * warp + lane) * 64 + c0 + j] = v[j];
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 64);
}
int main() {
  cudaSetDevice(0); cudaFree(0);
  std::vector<float> A(KR * 128), Q(32 * 64);
  for (int k = 0; k < KR; k++) for (int m = 0; m < 128; m++) A[k * 128 + m] = (float)((k * 7 + m * 3) % 17 - 8);
  for (int k = 0; k < 32; k++) for (int n = 0; n < 64; n++) Q[k * 64 + n] = (float)((n * 5 + k) % 13 - 6);
  float *dA, *dQ, *dO; cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dQ, Q.size() * 4); cudaMalloc(&dO, 128 * 64 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dQ, Q.data(), Q.size() * 4, cudaMemcpyHostToDevice);
  CUtensorMap tmA, tmQ;
  uint64_t dimsA[3] = {128, KR, 1}, strA[2] = {128 * 4, (uint64_t)KR * 128 * 4}; uint32_t boxA[3] = {32, KR, 1};
  uint64_t dimsQ[3] = {64, 32, 1}, strQ[2] = {64 * 4, 32 * 64 * 4}; uint32_t boxQ[3] = {32, 32, 1};
  if (agb_make_tmap(&tmA, dA, 3, dimsA, strA, boxA, true) || agb_make_tmap(&tmQ, dQ, 3, dimsQ, strQ, boxQ, true)) { printf("tmap failed\n"); return 1; }
  const int smem = 4 * BOXS + 2 * 4096 + 2048;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> O(128 * 64);
  const int shifts[] = {0, 1, 2, 3, 4, 5, 7, 8, 9, 12, 16};
  for (int s : shifts) {
    cudaMemset(dO, 0, O.size() * 4);
    probe<<<1, 128, smem>>>(tmA, tmQ, dO, s);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("shift %d: CUDA error %s\n", s, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int m = 0; m < 128; m++) for (int n = 0; n < 64; n++) {
      double ref = 0; for (int k = 0; k < 32; k++) ref += (double)A[(k + s) * 128 + m] * Q[k * 64 + n];
      double err = fabs(ref - O[m * 64 + n]); if (err > maxerr) maxerr = err; if (err > 1e-3) bad++;
    }
    printf("MN-major shift %3d (LBO %d): max_err %.3g bad %d/8192\n", s, BOXS, maxerr, bad);
  }
  return 0;
}
