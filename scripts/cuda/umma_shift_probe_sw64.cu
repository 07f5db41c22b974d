// (64-byte-row variant of umma_shift_probe.cu: K = 16 floats per row, TMA SWIZZLE_64B, UMMA layout SWIZZLE_64B = 4, SBO = 512)
// Stand-alone probe: can a K-major SWIZZLE_128B UMMA operand start at a row that is NOT a multiple of 8 (i.e. a start address
// that is 128-byte but not 1024-byte aligned)?  This is what a haloed implicit-GEMM conv needs: one TMA-loaded input window in
// shared memory, and every filter tap reads it through a descriptor shifted by (dy*pitch + dx) pixel rows.
//   variant 0: start address shifted, base_offset field = 0
//   variant 1: start address shifted, base_offset = (start >> 7) & 7   (PTX ISA: "matrix base offset")
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I../../rust-autograd_b200/csrc -o umma_shift_probe umma_shift_probe.cu -lcuda
#include "tc_common.cuh"
#include <stdio.h>
#include <vector>
#include <math.h>

void agb_set_error(const char*, ...) {}

#define ROWS 288          // rows of A resident in shared memory (18 KB)
#define RB 64             // bytes per row (16 floats)
__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out, int shift, int variant) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem; uint8_t* sB = smem + ((ROWS * RB + 1023) & ~1023);
  __shared__ uint64_t full, done; __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&full, 1); mbar_init(&done, 1); fence_barrier_init(); }
  if (warp == 1) { tmem_alloc(&tslot, 64); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tslot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&full, ROWS * RB + 64 * RB);
    // A is loaded as ONE box per 256-row limit: two loads of 144 rows
    tma_load_3d(sA, &tmA, &full, 0, 0, 0);
    tma_load_3d(sA + 144 * RB, &tmA, &full, 0, 144, 0);
    tma_load_3d(sB, &tmB, &full, 0, 0, 0);
    mbar_wait(&full, 0);
    tc_fence_after();
    constexpr uint32_t idesc = umma_idesc_tf32(128, 64, 0, 0);
    for (int k = 0; k < 2; k++) {
      const uint32_t a = smem_u32(sA) + shift * RB + k * 32;
      uint64_t dA = umma_smem_desc(a, 16, 512, 4);
      if (variant == 1) dA |= (uint64_t)((a >> 7) & 7) << 49;
      const uint64_t dB = umma_smem_desc(smem_u32(sB) + k * 32, 16, 512, 4);
      umma_tf32(tmem, dA, dB, idesc, k != 0);
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

static int mk(CUtensorMap* m, const void* base, const uint64_t* dims, const uint64_t* str, const uint32_t* box) {
  cuuint64_t d[3] = {dims[0], dims[1], dims[2]}, s[2] = {str[0], str[1]}; cuuint32_t b[3] = {box[0], box[1], box[2]}, e[3] = {1, 1, 1};
  return (int)agb_get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}
int main() {
  cudaSetDevice(0); cudaFree(0);
  std::vector<float> A(ROWS * 16), B(64 * 16);
  for (int r = 0; r < ROWS; r++) for (int k = 0; k < 16; k++) A[r * 16 + k] = (float)((r * 7 + k * 3) % 17 - 8);
  for (int n = 0; n < 64; n++) for (int k = 0; k < 16; k++) B[n * 16 + k] = (float)((n * 5 + k) % 13 - 6);
  float *dA, *dB, *dO; cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dO, 128 * 64 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  CUtensorMap tmA, tmB;
  uint64_t dimsA[3] = {16, ROWS, 1}, strA[2] = {RB, (uint64_t)ROWS * RB}; uint32_t boxA[3] = {16, 144, 1};
  uint64_t dimsB[3] = {16, 64, 1}, strB[2] = {RB, 64 * RB}; uint32_t boxB[3] = {16, 64, 1};
  if (mk(&tmA, dA, dimsA, strA, boxA) || mk(&tmB, dB, dimsB, strB, boxB)) { printf("tmap failed\n"); return 1; }
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, ROWS * 128 + 64 * 128 + 2048);
  std::vector<float> O(128 * 64);
  const int shifts[] = {0, 1, 2, 3, 4, 7, 8, 9, 13, 130, 131, 137};
  for (int variant = 0; variant < 2; variant++)
    for (int s : shifts) {
      cudaMemset(dO, 0, O.size() * 4);
      probe<<<1, 128, ROWS * 128 + 64 * 128 + 2048>>>(tmA, tmB, dO, s, variant);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("variant %d shift %d: CUDA error %s\n", variant, s, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0; int bad = 0;
      for (int r = 0; r < 128; r++) for (int n = 0; n < 64; n++) {
        double ref = 0; for (int k = 0; k < 16; k++) ref += (double)A[(r + s) * 16 + k] * B[n * 16 + k];
        double err = fabs(ref - O[r * 64 + n]); if (err > maxerr) maxerr = err; if (err > 1e-3) bad++;
      }
      printf("variant %d shift %3d: max_err %.3g bad %d/8192\n", variant, s, maxerr, bad);
    }
  return 0;
}
