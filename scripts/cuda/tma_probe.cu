// Stand-alone TMA probe: which tensor-map ranks / box shapes / swizzles does cp.async.bulk.tensor accept on this part?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_probe tma_probe.cu -lcuda ; run: ./tma_probe <case>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int R>
__global__ void k(const __grid_constant__ CUtensorMap tm, float* out, int bytes, int c0, int c1, int c2, int c3, int c4) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(&bar)), "r"(bytes) : "memory");
    if (R == 3) asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" :: "r"(s32(sm)), "l"(&tm), "r"(s32(&bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    if (R == 4) asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" :: "r"(s32(sm)), "l"(&tm), "r"(s32(&bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    if (R == 5) asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" :: "r"(s32(sm)), "l"(&tm), "r"(s32(&bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
    uint32_t ok = 0; long long t0 = clock64();
    while (!ok && clock64() - t0 < 200000000ll)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(&bar)) : "memory");
    out[0] = ok ? 1.f : -1.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[1 + i] = ((float*)sm)[i];
}
int main(int argc, char** argv) {
  int which = argc > 1 ? atoi(argv[1]) : 0;
  cudaSetDevice(0); cudaFree(0);
  const int W = 64, H = 48, C = 160, B = 3; size_t n = (size_t)W * H * C * B;
  float* h = (float*)malloc(n * 4); for (size_t i = 0; i < n; i++) h[i] = (float)(i % 100003);
  float *d, *out; cudaMalloc(&d, n * 4); cudaMemcpy(d, h, n * 4, cudaMemcpyHostToDevice); cudaMalloc(&out, 4 + 65536);
  CUtensorMap tm; cuuint64_t dims[5], str[4]; cuuint32_t box[5], es[5] = {1, 1, 1, 1, 1}; int rank; CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B;
  int c[5] = {0, 0, 0, 0, 0};
  switch (which) {
    case 0: rank = 3; dims[0] = W; dims[1] = H; dims[2] = (uint64_t)C * B; str[0] = W * 4; str[1] = (uint64_t)W * H * 4; box[0] = 32; box[1] = 1; box[2] = 128; c[0] = 0; c[1] = 5; c[2] = 7; break;
    case 1: rank = 4; dims[0] = W; dims[1] = H; dims[2] = C; dims[3] = B; str[0] = W * 4; str[1] = (uint64_t)W * H * 4; str[2] = (uint64_t)W * H * C * 4; box[0] = 32; box[1] = 1; box[2] = 128; box[3] = 1; c[1] = 5; c[2] = 7; c[3] = 1; break;
    case 2: rank = 4; dims[0] = W; dims[1] = H; dims[2] = C; dims[3] = B; str[0] = W * 4; str[1] = (uint64_t)W * H * 4; str[2] = (uint64_t)W * H * C * 4; box[0] = 32; box[1] = 2; box[2] = 64; box[3] = 1; c[1] = 5; c[2] = 7; c[3] = 1; break;
    case 3: rank = 4; sw = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B; dims[0] = W; dims[1] = C; dims[2] = H; dims[3] = B; str[0] = (uint64_t)W * H * 4; str[1] = W * 4; str[2] = (uint64_t)W * H * C * 4; box[0] = 32; box[1] = 32; box[2] = 4; box[3] = 1; c[0] = -1; c[1] = 32; c[2] = -1; c[3] = 2; break;
    case 4: rank = 4; sw = CU_TENSOR_MAP_SWIZZLE_NONE; dims[0] = W; dims[1] = H; dims[2] = C; dims[3] = B; str[0] = W * 4; str[1] = (uint64_t)W * H * 4; str[2] = (uint64_t)W * H * C * 4; box[0] = 32; box[1] = 1; box[2] = 128; box[3] = 1; c[1] = 5; c[2] = 7; c[3] = 1; break;
    case 5: rank = 4; dims[0] = W; dims[1] = H; dims[2] = C; dims[3] = B; str[0] = W * 4; str[1] = (uint64_t)W * H * 4; str[2] = (uint64_t)W * H * C * 4; box[0] = 32; box[1] = 1; box[2] = 128; box[3] = 1; c[0] = -1; c[1] = -1; c[2] = 100; c[3] = 2; break;
    case 6: case 7: case 8: case 9: rank = 4; dims[0] = W; dims[1] = H; dims[2] = C; dims[3] = B; str[0] = W * 4; str[1] = (uint64_t)W * H * 4; str[2] = (uint64_t)W * H * C * 4; box[0] = 32; box[1] = 1; box[2] = 128; box[3] = 1;
      c[0] = which == 6 ? 3 : which == 7 ? 4 : which == 8 ? -4 : 0; c[1] = which == 9 ? -1 : 5; c[2] = 7; c[3] = 1; break;
    case 10: rank = 4; sw = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B; dims[0] = W; dims[1] = C; dims[2] = H; dims[3] = B; str[0] = (uint64_t)W * H * 4; str[1] = W * 4; str[2] = (uint64_t)W * H * C * 4; box[0] = 32; box[1] = 32; box[2] = 4; box[3] = 1; c[0] = 0; c[1] = 32; c[2] = -1; c[3] = 2; break;
    case 11: rank = 4; sw = CU_TENSOR_MAP_SWIZZLE_NONE; dims[0] = W; dims[1] = H; dims[2] = C; dims[3] = B; str[0] = W * 4; str[1] = (uint64_t)W * H * 4; str[2] = (uint64_t)W * H * C * 4; box[0] = 32; box[1] = 1; box[2] = 128; box[3] = 1; c[0] = 3; c[1] = 5; c[2] = 7; c[3] = 1; break;
    default: return 2;
  }
  uint32_t bytes = 4; for (int i = 0; i < rank; i++) bytes *= box[i];
  CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("case %d: encode=%d bytes=%u\n", which, (int)r, bytes); if (r) return 1;
  cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000); cudaFuncSetAttribute(k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  if (rank == 3) k<3><<<1, 128, 66000>>>(tm, out, bytes, c[0], c[1], c[2], c[3], c[4]);
  else k<4><<<1, 128, 66000>>>(tm, out, bytes, c[0], c[1], c[2], c[3], c[4]);
  cudaError_t e = cudaDeviceSynchronize();
  float res[9]; cudaMemcpy(res, out, 36, cudaMemcpyDeviceToHost);
  printf("case %d: sync=%d (%s) done=%g first=%g %g %g %g\n", which, (int)e, cudaGetErrorString(e), res[0], res[1], res[2], res[3], res[4]);
  return e != cudaSuccess;
}
