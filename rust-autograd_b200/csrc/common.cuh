// common.cuh — shared internals of libagb200 (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <map>
#include <unordered_map>
#include <vector>
#include "../../include/agb200.h"

struct agb_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  // host-feed staging (agb_stage_*): a copy stream that runs H2D of the NEXT step's inputs under the current step's kernels
  cudaStream_t copy_stream = nullptr, d2h_stream = nullptr; cudaEvent_t stage_mark = nullptr, stage_done = nullptr; bool stage_pending = false, stage_marked = false;
  int math_mode = AGB_MATH_3XTF32;
  int64_t launches = 0;
  // stream-ordered caching arena: every block is only ever used on `stream`, so a freed block can be
  // handed out again immediately (work is ordered by the stream).
  std::multimap<size_t, void*> free_blocks;
  std::unordered_map<void*, size_t> block_size;   // every block we own (live or cached)
  std::unordered_map<void*, bool> is_live;
  size_t live_bytes = 0, cached_bytes = 0, peak_bytes = 0;
  // scratch for two-stage reductions / split-K / descriptor tables
  void* scratch = nullptr; size_t scratch_bytes = 0;
  void* scratch2 = nullptr; size_t scratch2_bytes = 0;
  std::vector<void*> retired_scratch;
  void* flush_buf = nullptr; size_t flush_bytes = 0;
  // optimizer descriptor-table cache (key = hash of pointer lists)
  std::unordered_map<uint64_t, void*> optim_tables;
  // device-side error flag (bad labels / indices)
  int* dev_err = nullptr;
  // nccl
  void* nccl_comm = nullptr;
  int rank = 0, world = 1;
  cudaStream_t comm_stream = nullptr; cudaEvent_t comm_ready = nullptr, comm_done = nullptr; bool comm_pending = false;   // bucketed all-reduce overlapped with the rest of backward
  bool capturing = false;
  // deterministic reductions (default on): split-K partial sums and per-channel side sums go to scratch and are added in a fixed order instead of
  // red.global.add / atomicAdd in arrival order, so a step is bit-reproducible run to run (the reference's loops are sequential: conv2d.rs:631-734)
  int deterministic = 1;
  // ReLU sign bits (1 bit per element, channels-last order: word (pixel * C + c) / 32, bit c % 32) travelling next to an activation: the conv entry
  // points set these for the launch they dispatch; kernels that can write / read the bits do and say so
  uint32_t* bits_out = nullptr; const uint32_t* mask_bits = nullptr; int bits_written = 0, mask_bits_used = 0;
  int pinned_graphs = 0;            // live agx_step graphs: the arena must not return blocks to the driver while they exist
  // Private memory of instantiated graphs.  A CUDA graph keeps the RAW addresses of every arena block its kernels touch; once the
  // capture ends those blocks would sit in the free list and the next eager allocation could receive one while a replay still writes
  // it.  Every block handed out during a capture is therefore recorded (`capture_blocks`); agb_graph_end moves them out of the free
  // list into the graph's reservation (blocks still live at that point follow when they are freed: `reserve_on_free`), and
  // agb_graph_destroy returns them.
  std::vector<void*> capture_blocks;
  struct GraphRes { agb_ctx* ctx; void* exec; std::vector<void*> blocks; int64_t kernel_nodes = 0; };
  std::unordered_map<void*, GraphRes*> reserve_on_free;
  // live profiler (agb_prof_*): event pairs per profiled entry-point call
  bool prof_on = false;
  struct ProfRec { int cls; cudaEvent_t a, b; double work; };
  std::vector<ProfRec> prof_recs;
  std::vector<cudaEvent_t> prof_pool;
};

// RAII bracket used by the profiled entry points: records an event pair around the launches issued in its scope
struct AgbProfScope {
  agb_ctx* ctx; int idx = -1;
  AgbProfScope(agb_ctx* c, int cls, double work) : ctx(c) {
    if (!c->prof_on || c->capturing) return;
    agb_ctx::ProfRec r; r.cls = cls; r.work = work;
    auto get = [&]() { cudaEvent_t e; if (!c->prof_pool.empty()) { e = c->prof_pool.back(); c->prof_pool.pop_back(); } else cudaEventCreate(&e); return e; };
    r.a = get(); r.b = get();
    cudaEventRecord(r.a, c->stream);
    idx = (int)c->prof_recs.size(); c->prof_recs.push_back(r);
  }
  void set_cls(int cls) { if (idx >= 0) ctx->prof_recs[idx].cls = cls; }   // the dispatcher re-labels the bracket with the kernel that ran
  ~AgbProfScope() { if (idx >= 0) cudaEventRecord(ctx->prof_recs[idx].b, ctx->stream); }
};

void agb_set_error(const char* fmt, ...);
int  agb_cuda_fail(cudaError_t e, const char* what, const char* file, int line);
int  agb_scratch(agb_ctx* ctx, size_t bytes, void** out);
int  agb_scratch2(agb_ctx* ctx, size_t bytes, void** out);      // a second buffer: partial sums of the deterministic reductions (the first may hold repacked operands of the same call)
// out[i] (+)= sum over k < nparts (in order) of part[k * stride + i]
int  agb_reduce_partials(agb_ctx* ctx, const float* part, float* out, int nparts, int64_t n, int64_t stride, int accumulate);
// partials laid out [nparts][tap][o][c] -> gw[o][c][tap]
int  agb_reduce_partials_wgrad(agb_ctx* ctx, const float* part, float* gw, int nparts, int O, int C, int T);
// many partials ([nparts][n], dense): 64 interleaved groups are added first, then the groups — both in a fixed order.  `part` must have room for
// agb_reduce_partials2_floats(nparts, n) floats (padding to a multiple of 64 partials + the [64][n] group sums)
static inline size_t agb_reduce_partials2_floats(int64_t nparts, int64_t n) { return (size_t)((nparts + 63) / 64 * 64 + 64) * (size_t)n; }
int  agb_reduce_partials2(agb_ctx* ctx, float* part, int64_t nparts, int64_t n, float* out, int accumulate);

#define AGB_CUDA(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) return agb_cuda_fail(_e, #x, __FILE__, __LINE__); } while (0)
#define AGB_CHECK(cond, code, ...) do { if (!(cond)) { agb_set_error(__VA_ARGS__); return (code); } } while (0)
#define AGB_TRY(x) do { int _r = (x); if (_r != AGB_OK) return _r; } while (0)
#define AGB_LAUNCHED(ctx) do { (ctx)->launches++; cudaError_t _e = cudaPeekAtLastError(); if (_e != cudaSuccess) return agb_cuda_fail(_e, "kernel launch", __FILE__, __LINE__); } while (0)

static inline int64_t agb_numel(const agb_tensor* t) {
  int64_t n = 1; for (int i = 0; i < t->rank; i++) n *= t->shape[i]; return n;
}
static inline bool agb_is_contig(const agb_tensor* t) {
  int64_t s = 1;
  for (int i = t->rank - 1; i >= 0; i--) {
    if (t->shape[i] != 1 && t->stride[i] != s) return false;
    s *= t->shape[i];
  }
  return true;
}
static inline int agb_grid_for(int64_t work_items, int threads, int sm_count, int per_sm = 8) {
  int64_t b = (work_items + threads - 1) / threads;
  int64_t cap = (int64_t)sm_count * per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// grid for a grid-stride kernel = ONE full wave of the blocks that are actually co-resident (registers / shared memory decide, not the
// 2048-thread limit): a grid of 8 blocks per SM on a kernel that fits 6 runs a second, quarter-full wave as long as the first.
template <class K>
static inline int agb_grid_occ(agb_ctx* ctx, K kernel, int64_t work_items, int threads, size_t smem = 0) {
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, smem) != cudaSuccess || nb < 1) nb = 1;
  return agb_grid_for(work_items, threads, ctx->sm_count, nb);
}

// ---- device helpers ----
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
