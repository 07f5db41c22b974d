// runtime.cu — context, stream-ordered caching arena, copies, events, CUDA graphs, NCCL plumbing.
// Replaces nothing numeric in the reference; it is the device-side counterpart of the evaluator's
// per-run heap allocations (`src/evaluation.rs:183-223`) and of `VariableEnvironment` storage
// (`src/variable.rs:152-155`): variables and intermediates live in HBM.
#include "common.cuh"
#include <stdlib.h>
#include <stdarg.h>
#include <algorithm>
#include <vector>
#include <dlfcn.h>

static thread_local char g_err[1024] = "";

void agb_set_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
int agb_cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  agb_set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  return AGB_ERR_CUDA;
}
extern "C" const char* agb_last_error(void) { return g_err; }

extern "C" int agb_device_count(int* out) {
  int n = 0; cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { *out = 0; return agb_cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__); }
  *out = n; return AGB_OK;
}

extern "C" int agb_init(int device, agb_ctx** out) {
  *out = nullptr;
  int n = 0;
  AGB_CUDA(cudaGetDeviceCount(&n));
  AGB_CHECK(n > 0 && device < n, AGB_ERR_CUDA, "agb_init: no CUDA device %d (count=%d); there is no CPU fallback", device, n);
  AGB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop; AGB_CUDA(cudaGetDeviceProperties(&prop, device));
  AGB_CHECK(prop.major == 10, AGB_ERR_CUDA, "agb_init: device is sm_%d%d; this library is built for sm_100a only", prop.major, prop.minor);
  agb_ctx* ctx = new agb_ctx();
  ctx->device = device; ctx->sm_count = prop.multiProcessorCount;
  if (const char* e = getenv("AGB_DETERMINISTIC")) ctx->deterministic = (e[0] == '0') ? 0 : 1;      // default on; agb_set_deterministic overrides
  AGB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  AGB_CUDA(cudaMalloc(&ctx->dev_err, sizeof(int)));
  AGB_CUDA(cudaMemsetAsync(ctx->dev_err, 0, sizeof(int), ctx->stream));
  *out = ctx;
  return AGB_OK;
}

extern "C" int agb_trim(agb_ctx* ctx) {
  if (ctx->pinned_graphs > 0) return AGB_OK;       // instantiated step graphs hold raw addresses of arena blocks
  cudaStreamSynchronize(ctx->stream);
  for (auto& kv : ctx->free_blocks) { cudaFree(kv.second); ctx->block_size.erase(kv.second); ctx->is_live.erase(kv.second); }
  ctx->free_blocks.clear(); ctx->cached_bytes = 0;
  return AGB_OK;
}

extern "C" int agb_destroy(agb_ctx* ctx) {
  if (!ctx) return AGB_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  agb_nccl_destroy(ctx);
  for (auto& kv : ctx->block_size) cudaFree(kv.first);
  for (auto& kv : ctx->optim_tables) cudaFree(kv.second);
  if (ctx->scratch) cudaFree(ctx->scratch);
  if (ctx->scratch2) cudaFree(ctx->scratch2);
  for (void* p : ctx->retired_scratch) cudaFree(p);
  if (ctx->flush_buf) cudaFree(ctx->flush_buf);
  if (ctx->dev_err) cudaFree(ctx->dev_err);
  if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); cudaStreamSynchronize(ctx->d2h_stream); cudaStreamDestroy(ctx->d2h_stream); cudaEventDestroy(ctx->stage_mark); cudaEventDestroy(ctx->stage_done); }
  cudaStreamDestroy(ctx->stream);
  delete ctx;
  return AGB_OK;
}

extern "C" int agb_sm_count(agb_ctx* ctx, int* out) { *out = ctx->sm_count; return AGB_OK; }
extern "C" int agb_set_math_mode(agb_ctx* ctx, int mode) {
  AGB_CHECK(mode >= 0 && mode <= 2, AGB_ERR_INVALID_DIMS, "agb_set_math_mode: bad mode %d", mode);
  ctx->math_mode = mode; return AGB_OK;
}
extern "C" int agb_get_math_mode(agb_ctx* ctx, int* mode) { *mode = ctx->math_mode; return AGB_OK; }
extern "C" int agb_launch_count(agb_ctx* ctx, int64_t* out) { *out = ctx->launches; return AGB_OK; }

// One process drives one GPU, but nothing stops a process from holding contexts on several devices: the calls that would
// silently act on the WRONG device (cudaMalloc lands on whichever device is current) make the context's device current first.  Kernel
// launches go to ctx->stream and fail loudly (invalid resource handle) if another device is current.
static inline void agb_use_device(agb_ctx* ctx) {
  int cur = -1;
  if (cudaGetDevice(&cur) != cudaSuccess || cur != ctx->device) cudaSetDevice(ctx->device);
}

static size_t round_block(size_t bytes) {
  if (bytes == 0) bytes = 1;
  if (bytes <= (1u << 20)) return (bytes + 511) & ~(size_t)511;              // 512 B granules
  return (bytes + ((1u << 21) - 1)) & ~(size_t)((1u << 21) - 1);             // 2 MiB granules
}

extern "C" int agb_alloc(agb_ctx* ctx, size_t bytes, void** out) {
  size_t sz = round_block(bytes);
  auto it = ctx->free_blocks.lower_bound(sz);
  // accept a cached block up to 25% larger than requested
  if (it != ctx->free_blocks.end() && it->first <= sz + sz / 4) {
    void* p = it->second; size_t got = it->first;
    ctx->free_blocks.erase(it);
    ctx->cached_bytes -= got; ctx->live_bytes += got; ctx->is_live[p] = true;
    if (ctx->live_bytes > ctx->peak_bytes) ctx->peak_bytes = ctx->live_bytes;
    if (ctx->capturing) ctx->capture_blocks.push_back(p);
    *out = p; return AGB_OK;
  }
  AGB_CHECK(!ctx->capturing, AGB_ERR_CUDA, "arena growth during graph capture; run the step eagerly (twice) first");
  agb_use_device(ctx);
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, sz);
  if (e != cudaSuccess) {   // give cached memory back and retry once
    cudaGetLastError();
    agb_trim(ctx);
    e = cudaMalloc(&p, sz);
    if (e != cudaSuccess) return agb_cuda_fail(e, "cudaMalloc", __FILE__, __LINE__);
  }
  ctx->block_size[p] = sz; ctx->is_live[p] = true; ctx->live_bytes += sz;
  if (ctx->live_bytes > ctx->peak_bytes) ctx->peak_bytes = ctx->live_bytes;
  *out = p; return AGB_OK;
}

extern "C" int agb_free(agb_ctx* ctx, void* ptr) {
  if (!ptr) return AGB_OK;
  auto it = ctx->block_size.find(ptr);
  AGB_CHECK(it != ctx->block_size.end(), AGB_ERR_INVALID_DIMS, "agb_free: pointer %p not owned by this context", ptr);
  AGB_CHECK(ctx->is_live[ptr], AGB_ERR_INVALID_DIMS, "agb_free: double free of %p", ptr);
  ctx->is_live[ptr] = false;
  ctx->live_bytes -= it->second;
  auto rs = ctx->reserve_on_free.find(ptr);
  if (rs != ctx->reserve_on_free.end()) { rs->second->blocks.push_back(ptr); ctx->reserve_on_free.erase(rs); return AGB_OK; }      // stays private to the graph that uses it
  ctx->cached_bytes += it->second;
  ctx->free_blocks.insert({it->second, ptr});
  return AGB_OK;
}

extern "C" int agb_mem_stats(agb_ctx* ctx, size_t* live, size_t* cached, size_t* peak) {
  if (live) *live = ctx->live_bytes; if (cached) *cached = ctx->cached_bytes; if (peak) *peak = ctx->peak_bytes;
  return AGB_OK;
}

static int scratch_grow(agb_ctx* ctx, void** buf, size_t* have, size_t bytes) {
  if (bytes > *have) {
    AGB_CHECK(!ctx->capturing, AGB_ERR_CUDA, "scratch growth during graph capture; run the step once eagerly first");
    agb_use_device(ctx);
    if (*buf) {
      AGB_CUDA(cudaStreamSynchronize(ctx->stream));
      if (ctx->pinned_graphs > 0) ctx->retired_scratch.push_back(*buf);      // an instantiated graph may hold this address: keep it alive until the context goes
      else AGB_CUDA(cudaFree(*buf));
    }
    size_t sz = bytes < (8u << 20) ? (8u << 20) : round_block(bytes);
    AGB_CUDA(cudaMalloc(buf, sz)); *have = sz;
  }
  return AGB_OK;
}
int agb_scratch(agb_ctx* ctx, size_t bytes, void** out) { AGB_TRY(scratch_grow(ctx, &ctx->scratch, &ctx->scratch_bytes, bytes)); *out = ctx->scratch; return AGB_OK; }
int agb_scratch2(agb_ctx* ctx, size_t bytes, void** out) { AGB_TRY(scratch_grow(ctx, &ctx->scratch2, &ctx->scratch2_bytes, bytes)); *out = ctx->scratch2; return AGB_OK; }

// ---- deterministic reductions -------------------------------------------------------------------------------------------
// block = 32 consecutive elements x 8 partial groups: thread (e, g) adds partials g, g + 8, g + 16, ... of element e, the eight group sums are
// added as a fixed tree.  (One thread per element walking all the partials was latency-bound: 23 us for the 2560 outputs of the classifier GEMM.)
// (pt > 0: the partials are laid out [tap][o][c] — coalesced drain stores of the filter-gradient kernels — and the sum goes to out[(o * pc + c) * pt + tap])
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ part, float* __restrict__ out, int nparts, int64_t n, int64_t stride, int accumulate, int po, int pc, int pt) {
  __shared__ float sm[8][33];
  const int e = threadIdx.x & 31, g = threadIdx.x >> 5;
  for (int64_t base = (int64_t)blockIdx.x * 32; base < n; base += (int64_t)gridDim.x * 32) {
    const int64_t i = base + e;
    float s0 = 0.0f, s1 = 0.0f;
    if (i < n) {
      const float* p = part + i;
      int k = g;
      for (; k + 8 < nparts; k += 16) { s0 += __ldg(p + (int64_t)k * stride); s1 += __ldg(p + (int64_t)(k + 8) * stride); }
      if (k < nparts) s0 += __ldg(p + (int64_t)k * stride);
    }
    sm[g][e] = s0 + s1;
    __syncthreads();
    if (g == 0 && i < n) {
      const float s = ((sm[0][e] + sm[1][e]) + (sm[2][e] + sm[3][e])) + ((sm[4][e] + sm[5][e]) + (sm[6][e] + sm[7][e]));
      int64_t oi = i;
      if (pt > 0) { const int64_t oc = (int64_t)po * pc; const int tap = (int)(i / oc); const int64_t r = i - tap * oc; oi = r * pt + tap; }      // r = o * pc + c
      out[oi] = accumulate ? out[oi] + s : s;
    }
    __syncthreads();
  }
}
int agb_reduce_partials(agb_ctx* ctx, const float* part, float* out, int nparts, int64_t n, int64_t stride, int accumulate) {
  if (n <= 0) return AGB_OK;
  int64_t blocks = (n + 31) / 32; const int64_t cap = (int64_t)ctx->sm_count * 16; if (blocks > cap) blocks = cap;
  reduce_partials_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(part, out, nparts, n, stride, accumulate, 0, 0, 0);
  AGB_LAUNCHED(ctx);
  return AGB_OK;
}
int agb_reduce_partials_wgrad(agb_ctx* ctx, const float* part, float* gw, int nparts, int O, int C, int T) {
  const int64_t n = (int64_t)O * C * T;
  if (n <= 0) return AGB_OK;
  int64_t blocks = (n + 31) / 32; const int64_t cap = (int64_t)ctx->sm_count * 16; if (blocks > cap) blocks = cap;
  reduce_partials_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(part, gw, nparts, n, n, 0, O, C, T);
  AGB_LAUNCHED(ctx);
  return AGB_OK;
}
int agb_reduce_partials2(agb_ctx* ctx, float* part, int64_t nparts, int64_t n, float* out, int accumulate) {
  if (nparts <= 64) return agb_reduce_partials(ctx, part, out, (int)nparts, n, n, accumulate);
  const int64_t pad = (nparts + 63) / 64 * 64;
  if (pad > nparts) AGB_TRY(agb_memset0(ctx, part + nparts * n, (size_t)(pad - nparts) * n * sizeof(float)));
  float* grp = part + pad * n;
  AGB_TRY(agb_reduce_partials(ctx, part, grp, (int)(pad / 64), 64 * n, 64 * n, 0));        // group g = partials g, g + 64, g + 128, ...
  return agb_reduce_partials(ctx, grp, out, 64, n, n, accumulate);
}
extern "C" int agb_set_deterministic(agb_ctx* ctx, int on) { ctx->deterministic = on ? 1 : 0; return AGB_OK; }
extern "C" int agb_get_deterministic(agb_ctx* ctx, int* on) { *on = ctx->deterministic; return AGB_OK; }

extern "C" int agb_host_alloc(size_t bytes, void** out) { AGB_CUDA(cudaMallocHost(out, bytes ? bytes : 1)); return AGB_OK; }
extern "C" int agb_host_free(void* p) { if (p) AGB_CUDA(cudaFreeHost(p)); return AGB_OK; }

extern "C" int agb_h2d(agb_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if (bytes) AGB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return AGB_OK;
}
// ---- host-feed staging: double-buffered inputs, copy of step i+1 overlapped with the kernels of step i --------------------------
static int stage_init(agb_ctx* ctx) {
  if (ctx->copy_stream) return AGB_OK;
  AGB_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  AGB_CUDA(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
  AGB_CUDA(cudaEventCreateWithFlags(&ctx->stage_mark, cudaEventDisableTiming));
  AGB_CUDA(cudaEventCreateWithFlags(&ctx->stage_done, cudaEventDisableTiming));
  return AGB_OK;
}
extern "C" int agb_stage_mark(agb_ctx* ctx) {
  AGB_TRY(stage_init(ctx));
  AGB_CUDA(cudaEventRecord(ctx->stage_mark, ctx->stream)); ctx->stage_marked = true;
  return AGB_OK;
}
extern "C" int agb_stage_h2d(agb_ctx* ctx, void* dst, const void* src, size_t bytes) {
  AGB_TRY(stage_init(ctx));
  if (ctx->stage_marked) AGB_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->stage_mark, 0));      // readers of dst enqueued before the mark
  if (bytes) AGB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
  AGB_CUDA(cudaEventRecord(ctx->stage_done, ctx->copy_stream)); ctx->stage_pending = true;
  return AGB_OK;
}
// D2H of a result on its own stream (separate from the H2D staging stream, so a result copy that waits for step i never delays the
// input copy of step i+1), ordered after everything enqueued on the compute stream so far.  dst must be pinned: a pageable
// destination would block the host until the copy runs.  *done receives an event (agb_event_sync / agb_event_destroy).
extern "C" int agb_stage_d2h(agb_ctx* ctx, void* dst_pinned, const void* src, size_t bytes, void** done) {
  AGB_TRY(stage_init(ctx));
  cudaEvent_t e = nullptr, d = nullptr;
  AGB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  AGB_CUDA(cudaEventRecord(e, ctx->stream));
  AGB_CUDA(cudaStreamWaitEvent(ctx->d2h_stream, e, 0));
  AGB_CUDA(cudaEventDestroy(e));                       // released by the driver once the wait has consumed it
  if (bytes) AGB_CUDA(cudaMemcpyAsync(dst_pinned, src, bytes, cudaMemcpyDeviceToHost, ctx->d2h_stream));
  AGB_CUDA(cudaEventCreateWithFlags(&d, cudaEventDisableTiming | cudaEventBlockingSync));
  AGB_CUDA(cudaEventRecord(d, ctx->d2h_stream));
  *done = d;
  return AGB_OK;
}
extern "C" int agb_event_sync(void* ev) { AGB_CUDA(cudaEventSynchronize((cudaEvent_t)ev)); return AGB_OK; }
extern "C" int agb_stage_wait(agb_ctx* ctx) {
  if (ctx->stage_pending) { AGB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->stage_done, 0)); ctx->stage_pending = false; }
  return AGB_OK;
}
extern "C" int agb_d2h(agb_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if (bytes) AGB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  AGB_CUDA(cudaStreamSynchronize(ctx->stream));
  return AGB_OK;
}
extern "C" int agb_d2d(agb_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if (bytes) AGB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return AGB_OK;
}
extern "C" int agb_memset0(agb_ctx* ctx, void* dst, size_t bytes) {
  if (bytes) AGB_CUDA(cudaMemsetAsync(dst, 0, bytes, ctx->stream));
  return AGB_OK;
}
extern "C" int agb_sync(agb_ctx* ctx) {
  AGB_CUDA(cudaStreamSynchronize(ctx->stream));
  int flag = 0;
  AGB_CUDA(cudaMemcpy(&flag, ctx->dev_err, sizeof(int), cudaMemcpyDeviceToHost));
  if (flag) {
    cudaMemsetAsync(ctx->dev_err, 0, sizeof(int), ctx->stream);      // ordered with the kernels that set it (the legacy stream is not)
    cudaStreamSynchronize(ctx->stream);
    agb_set_error("device-side index check failed (code %d): label / gather index out of range", flag);
    return AGB_ERR_OUT_OF_BOUNDS;
  }
  return AGB_OK;
}

__global__ void flush_kernel(float4* p, size_t n4) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) p[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}
extern "C" int agb_flush_l2(agb_ctx* ctx) {
  const size_t bytes = 256u << 20;   // 256 MiB > 126 MB L2
  if (!ctx->flush_buf) agb_use_device(ctx);
  if (!ctx->flush_buf) { AGB_CUDA(cudaMalloc(&ctx->flush_buf, bytes)); ctx->flush_bytes = bytes; }
  flush_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>((float4*)ctx->flush_buf, bytes / 16);
  AGB_CUDA(cudaPeekAtLastError());
  return AGB_OK;
}

// ---- live profiler ----
extern "C" int agb_prof_enable(agb_ctx* ctx, int on) { ctx->prof_on = on != 0; return AGB_OK; }
extern "C" int agb_prof_reset(agb_ctx* ctx) {
  AGB_CUDA(cudaStreamSynchronize(ctx->stream));
  for (auto& r : ctx->prof_recs) { ctx->prof_pool.push_back(r.a); ctx->prof_pool.push_back(r.b); }
  ctx->prof_recs.clear(); return AGB_OK;
}
extern "C" int agb_prof_collect(agb_ctx* ctx, int cls, double* total_ms, int64_t* calls, double* work) {
  AGB_CUDA(cudaStreamSynchronize(ctx->stream));
  double t = 0, w = 0; int64_t n = 0;
  for (auto& r : ctx->prof_recs) if (r.cls == cls) { float ms = 0; AGB_CUDA(cudaEventElapsedTime(&ms, r.a, r.b)); t += ms; w += r.work; n++; }
  if (total_ms) *total_ms = t; if (calls) *calls = n; if (work) *work = w;
  return AGB_OK;
}

// ---- events ----
extern "C" int agb_event_create(void** ev) { cudaEvent_t e; AGB_CUDA(cudaEventCreate(&e)); *ev = e; return AGB_OK; }
extern "C" int agb_event_destroy(void* ev) { AGB_CUDA(cudaEventDestroy((cudaEvent_t)ev)); return AGB_OK; }
extern "C" int agb_event_record(agb_ctx* ctx, void* ev) { AGB_CUDA(cudaEventRecord((cudaEvent_t)ev, ctx->stream)); return AGB_OK; }
extern "C" int agb_event_elapsed_ms(void* a, void* b, float* ms) {
  AGB_CUDA(cudaEventSynchronize((cudaEvent_t)b));
  AGB_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)a, (cudaEvent_t)b));
  return AGB_OK;
}

// ---- CUDA graphs ----
extern "C" int agb_graph_begin(agb_ctx* ctx) {
  AGB_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
  ctx->capturing = true; ctx->capture_blocks.clear(); return AGB_OK;
}
// The handle owns the executable graph AND the arena blocks its kernels address (see agb_ctx::capture_blocks): they leave the free list
// here and come back in agb_graph_destroy, so no later allocation can alias memory a replay writes.
extern "C" int agb_graph_end(agb_ctx* ctx, void** graph_exec) {
  cudaGraph_t g = nullptr; ctx->capturing = false;
  if (graph_exec == nullptr) { cudaStreamEndCapture(ctx->stream, &g); if (g) cudaGraphDestroy(g); cudaGetLastError(); ctx->capture_blocks.clear(); return AGB_OK; }     // abort
  AGB_CUDA(cudaStreamEndCapture(ctx->stream, &g));
  cudaGraphExec_t ge = nullptr;
  cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
  int64_t kernel_nodes = 0;
  {   // kernels per replay, for agb_launch_count (a replayed step launches them all without passing through AGB_LAUNCHED)
    size_t nn = 0;
    if (cudaGraphGetNodes(g, nullptr, &nn) == cudaSuccess && nn > 0) {
      std::vector<cudaGraphNode_t> nodes(nn);
      if (cudaGraphGetNodes(g, nodes.data(), &nn) == cudaSuccess)
        for (size_t i = 0; i < nn; i++) { cudaGraphNodeType ty; if (cudaGraphNodeGetType(nodes[i], &ty) == cudaSuccess && ty == cudaGraphNodeTypeKernel) kernel_nodes++; }
    }
    cudaGetLastError();
  }
  cudaGraphDestroy(g);
  if (e != cudaSuccess) { ctx->capture_blocks.clear(); return agb_cuda_fail(e, "cudaGraphInstantiate", __FILE__, __LINE__); }
  auto* res = new agb_ctx::GraphRes{ctx, (void*)ge, {}};
  res->kernel_nodes = kernel_nodes;
  std::sort(ctx->capture_blocks.begin(), ctx->capture_blocks.end());
  ctx->capture_blocks.erase(std::unique(ctx->capture_blocks.begin(), ctx->capture_blocks.end()), ctx->capture_blocks.end());
  for (void* p : ctx->capture_blocks) {
    if (ctx->reserve_on_free.count(p)) continue;                                  // already promised to an older graph
    if (ctx->is_live[p]) { ctx->reserve_on_free[p] = res; continue; }            // still held by the caller: joins the reservation when freed
    const size_t sz = ctx->block_size[p];
    auto range = ctx->free_blocks.equal_range(sz);
    for (auto it = range.first; it != range.second; ++it) if (it->second == p) { ctx->free_blocks.erase(it); ctx->cached_bytes -= sz; res->blocks.push_back(p); break; }
  }
  ctx->capture_blocks.clear();
  *graph_exec = res; return AGB_OK;
}
extern "C" int agb_arena_pin(agb_ctx* ctx, int delta) { ctx->pinned_graphs += delta; if (ctx->pinned_graphs < 0) ctx->pinned_graphs = 0; return AGB_OK; }
extern "C" int agb_graph_launch(agb_ctx* ctx, void* graph_exec) {
  AGB_CUDA(cudaGraphLaunch((cudaGraphExec_t)((agb_ctx::GraphRes*)graph_exec)->exec, ctx->stream));
  ctx->launches += ((agb_ctx::GraphRes*)graph_exec)->kernel_nodes;
  return AGB_OK;
}
extern "C" int agb_graph_destroy(void* graph_exec) {
  if (!graph_exec) return AGB_OK;
  auto* res = (agb_ctx::GraphRes*)graph_exec; agb_ctx* ctx = res->ctx;
  cudaStreamSynchronize(ctx->stream);                                             // a replay may still be running on the blocks given back below
  cudaError_t e = cudaGraphExecDestroy((cudaGraphExec_t)res->exec);
  for (auto it = ctx->reserve_on_free.begin(); it != ctx->reserve_on_free.end();) { if (it->second == res) it = ctx->reserve_on_free.erase(it); else ++it; }
  for (void* p : res->blocks) { const size_t sz = ctx->block_size[p]; ctx->cached_bytes += sz; ctx->free_blocks.insert({sz, p}); }
  delete res;
  if (e != cudaSuccess) return agb_cuda_fail(e, "cudaGraphExecDestroy", __FILE__, __LINE__);
  return AGB_OK;
}

// ---- NCCL (dlopen'ed so the library loads on boxes without it; torch's bundled copy is reused when
//      torch is already imported in the process) ----
typedef struct { char internal[128]; } nccl_uid;
typedef int (*fn_ncclGetUniqueId)(nccl_uid*);
typedef int (*fn_ncclCommInitRank)(void**, int, nccl_uid, int);
typedef int (*fn_ncclAllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_ncclCommDestroy)(void*);
typedef const char* (*fn_ncclGetErrorString)(int);
static struct {
  void* h = nullptr; fn_ncclGetUniqueId uid; fn_ncclCommInitRank init; fn_ncclAllReduce ar; fn_ncclCommDestroy destroy;
  fn_ncclGetErrorString errstr;
} g_nccl;

static int nccl_load() {
  if (g_nccl.h) return AGB_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
  for (int i = 0; names[i] && !g_nccl.h; i++) g_nccl.h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  AGB_CHECK(g_nccl.h, AGB_ERR_NCCL, "dlopen(libnccl.so.2) failed: %s", dlerror());
  g_nccl.uid = (fn_ncclGetUniqueId)dlsym(g_nccl.h, "ncclGetUniqueId");
  g_nccl.init = (fn_ncclCommInitRank)dlsym(g_nccl.h, "ncclCommInitRank");
  g_nccl.ar = (fn_ncclAllReduce)dlsym(g_nccl.h, "ncclAllReduce");
  g_nccl.destroy = (fn_ncclCommDestroy)dlsym(g_nccl.h, "ncclCommDestroy");
  g_nccl.errstr = (fn_ncclGetErrorString)dlsym(g_nccl.h, "ncclGetErrorString");
  AGB_CHECK(g_nccl.uid && g_nccl.init && g_nccl.ar && g_nccl.destroy, AGB_ERR_NCCL, "libnccl is missing symbols");
  return AGB_OK;
}
#define AGB_NCCL(x) do { int _r = (x); if (_r != 0) { agb_set_error("NCCL error %d (%s): %s", _r, g_nccl.errstr ? g_nccl.errstr(_r) : "?", #x); return AGB_ERR_NCCL; } } while (0)

extern "C" int agb_nccl_unique_id(void* id128) {
  AGB_TRY(nccl_load());
  AGB_NCCL(g_nccl.uid((nccl_uid*)id128));
  return AGB_OK;
}
extern "C" int agb_nccl_init(agb_ctx* ctx, int rank, int world, const void* id128) {
  AGB_TRY(nccl_load());
  AGB_CUDA(cudaSetDevice(ctx->device));
  nccl_uid id; memcpy(&id, id128, sizeof(id));
  AGB_NCCL(g_nccl.init(&ctx->nccl_comm, world, id, rank));
  ctx->rank = rank; ctx->world = world;
  return AGB_OK;
}
extern "C" int agb_allreduce_sum(agb_ctx* ctx, float* buf, int64_t n) {
  if (ctx->world <= 1) return AGB_OK;
  AGB_CHECK(ctx->nccl_comm, AGB_ERR_NCCL, "agb_allreduce_sum: agb_nccl_init was not called");
  // ncclFloat32 = 7, ncclSum = 0
  AGB_NCCL(g_nccl.ar(buf, buf, (size_t)n, 7, 0, ctx->nccl_comm, ctx->stream));
  ctx->launches++;
  return AGB_OK;
}
// Bucketed, overlapped form: the all-reduce of `buf` runs on the context's communication stream, ordered after everything enqueued on
// the compute stream so far (the kernels that produced the bucket); the compute stream carries on with the rest of the backward pass.
// agb_allreduce_wait makes the compute stream wait for every bucket issued so far (before the optimizer kernel reads the sums).
extern "C" int agb_allreduce_sum_async(agb_ctx* ctx, float* buf, int64_t n) {
  if (ctx->world <= 1) return AGB_OK;
  AGB_CHECK(ctx->nccl_comm, AGB_ERR_NCCL, "agb_allreduce_sum_async: agb_nccl_init was not called");
  if (!ctx->comm_stream) {
    AGB_CUDA(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
    AGB_CUDA(cudaEventCreateWithFlags(&ctx->comm_ready, cudaEventDisableTiming));
    AGB_CUDA(cudaEventCreateWithFlags(&ctx->comm_done, cudaEventDisableTiming));
  }
  AGB_CUDA(cudaEventRecord(ctx->comm_ready, ctx->stream));
  AGB_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->comm_ready, 0));
  AGB_NCCL(g_nccl.ar(buf, buf, (size_t)n, 7, 0, ctx->nccl_comm, ctx->comm_stream));
  AGB_CUDA(cudaEventRecord(ctx->comm_done, ctx->comm_stream));      // stream order: the latest record covers every earlier bucket
  ctx->comm_pending = true; ctx->launches++;
  return AGB_OK;
}
extern "C" int agb_allreduce_wait(agb_ctx* ctx) {
  if (ctx->comm_pending) { AGB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->comm_done, 0)); ctx->comm_pending = false; }
  return AGB_OK;
}
extern "C" int agb_nccl_destroy(agb_ctx* ctx) {
  if (ctx->comm_stream) { cudaStreamSynchronize(ctx->comm_stream); cudaStreamDestroy(ctx->comm_stream); cudaEventDestroy(ctx->comm_ready); cudaEventDestroy(ctx->comm_done); ctx->comm_stream = nullptr; }
  if (ctx->nccl_comm && g_nccl.destroy) { g_nccl.destroy(ctx->nccl_comm); ctx->nccl_comm = nullptr; }
  return AGB_OK;
}
