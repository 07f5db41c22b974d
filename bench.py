"""bench.py — headline benchmark: CNN training samples/s on B200 (BASELINE.json metric), VGG-style conv stack
(configs[3]: 64-256 ch, 3x128x128 synthetic images, batch 256 per GPU, Adam, data-parallel with NCCL gradient all-reduce).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode tf32|3xtf32|fp32] [--batch B]

One process per GPU (torchrun for N > 1).  Prints ONE JSON line on rank 0.
  value     whole-job samples/s with the batch already resident in HBM (CUDA events on the library's own stream, max over ranks)
  e2e       the same step driven through the public graph API with HOST feeds: H2D of the batch and D2H of the loss inside the timed region
  roofline  the dominant kernel class (conv implicit GEMM) timed live with CUDA event pairs around every call (agb_prof_*)
  cpu_baseline  the f32 C restatement of the reference's default-build CPU path (oracle/cpu_ref.c: im2col + one sgemm per sample in parallel
            over samples, sequential filter gradient, single-thread elementwise / pooling / Adam) on a bounded sample of the same workload
  modes     the same step in the f32-faithful 3xTF32 mode (the mode that matches the reference's f32 arithmetic to 1e-5)
  parity    loss / gradients of a batch-4 sample of the same network on the GPU (both modes) against the numpy oracle
  micro     GEMM 8192^3 (TF32, 3xTF32), reduce_sum / softmax over 2^28 elements, Adam 2^26: TFLOP/s, GB/s and roofline fractions
  configs   BASELINE configs[0..2] (MLP-MNIST, CNN-MNIST, LSTM LM): us per step, eager and replayed as a step graph
--impl reference times that C restatement alone with --steps / --warmup honoured (the Rust crate cannot be built in this image: no
cargo/rustc; kind = "port").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "vgg_stack_3x128x128_b256_per_gpu"
METRIC = "cnn_train_samples_per_s"


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}
    f = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(f):
        try:
            p.update(json.load(open(f)))
            p["src"] = "measured"
        except Exception:
            pass
    return p


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, False, []

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference arm (f32 C port on host cores)
def _cpu_trainer(batch):
    from oracle import cpu_ref as CR
    from rust_autograd_b200 import workloads as W
    rng = np.random.default_rng(0)
    tr = CR.VggTrainer(CR.vgg_params(rng), W.VGG_LAYERS, 128)
    x = rng.standard_normal((batch, 3, 128, 128)).astype(np.float32)
    y = rng.integers(0, 10, (batch, 1)).astype(np.float32)
    return CR, tr, x, y


def run_reference(args, rank):
    """The reference's own CPU algorithm for the path (f32, oracle/cpu_ref.c), all host threads it can use (one task per sample for
    conv / conv_transpose like rayon in the reference; everything else single-threaded like the reference), on a `--cpu-batch`-sample
    slice of the same VGG step.  --steps / --warmup are honoured; a wall-clock cap (REF_BUDGET_S) only trims a run that would not end
    within a few minutes, and says so."""
    if rank != 0:
        return
    CR, tr, x, y = _cpu_trainer(args.cpu_batch)
    budget = float(os.environ.get("REF_BUDGET_S", "240"))
    t_begin = time.time()
    warm_done = 0
    for _ in range(max(args.warmup, 1)):
        tr.step(x, y)
        warm_done += 1
        if time.time() - t_begin > budget / 4:
            break
    per = (time.time() - t_begin) / warm_done
    steps = max(1, min(args.steps, int((budget - (time.time() - t_begin)) / max(per, 1e-6))))
    t0 = time.time()
    loss = None
    for _ in range(steps):
        loss, _ = tr.step(x, y)
    dt = (time.time() - t0) / steps
    cores = CR.load().cr_threads()
    sample = ("f32 C port of the reference CPU path (%s), batch %d of the VGG stack per step, %d timed step(s) after %d warm-up%s"
              % (CR.SGEMM_BACKEND, args.cpu_batch, steps, warm_done, "" if steps == args.steps else " (trimmed from --steps %d by the %.0f s budget)" % (args.steps, budget)))
    v = args.cpu_batch / dt
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "samples/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm_done,
                      "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": WORKLOAD, "sample_batch": args.cpu_batch, "optimizer": "adam", "last_loss": loss},
                      "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
                      "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


def cpu_baseline(batch, steps=2):
    CR, tr, x, y = _cpu_trainer(batch)
    tr.step(x, y)
    t0 = time.time()
    for _ in range(steps):
        tr.step(x, y)
    dt = (time.time() - t0) / steps
    return {"value": batch / dt, "unit": "samples/s", "cores": CR.load().cr_threads(), "kind": "port",
            "sample": "f32 C port of the reference CPU path (%s; conv parallel over samples, the rest single-threaded like the reference), batch %d of the same "
                      "VGG stack, %d timed step(s) after 1 warm-up, %.1f s" % (CR.SGEMM_BACKEND, batch, steps, dt * steps)}


def parity_block(ag, ffi, lib, device, batch=4):
    """The same network on a batch-`batch` sample, GPU (3xTF32 and TF32) against the numpy oracle through the protocol of oracle/parity.py:
    discrete decisions (ReLU masks, pool argmaxes) agree except at verified near-ties, and loss / logits / all 16 gradients agree with the
    oracle evaluated under the device's decisions.  `max_grad_rel_unforced` is the plain comparison, where one near-tie flip shows as 1e-3."""
    from oracle import parity as P
    rng = np.random.default_rng(17)
    x = rng.standard_normal((batch, 3, 128, 128)).astype(np.float32)
    y = rng.integers(0, 10, (batch, 1)).astype(np.float32)
    res = {"sample_batch": batch, "oracle": "oracle/ref_graph.py (numpy, f64 accumulation) under the device's ReLU / max-pool decisions (oracle/parity.py)"}
    cache = None
    for mode, nm in ((0, "3xtf32"), (1, "tf32")):
        r, cache = P.vgg_parity(ag, lambda env, m: ffi.check(lib.agb_set_math_mode(env.agb_ctx(), m)), mode, x, y, device=device, ref_unforced=cache)
        res[nm] = {k: r[k] for k in ("loss_rel", "logits_rel", "max_grad_rel", "max_grad_rel_unforced", "decisions")}
    return res


# ------------------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--mode", default="tf32", choices=["tf32", "3xtf32", "fp32"])
    ap.add_argument("--batch", type=int, default=256, help="per-GPU batch")
    ap.add_argument("--cpu-batch", type=int, default=32, help="samples per step of the CPU reference arm / cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the modes / parity / micro / configs blocks (ncu captures)")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import ctypes as C
    import torch
    import torch.distributed as dist
    from rust_autograd_b200 import autograd as ag, ffi, workloads as W
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = ffi.load_library()
    env = ag.VariableEnvironment(local)
    ctx = env.agb_ctx()
    ffi.check(lib.agb_set_math_mode(ctx, {"3xtf32": 0, "tf32": 1, "fp32": 2}[args.mode]))
    if world > 1:          # ncclUniqueId from rank 0 to everybody over torch.distributed (plumbing only)
        idbuf = C.create_string_buffer(128)
        if rank == 0:
            ffi.check(lib.agb_nccl_unique_id(idbuf))
        t = torch.frombuffer(bytearray(idbuf.raw), dtype=torch.uint8).cuda()
        dist.broadcast(t, 0)
        env.set_data_parallel(rank, world, bytes(t.cpu().numpy().tobytes()))

    rng = np.random.default_rng(0)         # identical replicas on every rank
    W.vgg_init(env, rng)
    adam = ag.optimizers.Adam.default("adam", env.default_namespace().current_var_ids(), env)
    B = args.batch
    drng = np.random.default_rng(1000 + rank)   # each rank feeds its own shard of the global batch
    n_host = 4
    xs = [drng.standard_normal((B, 3, 128, 128)).astype(np.float32) for _ in range(n_host)]
    ys = [drng.integers(0, 10, (B, 1)).astype(np.float32) for _ in range(n_host)]
    # pinned host staging for the e2e leg
    pin = []
    for a in xs + ys:
        p = C.c_void_p()
        ffi.check(lib.agb_host_alloc(a.nbytes, C.byref(p)))
        C.memmove(p, a.ctypes.data, a.nbytes)
        pin.append(np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=a.shape))
    xs_p, ys_p = pin[:n_host], pin[n_host:]
    # device-resident copies for the `value` leg
    dev_x, dev_y = [], []
    for a in xs:
        p = C.c_void_p(); ffi.check(lib.agb_alloc(ctx, a.nbytes, C.byref(p))); ffi.check(lib.agb_h2d(ctx, p, a.ctypes.data, a.nbytes)); dev_x.append(ag.DeviceArray(p.value, a.shape))
    for a in ys:
        p = C.c_void_p(); ffi.check(lib.agb_alloc(ctx, a.nbytes, C.byref(p))); ffi.check(lib.agb_h2d(ctx, p, a.ctypes.data, a.nbytes)); dev_y.append(ag.DeviceArray(p.value, a.shape))
    ffi.check(lib.agb_sync(ctx))

    g = ag.Context(env)                   # the graph is built once; the training loop lives outside (README: "move the loop out")
    loss, _ = W.vgg_loss(ag, g)
    params, grads = ag.optimizers.grad_helper([loss], g.default_namespace())
    upd = adam.get_update_op(params, grads, g)

    def step_resident(i):
        g.evaluator().push(loss).push(upd).feed("x", dev_x[i % n_host]).feed("y", dev_y[i % n_host]).run_async()

    # e2e through the public graph API: every step's inputs start in pinned host memory and are copied host -> device inside the timed
    # region (HostPrefetcher: the copy of step i+1 runs on a second stream under the kernels of step i); every step's loss is read
    # back to the host (run_deferred: the read of step i completes while step i+1 is already running, the last one inside the region)
    pf = ag.HostPrefetcher(env, [xs_p[0].shape, ys_p[0].shape])
    pf.stage([xs_p[0], ys_p[0]])
    e2e_state = {"pending": None, "losses": []}

    def step_e2e(i):
        x, y = pf.acquire()
        d = g.evaluator().push(loss).push(upd).feed("x", x).feed("y", y).run_deferred()
        pf.stage([xs_p[(i + 1) % n_host], ys_p[(i + 1) % n_host]])
        e2e_drain()
        e2e_state["pending"] = d

    def e2e_drain():
        if e2e_state["pending"] is not None:
            e2e_state["losses"].append(float(np.asarray(e2e_state["pending"].get()[0].unwrap()).ravel()[0]))
            e2e_state["pending"] = None

    def barrier():
        ffi.check(lib.agb_sync(ctx))
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, warm, fin=None):
        for i in range(warm):
            fn(i)
        if fin:
            fin()
        barrier()
        ev0, ev1 = C.c_void_p(), C.c_void_p()
        ffi.check(lib.agb_event_create(C.byref(ev0))); ffi.check(lib.agb_event_create(C.byref(ev1)))
        l0 = C.c_int64(); ffi.check(lib.agb_launch_count(ctx, C.byref(l0)))
        t0 = time.time()
        ffi.check(lib.agb_event_record(ctx, ev0))
        for i in range(steps):
            fn(warm + i)
        if fin:
            fin()              # e2e: the last step's loss is read inside the timed region too
        ffi.check(lib.agb_event_record(ctx, ev1))
        ms = C.c_float(); ffi.check(lib.agb_event_elapsed_ms(ev0, ev1, C.byref(ms)))
        barrier()
        wall = (time.time() - t0) * 1e3
        l1 = C.c_int64(); ffi.check(lib.agb_launch_count(ctx, C.byref(l1)))
        dev_ms = float(ms.value)
        if world > 1:
            t = torch.tensor([dev_ms, wall], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); dev_ms, wall = float(t[0]), float(t[1])
        return dev_ms, wall, l1.value - l0.value

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- leg 1: inputs resident in HBM (value); then the same steps once more with the live kernel profiler on (event pairs around
    # every profiled call: that pass feeds `roofline`, not `value`)
    ffi.check(lib.agb_prof_reset(ctx))
    dev_ms, wall_ms, launches = timed(step_resident, args.steps, args.warmup)
    env.set_plan_cache(False)             # the per-call event pairs need eager launches (a replayed step graph is one opaque launch)
    ffi.check(lib.agb_prof_enable(ctx, 1)); ffi.check(lib.agb_prof_reset(ctx))
    prof_ms, _, _ = timed(step_resident, args.steps, 0)
    ffi.check(lib.agb_prof_enable(ctx, 0))
    env.set_plan_cache(True)
    prof = {}
    names = ["gemm", "conv_fprop", "conv_dgrad", "conv_wgrad", "ewise", "reduce", "softmax", "pool", "optim",
             "conv_small_c_fprop", "conv_small_c_wgrad", "conv_simt"]
    for cls, nm in enumerate(names):
        t, n, w = C.c_double(), C.c_int64(), C.c_double()
        ffi.check(lib.agb_prof_collect(ctx, cls, C.byref(t), C.byref(n), C.byref(w)))
        if n.value:
            prof[nm] = {"ms": t.value, "calls": n.value, "work": w.value}
    ffi.check(lib.agb_prof_reset(ctx))
    # ---- leg 2: end to end through the public API with host feeds
    e2e_state["losses"] = []
    e2e_ms, e2e_wall, _ = timed(step_e2e, args.steps, 1, e2e_drain)
    e2e_time = max(e2e_ms, e2e_wall)       # the D2H of the loss serialises host and device: wall time is the honest figure
    timed_losses = e2e_state["losses"][1:]  # [0] belongs to the warm-up step
    # ---- leg 3: the same resident step in the other tensor-core mode (3xTF32 = the f32-faithful mode; TF32 when --mode 3xtf32)
    other = None
    if not args.no_extras and args.mode in ("tf32", "3xtf32"):
        other_mode = "3xtf32" if args.mode == "tf32" else "tf32"
        ffi.check(lib.agb_set_math_mode(ctx, {"3xtf32": 0, "tf32": 1}[other_mode]))
        o_steps = max(3, args.steps // 2)
        o_ms, _, _ = timed(step_resident, o_steps, 3)
        other = {"mode": other_mode, "steps": o_steps, "ms_per_step": o_ms / o_steps, "value": world * B * o_steps / (o_ms / 1e3), "unit": "samples/s"}
        ffi.check(lib.agb_set_math_mode(ctx, {"3xtf32": 0, "tf32": 1, "fp32": 2}[args.mode]))
    sampler.stop_flag = True
    # ---- data-parallel invariant: every rank holds bit-identical variables after the timed steps
    dp_check = None
    if world > 1:
        import zlib
        n_vars = len(env.default_namespace().current_var_ids())
        crc = 0
        for vid in range(n_vars):
            crc = zlib.crc32(np.ascontiguousarray(env.get_array_by_id(vid)).tobytes(), crc)
        t = torch.tensor([crc], dtype=torch.int64, device="cuda")
        allc = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allc, t)
        dp_check = {"weights_crc32_equal_across_ranks": len({int(c.item()) for c in allc}) == 1, "variables": n_vars}

    if rank == 0:
        pk = peaks()
        step_ms = max(dev_ms, 0.0) / args.steps
        value = world * B * args.steps / (dev_ms / 1e3)
        e2e_value = world * B * args.steps / (e2e_time / 1e3)
        conv = {k: v for k, v in prof.items() if k.startswith("conv")}
        roof = None
        # dominant kernel: the tcgen05 implicit-GEMM conv (tc_tile_persist_kernel<ConvFpropPol<TN>>, conv_rows_kernel for wide maps) that serves Conv2D fprop and (flipped filter)
        # Conv2DTranspose / dgrad.  Classes conv_fprop + conv_dgrad hold exactly its launches (small-C and SIMT paths are re-labelled).
        dom = [conv[k] for k in ("conv_fprop", "conv_dgrad") if k in conv]
        if dom:
            d_ms = sum(v["ms"] for v in dom); d_w = sum(v["work"] for v in dom); d_n = sum(v["calls"] for v in dom)
            tot_ms = sum(v["ms"] for v in conv.values()); tot_w = sum(v["work"] for v in conv.values())
            achieved = d_w / (d_ms / 1e3) / 1e12
            peak_div = 3.0 if args.mode == "3xtf32" else 1.0          # 3xTF32 issues three MMAs per product: useful-FLOP ceiling = 1/3
            tf32_peak = pk["bf16_tflops_sustained"] / 2.0 / peak_div
            tf32_burst = pk["bf16_tflops"] / 2.0 / peak_div
            traffic, traffic_src = None, None
            tpath = os.path.join(ROOT, "profiles", "ncu_conv_traffic.json")     # dram bytes per launch from the committed ncu --set full capture
            if os.path.exists(tpath) and world == 1:
                tj = json.load(open(tpath))
                if tj.get("workload") == WORKLOAD and tj.get("math_mode") == args.mode:
                    traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
            roof = {"bound": "tensor", "kernel": "tc_tile_persist_kernel<ConvFpropPol> + conv_rows_kernel + conv_cols_kernel (tcgen05 implicit-GEMM conv: fprop + dgrad, %s)" % args.mode,
                    "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s", "frac": achieved / tf32_peak, "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": "%s bf16 sustained / 2 (dense TF32 = half the bf16 rate; the kernel is timed inside a long step), MEASURED_PEAKS.json" % pk["src"],
                    "peak_burst": tf32_burst, "frac_burst": achieved / tf32_burst,
                    "per_launch_ms": d_ms / d_n, "flops_per_launch": d_w / d_n, "launches_per_step": d_n / args.steps,
                    "share_of_step": d_ms / max(prof_ms, 1e-9),
                    "all_conv": {"achieved": tot_w / (tot_ms / 1e3) / 1e12, "share_of_step": tot_ms / max(prof_ms, 1e-9)},
                    "classes": {k: {"ms_per_step": v["ms"] / args.steps, "calls_per_step": v["calls"] / args.steps} for k, v in prof.items()}}
        out = {"metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (%s tensor-core contractions)" % args.mode, "data": "synthetic",
               "config": {"workload": WORKLOAD, "per_gpu_batch": B, "global_batch": B * world, "optimizer": "adam", "parallelism": "dp%d" % world,
                          "l2": "inputs and activations (>= 1 GB per layer) exceed the 126 MB L2", "math_mode": args.mode,
                          "flops_per_step_per_gpu": W.vgg_flops_per_sample() * B},
               "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": int(xs[0].nbytes + ys[0].nbytes), "d2h_bytes_per_step": 4,
                       "ms_per_step": e2e_time / args.steps},
               "gpu_launches": int(launches), "plan_cache": env.plan_stats(), "clocks": sampler.summary(), "roofline": roof,
               "model_tflops": W.vgg_flops_per_sample() * B / (step_ms / 1e3) / 1e12,
               "losses": {"first": timed_losses[0] if timed_losses else None, "last": timed_losses[-1] if timed_losses else None, "n": len(timed_losses),
                          "note": "rank 0's loss of the first / last timed e2e step (random labels: the value only shows the step trains and is finite)"}}
        if other is not None:
            out["modes"] = {other["mode"]: other}
        if dp_check is not None:
            out["dp_check"] = dp_check
        if world == 1 and not args.no_extras:
            ffi.check(lib.agb_sync(ctx)); pf.close(); g.close(); env.close()      # release the training arena before the side blocks allocate their own
            for key, fn in (("parity", lambda: parity_block(ag, ffi, lib, local)),
                            ("micro", lambda: __import__("scripts.bench_micro", fromlist=["run"]).run(pk, local)),
                            ("configs", lambda: __import__("scripts.bench_configs", fromlist=["small_configs"]).small_configs(local))):
                try:
                    out[key] = fn()
                except Exception as e:      # side blocks never take the headline down
                    out[key] = {"error": repr(e)}
        if not args.no_cpu_baseline and world == 1:
            try:
                out["cpu_baseline"] = cpu_baseline(args.cpu_batch)
            except Exception as e:      # the baseline is reported, never the product path
                out["cpu_baseline"] = {"error": repr(e)}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
