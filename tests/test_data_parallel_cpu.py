"""CPU test of the data-parallel arithmetic with two `gloo` ranks: every rank evaluates the SAME graph on its
shard of the batch (oracle backend), gradients are summed with an all-reduce and applied scaled by 1/world — exactly what the
CUDA engine's end-of-run flush does with NCCL (engine/ops_nn.cc: flush_pending_updates).  The result must equal a
single-process step on the whole batch (sum of shard means / world = global mean for equal shards)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _grads_and_update(xb, yb, world=1, rank=0, allreduce=None):
    sys.path.insert(0, ROOT)
    from oracle import ref_graph as OG, ref_ops as R
    from rust_autograd_b200 import workloads as W
    env = OG.VariableEnvironment()
    W.mlp_init(env, np.random.default_rng(0))
    shard = slice(rank * len(xb) // world, (rank + 1) * len(xb) // world)

    def body(g):
        loss, _ = W.mlp_loss(OG, g)
        params, grads = OG.optimizers.grad_helper([loss], g.default_namespace())
        return [p.attrs["vid"] for p in params], [r.unwrap() for r in g.evaluator().extend(grads).feed("x", xb[shard]).feed("y", yb[shard]).run()]
    vids, grads = env.run(body)
    if allreduce is not None:
        grads = [allreduce(gr) / world for gr in grads]          # NCCL sum, then grad_scale = 1/world inside the optimizer kernel
    out = []
    for vid, gr in zip(vids, grads):
        p = env.get_array_by_id(vid)
        out.append(R.adam_update(p, gr, np.zeros_like(p), np.zeros_like(p), np.float32(1.0))[0])
    return out


def _worker(rank, world, port, xb, yb, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def allreduce(a):
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32).copy())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.numpy()
    out = _grads_and_update(xb, yb, world, rank, allreduce)
    if rank == 0:
        q.put([o.copy() for o in out])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_data_parallel_equals_single_process():
    rng = np.random.default_rng(1)
    xb = rng.uniform(size=(64, 784)).astype(np.float32)
    yb = rng.integers(0, 10, (64, 1)).astype(np.float32)
    ref = _grads_and_update(xb, yb)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, xb, yb, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for a, b in zip(got, ref):
        np.testing.assert_allclose(a, b, rtol=2e-5, atol=1e-7)
