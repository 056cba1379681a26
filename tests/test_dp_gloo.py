"""Data-parallel logic on CPU (gloo, world_size 2): batch-sharding the path and averaging the gradients over ranks
(what DDP's all-reduce does) reproduces the full-batch gradient — samples are independent, so the
path shards with no data-path collective.  Uses the CPU oracle as the per-rank compute."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cases as C
from oracle import mmoe_oracle as O


def _grads(sd, ev, y):
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    lg, lb = O.two_task_mmoe(sd, ev)
    loss = O.bce_with_logits(lg, y, O.POS_WEIGHT_GOOD) + O.bce_with_logits(lb, y, O.POS_WEIGHT_BEST)
    loss.backward()
    return {k: v.grad for k, v in sd.items()}


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = C.CASES_BY_NAME["head_b16"]
    sd = {k: v.double() for k, v in case.state_dict().items()}
    (ev,) = case.inputs()
    y = (ev[:, 0, 0] > 0).double()
    shard = slice(rank * 8, (rank + 1) * 8)
    g = _grads(sd, ev[shard].double(), y[shard])
    for k in sorted(g):
        dist.all_reduce(g[k], op=dist.ReduceOp.SUM)
        g[k] /= world
    if rank == 0:
        full = _grads(sd, ev.double(), y)
        q.put(max(float((g[k] - full[k]).abs().max() / full[k].abs().max().clamp(min=1e-30)) for k in g))
    dist.destroy_process_group()


def test_sharded_gradients_average_to_full_batch_gradient():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-12


def _sync_worker(rank, world, port, q):
    """The native exchange (functional._GradSync) on CPU tensors: flat buffers and sub-spans are averaged over ranks."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mmoe_multimodal_rec_b200 as pkg
    Fn = pkg.functional
    sync = Fn.enable_grad_allreduce()
    params = [torch.zeros(3, 5), torch.zeros(7), torch.zeros(2, 2)]
    views, _ptrs, flat = Fn._alloc_grads(params, [True, False, True])
    assert views[1] is None and flat.numel() >= 15 + 4
    views[0].fill_(float(rank + 1))
    views[2].fill_(10.0 * (rank + 1))
    Fn._sync_grads(Fn._span(flat, views, 0, 1))       # one stage's span ...
    Fn._sync_grads(Fn._span(flat, views, 1, 3))       # ... and the rest (skips the unused parameter)
    assert len(sync.pending) == 2
    Fn.wait_grad_allreduce()
    ok = bool(torch.all(views[0] == 1.5)) and bool(torch.all(views[2] == 15.0)) and not sync.pending
    Fn.disable_grad_allreduce()
    Fn._sync_grads(flat)                               # off: no collective issued
    if rank == 0:
        q.put(ok)
    dist.destroy_process_group()


def test_native_flat_buffer_gradient_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_sync_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
