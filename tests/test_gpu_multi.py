"""Multi-GPU gradient equality (needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`).

Two ranks, each with half of a batch, must end a backward pass with exactly the gradient a single GPU computes on the
whole batch — for EVERY parameter of EVERY module of the v1 path — through both exchange mechanisms:
  * the package's native flat-buffer exchange (functional.enable_grad_allreduce), including a second, accumulating step;
  * torch DistributedDataParallel wrappers around the drop-in modules, the way the reference's train.py:133-139 builds them
    (head called through ``.module`` as train.py:251 does, so its gradient stays local — SURVEY.md §0 quirk 1).
Eval-mode arithmetic (dropout off) so the comparison is exact up to fp32 summation order.
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

S, D, NTOK = 64, 768, 197


def _batch(B, seed):
    g = torch.Generator().manual_seed(seed)
    lens_u = torch.randint(1, S + 1, (B,), generator=g)
    lens_i = torch.randint(1, S + 1, (B,), generator=g)
    ar = torch.arange(S)[None]
    return {
        "u_sent": torch.randn(B, S, D, generator=g), "i_sent": torch.randn(B, S, D, generator=g),
        "u_mask": ar >= lens_u[:, None], "i_mask": ar >= lens_i[:, None],
        "u_doc": torch.randn(B, D, generator=g), "i_doc": torch.randn(B, D, generator=g),
        "img": torch.randn(B, NTOK, D, generator=g),
        "y_good": (torch.rand(B, generator=g) < 0.5).float(), "y_best": (torch.rand(B, generator=g) < 0.5).float(),
    }


def _build(dev, fused=False):
    import mmoe_multimodal_rec_b200 as pkg
    from parity_util import FakeBackbone
    M = pkg.modules
    torch.manual_seed(4321)
    pkg.functional.set_flat_parameters(fused)       # fused-parameter mode: one nn.Parameter per module (modules._Native)
    try:
        mods = {"img": M.ItemImageExpert(FakeBackbone(), pool_type="mean"), "cross": M.RobustTextCrossExpert(),
                "concat_ui": M.EnhancedCrossFuse(), "concat_ti": M.EnhancedCrossFuse(), "head": M.TwoTaskMMoE()}
    finally:
        pkg.functional.set_flat_parameters(False)
    for m in mods.values():
        m.to(dev).eval()
    return mods


def _step(mods, call, b, autocast_dtype):
    """train.py:244-254 on the drop-ins; `call` maps a module name to the callable to use (module, DDP wrapper, .module)."""
    import contextlib
    import torch.nn.functional as F
    ctx = torch.autocast("cuda", dtype=autocast_dtype) if autocast_dtype is not None else contextlib.nullcontext()
    with ctx:
        img_vec = call["img"](b["img"], trainable=False)
        ui = call["cross"](b["u_sent"], b["u_mask"], b["i_sent"], b["i_mask"])
        xui = call["concat_ui"](b["u_doc"], img_vec)
        xti = call["concat_ti"](b["i_doc"], img_vec)
        ev = torch.stack([b["u_doc"], b["i_doc"], img_vec, ui, xui, xti], dim=1)
        lg, lb = call["head"](ev)
        loss = F.binary_cross_entropy_with_logits(lg.float(), b["y_good"]) + F.binary_cross_entropy_with_logits(lb.float(), b["y_best"])
    loss.backward()
    return loss


def _grads(mods):
    out = {}
    for k, m in mods.items():
        named = m.named_gradients() if hasattr(m, "named_gradients") else [(n, p.grad) for n, p in m.named_parameters()]
        out.update({f"{k}.{n}": g.detach().clone() for n, g in named if g is not None})
    return out


def _worker(rank, world, port, mode, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import mmoe_multimodal_rec_b200 as pkg
    Fn = pkg.functional
    dtype = {"fp32": None, "bf16": torch.bfloat16}[mode.split("/")[1]]
    how = mode.split("/")[0]
    fused = how.endswith("flat")
    how = how[:-4] if fused else how
    B = 16
    full = {k: v.to(dev) for k, v in _batch(B, 7).items()}
    shard = {k: v[rank * (B // world):(rank + 1) * (B // world)].contiguous() for k, v in full.items()}
    mods = _build(dev)
    plain = {k: m for k, m in mods.items()}
    # ---- single-GPU full-batch gradient (every rank computes it; identical weights by construction) ----
    _step(mods, plain, full, dtype)
    ref = _grads(mods)
    for m in mods.values():
        m.zero_grad(set_to_none=True)
    problems = []
    if fused:
        # the reference gradient above comes from the per-tensor modules; the exchange is tested on fused-parameter
        # modules that received the same weights through load_state_dict (reference keys)
        fmods = _build(dev, fused=True)
        for k in mods:
            fmods[k].load_state_dict(mods[k].state_dict())
        n_par = {k: len(list(m.parameters())) for k, m in fmods.items()}
        if any(n_par[k] != 1 for k in ("cross", "concat_ui", "concat_ti", "head")):
            problems.append(f"fused modules expose {n_par} parameters")
        mods = fmods
        plain = {k: m for k, m in mods.items()}

    def compare(tag, scale=1.0, skip_prefix=()):
        got = _grads(mods)
        for k, r in ref.items():
            if k.startswith(skip_prefix):
                continue
            if k not in got:
                problems.append(f"{tag}: {k} has no grad")
                continue
            err = float((got[k].double() - scale * r.double()).abs().max()) / max(float(r.abs().max()) * scale, 1e-30)
            tol = 1e-5 if dtype is None else 1e-2     # bf16: batch-size dependent tiling changes fp32 summation order before 16-bit roundings
            if not err <= tol:
                problems.append(f"{tag}: {k} err {err:.2e}")
            other = got[k].clone()
            dist.broadcast(other, src=0)
            if not torch.equal(other, got[k]):
                problems.append(f"{tag}: {k} differs between ranks")

    if how == "native":
        Fn.enable_grad_allreduce()
        _step(mods, plain, shard, dtype)
        compare("native")
        _step(mods, plain, shard, dtype)                      # accumulate on top (no zero_grad)
        compare("native/accumulate", 2.0)
        for m in mods.values():
            m.zero_grad(set_to_none=False)
        _step(mods, plain, shard, dtype)
        compare("native/zeroed-in-place")
        # no_sync-style accumulation + one reduce of the accumulated buffers
        for m in mods.values():
            m.zero_grad(set_to_none=True)
        Fn.set_grad_sync(False)
        _step(mods, plain, shard, dtype)
        _step(mods, plain, shard, dtype)
        Fn.set_grad_sync(True)
        Fn.allreduce_accumulated(list(mods.values()))
        compare("native/no_sync+allreduce_accumulated", 2.0)
        Fn.disable_grad_allreduce()
    else:
        from torch.nn.parallel import DistributedDataParallel as DDP
        wrapped = {k: DDP(mods[k], device_ids=[rank]) for k in ("cross", "concat_ui", "concat_ti", "head")}     # train.py:136-139
        call = {"img": mods["img"], "cross": wrapped["cross"], "concat_ui": wrapped["concat_ui"], "concat_ti": wrapped["concat_ti"],
                "head": wrapped["head"].module}                                                                # train.py:251
        _step(mods, call, shard, dtype)
        # head and img are not synchronised by train.py (quirk 1 / img expert not wrapped): compare the wrapped modules
        compare("ddp", skip_prefix=("head.", "img."))
        for m in mods.values():
            m.zero_grad(set_to_none=True)
        with wrapped["cross"].no_sync(), wrapped["concat_ui"].no_sync(), wrapped["concat_ti"].no_sync():        # train.py:266-274
            _step(mods, call, shard, dtype)
        _step(mods, call, shard, dtype)
        compare("ddp/grad_accum 2", 2.0, skip_prefix=("head.", "img."))
    torch.cuda.synchronize()
    if rank == 0:
        q.put(problems)
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["native/fp32", "native/bf16", "ddp/fp32", "ddp/bf16", "ddpflat/fp32", "ddpflat/bf16", "nativeflat/bf16"])
def test_two_rank_gradients_equal_full_batch_gradient(mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 35500 + (os.getpid() + hash(mode)) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    problems = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert not problems, problems[:12]
