"""Fused-parameter mode (modules._Native.fuse_parameters) on CPU: the module exposes ONE nn.Parameter, keeps the
reference's state_dict keys / order / shapes in both directions, survives .to() / deepcopy / an optimizer step with the
per-name views still aliasing the parameter, and hands the flat gradient buffer of the backward to autograd as that
parameter's gradient — locally, under torch DistributedDataParallel (gloo, world_size 2) and under the native exchange.
The CUDA library is the stub of test_grad_exchange_cpu.py (it writes (rank+1)*(index+1) into gradient buffer `index`)."""
import copy
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import mmoe_multimodal_rec_b200 as pkg
from test_grad_exchange_cpu import _StubLib

Fn, M, H = pkg.functional, pkg.modules, pkg.modules_home
D = 64


def _build(fused):
    Fn.set_flat_parameters(fused)
    try:
        torch.manual_seed(0)
        return {
            "cross": M.RobustTextCrossExpert(d=D, n_layer=2, n_head=8),
            "cross_home": H.RobustTextCrossExpert(d=D, n_layer=2, n_head=8),
            "fuse": M.EnhancedCrossFuse(d=D, n_head=8, depth=2),
            "fuse_home": H.EnhancedCrossFuse(d=D, n_head=8, depth=2),
            "head": M.TwoTaskMMoE(expert_dim=D, n_expert=6, tower_hidden=32),
            "home_head": H.HOME_MMoE_Complete(num_input_experts=6, expert_dim=D, tower_hidden=32),
        }
    finally:
        Fn.set_flat_parameters(False)


def _unused(name):
    return {"cross_home": ("norm.", "mlp."), "fuse_home": ("proj.",)}.get(name, ())


def test_state_dict_is_the_reference_layout():
    plain, fused = _build(False), _build(True)
    for name in plain:
        a, b = plain[name].state_dict(), fused[name].state_dict()
        assert list(a.keys()) == list(b.keys()), name
        for k in a:
            assert a[k].shape == b[k].shape and torch.equal(a[k], b[k]), (name, k)
        ps = dict(fused[name].named_parameters())
        assert "_flat_param" in ps
        rest = sorted(k for k in ps if k != "_flat_param")
        assert rest == sorted(k for k in a if k.startswith(_unused(name))) if _unused(name) else rest == []
        # nested in a container: prefixed keys, reference order
        box = torch.nn.ModuleDict({"x": torch.nn.Linear(2, 2), "m": fused[name]})
        assert [k for k in box.state_dict() if k.startswith("m.")] == ["m." + k for k in a.keys()]


def test_load_state_dict_round_trip_and_errors():
    plain, fused = _build(False), _build(True)
    for name, m in fused.items():
        src = {k: torch.randn_like(v) for k, v in plain[name].state_dict().items()}
        res = m.load_state_dict(src)
        assert not res.missing_keys and not res.unexpected_keys
        got = m.state_dict()
        for k, v in src.items():
            assert torch.equal(got[k], v), (name, k)
        plain[name].load_state_dict(got)                        # and back into the unfused module
        k0 = next(k for k in src if not k.startswith(_unused(name) or ("\0",)))
        short = {k: v for k, v in src.items() if k != k0}
        with pytest.raises(RuntimeError, match="Missing key"):
            m.load_state_dict(short)
        assert k0 in m.load_state_dict(short, strict=False).missing_keys
        bad = dict(src)
        bad[k0] = torch.zeros(3, 5)
        with pytest.raises(RuntimeError, match="size mismatch"):
            m.load_state_dict(bad)
        extra = dict(src)
        extra["nope.weight"] = torch.zeros(1)
        with pytest.raises(RuntimeError, match="Unexpected key"):
            m.load_state_dict(extra)


def test_views_follow_the_parameter():
    fused = _build(True)
    m = fused["fuse"]
    flat = m._flat_param
    w = m.layers[0].linear1.weight
    assert not isinstance(w, torch.nn.Parameter) and w.shape == (4 * D, D)
    v0 = w._version
    opt = torch.optim.SGD(m.parameters(), lr=1.0)
    flat.grad = torch.ones_like(flat)
    before = w.clone()
    opt.step()
    assert torch.equal(m.layers[0].linear1.weight, before - 1.0)          # same storage
    assert m.layers[0].linear1.weight._version > v0                       # and the 16-bit weight cache sees the update
    assert torch.equal(m.state_dict()["layers.0.linear1.weight"], before - 1.0)
    # padding between tensors stays out of the state dict and is zero
    lay = m.fused_layout()
    assert sum(int(torch.tensor(s).prod()) for _, s in lay.values()) <= flat.numel()
    m2 = copy.deepcopy(m)
    m2._flat_param.data.add_(1.0)
    assert torch.equal(m2.layers[0].linear1.weight, m.layers[0].linear1.weight + 1.0)
    assert m2.layers[0].linear1.weight.data_ptr() != m.layers[0].linear1.weight.data_ptr()
    m3 = m.to(torch.float32).to("cpu")
    m3._params()
    assert m3.layers[0].linear1.weight.untyped_storage().data_ptr() == m3._flat_param.untyped_storage().data_ptr()
    m.requires_grad_(False)
    assert not m._flat_param.requires_grad
    # whole-module pickling (torch.save(module)): caches stay behind, views are rebuilt on first use
    import io
    m._pack().tensors  # noqa: B018  (make sure a pack exists)
    buf = io.BytesIO()
    torch.save(m2, buf)
    buf.seek(0)
    m4 = torch.load(buf, weights_only=False)
    m4._params()
    assert torch.equal(m4.state_dict()["layers.0.linear1.weight"], m2.state_dict()["layers.0.linear1.weight"])
    m4._flat_param.data.mul_(2.0)
    assert torch.equal(m4.layers[0].linear1.weight, 2.0 * m2.layers[0].linear1.weight)


# ----------------------------------------------------------------------------------------------
def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    stub = _StubLib(rank)
    Fn.lib = lambda: stub
    Fn._require_cuda = lambda *a: None
    Fn._stream = lambda: 0
    errors = []

    def ck(cond, what):
        if not cond:
            errors.append(what)

    mods = _build(True)
    mods.pop("home_head")                                     # HeadFn is covered by "head"

    def run(name, m, call=None):
        names, tensors = m._names_and_tensors()
        stub.register(len(names), [t.numel() for t in tensors])
        stub._n_current = len(names)
        f = call or m
        B, S = 3, 4
        if name.startswith("cross"):
            msk = torch.zeros(B, S, dtype=torch.bool)
            out = f(torch.randn(B, S, D, requires_grad=True), msk, torch.randn(B, S, D, requires_grad=True), msk)
        elif name.startswith("fuse"):
            out = f(torch.randn(B, D, requires_grad=True), torch.randn(B, D))
        else:
            lg, lb = f(torch.randn(B, 6, D, requires_grad=True))
            out = lg + lb
        out.backward(torch.zeros_like(out))

    def check(name, m, scale, what):
        for i, (n, g) in enumerate(m.named_gradients()):
            if n.startswith(_unused(name) or ("\0",)):
                ck(g is None, f"{what}/{name}: {n} should have no grad")
            else:
                ck(g is not None and bool(torch.all(g == scale * (i + 1))),
                   f"{what}/{name}: {n} expected {scale * (i + 1)}, got {None if g is None else g.flatten()[:2].tolist()}")

    avg = (1 + world) / 2.0
    # local
    for name, m in mods.items():
        run(name, m)
        ck(m._flat_param.grad is not None and m._flat_param.grad.shape == m._flat_param.shape, f"local/{name}: no fused gradient")
        check(name, m, rank + 1, "local")
        run(name, m)                                          # accumulation
        check(name, m, 2 * (rank + 1), "local/accumulate")
        m.zero_grad(set_to_none=True)
    # native exchange
    Fn.enable_grad_allreduce()
    for name, m in mods.items():
        run(name, m)
        check(name, m, avg, "native")
        m.zero_grad(set_to_none=True)
    Fn.disable_grad_allreduce()
    # the scripts' way: one DistributedDataParallel wrapper per module (train.py:136-139; HoME: find_unused_parameters)
    from torch.nn.parallel import DistributedDataParallel as DDP
    for name, m in mods.items():
        w = DDP(m, find_unused_parameters=bool(_unused(name)))
        ck(len([p for p in w.parameters()]) == 1 + len([n for n, _ in m.named_parameters() if n != "_flat_param"]), f"ddp/{name}: parameter count")
        for _ in range(2):
            run(name, m, call=w)
            check(name, m, avg, "ddp")
            m.zero_grad(set_to_none=True)
    if rank == 0:
        q.put(errors)
    dist.destroy_process_group()


def test_fused_gradients_local_native_and_ddp():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 35500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    errors = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert not errors, errors[:10]
