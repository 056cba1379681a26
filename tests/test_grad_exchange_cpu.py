"""The gradient hand-over of the autograd wrappers, end to end through torch autograd, on CPU (gloo, world_size 2).

The CUDA library is replaced by a stub that only writes recognisable values into the gradient buffers it is handed
(rank- and parameter-dependent, ACCUMULATING like the real kernels do), so what is tested is exactly the part the
round-1 review found broken: which buffer ends up in ``p.grad`` and whether it is the all-reduced one.

  * native exchange on : every ``.grad`` of every module equals the average over ranks, is identical on both ranks,
    aliases the flat buffer the collective ran on (no copy), also after a second (accumulating) backward and after
    ``zero_grad(set_to_none=False)``; unused HoME parameters keep ``grad is None``;
  * native exchange off: gradients are the local ones and AccumulateGrad adopted the views (no clone), also on the
    staged cross-expert path.
"""
import ctypes as C
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _StubLib:
    """Stands in for libmmoe_b200.so: size queries return small numbers, forwards do nothing, backwards add
    (rank + 1) * (index + 1) to every gradient buffer of the stage they are asked to run."""

    def __init__(self, rank):
        self.rank = rank
        self.numels = {}          # parameter sizes of the module under test, keyed by its parameter count
        self.calls = []

    def register(self, n_params, numels):
        self.numels[n_params] = list(numels)

    # -- helpers
    def _fill(self, call, idx):
        c = call._obj
        numels = self.numels[self._n_current]
        for i in idx:
            addr = c.grads[i]
            if not addr:
                continue
            buf = np.ctypeslib.as_array((C.c_float * numels[i]).from_address(addr))
            buf += float((self.rank + 1) * (i + 1))

    def __getattr__(self, name):
        if name.endswith("_saved_bytes") or name.endswith("_workspace_bytes"):
            return lambda *a: 64
        if name == "mmoe_last_error":
            return lambda: b"stub"
        if name in ("mmoe_head_fwd", "mmoe_home_fwd", "mmoe_cross_fwd", "mmoe_fuse_fwd", "mmoe_img_pool_fwd", "mmoe_cast_f32"):
            return lambda *a: 0
        if name in ("mmoe_head_bwd", "mmoe_home_bwd", "mmoe_fuse_bwd", "mmoe_cross_bwd"):
            def bwd(call, *a):
                self._fill(call, range(self._n_current))
                return 0
            return bwd
        if name == "mmoe_cross_bwd_stage":
            def stage_bwd(call, cfg, stage, *a):
                n_layer = cfg._obj.n_layer
                if stage == 0:
                    idx = [0] + list(range(1 + 24 * n_layer, self._n_current))
                elif stage >= 200:
                    l = stage - 200
                    idx = range(1 + 12 * n_layer + 12 * l, 1 + 12 * n_layer + 12 * (l + 1))
                else:
                    l = stage - 100
                    idx = range(1 + 12 * l, 1 + 12 * (l + 1))
                self.calls.append(stage)
                self._fill(call, idx)
                return 0
            return stage_bwd
        raise AttributeError(name)


def _expected(params, used, scale):
    return [None if not u else scale * (i + 1) for i, (p, u) in enumerate(zip(params, used))]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mmoe_multimodal_rec_b200 as pkg
    Fn, M, H = pkg.functional, pkg.modules, pkg.modules_home
    stub = _StubLib(rank)
    Fn.lib = lambda: stub
    Fn._require_cuda = lambda *a: None
    Fn._stream = lambda: 0
    errors = []

    def ck(cond, what):
        if not cond:
            errors.append(what)

    torch.manual_seed(0)
    d = 64
    Fn.set_flat_parameters(False)          # this file tests the per-tensor parameter mode (fused mode: test_flat_params_cpu.py)
    mods = {
        "cross": M.RobustTextCrossExpert(d=d, n_layer=2, n_head=8),
        "cross_home": H.RobustTextCrossExpert(d=d, n_layer=2, n_head=8),
        "fuse": M.EnhancedCrossFuse(d=d, n_head=8, depth=2),
        "fuse_home": H.EnhancedCrossFuse(d=d, n_head=8, depth=2),
        "head": M.TwoTaskMMoE(expert_dim=d, n_expert=6, tower_hidden=32),
    }

    def used_of(name, m):
        names = [n for n, _ in m.named_parameters()]
        if name == "cross_home":
            return [not (n.startswith("norm.") or n.startswith("mlp.")) for n in names]
        if name == "fuse_home":
            return [not n.startswith("proj.") for n in names]
        return [True] * len(names)

    def run(name, m):
        params = list(m.parameters())
        stub.register(len(params), [p.numel() for p in params])
        stub._n_current = len(params)
        B, S = 3, 4
        if name.startswith("cross"):
            u = torch.randn(B, S, d, requires_grad=True)
            i = torch.randn(B, S, d, requires_grad=True)
            msk = torch.zeros(B, S, dtype=torch.bool)
            out = m(u, msk, i, msk)
        elif name.startswith("fuse"):
            out = m(torch.randn(B, d, requires_grad=True), torch.randn(B, d))
        else:
            lg, lb = m(torch.randn(B, 6, d, requires_grad=True))
            out = lg + lb
        out.backward(torch.zeros_like(out))

    def grads_equal(m, expect, what):
        for (n, p), e in zip(m.named_parameters(), expect):
            if e is None:
                ck(p.grad is None, f"{what}: {n} should have no grad")
            else:
                ck(p.grad is not None and bool(torch.all(p.grad == e)), f"{what}: {n} expected {e}, got "
                   f"{None if p.grad is None else p.grad.flatten()[:2].tolist()}")

    def aliases_one_buffer(m, used):
        """consecutive gradients sit at the spacing of the flat layout (numel rounded up to 64 floats): nobody cloned them"""
        ps = [p for p, u in zip(m.parameters(), used) if u]
        off = 0
        base = ps[0].grad.data_ptr()
        ok = True
        for p in ps:
            ok &= (p.grad.data_ptr() - base) == off * 4
            off += (p.numel() + 63) // 64 * 64
        return ok

    # ---------------- exchange off: local gradients, adopted without a copy ----------------
    for name, m in mods.items():
        used = used_of(name, m)
        run(name, m)
        grads_equal(m, _expected(list(m.parameters()), used, rank + 1), f"off/{name}")
        # (also on the staged cross-expert path: each stage's span lies inside the one shared flat buffer)
        ck(aliases_one_buffer(m, used), f"off/{name}: gradients were cloned")
        m.zero_grad(set_to_none=True)
    ck(stub.calls[:5] == [0, 201, 200, 101, 100], f"stage order {stub.calls[:5]}")

    # ---------------- exchange on ----------------
    sync = Fn.enable_grad_allreduce()
    avg = (1 + world) / 2.0
    for name, m in mods.items():
        used = used_of(name, m)
        params = list(m.parameters())
        run(name, m)
        ck(not sync.pending and not sync.deferred, f"on/{name}: exchange not finished at the end of backward")
        grads_equal(m, _expected(params, used, avg), f"on/{name}")
        ck(aliases_one_buffer(m, used), f"on/{name}: .grad does not alias the reduced buffer")
        ptrs = [p.grad.data_ptr() for p, u in zip(params, used) if u]
        # second backward without zero_grad: accumulates into the same buffers
        run(name, m)
        grads_equal(m, _expected(params, used, 2 * avg), f"on/{name}/accumulate")
        ck(ptrs == [p.grad.data_ptr() for p, u in zip(params, used) if u], f"on/{name}: accumulation re-allocated .grad")
        # zero_grad(set_to_none=False) keeps the tensors
        m.zero_grad(set_to_none=False)
        run(name, m)
        grads_equal(m, _expected(params, used, avg), f"on/{name}/zeroed")
        # identical across ranks, bit for bit
        for p, u in zip(params, used):
            if u:
                other = p.grad.clone()
                dist.broadcast(other, src=0)
                ck(bool(torch.equal(other, p.grad)), f"on/{name}: ranks disagree")
        # local accumulation (no_sync equivalent) then one reduce of the accumulated buffers
        m.zero_grad(set_to_none=True)
        Fn.set_grad_sync(False)
        run(name, m)
        run(name, m)
        grads_equal(m, _expected(params, used, 2 * (rank + 1)), f"on/{name}/no_sync")
        Fn.set_grad_sync(True)
        Fn.allreduce_accumulated([m])
        grads_equal(m, _expected(params, used, 2 * avg), f"on/{name}/allreduce_accumulated")
        m.zero_grad(set_to_none=True)
    Fn.disable_grad_allreduce()
    if rank == 0:
        q.put(errors)
    dist.destroy_process_group()


def test_grad_handover_and_native_exchange_through_autograd():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    errors = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert not errors, errors[:10]
