"""SURVEY.md §8f rows on the GPU: the native kernels (through the C ABI / the package's wrappers) against the oracle
(oracle/next_rows.py) and the golden vectors the reference's own code produced (tests/golden/next_*.pt)."""
import ctypes as Ct
import types

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import make_golden_next as G
from oracle import next_rows as N
from oracle import synth

pytestmark = pytest.mark.gpu


def nerr(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    if not torch.isfinite(a).all():
        return float("inf")
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)


def _pkg():
    import mmoe_multimodal_rec_b200 as pkg
    return pkg


def _wrappers(w, b, rm, rv, train, p=0.0):
    HW = _pkg().home_wrap
    ws = []
    for e in range(len(w)):
        m = HW.HomeExpertWrapper(768, dropout_p=p).cuda()
        with torch.no_grad():
            m.norm.weight.copy_(w[e]); m.norm.bias.copy_(b[e]); m.norm.running_mean.copy_(rm[e]); m.norm.running_var.copy_(rv[e])
        ws.append(m.train(train))
    return ws


# ------------------------------------------------------------------------------------------ (f)-1
@pytest.mark.parametrize("mode", ["train", "eval"])
def test_fused_wrapper_stack_matches_reference_golden(mode):
    g = load_golden("next_wrapper_b16")[mode]
    xs, w, b, rm, rv, cot = G.wrapper_inputs()
    ws = _wrappers(w, b, rm, rv, mode == "train")
    stack = _pkg().home_wrap.FusedHomeExpertStack(ws)
    xin = [x.cuda().requires_grad_(True) for x in xs]
    out = stack(*xin)
    out.backward(cot.cuda())
    assert nerr(out, g["out"]) <= 1e-5
    for e in range(6):
        assert nerr(xin[e].grad, g["dx"][e]) <= 1e-4, e
        assert nerr(ws[e].norm.weight.grad, g["dgamma"][e]) <= 1e-4 and nerr(ws[e].norm.bias.grad, g["dbeta"][e]) <= 1e-4
        assert nerr(ws[e].norm.running_mean, g["running_mean"][e]) <= 1e-5 and nerr(ws[e].norm.running_var, g["running_var"][e]) <= 1e-5
        assert int(ws[e].norm.num_batches_tracked) == g["num_batches_tracked"][e]
    # a single wrapper called on its own (the script's call pattern) gives the same column of the stack
    w0 = _wrappers(w[:1], b[:1], rm[:1], rv[:1], mode == "train")[0]
    assert nerr(w0(xs[0].cuda()), g["out"][:, 0]) <= 1e-5


def test_fused_wrapper_stack_train_dropout_and_benchmark_batch():
    """Dropout active (p = 0.1) at B = 512: the kernel's keep-mask (site 0 over the stacked [B,6,d] index) fed to the oracle."""
    import parity_util as PU
    pkg = _pkg()
    L = pkg.lib()
    B, d, n, p = 512, 768, 6, 0.1
    xs = [synth.normal(401, (B, d), e) * (1 + 0.2 * e) for e in range(n)]
    w = [1.0 + 0.2 * synth.uniform_pm1(401, (d,), 30 + e) for e in range(n)]
    b = [0.2 * synth.uniform_pm1(401, (d,), 50 + e) for e in range(n)]
    rm = [torch.zeros(d) for _ in range(n)]
    rv = [torch.ones(d) for _ in range(n)]
    cot = synth.normal(402, (B, n, d), 1)
    ws = _wrappers(w, b, rm, rv, True, p)
    stack = pkg.home_wrap.FusedHomeExpertStack(ws)
    xin = [x.cuda().requires_grad_(True) for x in xs]
    seed = PU.call_seed(77)
    torch.manual_seed(77)
    out = stack(*xin)
    out.backward(cot.cuda())
    k0, k1 = Ct.c_uint32(), Ct.c_uint32()
    assert L.mmoe_site_keys(seed, 0, Ct.byref(k0), Ct.byref(k1)) == 0
    keep = torch.empty(B * n * d, dtype=torch.uint8, device="cuda")
    assert L.mmoe_dropout_mask(k0.value, k1.value, p, keep.numel(), keep.data_ptr(), torch.cuda.current_stream().cuda_stream) == 0
    keep = keep.view(B, n, d).double()
    assert abs(float(keep.mean()) - 0.9) < 2e-3
    xo = [x.double().cuda().requires_grad_(True) for x in xs]
    wo = [t.double().cuda().requires_grad_(True) for t in w]
    bo = [t.double().cuda().requires_grad_(True) for t in b]
    o, _, _ = N.home_wrapper_stack(xo, wo, bo, [t.double().cuda() for t in rm], [t.double().cuda() for t in rv], True,
                                   drop=lambda site, t: t * keep / (1 - p))
    o.backward(cot.double().cuda())
    assert nerr(out, o) <= 1e-5
    for e in range(n):
        assert nerr(xin[e].grad, xo[e].grad) <= 1e-4 and nerr(ws[e].norm.weight.grad, wo[e].grad) <= 1e-4 and nerr(ws[e].norm.bias.grad, bo[e].grad) <= 1e-4


# ------------------------------------------------------------------------------------------ (f)-4
def test_two_task_bce_matches_reference_golden():
    g = load_golden("next_losses")["bce2"]
    B = 64
    lg, lb = synth.normal(311, (B,), 1) * 2.0, synth.normal(311, (B,), 2) * 2.0
    yg = torch.from_numpy((synth.uniform01(311, B, 3) < 0.5).astype(np.float32))
    yb = torch.from_numpy((synth.uniform01(311, B, 4) < 0.3).astype(np.float32))
    a, b = lg.cuda().requires_grad_(True), lb.cuda().requires_grad_(True)
    loss = _pkg().losses.TwoTaskBCEWithLogits()(a, b, yg.cuda(), yb.cuda())
    (3.0 * loss).backward()
    assert abs(float(loss) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert nerr(a.grad, 3.0 * g["dgood"]) <= 1e-5 and nerr(b.grad, 3.0 * g["dbest"]) <= 1e-5


@pytest.mark.parametrize("mode", ["fp32", "bf16", "fp16"])
def test_info_nce_matches_reference_golden_and_oracle(mode):
    import parity_util as PU
    g = load_golden("next_losses")["info_nce"]
    tol = PU.TOL[mode]
    t = {k: synth.normal(312, (16, 768), i).cuda().requires_grad_(True) for i, k in enumerate(("ui", "idoc", "udoc", "proj"))}
    with PU.autocast_ctx(mode):
        losses = _pkg().losses.info_nce_losses([(t["ui"], t["idoc"]), (t["udoc"], t["proj"]), (t["idoc"], t["proj"])])
    (0.5 * losses[0] + 0.7 * losses[1] + 1.3 * losses[2]).backward()
    assert nerr(losses, g["loss"]) <= tol
    for k in t:
        assert nerr(t[k].grad, g["grads"][k]) <= tol, k
    # B = 512 (the training batch), against the fp64 oracle on the GPU
    B = 512
    a0, p0 = synth.normal(313, (B, 768), 1), synth.normal(313, (B, 768), 2) + 0.5 * synth.normal(313, (B, 768), 1)
    a, p = a0.cuda().requires_grad_(True), p0.cuda().requires_grad_(True)
    with PU.autocast_ctx(mode):
        l = _pkg().losses.info_nce_losses([(a, p)])[0]
    l.backward()
    ao, po = a0.double().cuda().requires_grad_(True), p0.double().cuda().requires_grad_(True)
    lo = N.info_nce(ao, po)
    lo.backward()
    assert abs(float(l) - float(lo)) <= tol * abs(float(lo))
    assert nerr(a.grad, ao.grad) <= tol and nerr(p.grad, po.grad) <= tol


@pytest.mark.parametrize("n", [1, 7, 2048, 4099, 22281, 65536, 100003])
def test_device_auc_equals_rank_statistic(n):
    """n = 22281 is the reference's validation split (README.md:25-27); ties included; bit-for-bit the rank statistic."""
    s = np.round(synth.normal(500 + n % 97, (n,), 1).numpy() * 8.0) / 8.0 if n % 2 else synth.normal(500 + n % 97, (n,), 1).numpy()
    y = (synth.uniform01(501 + n % 97, n, 3) < 0.35).astype(np.float32)
    got = _pkg().losses.roc_auc(torch.from_numpy(s.astype(np.float32)).cuda(), torch.from_numpy(y).cuda())
    ref = N.roc_auc(s.astype(np.float32), y)
    if np.isnan(ref):
        assert torch.isnan(got)
    else:
        assert abs(float(got) - ref) < 1e-12, (float(got), ref)


def test_device_auc_matches_sklearn_golden():
    g = load_golden("next_losses")["auc"]
    n = 5000
    s = np.round(synth.normal(313, (n,), 1).numpy() * 4.0) / 4.0 + (synth.uniform01(313, n, 2) < 0.5) * 0.5
    y = (synth.uniform01(313, n, 3) < 0.35).astype(np.float32)
    got = _pkg().losses.roc_auc(torch.from_numpy(s.astype(np.float32)).cuda(), torch.from_numpy(y).cuda())
    assert abs(float(got) - g["ties"]) < 1e-12


# ------------------------------------------------------------------------------------------ X1 / (f)-3
@pytest.mark.parametrize("mode", ["fp32", "bf16", "fp16"])
def test_patch_projection_from_bytes_and_from_images_matches_hf_golden(mode):
    import parity_util as PU
    g = load_golden("next_patches")
    raw = G.patch_bytes(g["B"])
    w, bias = G.patch_weights()
    conv = torch.nn.Conv2d(3, 768, 16, 16)
    with torch.no_grad():
        conv.weight.copy_(w); conv.bias.copy_(bias)
    pe = _pkg().ingest.NativePatchEmbeddings(conv.cuda())
    raw_t = torch.from_numpy(raw).cuda()
    with PU.autocast_ctx(mode):
        from_bytes = pe(raw_t)
        from_images = pe(_pkg().ingest.decode_patch_bytes(raw_t))
    tol = PU.TOL[mode]
    assert from_bytes.shape == (g["B"], 196, 768)
    assert nerr(from_bytes.float(), g["tokens"]) <= tol and nerr(from_images.float(), g["tokens"]) <= tol
    # the float image the device-side decode builds is the one decode_sample builds
    img = _pkg().ingest.decode_patch_bytes(raw_t)
    assert nerr(img, torch.stack([N.unpatchify_normalise(raw[b]) for b in range(g["B"])])) <= 1e-6


def test_patch_projection_inside_hf_vit_at_batch_512_rows():
    """install_native_patch_embeddings on a random-init HF ViTModel: same last_hidden_state as the stock model (bf16 tolerance),
    M = 64*196 = 12544 rows -> the CTA-pair GEMM."""
    from transformers import ViTConfig, ViTModel
    torch.manual_seed(0)
    vit = ViTModel(ViTConfig(num_hidden_layers=1)).cuda().eval()
    images = torch.randn(64, 3, 224, 224, device="cuda")
    with torch.no_grad():
        ref = vit(pixel_values=images).last_hidden_state
        _pkg().ingest.install_native_patch_embeddings(vit)
        L = _pkg().lib()
        L.mmoe_launch_trace(1)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            got = vit(pixel_values=images).last_hidden_state
        n = L.mmoe_launch_trace_read(None, 0)
        buf = (Ct.c_int32 * (4 * max(n, 1)))()
        L.mmoe_launch_trace_read(buf, n)
        L.mmoe_launch_trace(0)
    assert n >= 1 and buf[1] == 2            # the patch projection ran on the pair kernel
    assert nerr(got.float(), ref) <= 2e-2


def test_item_image_expert_accepts_patch_bytes():
    """ItemImageExpert fed with patch.bin bytes == fed with the float image decode_sample would have built from them."""
    from transformers import ViTConfig, ViTModel
    torch.manual_seed(1)
    M = _pkg().modules
    expert = M.ItemImageExpert(ViTModel(ViTConfig(num_hidden_layers=1)), pool_type="mean").cuda().eval()
    raw = torch.from_numpy(G.patch_bytes(4, seed=777)).cuda()
    with torch.no_grad():
        ref = expert(_pkg().ingest.decode_patch_bytes(raw))                  # stock HF conv path
        with torch.autocast("cuda", dtype=torch.bfloat16):
            got = expert(raw)                                                  # native byte path
    assert got.shape == (4, 768) and nerr(got, ref) <= 2e-2


# ------------------------------------------------------------------------------------------ (f)-2
@pytest.mark.parametrize("which", ["model", "model_HoME"])
def test_sentence_gather_matches_reference_golden(which):
    g = load_golden("next_gather")[which]
    h, c2s, pos, S = G.gather_inputs()
    from mmoe_multimodal_rec_b200 import text_data as TD
    enc = G._FakeEncoder(h.clone()).cuda()
    cls = TD.TextExpert if which == "model" else TD.TextExpertHoME
    te = cls(enc, types.SimpleNamespace(pad_token_id=0)).cuda().eval()
    if which == "model":
        with torch.no_grad():
            te.norm.weight.copy_(g["norm_w"]); te.norm.bias.copy_(g["norm_b"])
    ids = [[1] * h.shape[1] for _ in c2s]
    sent, mask, doc = te(ids, c2s, pos, S, trainable=True) if which == "model" else te(ids, c2s, pos, S)
    cs, cd = synth.normal(333, tuple(sent.shape), 1), synth.normal(333, tuple(doc.shape), 2)
    torch.autograd.backward([sent, doc], [cs.cuda(), cd.cuda()])
    assert nerr(sent, g["sent"]) <= 1e-5 and nerr(doc, g["doc"]) <= 1e-5 and torch.equal(mask.cpu(), g["mask"])
    assert nerr(enc.h.grad, g["dh"]) <= 1e-4
    if which == "model":
        assert nerr(te.norm.weight.grad, g["dgamma"]) <= 1e-4 and nerr(te.norm.bias.grad, g["dbeta"]) <= 1e-4


def test_sentence_gather_train_mode_and_16bit_hidden_states():
    """Dropout sites 0 (sentence rows) and 1 (doc vectors) fed to the oracle; bf16 hidden states as under autocast; a
    realistic ragged batch (B = 64, up to 4 chunks of up to 24 sentences, 64 slots)."""
    import parity_util as PU
    pkg = _pkg()
    L = pkg.lib()
    rng = np.random.RandomState(5)
    B, S, d, seq, p = 64, 64, 768, 384, 0.1
    c2s, pos = [], []
    for b in range(B):
        for _ in range(rng.randint(0, 5) if b != 3 else 0):           # sample 3 has no chunk; the last sample always has one
            c2s.append(b)
            k = rng.randint(1, 25)
            pos.append(sorted(rng.choice(np.arange(1, seq), size=k, replace=False).tolist()) + [-1] * (24 - k))
    c2s.append(B - 1); pos.append([1, 2, 3] + [-1] * 21)
    h = synth.normal(601, (len(c2s), seq, d), 1).to(torch.bfloat16)
    norm = torch.nn.LayerNorm(d).cuda()
    with torch.no_grad():
        norm.weight.copy_(1.0 + 0.2 * synth.uniform_pm1(602, (d,), 1)); norm.bias.copy_(0.2 * synth.uniform_pm1(602, (d,), 2))
    hc = h.cuda().requires_grad_(True)
    seed = PU.call_seed(91)
    torch.manual_seed(91)
    sent, mask, doc = pkg.ingest.sentence_gather(hc, c2s, pos, S, norm, p, True)
    cs, cd = synth.normal(603, tuple(sent.shape), 1), synth.normal(603, tuple(doc.shape), 2)
    torch.autograd.backward([sent, doc], [cs.cuda(), cd.cuda()])

    def keep(site, numel):
        k0, k1 = Ct.c_uint32(), Ct.c_uint32()
        assert L.mmoe_site_keys(seed, site, Ct.byref(k0), Ct.byref(k1)) == 0
        m = torch.empty(numel, dtype=torch.uint8, device="cuda")
        assert L.mmoe_dropout_mask(k0.value, k1.value, p, numel, m.data_ptr(), torch.cuda.current_stream().cuda_stream) == 0
        return m.cpu().double()
    ks, kd = keep(0, sent.numel()).view(sent.shape), keep(1, doc.numel()).view(doc.shape)
    ho = h.double().requires_grad_(True)
    wo, bo = norm.weight.detach().double().cpu().requires_grad_(True), norm.bias.detach().double().cpu().requires_grad_(True)
    so, mo, do = N.sentence_gather(ho, c2s, pos, S, wo, bo, drop=lambda site, t: t * (ks if site == "sent" else kd) / (1 - p))
    torch.autograd.backward([so, do], [cs.double(), cd.double()])
    assert torch.equal(mask.cpu(), mo)
    assert nerr(sent, so) <= 1e-5 and nerr(doc, do) <= 1e-5
    assert nerr(hc.grad.float(), ho.grad) <= 1e-2                      # gradient handed back in bf16, the dtype of h
    assert nerr(norm.weight.grad, wo.grad) <= 1e-4 and nerr(norm.bias.grad, bo.grad) <= 1e-4


# ------------------------------------------------------------------------------------------ out-of-bounds writes
def test_new_kernels_do_not_write_outside_their_outputs():
    """compute-sanitizer is closed on this pool, so the new kernels are called through the C ABI with every output embedded
    in a larger buffer of sentinels (ragged sizes on purpose); the bands around the outputs must come back untouched."""
    import mmoe_multimodal_rec_b200 as pkg
    from mmoe_multimodal_rec_b200 import _lib as LB
    L = pkg.lib()
    st = torch.cuda.current_stream().cuda_stream
    PAD = 4096

    class Guard:
        def __init__(self, numel, dtype):
            self.buf = torch.full((numel + 2 * PAD,), 77, dtype=dtype, device="cuda")
            self.view = self.buf[PAD:PAD + numel]
            self.n = numel

        def ok(self):
            return bool((self.buf[:PAD] == 77).all()) and bool((self.buf[PAD + self.n:] == 77).all())

    def ptrs(ts):
        return LB.ptr_array([t.data_ptr() for t in ts])
    # 6-way BN-SiLU-dropout-stack, B = 37
    B, n, d = 37, 6, 768
    xs = [torch.randn(B, d, device="cuda") for _ in range(n)]
    prm = [torch.ones(d, device="cuda") if i % 2 == 0 else torch.zeros(d, device="cuda") for i in range(2 * n)]
    rm, rv = [torch.zeros(d, device="cuda") for _ in range(n)], [torch.ones(d, device="cuda") for _ in range(n)]
    out, sm, sr = Guard(B * n * d, torch.float32), Guard(n * d, torch.float32), Guard(n * d, torch.float32)
    c = LB.Call(); c.dtype, c.B, c.training, c.drop_p, c.seed, c.stream = 0, B, 1, 0.1, 5, st
    pa = ptrs(prm); c.params = Ct.cast(pa, Ct.POINTER(Ct.c_void_p))
    assert L.mmoe_bn_silu_stack_fwd(Ct.byref(c), n, d, ptrs(xs), out.view.data_ptr(), sm.view.data_ptr(), sr.view.data_ptr(), ptrs(rm), ptrs(rv), 0.1, 1e-5) == 0
    dxs = [Guard(B * d, torch.float32) for _ in range(n)]
    grads = [Guard(d, torch.float32) for _ in range(2 * n)]
    ga = ptrs([g.view for g in grads]); c.grads = Ct.cast(ga, Ct.POINTER(Ct.c_void_p))
    dout = torch.randn(B, n, d, device="cuda")
    assert L.mmoe_bn_silu_stack_bwd(Ct.byref(c), n, d, ptrs(xs), dout.data_ptr(), sm.view.data_ptr(), sr.view.data_ptr(), ptrs(rm), ptrs(rv),
                                    ptrs([g.view for g in dxs]), 1e-5) == 0
    torch.cuda.synchronize()
    assert out.ok() and sm.ok() and sr.ok() and all(g.ok() for g in dxs) and all(g.ok() for g in grads)
    # sentence gather, B = 5, S = 7 (not a multiple of the 8 warps), d = 768
    Bs, S = 5, 7
    h = torch.randn(11 * 13, d, device="cuda")
    src = torch.randint(-1, 11 * 13, (Bs, S), dtype=torch.int32, device="cuda")
    sent, mask, doc = Guard(Bs * S * d, torch.float32), Guard(Bs * S, torch.uint8), Guard(Bs * d, torch.float32)
    pre, stats = Guard(Bs * d, torch.float32), Guard(Bs * (S + 1) * 2, torch.float32)
    c2 = LB.Call(); c2.dtype, c2.B, c2.training, c2.drop_p, c2.seed, c2.stream = 0, Bs, 1, 0.1, 9, st
    nw = [torch.ones(d, device="cuda"), torch.zeros(d, device="cuda")]
    pa2 = ptrs(nw); c2.params = Ct.cast(pa2, Ct.POINTER(Ct.c_void_p))
    assert L.mmoe_sent_gather_fwd(Ct.byref(c2), S, d, h.data_ptr(), 0, src.data_ptr(), sent.view.data_ptr(), mask.view.data_ptr(), doc.view.data_ptr(),
                                  pre.view.data_ptr(), stats.view.data_ptr()) == 0
    dh, dgw = Guard(h.numel(), torch.float32), [Guard(d, torch.float32), Guard(d, torch.float32)]
    dh.view.zero_(); [g.view.zero_() for g in dgw]
    ga2 = ptrs([g.view for g in dgw]); c2.grads = Ct.cast(ga2, Ct.POINTER(Ct.c_void_p))
    ds, dd = torch.randn(Bs, S, d, device="cuda"), torch.randn(Bs, d, device="cuda")
    assert L.mmoe_sent_gather_bwd(Ct.byref(c2), S, d, h.data_ptr(), 0, src.data_ptr(), mask.view.data_ptr(), pre.view.data_ptr(), stats.view.data_ptr(),
                                  ds.data_ptr(), dd.data_ptr(), dh.view.data_ptr()) == 0
    torch.cuda.synchronize()
    assert sent.ok() and mask.ok() and doc.ok() and pre.ok() and stats.ok() and dh.ok() and all(g.ok() for g in dgw)
    # patch bytes -> operand, float image -> operand (3 images), bf16
    raw = torch.randint(0, 256, (3 * 196, 768), dtype=torch.uint8, device="cuda")
    op1, op2 = Guard(3 * 196 * 768, torch.bfloat16), Guard(3 * 196 * 768, torch.bfloat16)
    assert L.mmoe_patch_u8_to_operand(raw.data_ptr(), op1.view.data_ptr(), 3 * 196, 768, 1, st) == 0
    img = torch.randn(3, 3, 224, 224, device="cuda")
    assert L.mmoe_patchify(img.data_ptr(), op2.view.data_ptr(), 3, 3, 224, 224, 16, 1, st) == 0
    # AUC workspace, n = 4099
    nb = L.mmoe_auc_workspace_bytes(4099)
    work, res = Guard(nb // 8, torch.int64), Guard(1, torch.float64)
    sc, lb = torch.randn(4099, device="cuda"), (torch.rand(4099, device="cuda") < 0.5).float()
    assert L.mmoe_auc(sc.data_ptr(), lb.data_ptr(), 4099, work.view.data_ptr(), nb, res.view.data_ptr(), st) == 0
    # BCE
    lg = Guard(2 * 37, torch.float32); lossb = Guard(1, torch.float32); lossb.view.zero_()
    x = torch.randn(2 * 37, device="cuda"); yy = (torch.rand(37, device="cuda") < 0.5).float()
    assert L.mmoe_bce2_fwd_bwd(x.data_ptr(), yy.data_ptr(), yy.data_ptr(), 1.0, 2.0, 37, lossb.view.data_ptr(), lg.view.data_ptr(), 1.0, st) == 0
    torch.cuda.synchronize()
    assert op1.ok() and op2.ok() and work.ok() and res.ok() and lg.ok() and lossb.ok()
    assert torch.equal(op1.view.view(3 * 196, 768).float(), raw.float())
