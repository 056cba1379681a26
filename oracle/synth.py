"""Deterministic synthetic weights and inputs for the parity suite.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (the package
``mmoe-multimodal-rec_b200/``, ``model.py``, ``model_HoME.py``) may import this
file; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline leg do.

The generator is a counter-based splitmix64 hash, written with numpy integer
arithmetic only, so the very same float32 tensors are produced in the
development container (where the golden vectors are made from the real
reference) and on the GPU box (where ``/root/reference`` does not exist).  It
does not depend on torch's or numpy's RNG stream definitions.

Shapes / state-dict keys follow the reference's modules:
``TwoTaskMMoE`` (model.py:527-577), ``RobustTextCrossExpert`` (model.py:386-451),
``EnhancedCrossFuse`` (model.py:454-507), ``HOME_MMoE_Complete``
(model_HoME.py:530-638), ``ItemImageExpert`` (model.py:343-385),
``ImageExpertWithProjection`` (model_HoME.py:373-399).
"""
from __future__ import annotations

import zlib
from collections import OrderedDict

import numpy as np
import torch

_M64 = (1 << 64) - 1


def _splitmix(z: np.ndarray) -> np.ndarray:
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def uniform01(seed: int, n: int, stream: int = 0) -> np.ndarray:
    """n float64 values in [0, 1), a pure function of (seed, stream, index)."""
    base = (seed * 0x9E3779B97F4A7C15 + stream * 0xD1B54A32D192ED03 + 0x2545F4914F6CDD1D) & _M64
    with np.errstate(over="ignore"):
        z = np.arange(n, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(base)
        z = _splitmix(_splitmix(z))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / float(1 << 53))


def uniform_pm1(seed: int, shape, stream: int = 0) -> torch.Tensor:
    n = int(np.prod(shape)) if len(shape) else 1
    u = uniform01(seed, n, stream) * 2.0 - 1.0
    return torch.from_numpy(u.astype(np.float32)).reshape(tuple(shape))


def normal(seed: int, shape, stream: int = 0) -> torch.Tensor:
    """Approximately N(0,1) float32 (Box-Muller on the hashed uniforms)."""
    n = int(np.prod(shape)) if len(shape) else 1
    u1 = uniform01(seed, n, 2 * stream + 1)
    u2 = uniform01(seed, n, 2 * stream + 2)
    r = np.sqrt(-2.0 * np.log(1.0 - u1))
    z = r * np.cos(2.0 * np.pi * u2)
    return torch.from_numpy(z.astype(np.float32)).reshape(tuple(shape))


def randint(seed: int, n: int, lo: int, hi: int, stream: int = 0) -> np.ndarray:
    """n integers in [lo, hi] inclusive."""
    u = uniform01(seed, n, stream)
    return (lo + np.floor(u * (hi - lo + 1))).astype(np.int64).clip(lo, hi)


def _stream_of(key: str) -> int:
    return zlib.crc32(key.encode("utf-8")) & 0x7FFFFFFF


def fill_state_dict(shapes: "OrderedDict[str, tuple]", seed: int) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic, non-degenerate values for every tensor of a state dict.

    * 2-D weights: U(-1,1) * 1.5/sqrt(fan_in)   (slightly hotter than torch's
      default so attention / gates are not near-uniform)
    * 1-D ``*.weight`` (LayerNorm gains): 1 + 0.2*U(-1,1)
    * 1-D ``*.bias`` / ``in_proj_bias``: 0.2*U(-1,1)
    * ``pool.query`` [1,1,d]: N(0,1) (the reference inits N(0,1)/sqrt(d),
      model.py:196; unit scale makes the pooling softmax non-trivial)
    * scalar ``gate`` [1]: 0.3 (reference init 0.5, model.py:411)
    """
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for key, shape in shapes.items():
        shape = tuple(shape)
        st = _stream_of(key)
        if key.endswith("num_batches_tracked"):
            out[key] = torch.zeros(shape, dtype=torch.int64)
        elif len(shape) == 2:
            fan_in = shape[1]
            out[key] = uniform_pm1(seed, shape, st) * (1.5 / float(np.sqrt(fan_in)))
        elif len(shape) == 3:  # pool.query
            out[key] = normal(seed, shape, st)
        elif len(shape) == 1 and shape[0] == 1 and key.split(".")[-1] == "gate":
            out[key] = torch.full(shape, 0.3)
        elif len(shape) == 1 and key.endswith("weight"):
            out[key] = 1.0 + 0.2 * uniform_pm1(seed, shape, st)
        else:
            out[key] = 0.2 * uniform_pm1(seed, shape, st)
    return out


# ----------------------------------------------------------------------------
# state-dict shape tables (must equal the reference modules' state_dict();
# tests/test_oracle_vs_reference.py checks this against /root/reference)
# ----------------------------------------------------------------------------

def _encoder_layer_shapes(prefix: str, d: int, ff: int) -> "OrderedDict[str, tuple]":
    s = OrderedDict()
    s[prefix + "self_attn.in_proj_weight"] = (3 * d, d)
    s[prefix + "self_attn.in_proj_bias"] = (3 * d,)
    s[prefix + "self_attn.out_proj.weight"] = (d, d)
    s[prefix + "self_attn.out_proj.bias"] = (d,)
    s[prefix + "linear1.weight"] = (ff, d)
    s[prefix + "linear1.bias"] = (ff,)
    s[prefix + "linear2.weight"] = (d, ff)
    s[prefix + "linear2.bias"] = (d,)
    s[prefix + "norm1.weight"] = (d,)
    s[prefix + "norm1.bias"] = (d,)
    s[prefix + "norm2.weight"] = (d,)
    s[prefix + "norm2.bias"] = (d,)
    return s


def cross_expert_shapes(d: int = 768, n_layer: int = 2) -> "OrderedDict[str, tuple]":
    """RobustTextCrossExpert.state_dict() (model.py:387-424)."""
    s = OrderedDict()
    s["gate"] = (1,)
    for name in ("self_user", "self_item"):
        for l in range(n_layer):
            s.update(_encoder_layer_shapes(f"{name}.{l}.", d, 4 * d))
    s["cross_attn.in_proj_weight"] = (3 * d, d)
    s["cross_attn.in_proj_bias"] = (3 * d,)
    s["cross_attn.out_proj.weight"] = (d, d)
    s["cross_attn.out_proj.bias"] = (d,)
    s["pool.query"] = (1, 1, d)
    s["norm.weight"] = (d,)
    s["norm.bias"] = (d,)
    s["mlp.0.weight"] = (4 * d, d)
    s["mlp.0.bias"] = (4 * d,)
    s["mlp.3.weight"] = (d, 4 * d)
    s["mlp.3.bias"] = (d,)
    return s


def cross_fuse_shapes(d: int = 768, depth: int = 2) -> "OrderedDict[str, tuple]":
    """EnhancedCrossFuse.state_dict() (model.py:456-489)."""
    s = OrderedDict()
    for l in range(depth):
        s.update(_encoder_layer_shapes(f"layers.{l}.", d, 4 * d))
    s["res_proj.0.weight"] = (d, 2 * d)
    s["res_proj.0.bias"] = (d,)
    s["res_proj.1.weight"] = (d,)
    s["res_proj.1.bias"] = (d,)
    s["gate.0.weight"] = (d // 2, 2 * d)
    s["gate.0.bias"] = (d // 2,)
    s["gate.2.weight"] = (1, d // 2)
    s["gate.2.bias"] = (1,)
    s["proj.0.weight"] = (d,)
    s["proj.0.bias"] = (d,)
    s["proj.1.weight"] = (d, d)
    s["proj.1.bias"] = (d,)
    return s


def mmoe_head_shapes(d: int = 768, n_expert: int = 6, hidden: int = 256) -> "OrderedDict[str, tuple]":
    """TwoTaskMMoE.state_dict() (model.py:532-559)."""
    s = OrderedDict()
    for t in ("good", "best"):
        s[f"gate_{t}.fc.weight"] = (n_expert, d)
        s[f"gate_{t}.fc.bias"] = (n_expert,)
    for t in ("good", "best"):
        s[f"tower_{t}.0.weight"] = (d,)
        s[f"tower_{t}.0.bias"] = (d,)
        s[f"tower_{t}.1.weight"] = (hidden, d)
        s[f"tower_{t}.1.bias"] = (hidden,)
        s[f"tower_{t}.4.weight"] = (hidden // 2, hidden)
        s[f"tower_{t}.4.bias"] = (hidden // 2,)
        s[f"tower_{t}.7.weight"] = (1, hidden // 2)
        s[f"tower_{t}.7.bias"] = (1,)
    return s


def home_head_shapes(n_in: int = 6, d: int = 768, n_shared: int = 4, n_task: int = 2,
                     tower_hidden: int = 256, expert_hidden: int = 1024) -> "OrderedDict[str, tuple]":
    """HOME_MMoE_Complete.state_dict() (model_HoME.py:534-588)."""
    s = OrderedDict()
    s["input_projection.0.weight"] = (d, n_in * d)
    s["input_projection.0.bias"] = (d,)
    s["input_projection.1.weight"] = (d,)
    s["input_projection.1.bias"] = (d,)

    def experts(prefix, n):
        for i in range(n):
            s[f"{prefix}.{i}.0.weight"] = (expert_hidden, d)
            s[f"{prefix}.{i}.0.bias"] = (expert_hidden,)
            s[f"{prefix}.{i}.3.weight"] = (d, expert_hidden)
            s[f"{prefix}.{i}.3.bias"] = (d,)

    experts("meta_experts", n_shared)
    experts("task_experts_good", n_task)
    experts("task_experts_best", n_task)
    for name, n in (("fg_meta", n_shared), ("fg_good", n_task), ("fg_best", n_task)):
        s[f"{name}.gate.weight"] = (d * n, d)
        s[f"{name}.gate.bias"] = (d * n,)
    for name in ("sg_meta", "sg_good", "sg_best"):
        s[f"{name}.gate.0.weight"] = (d, d)
        s[f"{name}.gate.0.bias"] = (d,)
    for t in ("good", "best"):
        s[f"gate_{t}.fc.weight"] = (n_shared + n_task, d)
        s[f"gate_{t}.fc.bias"] = (n_shared + n_task,)
    for t in ("good", "best"):
        s[f"tower_{t}.0.weight"] = (d,)
        s[f"tower_{t}.0.bias"] = (d,)
        s[f"tower_{t}.1.weight"] = (tower_hidden, d)
        s[f"tower_{t}.1.bias"] = (tower_hidden,)
        s[f"tower_{t}.4.weight"] = (1, tower_hidden)
        s[f"tower_{t}.4.bias"] = (1,)
    return s


def image_wrapper_shapes(d: int = 768) -> "OrderedDict[str, tuple]":
    """ItemImageExpert's own parameters (model.py:364); the backbone is HF."""
    return OrderedDict([("norm.weight", (d,)), ("norm.bias", (d,))])


def image_projection_shapes(d: int = 768, proj: int = 768) -> "OrderedDict[str, tuple]":
    """ImageExpertWithProjection.projection_head (model_HoME.py:383-387)."""
    return OrderedDict([
        ("projection_head.0.weight", (2 * d, d)), ("projection_head.0.bias", (2 * d,)),
        ("projection_head.2.weight", (proj, 2 * d)), ("projection_head.2.bias", (proj,)),
    ])


# ----------------------------------------------------------------------------
# inputs
# ----------------------------------------------------------------------------

def sentence_batch(seed: int, B: int, S: int = 64, d: int = 768, stream: int = 0, min_len: int = 1):
    """Sentence vectors ~N(0,1) [B,S,d] and a bool key-padding mask [B,S]
    (True = padded; at least ``min_len`` valid sentences per row, cf. model.py:313-314)."""
    x = normal(seed, (B, S, d), 100 + stream)
    lens = randint(seed, B, min_len, S, 200 + stream)
    lens[0] = S                      # edge case: no padding at all
    if B > 1:
        lens[1] = min_len            # edge case: a single valid sentence
    mask = torch.from_numpy(np.arange(S)[None, :] >= lens[:, None])
    return x, mask


def cross_inputs(seed: int, B: int, S: int = 64, d: int = 768):
    u, um = sentence_batch(seed, B, S, d, stream=0)
    i, im = sentence_batch(seed, B, S, d, stream=1)
    return u, um, i, im


def doc_vectors(seed: int, B: int, d: int = 768, stream: int = 0) -> torch.Tensor:
    return normal(seed, (B, d), 300 + stream)


def expert_vecs(seed: int, B: int, n: int = 6, d: int = 768) -> torch.Tensor:
    return normal(seed, (B, n, d), 400)


def labels(seed: int, B: int, stream: int = 0) -> torch.Tensor:
    return torch.from_numpy((uniform01(seed, B, 500 + stream) < 0.5).astype(np.float32))
