"""Deterministic synthetic weights and inputs (no checkpoints or datasets exist offline).

The generator is keyed by parameter NAME (not by module construction order), so the same
``state_dict`` is reproduced on any box with the same torch build — fixtures under
``tests/golden`` therefore only store outputs plus a checksum of the weights.

Distributions follow PyTorch's default initialisers used by the reference modules
(``nn.Conv1d`` / ``nn.Linear``: U(-1/sqrt(fan_in), 1/sqrt(fan_in)); norms: weight 1, bias 0;
``AttentionPooling.positional_embedding``: N(0,1)/sqrt(dim), reference ``embeddings.py:505``).
Norm affine parameters are additionally perturbed so that parity tests exercise gamma/beta.
"""
from __future__ import annotations

import hashlib
import math
from typing import Dict

import torch

from .arch import UNetConfig, param_shapes


def _seed_for(name: str, seed: int) -> int:
    h = hashlib.sha256(f"{seed}:{name}".encode()).digest()
    return int.from_bytes(h[:7], "little")


def make_state_dict(cfg: UNetConfig, seed: int = 0, perturb_norms: bool = True) -> Dict[str, torch.Tensor]:
    sd: Dict[str, torch.Tensor] = {}
    shapes = param_shapes(cfg)
    for name, shape in shapes.items():
        g = torch.Generator().manual_seed(_seed_for(name, seed))
        leaf = name.rsplit(".", 1)[-1]
        owner = name.rsplit(".", 1)[0]
        is_norm = (len(shape) == 1 and owner.split(".")[-1].startswith(("norm", "conv_norm_out")))
        if name.endswith("positional_embedding"):
            t = torch.randn(shape, generator=g) / math.sqrt(shape[-1])
        elif is_norm:
            if leaf == "weight":
                t = torch.ones(shape)
                if perturb_norms:
                    t = t + 0.1 * torch.randn(shape, generator=g)
            else:
                t = torch.zeros(shape)
                if perturb_norms:
                    t = 0.1 * torch.randn(shape, generator=g)
        else:
            wshape = shapes[owner + ".weight"]
            fan_in = 1
            for d in wshape[1:]:
                fan_in *= d
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        sd[name] = t.to(torch.float32).contiguous()
    return sd


def make_pre_state_dict(cfg: dict, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Synthetic `Pre_model` weights keyed by parameter name: norms 1 + 0.1 N(0,1) / 0.1 N(0,1), everything else
    U(-b, b) with b = fan_in^-1/2 (ConvTBC weights [k, c_in, c_out]: fan_in = k c_in)."""
    from .pre_model import pre_param_shapes
    sd: Dict[str, torch.Tensor] = {}
    for name, shape in pre_param_shapes(cfg).items():
        g = torch.Generator().manual_seed(_seed_for(name, seed))
        owner, leaf = name.rsplit(".", 1)
        if len(shape) == 1 and "norm" in owner.rsplit(".", 1)[-1]:
            t = (1.0 if leaf == "weight" else 0.0) + 0.1 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for d in (shape[1:] if len(shape) > 1 else shape):
                fan_in *= d
            if name.endswith("conv.weight") and len(shape) == 3 and "spk_proj" not in name:
                fan_in = shape[0] * shape[1]
            t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(max(fan_in, 1))
        sd[name] = t.to(torch.float32).contiguous()
    return sd


def make_pre_inputs(B: int, T: int, S: int, content_ch: int = 256, ragged: bool = False, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Synthetic `Pre_model.infer` inputs: c ~ N(0,1) [B, 256, T] (ContentVec features), refer ~ N(0,1) [B, 100, S] (mel prompt)."""
    g = torch.Generator().manual_seed(seed + 11)
    c = torch.randn((B, content_ch, T), generator=g)
    refer = torch.randn((B, 100, S), generator=g)
    if ragged:
        lengths = torch.tensor([max(1, T - 37 * (i % 8)) for i in range(B)], dtype=torch.int64)
        refer_lengths = torch.tensor([max(1, S - 17 * (i % 8)) for i in range(B)], dtype=torch.int64)
    else:
        lengths = torch.full((B,), T, dtype=torch.int64)
        refer_lengths = torch.full((B,), S, dtype=torch.int64)
    return dict(c=c, refer=refer, lengths=lengths, refer_lengths=refer_lengths)


def state_dict_checksum(sd: Dict[str, torch.Tensor]) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()[:16]


def make_inputs(B: int, T: int, S: int, latent_ch: int = 100, content_ch: int = 256,
                ragged: bool = False, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Synthetic denoiser inputs per SURVEY.md §8(d): content ~ N(0,1) [T,B,256] (seed+1),
    prompt ~ N(0,1) [S,B,256] (seed+2), x_T ~ N(0,1) [B,100,T] (seed+3); ragged prompt lengths
    S - 17*i (i = sample index mod 8), floored at 1, for parity runs."""
    def gen(k):
        return torch.Generator().manual_seed(seed + k)
    content = torch.randn((T, B, content_ch), generator=gen(1))
    prompt = torch.randn((S, B, content_ch), generator=gen(2))
    x = torch.randn((B, latent_ch, T), generator=gen(3))
    lengths = torch.full((B,), T, dtype=torch.int64)
    if ragged:
        refer_lengths = torch.tensor([max(1, S - 17 * (i % 8)) for i in range(B)], dtype=torch.int64)
    else:
        refer_lengths = torch.full((B,), S, dtype=torch.int64)
    return dict(content=content, prompt=prompt, x=x, lengths=lengths, refer_lengths=refer_lengths)


def linear_betas(timesteps: int = 1000) -> torch.Tensor:
    """fp32 copy of the reference's linear beta schedule (``model.py:426-433, 473``):
    float64 linspace cast to float32 by ``register_buffer``."""
    scale = 1000 / timesteps
    return torch.linspace(scale * 0.0001, scale * 0.02, timesteps, dtype=torch.float64).to(torch.float32)
