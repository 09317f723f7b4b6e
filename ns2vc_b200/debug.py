"""Diagnostics for the parity tests: run a forward and collect the engine's per-op activations."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from .arch import level_lengths


def forward_with_taps(unet, sample: torch.Tensor, timestep, ehs: torch.Tensor,
                      mask: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """Returns (output [B,Cout,T], {op name: activation [B,C,T_level]}) — activations converted from the
    engine's token-major layout to the reference's channel-major layout."""
    L = _lib.lib()
    dev = sample.device
    with torch.no_grad():
        unet(sample, timestep, ehs, encoder_attention_mask=mask)      # builds the program for this shape
        h = unet.engine(dev)
        B, _, T = sample.shape
        Tl = level_lengths(T, len(unet.cfg.block_out_channels))
        n = L.ns2vc_unet_num_taps(h)
        bufs, names = [], []
        for i in range(n):
            name, lvl, ch = C.c_char_p(), C.c_int(), C.c_int()
            _lib.check(L.ns2vc_unet_tap_info(h, i, C.byref(name), C.byref(lvl), C.byref(ch)))
            buf = torch.empty((B, Tl[lvl.value], ch.value), dtype=torch.float32, device=dev)
            _lib.check(L.ns2vc_unet_set_tap(h, i, buf.data_ptr()))
            bufs.append(buf)
            names.append(name.value.decode())
        try:
            out = unet(sample, timestep, ehs, encoder_attention_mask=mask).sample
            torch.cuda.synchronize(dev)
        finally:
            for i in range(n):
                L.ns2vc_unet_set_tap(h, i, None)
    return out, {k: v.permute(0, 2, 1).contiguous() for k, v in zip(names, bufs)}


def profile_forward(unet, steps_fn, device, dump_csv: Optional[str] = None) -> Dict[str, Tuple[float, int]]:
    """Run ``steps_fn()`` with per-launch CUDA-event timing on; returns {kind: (total ms, launches)}."""
    L = _lib.lib()
    h = unet.engine(device)
    _lib.check(L.ns2vc_unet_profile_reset(h))
    _lib.check(L.ns2vc_unet_set_profiling(h, 1))
    try:
        steps_fn()
        torch.cuda.synchronize(device)
    finally:
        _lib.check(L.ns2vc_unet_set_profiling(h, 0))
    out = {}
    for k in range(L.ns2vc_profile_num_kinds()):
        ms, n = C.c_double(), C.c_longlong()
        _lib.check(L.ns2vc_unet_profile_read(h, k, C.byref(ms), C.byref(n)))
        if n.value:
            out[L.ns2vc_profile_kind_name(k).decode()] = (ms.value, n.value)
    if dump_csv:
        _lib.check(L.ns2vc_unet_profile_dump(h, dump_csv.encode()))
    _lib.check(L.ns2vc_unet_profile_reset(h))
    return out
