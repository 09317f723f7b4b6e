"""BASELINE.json configs[0] exactly, and the DDIM path, through the UNMODIFIED reference (build container only):

  cfg1 : UNet1DConditionModel with layers_per_block=1, C=100, T=128, B=1, S=64, ten DDPM ``p_sample`` steps t = 999 .. 990 with
         fixed per-step noise (reference model.py:535-542 on top of Diffusion_Encoder.forward :403-415)
  ddim : ``NaturalSpeech2.ddim_sample`` (model.py:563-603; sampling_timesteps = 6, eta = 0) on the shipped 66 M denoiser, B=2, T=72

    python oracle/make_golden_cfg1.py      # writes tests/golden/cfg1_p_sample.pt, ddim.pt; asserts oracle == reference
"""
from __future__ import annotations

import json
import os
import sys
from unittest.mock import MagicMock

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("NS2VC_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)
sys.path.insert(0, REF)

from ns2vc_b200.arch import UNetConfig, ns2vc_denoiser_config  # noqa: E402
from ns2vc_b200.synth import make_inputs, make_state_dict, state_dict_checksum  # noqa: E402
from oracle import sampler_oracle, unet_oracle  # noqa: E402
from oracle.make_golden import ref_unet  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")


def cfg1_config() -> UNetConfig:
    c = ns2vc_denoiser_config()
    return UNetConfig(in_channels=c.in_channels, out_channels=c.out_channels, block_out_channels=c.block_out_channels,
                      layers_per_block=1, norm_num_groups=c.norm_num_groups, cross_attention_dim=c.cross_attention_dim,
                      num_heads=c.num_heads, addition_embed_type=c.addition_embed_type,
                      resnet_time_scale_shift=c.resnet_time_scale_shift)


@torch.no_grad()
def main():
    for name in ("matplotlib", "matplotlib.pyplot", "vocos", "accelerate", "librosa", "soundfile", "tensorboardX"):
        sys.modules.setdefault(name, MagicMock())
    import model as ref_model
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    cfg_json = json.load(open(os.path.join(REF, "config.json")))
    torch.manual_seed(0)
    ns2 = ref_model.NaturalSpeech2(cfg_json).eval()
    full = ns2vc_denoiser_config()
    sd_full = make_state_dict(full, seed=0)
    unet_full = ns2.diff_model.unet
    unet_full.load_state_dict(sd_full, strict=True)

    # ---------------------------------------------------------------- cfg1: 1-layer UNet, ten p_sample steps
    c1 = cfg1_config()
    sd1 = make_state_dict(c1, seed=0)
    u1 = ref_unet(c1)
    u1.load_state_dict(sd1, strict=True)
    ns2.diff_model.unet = u1                                  # the only change: the denoiser module the reference code drives
    inp = make_inputs(1, 128, 64, seed=40)
    data = (inp["content"], inp["prompt"], inp["lengths"], inp["refer_lengths"])
    x = inp["x"]
    ddpm = sampler_oracle.OracleDDPM(1000)
    ofn = lambda xx, tt: unet_oracle.denoiser_forward(sd1, c1, xx, inp["content"], inp["prompt"], inp["refer_lengths"], tt)
    xo, xs = x.clone(), []
    for i, t in enumerate(range(999, 989, -1)):
        noise = torch.randn(x.shape, generator=torch.Generator().manual_seed(200 + i))
        orig = torch.randn_like
        torch.randn_like = lambda _x, _n=noise: _n
        try:
            x, _ = ns2.p_sample(x, t, data)
        finally:
            torch.randn_like = orig
        xs.append(x)
        xo = ddpm.p_sample(ofn, xo, t, noise)
    d = (xo - x).abs().max().item()
    print(f"cfg1 p_sample x10 oracle vs reference: {d:.3e}")
    assert d <= 1e-5, d
    torch.save(dict(checksum=state_dict_checksum(sd1), xs=xs, seed_inputs=40, noise_seed0=200), os.path.join(GOLD, "cfg1_p_sample.pt"))

    # ---------------------------------------------------------------- ddim_sample of the reference, 6 steps, eta 0
    ns2.diff_model.unet = unet_full
    inp2 = make_inputs(2, 72, 24, ragged=True, seed=41)
    ns2.sampling_timesteps = 6
    xT = inp2["x"]
    ns2.pre_model.infer = lambda data, auto_predict_f0=True: (inp2["content"], inp2["prompt"])
    orig_randn, orig_like = torch.randn, torch.randn_like
    torch.randn = lambda *a, **k: xT.clone()
    torch.randn_like = lambda t: torch.zeros_like(t)
    try:
        out = ns2.ddim_sample(None, None, inp2["lengths"], inp2["refer_lengths"], None, None)
    finally:
        torch.randn, torch.randn_like = orig_randn, orig_like
    ofn2 = lambda xx, tt: unet_oracle.denoiser_forward(sd_full, full, xx, inp2["content"], inp2["prompt"], inp2["refer_lengths"], tt)
    want = sampler_oracle.ddim_sample(ofn2, ns2.alphas_cumprod, xT, 1000, 6)
    d = (want - out).abs().max().item()
    print(f"ddim_sample (6 steps) oracle vs reference: {d:.3e}")
    assert d <= 1e-5, d
    torch.save(dict(out=out, alphas_cumprod=ns2.alphas_cumprod.clone(), seed_inputs=41, steps=6), os.path.join(GOLD, "ddim.pt"))
    print({f: os.path.getsize(os.path.join(GOLD, f)) for f in ("cfg1_p_sample.pt", "ddim.pt")})


if __name__ == "__main__":
    main()
