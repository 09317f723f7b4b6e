"""Generate tests/golden/*.pt by running the UNMODIFIED reference (imported from /root/reference,
build container only) and pin the oracle restatements against it.

    python oracle/make_golden.py          # writes fixtures, asserts oracle == reference

The reference holds no tests or golden vectors of its own, so these fixtures —
outputs of the reference itself on deterministic synthetic weights/inputs — are what pins the
oracle.  Weights come from ns2vc_b200.synth (keyed by parameter name), so fixtures store only
outputs plus a weight checksum.
"""
from __future__ import annotations

import os
import sys
from unittest.mock import MagicMock

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("NS2VC_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)
sys.path.insert(0, REF)

from ns2vc_b200.arch import UNetConfig, ns2vc_denoiser_config, param_shapes  # noqa: E402
from ns2vc_b200.synth import make_state_dict, make_inputs, state_dict_checksum, linear_betas  # noqa: E402
from oracle import unet_oracle, sampler_oracle  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")


def tiny_config() -> UNetConfig:
    return UNetConfig(in_channels=36, out_channels=20, block_out_channels=(32, 64, 64, 96), norm_num_groups=8,
                      cross_attention_dim=16, num_heads=8, addition_embed_type="text", addition_embed_type_num_heads=4,
                      resnet_time_scale_shift="scale_shift")


def ref_unet(cfg: UNetConfig):
    from unet1d.unet_1d_condition import UNet1DConditionModel
    m = UNet1DConditionModel(
        in_channels=cfg.in_channels, out_channels=cfg.out_channels, block_out_channels=cfg.block_out_channels,
        layers_per_block=list(cfg.layers_per_block), norm_num_groups=cfg.norm_num_groups, cross_attention_dim=cfg.cross_attention_dim,
        attention_head_dim=cfg.num_heads, addition_embed_type=cfg.addition_embed_type,
        addition_embed_type_num_heads=cfg.addition_embed_type_num_heads,
        resnet_time_scale_shift=cfg.resnet_time_scale_shift).eval()
    return m


def denoiser_closure(unet, inp):
    """What NaturalSpeech2.sample_fun -> Diffusion_Encoder.forward does (model.py:520-526, 403-415)."""
    from modules import commons
    content, prompt, plen = inp["content"], inp["prompt"], inp["refer_lengths"]

    def fn(x, t, **kw):
        assert torch.isnan(x).any() == False  # noqa: E712
        p = prompt.permute(1, 0, 2)
        c = content.permute(1, 2, 0)
        xin = torch.cat([x, c], dim=1)
        mask = commons.sequence_mask(plen, p.size(1)).to(torch.bool)
        return unet(xin, t, p, encoder_attention_mask=mask).sample
    return fn


@torch.no_grad()
def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    report = {}

    # ------------------------------------------------------------------ structure
    full = ns2vc_denoiser_config()
    n_params = sum(int(torch.tensor(s).prod()) for s in param_shapes(full).values())
    assert n_params == 66076900, n_params                        # demo.ipynb:448 "diff params"
    m_full = ref_unet(full)
    assert list(m_full.state_dict().keys()) == list(param_shapes(full).keys())
    assert all(tuple(v.shape) == param_shapes(full)[k] for k, v in m_full.state_dict().items())

    # ------------------------------------------------------------------ UNet forward, tiny config (+ taps)
    tiny = tiny_config()
    sd_t = make_state_dict(tiny, seed=0)
    m_t = ref_unet(tiny)
    m_t.load_state_dict(sd_t, strict=True)
    inp = make_inputs(2, 37, 11, latent_ch=20, content_ch=16, ragged=True, seed=10)
    inp["refer_lengths"] = torch.tensor([11, 7])
    t_frac = torch.tensor([500.5, 37.25])
    xin = torch.cat([inp["x"], inp["content"].permute(1, 2, 0)], 1)
    ehs = inp["prompt"].permute(1, 0, 2).contiguous()
    mask = unet_oracle.sequence_mask(inp["refer_lengths"], 11)
    ref_out = m_t(xin, t_frac, ehs, encoder_attention_mask=mask).sample
    taps = {}
    ora_out = unet_oracle.unet_forward(sd_t, tiny, xin, t_frac, ehs, mask, tap=lambda n, v: taps.__setitem__(n, v.clone()))
    d = (ref_out - ora_out).abs().max().item()
    report["tiny_forward_oracle_vs_ref"] = d
    assert d <= 1e-6, d
    # no-mask / integer-timestep variant
    ref_nomask = m_t(xin, torch.tensor([999, 0]), ehs).sample
    ora_nomask = unet_oracle.unet_forward(sd_t, tiny, xin, torch.tensor([999, 0]), ehs, None)
    assert (ref_nomask - ora_nomask).abs().max().item() <= 1e-6
    torch.save(dict(checksum=state_dict_checksum(sd_t), out=ref_out, out_nomask=ref_nomask, t=t_frac,
                    taps={k: v for k, v in taps.items()}), os.path.join(GOLD, "tiny_forward.pt"))

    # ------------------------------------------------------------------ UNet forward, full config (odd T, ragged mask)
    sd_f = make_state_dict(full, seed=0)
    m_full.load_state_dict(sd_f, strict=True)
    inp_f = make_inputs(2, 131, 48, ragged=True, seed=20)
    xin_f = torch.cat([inp_f["x"], inp_f["content"].permute(1, 2, 0)], 1)
    ehs_f = inp_f["prompt"].permute(1, 0, 2).contiguous()
    mask_f = unet_oracle.sequence_mask(inp_f["refer_lengths"], 48)
    ref_f = m_full(xin_f, t_frac, ehs_f, encoder_attention_mask=mask_f).sample
    ora_f = unet_oracle.unet_forward(sd_f, full, xin_f, t_frac, ehs_f, mask_f)
    d = (ref_f - ora_f).abs().max().item()
    report["full_forward_oracle_vs_ref"] = d
    assert d <= 1e-6, d
    torch.save(dict(checksum=state_dict_checksum(sd_f), out=ref_f, t=t_frac), os.path.join(GOLD, "full_forward.pt"))

    # ------------------------------------------------------------------ schedule scalars (bit-exact)
    from sampler import dpm_solver as ref_dpm, uni_pc as ref_upc
    betas = linear_betas(1000)
    ns_ref = ref_dpm.NoiseScheduleVP("discrete", betas=betas)
    sch = sampler_oracle.OracleSchedule(betas)
    sched = {}
    for steps in (10, 30, 50):
        ts = torch.linspace(1.0, 1.0 / 1000, steps + 1)
        la = torch.stack([ns_ref.marginal_log_mean_coeff(t) for t in ts]).reshape(-1)
        sg = torch.stack([ns_ref.marginal_std(t) for t in ts]).reshape(-1)
        lm = torch.stack([ns_ref.marginal_lambda(t) for t in ts]).reshape(-1)
        assert torch.equal(la, sch.log_alpha_t(ts)) and torch.equal(sg, sch.sigma(ts)) and torch.equal(lm, sch.lam(ts))
        sched[steps] = dict(ts=ts, log_alpha=la, sigma=sg, lam=lm)
    lam_q = torch.linspace(-5.0, 5.0, 23)
    sched["inverse_lambda"] = dict(lam=lam_q, t=ns_ref.inverse_lambda(lam_q))
    torch.save(sched, os.path.join(GOLD, "schedule.pt"))

    # ------------------------------------------------------------------ samplers with a cheap analytic model
    def toy_model(x, t, **kw):                         # x_start-type "network": smooth, t-dependent
        return torch.tanh(x) * 0.7 + 0.1 * torch.sin(t / 100.0)[:, None, None]
    xT = torch.randn((2, 5, 33), generator=torch.Generator().manual_seed(5))
    toy = {}
    for algo in ("dpmsolver++", "dpmsolver"):
        for method, order, steps, skip in (("multistep", 2, 12, "time_uniform"), ("multistep", 3, 15, "logSNR"),
                                           ("multistep", 2, 6, "time_quadratic"), ("singlestep", 3, 11, "time_uniform"),
                                           ("singlestep", 2, 9, "logSNR"), ("multistep", 1, 5, "time_uniform")):
            for stype in ("dpmsolver", "taylor"):
                fn = ref_dpm.model_wrapper(toy_model, ns_ref, model_type="x_start")
                out = ref_dpm.DPM_Solver(fn, ns_ref, algorithm_type=algo).sample(
                    xT, steps=steps, order=order, skip_type=skip, method=method, solver_type=stype)
                toy[f"dpm|{algo}|{method}|{order}|{steps}|{skip}|{stype}"] = out
    ns_upc = ref_upc.NoiseScheduleVP("discrete", betas=betas)
    for variant in ("bh1", "bh2", "vary_coeff"):
        for order, steps in ((2, 8), (3, 9), (1, 4)):
            for algo in ("data_prediction", "noise_prediction"):
                fn = ref_upc.model_wrapper(toy_model, ns_upc, model_type="x_start")
                # the reference's uni_pc.model_wrapper broadcasts alpha_t [B] against [B,C,T] without
                # expand_dims (uni_pc.py:191): it only works for B == 1 (NS2VC's infer.py case), so the
                # reference is run per sample and the results stacked
                out = torch.cat([ref_upc.UniPC(fn, ns_upc, algorithm_type=algo, variant=variant).sample(
                    xT[i:i + 1], steps=steps, order=order, skip_type="time_uniform", method="multistep") for i in range(xT.shape[0])])
                toy[f"unipc|{variant}|{order}|{steps}|{algo}"] = out
    # oracle loops vs reference (bit-exact)
    o = sampler_oracle.dpmpp_2m(toy_model, sch, xT, 12)
    assert torch.equal(o, toy["dpm|dpmsolver++|multistep|2|12|time_uniform|dpmsolver"]), (o - toy["dpm|dpmsolver++|multistep|2|12|time_uniform|dpmsolver"]).abs().max()
    o = sampler_oracle.unipc_bh(toy_model, sch, xT, 8, "bh2")
    d = (o - toy["unipc|bh2|2|8|data_prediction"]).abs().max().item()
    report["unipc_oracle_vs_ref"] = d
    assert d <= 1e-6, d
    torch.save(dict(xT=xT, out=toy), os.path.join(GOLD, "toy_samplers.pt"))

    # ------------------------------------------------------------------ samplers over the tiny reference UNet
    fn_t = denoiser_closure(m_t, inp)
    x0 = inp["x"]
    res = {}
    mf = ref_dpm.model_wrapper(fn_t, ns_ref, model_type="x_start", model_kwargs={})
    res["dpmpp2m_12"] = ref_dpm.DPM_Solver(mf, ns_ref, algorithm_type="dpmsolver++").sample(
        x0, steps=12, order=2, skip_type="time_uniform", method="multistep")
    outs = []
    for i in range(x0.shape[0]):                       # per sample: see the B == 1 note above
        inp_i = dict(content=inp["content"][:, i:i + 1], prompt=inp["prompt"][:, i:i + 1], refer_lengths=inp["refer_lengths"][i:i + 1])
        mf = ref_upc.model_wrapper(denoiser_closure(m_t, inp_i), ns_upc, model_type="x_start", model_kwargs={})
        outs.append(ref_upc.UniPC(mf, ns_upc, variant="bh2").sample(x0[i:i + 1], steps=8, order=2, skip_type="time_uniform", method="multistep"))
    res["unipc_bh2_8"] = torch.cat(outs)
    ofn = lambda x, t: unet_oracle.denoiser_forward(sd_t, tiny, x, inp["content"], inp["prompt"], inp["refer_lengths"], t)
    d1 = (sampler_oracle.dpmpp_2m(ofn, sch, x0, 12) - res["dpmpp2m_12"]).abs().max().item()
    d2 = (sampler_oracle.unipc_bh(ofn, sch, x0, 8) - res["unipc_bh2_8"]).abs().max().item()
    report["tiny_dpm_oracle_vs_ref"], report["tiny_unipc_oracle_vs_ref"] = d1, d2
    assert d1 <= 1e-5 and d2 <= 1e-5, (d1, d2)
    torch.save(res, os.path.join(GOLD, "tiny_samplers.pt"))

    # ------------------------------------------------------------------ DDPM p_sample through the reference model.py
    for name in ("matplotlib", "matplotlib.pyplot", "vocos", "accelerate", "librosa", "soundfile", "tensorboardX"):
        sys.modules.setdefault(name, MagicMock())
    import json
    import model as ref_model
    cfg_json = json.load(open(os.path.join(REF, "config.json")))
    torch.manual_seed(0)
    ns2 = ref_model.NaturalSpeech2(cfg_json).eval()
    ns2.diff_model.unet.load_state_dict(sd_f, strict=True)
    inp_p = make_inputs(1, 64, 32, seed=30)
    data = (inp_p["content"], inp_p["prompt"], inp_p["lengths"], inp_p["refer_lengths"])
    x = inp_p["x"]
    noises = [torch.randn(x.shape, generator=torch.Generator().manual_seed(100 + i)) for i in range(3)]
    xs = []
    ddpm = sampler_oracle.OracleDDPM(1000)
    xo = x.clone()
    ofull = lambda xx, tt: unet_oracle.denoiser_forward(sd_f, full, xx, inp_p["content"], inp_p["prompt"], inp_p["refer_lengths"], tt)
    for i, t in enumerate((999, 998, 997)):
        orig = torch.randn_like
        torch.randn_like = lambda _x, _n=noises[i]: _n
        try:
            x, _ = ns2.p_sample(x, t, data)
        finally:
            torch.randn_like = orig
        xs.append(x)
        xo = ddpm.p_sample(ofull, xo, t, noises[i])
    d = (xo - x).abs().max().item()
    report["p_sample_oracle_vs_ref"] = d
    assert d <= 1e-5, d
    bufs = {k: getattr(ns2, k).clone() for k in ("betas", "posterior_mean_coef1", "posterior_mean_coef2", "posterior_log_variance_clipped")}
    assert torch.equal(bufs["posterior_mean_coef1"], ddpm.coef1) and torch.equal(bufs["posterior_log_variance_clipped"], ddpm.log_var)
    torch.save(dict(xs=xs, buffers=bufs), os.path.join(GOLD, "p_sample.pt"))

    # ------------------------------------------------------------------ bit-exact index/mask ops
    import torch.nn.functional as F
    idx = {}
    for tin, tout in ((17, 33), (33, 66), (66, 131), (128, 256), (125, 250), (128, 255), (5, 5), (3, 8), (7, 20), (250, 500), (16, 31)):
        src = torch.arange(tin, dtype=torch.float32).view(1, 1, tin)
        idx[(tin, tout)] = F.interpolate(src, size=(tout,), mode="nearest").view(-1).to(torch.int32)
    torch.save(idx, os.path.join(GOLD, "nearest_index.pt"))

    for k, v in report.items():
        print(f"{k}: {v:.3e}")
    sizes = {f: os.path.getsize(os.path.join(GOLD, f)) for f in sorted(os.listdir(GOLD))}
    print("fixtures:", sizes)


if __name__ == "__main__":
    main()
