"""The oracle restatements against the fixtures produced by the reference itself
(oracle/make_golden.py).  CPU only."""
import torch

from conftest import tiny_config, tiny_inputs
from ns2vc_b200.arch import ns2vc_denoiser_config, param_shapes
from ns2vc_b200.synth import make_state_dict, make_inputs, state_dict_checksum, linear_betas
from oracle import unet_oracle, sampler_oracle


def test_param_count_known_answer():
    # demo.ipynb:448 prints "diff params: 66076900"
    shapes = param_shapes(ns2vc_denoiser_config())
    assert len(shapes) == 701
    assert sum(int(torch.tensor(s).prod()) for s in shapes.values()) == 66076900


def test_tiny_forward_and_taps(gold):
    g = gold("tiny_forward.pt")
    cfg = tiny_config()
    sd = make_state_dict(cfg, 0)
    assert state_dict_checksum(sd) == g["checksum"]
    inp = tiny_inputs()
    x = torch.cat([inp["x"], inp["content"].permute(1, 2, 0)], 1)
    ehs = inp["prompt"].permute(1, 0, 2).contiguous()
    mask = unet_oracle.sequence_mask(inp["refer_lengths"], 11)
    taps = {}
    out = unet_oracle.unet_forward(sd, cfg, x, g["t"], ehs, mask, tap=lambda n, v: taps.__setitem__(n, v))
    assert torch.allclose(out, g["out"], rtol=0, atol=1e-6)
    assert set(taps) == set(g["taps"])
    for k in taps:
        assert torch.allclose(taps[k], g["taps"][k], rtol=0, atol=1e-6), k
    out2 = unet_oracle.unet_forward(sd, cfg, x, torch.tensor([999, 0]), ehs, None)
    assert torch.allclose(out2, g["out_nomask"], rtol=0, atol=1e-6)


def test_full_forward(gold):
    g = gold("full_forward.pt")
    cfg = ns2vc_denoiser_config()
    sd = make_state_dict(cfg, 0)
    assert state_dict_checksum(sd) == g["checksum"]
    inp = make_inputs(2, 131, 48, ragged=True, seed=20)
    out = unet_oracle.denoiser_forward(sd, cfg, inp["x"], inp["content"], inp["prompt"], inp["refer_lengths"], g["t"])
    assert torch.allclose(out, g["out"], rtol=0, atol=2e-6)


def test_schedule_bit_exact(gold):
    g = gold("schedule.pt")
    sch = sampler_oracle.OracleSchedule(linear_betas(1000))
    for steps in (10, 30, 50):
        e = g[steps]
        assert torch.equal(sch.log_alpha_t(e["ts"]), e["log_alpha"])
        assert torch.equal(sch.sigma(e["ts"]), e["sigma"])
        assert torch.equal(sch.lam(e["ts"]), e["lam"])


def test_sampler_oracles(gold):
    g = gold("toy_samplers.pt")
    sch = sampler_oracle.OracleSchedule(linear_betas(1000))

    def toy(x, t, **kw):
        return torch.tanh(x) * 0.7 + 0.1 * torch.sin(t / 100.0)[:, None, None]
    o = sampler_oracle.dpmpp_2m(toy, sch, g["xT"], 12)
    assert torch.equal(o, g["out"]["dpm|dpmsolver++|multistep|2|12|time_uniform|dpmsolver"])
    o = sampler_oracle.dpmpp_2m(toy, sch, g["xT"], 6, t_0=None)
    o = sampler_oracle.unipc_bh(toy, sch, g["xT"], 8, "bh2")
    assert torch.allclose(o, g["out"]["unipc|bh2|2|8|data_prediction"], rtol=0, atol=1e-6)
    o = sampler_oracle.unipc_bh(toy, sch, g["xT"], 8, "bh1")
    assert torch.allclose(o, g["out"]["unipc|bh1|2|8|data_prediction"], rtol=0, atol=1e-6)


def test_tiny_unet_samplers(gold):
    g = gold("tiny_samplers.pt")
    cfg = tiny_config()
    sd = make_state_dict(cfg, 0)
    inp = tiny_inputs()
    sch = sampler_oracle.OracleSchedule(linear_betas(1000))
    fn = lambda x, t: unet_oracle.denoiser_forward(sd, cfg, x, inp["content"], inp["prompt"], inp["refer_lengths"], t)
    assert torch.allclose(sampler_oracle.dpmpp_2m(fn, sch, inp["x"], 12), g["dpmpp2m_12"], rtol=0, atol=1e-5)
    assert torch.allclose(sampler_oracle.unipc_bh(fn, sch, inp["x"], 8), g["unipc_bh2_8"], rtol=0, atol=1e-5)


def test_p_sample(gold):
    g = gold("p_sample.pt")
    cfg = ns2vc_denoiser_config()
    sd = make_state_dict(cfg, 0)
    inp = make_inputs(1, 64, 32, seed=30)
    ddpm = sampler_oracle.OracleDDPM(1000)
    assert torch.equal(ddpm.coef1, g["buffers"]["posterior_mean_coef1"])
    assert torch.equal(ddpm.coef2, g["buffers"]["posterior_mean_coef2"])
    assert torch.equal(ddpm.log_var, g["buffers"]["posterior_log_variance_clipped"])
    assert torch.equal(ddpm.betas, g["buffers"]["betas"])
    fn = lambda x, t: unet_oracle.denoiser_forward(sd, cfg, x, inp["content"], inp["prompt"], inp["refer_lengths"], t)
    x = inp["x"]
    for i, t in enumerate((999, 998, 997)):
        noise = torch.randn(x.shape, generator=torch.Generator().manual_seed(100 + i))
        x = ddpm.p_sample(fn, x, t, noise)
        assert torch.allclose(x, g["xs"][i], rtol=0, atol=1e-5)


def test_cfg1_p_sample_ten_steps(gold):
    """BASELINE.json configs[0]: 1-layer UNet, C=100, T=128, B=1, S=64, ten DDPM steps t = 999 .. 990 (fixture produced by the
    reference's own p_sample, oracle/make_golden_cfg1.py)."""
    from oracle.make_golden_cfg1 import cfg1_config
    g = gold("cfg1_p_sample.pt")
    cfg = cfg1_config()
    sd = make_state_dict(cfg, 0)
    assert state_dict_checksum(sd) == g["checksum"]
    inp = make_inputs(1, 128, 64, seed=g["seed_inputs"])
    ddpm = sampler_oracle.OracleDDPM(1000)
    fn = lambda x, t: unet_oracle.denoiser_forward(sd, cfg, x, inp["content"], inp["prompt"], inp["refer_lengths"], t)
    x = inp["x"]
    for i, t in enumerate(range(999, 989, -1)):
        noise = torch.randn(x.shape, generator=torch.Generator().manual_seed(g["noise_seed0"] + i))
        x = ddpm.p_sample(fn, x, t, noise)
        assert torch.allclose(x, g["xs"][i], rtol=0, atol=1e-5), i


def test_ddim_sample(gold):
    g = gold("ddim.pt")
    cfg = ns2vc_denoiser_config()
    sd = make_state_dict(cfg, 0)
    inp = make_inputs(2, 72, 24, ragged=True, seed=g["seed_inputs"])
    fn = lambda x, t: unet_oracle.denoiser_forward(sd, cfg, x, inp["content"], inp["prompt"], inp["refer_lengths"], t)
    out = sampler_oracle.ddim_sample(fn, g["alphas_cumprod"], inp["x"], 1000, g["steps"])
    assert torch.allclose(out, g["out"], rtol=0, atol=1e-5)
