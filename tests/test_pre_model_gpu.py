"""Condition encoders on the GPU (`Pre_model.infer`) against fixtures written by the unmodified
reference (tests/golden/pre_model_*.pt, oracle/make_golden_pre.py) and against the pinned oracle at larger, ragged shapes.
Through the C-ABI (ns2vc_pre_*).  Tolerance: the north star's rtol 1e-3 / atol 1e-4 vs the CPU fp32 path; frames past an
utterance's length must be EXACTLY zero (model.py:142-144, 186-188)."""
import os

import pytest
import torch

from ns2vc_b200.pre_model import Pre_model
from oracle import pre_model_oracle as po

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-3, 1e-4


def close(a, b, rtol=RTOL, atol=ATOL):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    err = (a - b).abs()
    viol = (err > atol + rtol * b.abs()).float().mean().item()
    worst = (err / (atol + rtol * b.abs())).max().item()
    return viol == 0.0, f"max_abs={err.max().item():.3e} violations={viol:.3%} worst err/tol={worst:.2f} ref_rms={b.pow(2).mean().sqrt().item():.3e}"


def inputs(B, T, S, c_in, seed, dl=13, ds=7):
    g = torch.Generator().manual_seed(seed)
    c = torch.randn((B, c_in, T), generator=g)
    refer = torch.randn((B, 100, S), generator=g)
    lengths = torch.tensor([max(1, T - dl * i) for i in range(B)], dtype=torch.int64)
    refer_lengths = torch.tensor([max(1, S - ds * i) for i in range(B)], dtype=torch.int64)
    return c, refer, lengths, refer_lengths


def make(cfg, shapes=None, backend=None, seed=0):
    m = Pre_model(cfg)
    sd = po.synth_state_dict(shapes or {k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=seed)
    m.load_state_dict(sd, strict=True)
    m = m.to("cuda").eval()
    old = os.environ.get("NS2VC_GEMM_BACKEND")
    if backend:
        os.environ["NS2VC_GEMM_BACKEND"] = backend
    else:
        os.environ.pop("NS2VC_GEMM_BACKEND", None)
    try:
        m.engine(torch.device("cuda", 0))                     # the backend is fixed when the engine is created
    finally:
        if old is None:
            os.environ.pop("NS2VC_GEMM_BACKEND", None)
        else:
            os.environ["NS2VC_GEMM_BACKEND"] = old
    return m, sd


def data_of(c, refer, lengths, refer_lengths, dev="cuda"):
    return (c.to(dev), refer.to(dev), None, None, None, lengths.to(dev), refer_lengths.to(dev), None)


def check_padding(content, prompt, lengths, refer_lengths):
    for b in range(content.shape[1]):
        assert (content[int(lengths[b]):, b] == 0).all() and (prompt[int(refer_lengths[b]):, b] == 0).all()


@pytest.mark.parametrize("backend", ["simt", "tc"])
def test_tiny_fixture_every_layer(gold, backend):
    """Per-layer activations of the tiny configuration (hidden 64, head width 8 -> the v1 attention kernel; ragged lengths)."""
    g = gold("pre_model_tiny.pt")
    cfg = g["cfg"]
    m, _ = make(cfg, {k: tuple(v) for k, v in g["shapes"].items()}, backend=None if backend == "tc" else "simt")
    c, refer, lengths, refer_lengths = inputs(g["B"], g["T"], g["S"], cfg["phoneme_encoder"]["in_channels"], g["input_seed"])
    data = data_of(c, refer, lengths, refer_lengths)
    taps = m.taps(data)
    bad = []
    for k, ref in g["taps"].items():
        got = taps[k].cpu()
        ref = ref.squeeze(-1).unsqueeze(1) if k == "ref_enc" else ref.transpose(0, 1)      # reference: [B, 100, 1] / [T, B, C]
        ok, msg = close(got, ref)
        if not ok:
            bad.append(f"{k}: {msg}")
    assert not bad, "\n".join(bad)
    content, prompt = m.infer(data)
    assert content.shape == g["content"].shape and prompt.shape == g["prompt"].shape
    ok, msg = close(content, g["content"]); assert ok, "content " + msg
    ok, msg = close(prompt, g["prompt"]); assert ok, "prompt " + msg
    check_padding(content.cpu(), prompt.cpu(), lengths, refer_lengths)


def test_full_fixture(gold):
    """The shipped configuration (config.json:27-49; 34.9 M parameters) against the reference's outputs."""
    g = gold("pre_model_full.pt")
    cfg = g["cfg"]
    m, _ = make(cfg, {k: tuple(v) for k, v in g["shapes"].items()})
    c, refer, lengths, refer_lengths = inputs(g["B"], g["T"], g["S"], 256, g["input_seed"])
    content, prompt = m.infer(data_of(c, refer, lengths, refer_lengths))
    ok, msg = close(content, g["content"]); assert ok, "content " + msg
    ok, msg = close(prompt, g["prompt"]); assert ok, "prompt " + msg
    check_padding(content.cpu(), prompt.cpu(), lengths, refer_lengths)
    assert m.launch_count() > 0


FULL = {"phoneme_encoder": dict(in_channels=256, hidden_channels=256, out_channels=256, n_layers=6, p_dropout=0.2),
        "prompt_encoder": dict(in_channels=100, hidden_channels=256, out_channels=256, n_layers=6, p_dropout=0.2)}


@pytest.mark.parametrize("B,T,S", [(3, 300, 131), (1, 1100, 256), (2, 8, 1)])
def test_full_config_vs_oracle_shapes(B, T, S):
    """Ragged lengths over several row tiles (TMA-fed attention), T > 1024 (biased keys: the v1 attention kernel) and the
    smallest shapes, against the oracle (pinned to the reference by oracle/make_golden_pre.py)."""
    m, sd = make(FULL, seed=1)
    c, refer, lengths, refer_lengths = inputs(B, T, S, 256, seed=11, dl=57, ds=29)
    with torch.no_grad():
        ref_c, ref_p = po.pre_model_infer(sd, c, refer, lengths, refer_lengths, 6, 6)
    content, prompt = m.infer(data_of(c, refer, lengths, refer_lengths))
    ok, msg = close(content, ref_c); assert ok, f"content B={B} T={T} S={S}: " + msg
    ok, msg = close(prompt, ref_p); assert ok, f"prompt B={B} T={T} S={S}: " + msg
    check_padding(content.cpu(), prompt.cpu(), lengths, refer_lengths)
    # a second call on the cached program and a different shape on the same module
    content2, _ = m.infer(data_of(c, refer, lengths, refer_lengths))
    assert torch.equal(content2, content)


def test_padded_frames_of_the_input_do_not_leak():
    """Values (even NaN) in the padded frames of c / refer must not change anything: the reference zero-fills them first
    (ConvLayer.forward model.py:92-93) - except through ref_enc, which reads ALL prompt frames (model.py:362)."""
    m, _ = make(FULL, seed=2)
    c, refer, lengths, refer_lengths = inputs(2, 96, 40, 256, seed=5, dl=31, ds=0)
    base_c, base_p = m.infer(data_of(c, refer, lengths, refer_lengths))
    c2 = c.clone()
    c2[1, :, int(lengths[1]):] = float("nan")
    got_c, got_p = m.infer(data_of(c2, refer, lengths, refer_lengths))
    assert torch.equal(got_c, base_c) and torch.equal(got_p, base_p)


def test_one_prompt_broadcast_over_the_batch():
    """refer_padded with ONE row and B > 1 utterances (what `NaturalSpeech2.sample` produces from a 2-row prompt, model.py:610-611):
    the reference's modules broadcast it; the oracle does the same."""
    m, sd = make(FULL, seed=3)
    c, refer, lengths, refer_lengths = inputs(3, 70, 33, 256, seed=9, dl=21, ds=5)
    refer1 = refer[:1]
    with torch.no_grad():
        ref_c, ref_p = po.pre_model_infer(sd, c, refer1, lengths, refer_lengths, 6, 6)
    content, prompt = m.infer(data_of(c, refer1, lengths, refer_lengths))
    assert content.shape == ref_c.shape and prompt.shape == ref_p.shape
    ok, msg = close(content, ref_c); assert ok, "content " + msg
    ok, msg = close(prompt, ref_p); assert ok, "prompt " + msg
