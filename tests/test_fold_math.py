"""Host-side check of the LayerNorm folding identity the GEMM epilogue relies on (csrc/engine.cu, EPI_LNFOLD):

    LN(x; gamma, beta) @ W.T + bias  ==  rstd * (x @ (gamma * W).T - mean * g) + (W @ beta + bias),   g = W @ gamma

(reference attention.py:83,102,118 nn.LayerNorm eps 1e-5 feeding attention_processor.py:128-143 / attention.py:291 linears).
"""
import torch
import torch.nn.functional as F


def test_layernorm_fold_identity_matches_layernorm_then_linear():
    g = torch.Generator().manual_seed(0)
    for C, N in ((128, 384), (384, 3072)):
        x = torch.randn(64, C, generator=g, dtype=torch.float64) * 1.7 + 0.3
        W = torch.randn(N, C, generator=g, dtype=torch.float64) / C ** 0.5
        gamma = 1 + 0.2 * torch.randn(C, generator=g, dtype=torch.float64)
        beta = 0.1 * torch.randn(C, generator=g, dtype=torch.float64)
        bias = 0.1 * torch.randn(N, generator=g, dtype=torch.float64)
        want = F.linear(F.layer_norm(x, (C,), gamma, beta, 1e-5), W, bias)
        mean = x.mean(-1, keepdim=True)
        var = (x * x).mean(-1, keepdim=True) - mean * mean            # what the producer's row sums give
        rstd = torch.rsqrt(var + 1e-5)
        got = rstd * (x @ (W * gamma).T - mean * (W @ gamma)) + (W @ beta + bias)
        assert torch.allclose(got, want, rtol=1e-10, atol=1e-10)


def test_layernorm_fold_fp32_error_is_far_inside_the_parity_budget():
    g = torch.Generator().manual_seed(1)
    C, N = 256, 768
    x = (torch.randn(512, C, generator=g) * 2.0 + 0.5)
    W = torch.randn(N, C, generator=g) / C ** 0.5
    gamma = 1 + 0.2 * torch.randn(C, generator=g)
    beta = 0.1 * torch.randn(C, generator=g)
    want = F.linear(F.layer_norm(x.double(), (C,), gamma.double(), beta.double(), 1e-5), W.double())
    s, q = x.double().sum(-1, keepdim=True), (x.double() ** 2).sum(-1, keepdim=True)   # double row sums, as accumulated on the device
    mean = s / C
    rstd = torch.rsqrt((q / C - mean * mean).float() + 1e-5)
    got = rstd * (x @ (W * gamma).T - mean.float() * (W @ gamma)) + W @ beta
    err = (got.double() - want).abs().max().item()
    assert err < 2e-5, err
