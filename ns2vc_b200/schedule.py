"""VP noise schedule and the denoiser-call wrapper shared by both samplers.

API mirror of the reference's ``sampler/dpm_solver.py`` (``NoiseScheduleVP`` :6-167,
``model_wrapper`` :170-334, ``interpolate_fn`` :1253-1292, ``expand_dims`` :1295) and of the
near-identical copies in ``sampler/uni_pc.py`` (:6-234; no ``numerical_clip_alpha`` there).
The schedule methods are a handful of one-line formulas whose op ORDER is part of the contract: every scalar
must be bit-equal to the reference's fp32 value (the fused samplers precompute all per-step coefficients from
them), so they follow the reference statement for statement; `interpolate_fn` (searchsorted instead of the
reference's sort/gather) and `WrappedModel` are restructured.  Checked against fixtures generated from the
reference in tests/test_samplers_cpu.py::test_schedule_bit_exact and tests/test_oracle_golden.py.
"""
from __future__ import annotations

import torch


def expand_dims(v: torch.Tensor, dims: int) -> torch.Tensor:
    """[N] -> [N,1,...,1] with ``dims`` dimensions (reference :1295-1304)."""
    return v[(...,) + (None,) * (dims - 1)]


def interpolate_fn(x: torch.Tensor, xp: torch.Tensor, yp: torch.Tensor) -> torch.Tensor:
    """Piecewise-linear y=f(x) through keypoints (xp ascending), linear extrapolation outside.

    x [N,C], xp/yp [C,K] -> [N,C].  The reference finds the segment by sorting x together with
    the keypoints (:1266-1290); the equivalent closed form is: with i = #{keypoints < x} (x sorts
    first among equal values), the segment is [max(i,1)-1 clipped to K-2, +1], and the value is
    ``y0 + (x - x0) * (y1 - y0) / (x1 - x0)`` evaluated in exactly that order.
    """
    N, C = x.shape
    K = xp.shape[1]
    xq = x.transpose(0, 1).contiguous()                       # [C,N]
    idx = torch.searchsorted(xp.contiguous(), xq, right=False)  # count of knots < x
    seg = (idx - 1).clamp(min=0, max=K - 2)
    x0 = torch.gather(xp, 1, seg)
    x1 = torch.gather(xp, 1, seg + 1)
    y0 = torch.gather(yp, 1, seg)
    y1 = torch.gather(yp, 1, seg + 1)
    out = y0 + (xq - x0) * (y1 - y0) / (x1 - x0)
    return out.transpose(0, 1)


class NoiseScheduleVP:
    """Forward-SDE wrapper (VP type): alpha_t, sigma_t, lambda_t = log(alpha_t/sigma_t) and its
    inverse, for discrete-time models (piecewise-linear log-alpha over t_i=(i+1)/N) or the
    continuous linear-beta VPSDE.  Same constructor and methods as the reference class."""

    def __init__(self, schedule="discrete", betas=None, alphas_cumprod=None,
                 continuous_beta_0=0.1, continuous_beta_1=20.0, dtype=torch.float32, clip_alpha=True):
        if schedule not in ("discrete", "linear"):
            raise ValueError("Unsupported noise schedule {}. The schedule needs to be 'discrete' or 'linear'".format(schedule))
        self.schedule = schedule
        self.T = 1.0
        if schedule == "discrete":
            if betas is not None:
                log_alphas = 0.5 * torch.log(1 - betas).cumsum(dim=0)
            else:
                assert alphas_cumprod is not None
                log_alphas = 0.5 * torch.log(alphas_cumprod)
            if clip_alpha:
                log_alphas = self.numerical_clip_alpha(log_alphas)
            self.log_alpha_array = log_alphas.reshape((1, -1)).to(dtype=dtype)
            self.total_N = self.log_alpha_array.shape[1]
            self.t_array = torch.linspace(0.0, 1.0, self.total_N + 1)[1:].reshape((1, -1)).to(dtype=dtype)
        else:
            self.total_N = 1000
            self.beta_0 = continuous_beta_0
            self.beta_1 = continuous_beta_1

    def numerical_clip_alpha(self, log_alphas, clipped_lambda=-5.1):
        """Drop the tail where the half-logSNR falls below ``clipped_lambda`` (reference :109-120;
        a no-op for the linear-beta schedule NS2VC trains with)."""
        log_sigmas = 0.5 * torch.log(1.0 - torch.exp(2.0 * log_alphas))
        lambs = log_alphas - log_sigmas
        idx = torch.searchsorted(torch.flip(lambs, [0]), clipped_lambda)
        if idx > 0:
            log_alphas = log_alphas[:-idx]
        return log_alphas

    def marginal_log_mean_coeff(self, t):
        if self.schedule == "discrete":
            return interpolate_fn(t.reshape((-1, 1)), self.t_array.to(t.device),
                                  self.log_alpha_array.to(t.device)).reshape((-1))
        return -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0

    def marginal_alpha(self, t):
        return torch.exp(self.marginal_log_mean_coeff(t))

    def marginal_std(self, t):
        return torch.sqrt(1.0 - torch.exp(2.0 * self.marginal_log_mean_coeff(t)))

    def marginal_lambda(self, t):
        log_mean_coeff = self.marginal_log_mean_coeff(t)
        log_std = 0.5 * torch.log(1.0 - torch.exp(2.0 * log_mean_coeff))
        return log_mean_coeff - log_std

    def inverse_lambda(self, lamb):
        if self.schedule == "linear":
            tmp = 2.0 * (self.beta_1 - self.beta_0) * torch.logaddexp(-2.0 * lamb, torch.zeros((1,)).to(lamb))
            Delta = self.beta_0 ** 2 + tmp
            return tmp / (torch.sqrt(Delta) + self.beta_0) / (self.beta_1 - self.beta_0)
        log_alpha = -0.5 * torch.logaddexp(torch.zeros((1,)).to(lamb.device), -2.0 * lamb)
        t = interpolate_fn(log_alpha.reshape((-1, 1)), torch.flip(self.log_alpha_array.to(lamb.device), [1]),
                           torch.flip(self.t_array.to(lamb.device), [1]))
        return t.reshape((-1,))


class WrappedModel:
    """Callable returned by :func:`model_wrapper`.  Behaves like the reference's closure
    ``model_fn(x, t_continuous) -> noise`` but keeps its ingredients inspectable, which is what
    lets the samplers take the fused CUDA fast path when the wrapped model is our denoiser."""

    def __init__(self, model, noise_schedule, model_type, model_kwargs, guidance_type, condition,
                 unconditional_condition, guidance_scale, classifier_fn, classifier_kwargs):
        assert model_type in ["noise", "x_start", "v", "score"]
        assert guidance_type in ["uncond", "classifier", "classifier-free"]
        self.model = model
        self.noise_schedule = noise_schedule
        self.model_type = model_type
        self.model_kwargs = model_kwargs
        self.guidance_type = guidance_type
        self.condition = condition
        self.unconditional_condition = unconditional_condition
        self.guidance_scale = guidance_scale
        self.classifier_fn = classifier_fn
        self.classifier_kwargs = classifier_kwargs

    def get_model_input_time(self, t_continuous):
        ns = self.noise_schedule
        if ns.schedule == "discrete":
            # [1/N, 1] -> [0, 1000*(N-1)/N]; fractional (reference :271-281)
            return (t_continuous - 1.0 / ns.total_N) * ns.total_N
        return t_continuous

    def noise_pred_fn(self, x, t_continuous, cond=None):
        ns = self.noise_schedule
        t_input = self.get_model_input_time(t_continuous)
        if cond is None:
            output = self.model(x, t_input, **self.model_kwargs)
        else:
            output = self.model(x, t_input, cond, **self.model_kwargs)
        if self.model_type == "noise":
            return output
        if self.model_type == "x_start":
            alpha_t, sigma_t = ns.marginal_alpha(t_continuous), ns.marginal_std(t_continuous)
            return (x - expand_dims(alpha_t, x.dim()) * output) / expand_dims(sigma_t, x.dim())
        if self.model_type == "v":
            alpha_t, sigma_t = ns.marginal_alpha(t_continuous), ns.marginal_std(t_continuous)
            return expand_dims(alpha_t, x.dim()) * output + expand_dims(sigma_t, x.dim()) * x
        sigma_t = ns.marginal_std(t_continuous)            # "score"
        return -expand_dims(sigma_t, x.dim()) * output

    def cond_grad_fn(self, x, t_input):
        with torch.enable_grad():
            x_in = x.detach().requires_grad_(True)
            log_prob = self.classifier_fn(x_in, t_input, self.condition, **self.classifier_kwargs)
            return torch.autograd.grad(log_prob.sum(), x_in)[0]

    def __call__(self, x, t_continuous):
        if self.guidance_type == "uncond":
            return self.noise_pred_fn(x, t_continuous)
        if self.guidance_type == "classifier":
            assert self.classifier_fn is not None
            t_input = self.get_model_input_time(t_continuous)
            cond_grad = self.cond_grad_fn(x, t_input)
            sigma_t = self.noise_schedule.marginal_std(t_continuous)
            noise = self.noise_pred_fn(x, t_continuous)
            return noise - self.guidance_scale * expand_dims(sigma_t, x.dim()) * cond_grad
        # classifier-free
        if self.guidance_scale == 1.0 or self.unconditional_condition is None:
            return self.noise_pred_fn(x, t_continuous, cond=self.condition)
        x_in = torch.cat([x] * 2)
        t_in = torch.cat([t_continuous] * 2)
        c_in = torch.cat([self.unconditional_condition, self.condition])
        noise_uncond, noise = self.noise_pred_fn(x_in, t_in, cond=c_in).chunk(2)
        return noise_uncond + self.guidance_scale * (noise - noise_uncond)


def model_wrapper(model, noise_schedule, model_type="noise", model_kwargs={}, guidance_type="uncond",
                  condition=None, unconditional_condition=None, guidance_scale=1.0,
                  classifier_fn=None, classifier_kwargs={}):
    """Same signature and semantics as the reference's ``model_wrapper`` (:170-334): turns a
    discrete- or continuous-time model of type noise / x_start / v / score, optionally guided,
    into a continuous-time noise-prediction function."""
    return WrappedModel(model, noise_schedule, model_type, model_kwargs, guidance_type, condition,
                        unconditional_condition, guidance_scale, classifier_fn, classifier_kwargs)
