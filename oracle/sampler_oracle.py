"""ORACLE (test infrastructure only — never imported by the product path).

Plain-loop CPU restatements of the sampling algorithms the reference iterates the denoiser with.
Each function follows the cited reference lines in the same fp32 op order; all are pinned against
the reference classes themselves by ``oracle/make_golden.py`` (fixtures in ``tests/golden``).

  noise schedule        sampler/dpm_solver.py:6-167, interpolate_fn :1253-1292
  x_start wrapper       sampler/dpm_solver.py:271-298 (model_wrapper.noise_pred_fn)
  DPM-Solver++(2M)      sampler/dpm_solver.py:433-442, 547-580, 796-852, 1171-1213
  UniPC-bh              sampler/uni_pc.py:471-588, 606-658
  DDPM p_sample         model.py:504-542 ; schedule buffers model.py:456-498
"""
from __future__ import annotations

from typing import Callable, List

import torch


class OracleSchedule:
    """Discrete VP schedule: piecewise-linear log-alpha over t_i = (i+1)/N."""

    def __init__(self, betas: torch.Tensor):
        self.log_alpha = (0.5 * torch.log(1 - betas).cumsum(dim=0)).to(torch.float32)   # :100
        self.N = self.log_alpha.shape[0]
        self.t_knots = torch.linspace(0.0, 1.0, self.N + 1)[1:].to(torch.float32)       # :107

    def log_alpha_t(self, t: torch.Tensor) -> torch.Tensor:
        """t: 1-D tensor.  Segment choice as in interpolate_fn's sort (:1266-1290): the query sorts
        before equal knots; out-of-range queries extrapolate the outermost segment."""
        out = torch.empty_like(t)
        xp, yp, K = self.t_knots, self.log_alpha, self.N
        for i in range(t.shape[0]):
            x = t[i]
            n_less = int((xp < x).sum())
            seg = min(max(n_less - 1, 0), K - 2)
            out[i] = yp[seg] + (x - xp[seg]) * (yp[seg + 1] - yp[seg]) / (xp[seg + 1] - xp[seg])
        return out

    def alpha(self, t):
        return torch.exp(self.log_alpha_t(t))

    def sigma(self, t):
        return torch.sqrt(1.0 - torch.exp(2.0 * self.log_alpha_t(t)))

    def lam(self, t):
        la = self.log_alpha_t(t)
        return la - 0.5 * torch.log(1.0 - torch.exp(2.0 * la))


def x0_model(x_start_fn: Callable, sch: OracleSchedule, x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """DPM_Solver.model_fn for an x_start network: model time (:278), x0 -> noise (:291-292),
    noise -> x0 (:437-439).  t: 0-d or 1-elem tensor."""
    B = x.shape[0]
    tc = t.reshape(-1)[:1].expand(B)
    t_in = (tc - 1.0 / sch.N) * sch.N
    out = x_start_fn(x, t_in)
    a, s = sch.alpha(tc)[:, None, None], sch.sigma(tc)[:, None, None]
    noise = (x - a * out) / s
    a1, s1 = sch.alpha(t.reshape(-1)[:1]), sch.sigma(t.reshape(-1)[:1])
    return (x - s1 * noise) / a1


def dpmpp_2m(x_start_fn: Callable, sch: OracleSchedule, x: torch.Tensor, steps: int, t_T: float = 1.0,
             t_0: float = None, lower_order_final: bool = True) -> torch.Tensor:
    t_0 = 1.0 / sch.N if t_0 is None else t_0
    ts = torch.linspace(t_T, t_0, steps + 1)
    m_hist: List[torch.Tensor] = [x0_model(x_start_fn, sch, x, ts[0])]
    t_hist = [ts[0]]
    for step in range(1, steps + 1):
        s, t = t_hist[-1].reshape(1), ts[step].reshape(1)
        if step < 2:
            order = 1
        elif lower_order_final and steps < 10:
            order = min(2, steps + 1 - step)
        else:
            order = 2
        h = sch.lam(t) - sch.lam(s)
        alpha_t = torch.exp(sch.log_alpha_t(t))
        phi_1 = torch.expm1(-h)
        if order == 1:
            x = sch.sigma(t) / sch.sigma(s) * x - alpha_t * phi_1 * m_hist[-1]
        else:
            h_0 = sch.lam(s) - sch.lam(t_hist[-2].reshape(1))
            r0 = h_0 / h
            D1_0 = (1.0 / r0) * (m_hist[-1] - m_hist[-2])
            x = (sch.sigma(t) / sch.sigma(s)) * x - (alpha_t * phi_1) * m_hist[-1] - 0.5 * (alpha_t * phi_1) * D1_0
        if step < steps:
            m_hist = (m_hist + [x0_model(x_start_fn, sch, x, ts[step])])[-2:]
            t_hist = (t_hist + [ts[step]])[-2:]
    return x


def unipc_bh(x_start_fn: Callable, sch: OracleSchedule, x: torch.Tensor, steps: int, variant: str = "bh2",
             t_T: float = 1.0, t_0: float = None) -> torch.Tensor:
    """UniPC multistep, order 2, data prediction, lower_order_final=True."""
    t_0 = 1.0 / sch.N if t_0 is None else t_0
    ts = torch.linspace(t_T, t_0, steps + 1)
    m_hist = [x0_model(x_start_fn, sch, x, ts[0])]
    t_hist = [ts[0]]
    for step in range(1, steps + 1):
        t = ts[step].reshape(1)
        order = step if step < 2 else min(2, steps + 1 - step)
        use_corr = step != steps or step < 2
        t0 = t_hist[-1].reshape(1)
        m0 = m_hist[-1]
        h = sch.lam(t) - sch.lam(t0)
        alpha_t = torch.exp(sch.log_alpha_t(t))
        rks, D1s = [], []
        for i in range(1, order):
            rk = (sch.lam(t_hist[-(i + 1)].reshape(1)) - sch.lam(t0)) / h
            rks.append(rk)
            D1s.append((m_hist[-(i + 1)] - m0) / rk)
        rks.append(1.0)
        rks = torch.tensor(rks)
        hh = -h
        h_phi_1 = torch.expm1(hh)
        B_h = hh if variant == "bh1" else torch.expm1(hh)
        R, b = [], []
        h_phi_k = h_phi_1 / hh - 1
        fact = 1
        for i in range(1, order + 1):
            R.append(torch.pow(rks, i - 1))
            b.append(h_phi_k * fact / B_h)
            fact *= (i + 1)
            h_phi_k = h_phi_k / hh - 1 / fact
        R, b = torch.stack(R), torch.cat(b)
        x_bar = sch.sigma(t) / sch.sigma(t0) * x - alpha_t * h_phi_1 * m0
        if D1s:
            rho_p = torch.tensor([0.5]) if order == 2 else torch.linalg.solve(R[:-1, :-1], b[:-1])
            pred = sum(rho_p[k] * D1s[k] for k in range(len(D1s)))
            x_t = x_bar - alpha_t * B_h * pred
        else:
            x_t = x_bar
        m_t = None
        if use_corr:
            rho_c = torch.tensor([0.5]) if order == 1 else torch.linalg.solve(R, b)
            m_t = x0_model(x_start_fn, sch, x_t, t)
            corr = sum(rho_c[k] * D1s[k] for k in range(len(D1s))) if D1s else 0
            x_t = x_bar - alpha_t * B_h * (corr + rho_c[-1] * (m_t - m0))
        x = x_t
        if step < steps:
            if m_t is None:
                m_t = x0_model(x_start_fn, sch, x, t)
            m_hist = (m_hist + [m_t])[-2:]
            t_hist = (t_hist + [ts[step]])[-2:]
    return x


class OracleDDPM:
    """Posterior buffers of NaturalSpeech2.__init__ (model.py:456-498) and p_sample (:535-542)."""

    def __init__(self, timesteps: int = 1000):
        scale = 1000 / timesteps
        betas = torch.linspace(scale * 0.0001, scale * 0.02, timesteps, dtype=torch.float64)
        alphas = 1.0 - betas
        ac = torch.cumprod(alphas, dim=0)
        ac_prev = torch.nn.functional.pad(ac[:-1], (1, 0), value=1.0)
        f32 = lambda v: v.to(torch.float32)
        self.betas = f32(betas)
        post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
        self.log_var = f32(torch.log(post_var.clamp(min=1e-20)))
        self.coef1 = f32(betas * torch.sqrt(ac_prev) / (1.0 - ac))
        self.coef2 = f32((1.0 - ac_prev) * torch.sqrt(alphas) / (1.0 - ac))

    def p_sample(self, x_start_fn: Callable, x: torch.Tensor, t: int, noise: torch.Tensor) -> torch.Tensor:
        B = x.shape[0]
        bt = torch.full((B,), t, dtype=torch.long)
        x0 = x_start_fn(x, bt)
        mean = self.coef1[bt][:, None, None] * x0 + self.coef2[bt][:, None, None] * x
        nz = noise if t > 0 else 0.0
        return mean + (0.5 * self.log_var[bt][:, None, None]).exp() * nz


def ddim_sample(x_start_fn: Callable, alphas_cumprod: torch.Tensor, x: torch.Tensor, total_timesteps: int, sampling_timesteps: int,
                eta: float = 0.0) -> torch.Tensor:
    """NaturalSpeech2.ddim_sample after pre_model.infer (reference model.py:570-603), eta = 0 (the reference's default,
    :446): integer timesteps linspace(-1, T-1, S+1) reversed; x0 = model(x, t); noise = predict_noise_from_start (:498-503);
    x <- sqrt(a_next) x0 + sqrt(1 - a_next) noise; the last pair (t, -1) returns x0."""
    assert eta == 0.0
    times = torch.linspace(-1, total_timesteps - 1, steps=sampling_timesteps + 1)
    times = list(reversed(times.int().tolist()))
    sqrt_recip = torch.sqrt(1.0 / alphas_cumprod).to(torch.float32)
    sqrt_recipm1 = torch.sqrt(1.0 / alphas_cumprod - 1).to(torch.float32)
    ac32 = alphas_cumprod.to(torch.float32)
    B = x.shape[0]
    for time, time_next in zip(times[:-1], times[1:]):
        bt = torch.full((B,), time, dtype=torch.long)
        x0 = x_start_fn(x, bt)
        pred_noise = (sqrt_recip[bt][:, None, None] * x - x0) / sqrt_recipm1[bt][:, None, None]
        if time_next < 0:
            x = x0
            continue
        alpha, alpha_next = ac32[time], ac32[time_next]
        sigma = eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
        c = (1 - alpha_next - sigma ** 2).sqrt()
        x = x0 * alpha_next.sqrt() + c * pred_noise
    return x
