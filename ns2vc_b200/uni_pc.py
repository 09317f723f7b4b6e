"""UniPC sampler — drop-in for the reference's ``sampler/uni_pc.py``.

Public surface follows the reference (``UniPC`` :236-672; constructor and ``sample()``
signatures as used at ``model.py:655-686``).  Only ``method='multistep'`` exists in the
reference's ``sample`` (:606-660) and here.  Arithmetic order of the B(h) update
(:471-588) is preserved so the fused CUDA path (``fused.try_fused_unipc``) and this generic
path agree with the reference to fp32 rounding.

Not carried over: the 'cosine' continuous schedule of the reference's uni_pc copy of
NoiseScheduleVP (:88-100) — NS2VC only uses 'discrete'.
"""
from __future__ import annotations

import torch

from .schedule import NoiseScheduleVP as _NoiseScheduleVP, model_wrapper, interpolate_fn, expand_dims  # noqa: F401


class NoiseScheduleVP(_NoiseScheduleVP):
    """The uni_pc copy of the schedule never clips the log-alpha tail (reference uni_pc.py:78-87)."""

    def __init__(self, schedule="discrete", betas=None, alphas_cumprod=None, continuous_beta_0=0.1,
                 continuous_beta_1=20.0, dtype=torch.float32):
        if schedule == "cosine":
            raise ValueError("the 'cosine' schedule is not supported by ns2vc_b200 (NS2VC uses 'discrete')")
        super().__init__(schedule, betas, alphas_cumprod, continuous_beta_0, continuous_beta_1, dtype, clip_alpha=False)


class UniPC:
    def __init__(self, model_fn, noise_schedule, algorithm_type="data_prediction", correcting_x0_fn=None,
                 correcting_xt_fn=None, thresholding_max_val=1.0, dynamic_thresholding_ratio=0.995, variant="bh1"):
        assert algorithm_type in ["data_prediction", "noise_prediction"]
        self._wrapped = model_fn
        self.model = lambda x, t: model_fn(x, t.expand((x.shape[0])))
        self.noise_schedule = noise_schedule
        self.correcting_x0_fn = self.dynamic_thresholding_fn if correcting_x0_fn == "dynamic_thresholding" else correcting_x0_fn
        self.correcting_xt_fn = correcting_xt_fn
        self.dynamic_thresholding_ratio = dynamic_thresholding_ratio
        self.thresholding_max_val = thresholding_max_val
        self.variant = variant
        self.predict_x0 = algorithm_type == "data_prediction"

    def dynamic_thresholding_fn(self, x0, t=None):
        p = self.dynamic_thresholding_ratio
        s = torch.quantile(torch.abs(x0).reshape((x0.shape[0], -1)), p, dim=1)
        s = expand_dims(torch.maximum(s, self.thresholding_max_val * torch.ones_like(s)), x0.dim())
        return torch.clamp(x0, -s, s) / s

    def noise_prediction_fn(self, x, t):
        return self.model(x, t)

    def data_prediction_fn(self, x, t):
        noise = self.noise_prediction_fn(x, t)
        alpha_t, sigma_t = self.noise_schedule.marginal_alpha(t), self.noise_schedule.marginal_std(t)
        x0 = (x - sigma_t * noise) / alpha_t
        if self.correcting_x0_fn is not None:
            x0 = self.correcting_x0_fn(x0)
        return x0

    def model_fn(self, x, t):
        return self.data_prediction_fn(x, t) if self.predict_x0 else self.noise_prediction_fn(x, t)

    def get_time_steps(self, skip_type, t_T, t_0, N, device):
        if skip_type == "logSNR":
            lambda_T = self.noise_schedule.marginal_lambda(torch.tensor(t_T).to(device))
            lambda_0 = self.noise_schedule.marginal_lambda(torch.tensor(t_0).to(device))
            grid = torch.linspace(lambda_T.cpu().item(), lambda_0.cpu().item(), N + 1).to(device)
            return self.noise_schedule.inverse_lambda(grid)
        if skip_type == "time_uniform":
            return torch.linspace(t_T, t_0, N + 1).to(device)
        if skip_type == "time_quadratic":
            return torch.linspace(t_T ** 0.5, t_0 ** 0.5, N + 1).pow(2).to(device)
        raise ValueError("Unsupported skip_type {}, need to be 'logSNR' or 'time_uniform' or 'time_quadratic'".format(skip_type))

    def denoise_to_zero_fn(self, x, s):
        return self.data_prediction_fn(x, s)

    def multistep_uni_pc_update(self, x, model_prev_list, t_prev_list, t, order, **kwargs):
        if len(t.shape) == 0:
            t = t.view(-1)
        if "bh" in self.variant:
            return self.multistep_uni_pc_bh_update(x, model_prev_list, t_prev_list, t, order, **kwargs)
        assert self.variant == "vary_coeff"
        return self.multistep_uni_pc_vary_update(x, model_prev_list, t_prev_list, t, order, **kwargs)

    # -------------------------------------------------------------- shared step scalars
    def _step_setup(self, x, model_prev_list, t_prev_list, t, order):
        ns = self.noise_schedule
        assert order <= len(model_prev_list)
        t0 = t_prev_list[-1]
        lam0, lam_t = ns.marginal_lambda(t0), ns.marginal_lambda(t)
        m0 = model_prev_list[-1]
        sg0, sg_t = ns.marginal_std(t0), ns.marginal_std(t)
        la0, la_t = ns.marginal_log_mean_coeff(t0), ns.marginal_log_mean_coeff(t)
        h = lam_t - lam0
        rks, D1s = [], []
        for i in range(1, order):
            lam_i = ns.marginal_lambda(t_prev_list[-(i + 1)])
            rk = (lam_i - lam0) / h
            rks.append(rk)
            D1s.append((model_prev_list[-(i + 1)] - m0) / rk)
        rks.append(1.0)
        rks = torch.tensor(rks, device=x.device)
        hh = -h if self.predict_x0 else h
        h_phi_1 = torch.expm1(hh)
        if self.predict_x0:
            w = torch.exp(la_t)                               # alpha_t
            x_base = sg_t / sg0 * x - w * h_phi_1 * m0
        else:
            w = sg_t
            x_base = torch.exp(la_t - la0) * x - (w * h_phi_1) * m0
        return m0, rks, D1s, hh, h_phi_1, w, x_base

    def multistep_uni_pc_bh_update(self, x, model_prev_list, t_prev_list, t, order, x_t=None, use_corrector=True):
        m0, rks, D1s, hh, h_phi_1, w, x_base = self._step_setup(x, model_prev_list, t_prev_list, t, order)
        if self.variant == "bh1":
            B_h = hh
        elif self.variant == "bh2":
            B_h = torch.expm1(hh)
        else:
            raise NotImplementedError()
        R, b = [], []
        h_phi_k = h_phi_1 / hh - 1
        fact = 1
        for i in range(1, order + 1):
            R.append(torch.pow(rks, i - 1))
            b.append(h_phi_k * fact / B_h)
            fact *= (i + 1)
            h_phi_k = h_phi_k / hh - 1 / fact
        R = torch.stack(R)
        b = torch.cat(b)
        use_predictor = len(D1s) > 0 and x_t is None
        if len(D1s) > 0:
            D1s = torch.stack(D1s, dim=1)                     # [B,K,C,T]
            if x_t is None:
                rhos_p = torch.tensor([0.5], device=b.device) if order == 2 else torch.linalg.solve(R[:-1, :-1], b[:-1])
        else:
            D1s = None
        if use_corrector:
            rhos_c = torch.tensor([0.5], device=b.device) if order == 1 else torch.linalg.solve(R, b)
        model_t = None
        if x_t is None:
            pred_res = torch.einsum("k,bkct->bct", rhos_p, D1s) if use_predictor else 0
            x_t = x_base - w * B_h * pred_res
        if use_corrector:
            model_t = self.model_fn(x_t, t)
            corr_res = torch.einsum("k,bkct->bct", rhos_c[:-1], D1s) if D1s is not None else 0
            D1_t = model_t - m0
            x_t = x_base - w * B_h * (corr_res + rhos_c[-1] * D1_t)
        return x_t, model_t

    def multistep_uni_pc_vary_update(self, x, model_prev_list, t_prev_list, t, order, use_corrector=True):
        m0, rks, D1s, hh, h_phi_1, w, x_base = self._step_setup(x, model_prev_list, t_prev_list, t, order)
        K = len(rks)
        cols, col = [], torch.ones_like(rks)
        for k in range(1, K + 1):
            cols.append(col)
            col = col * rks / (k + 1)
        C = torch.stack(cols, dim=1)
        if len(D1s) > 0:
            D1s = torch.stack(D1s, dim=1)
            A_p = torch.linalg.inv(C[:-1, :-1])
        if use_corrector:
            A_c = torch.linalg.inv(C)
        h_phi_ks, fact, h_phi_k = [], 1, h_phi_1
        for k in range(1, K + 2):
            h_phi_ks.append(h_phi_k)
            h_phi_k = h_phi_k / hh - 1 / fact
            fact *= (k + 1)
        model_t = None
        x_t = x_base
        if len(D1s) > 0:
            for k in range(K - 1):
                x_t = x_t - w * h_phi_ks[k + 1] * torch.einsum("bkct,k->bct", D1s, A_p[k])
        if use_corrector:
            model_t = self.model_fn(x_t, t)
            D1_t = model_t - m0
            x_t = x_base
            k = 0
            for k in range(K - 1):
                x_t = x_t - w * h_phi_ks[k + 1] * torch.einsum("bkct,k->bct", D1s, A_c[k][:-1])
            x_t = x_t - w * h_phi_ks[K] * (D1_t * A_c[k][-1])
        return x_t, model_t

    # -------------------------------------------------------------- driver
    def sample(self, x, steps=20, t_start=None, t_end=None, order=2, skip_type="time_uniform", method="multistep",
               lower_order_final=True, denoise_to_zero=False, atol=0.0078, rtol=0.05, return_intermediate=False):
        t_0 = 1.0 / self.noise_schedule.total_N if t_end is None else t_end
        t_T = self.noise_schedule.T if t_start is None else t_start
        assert t_0 > 0 and t_T > 0, "Time range needs to be greater than 0. For discrete-time DPMs, it needs to be in [1 / N, 1], where N is the length of betas array"
        if method != "multistep":
            raise ValueError("Got wrong method {}".format(method))
        device = x.device

        if (order == 2 and self.variant == "bh2" and self.predict_x0 and lower_order_final and not return_intermediate
                and not denoise_to_zero and self.correcting_x0_fn is None and self.correcting_xt_fn is None
                and steps >= 3):
            from . import fused
            out = fused.try_fused_unipc(self, x, steps, skip_type, t_T, t_0)
            if out is not None:
                return out

        track = []

        def after(x, t, step):
            if self.correcting_xt_fn is not None:
                x = self.correcting_xt_fn(x, t, step)
            if return_intermediate:
                track.append(x)
            return x

        with torch.no_grad():
            assert steps >= order
            ts = self.get_time_steps(skip_type=skip_type, t_T=t_T, t_0=t_0, N=steps, device=device)
            assert ts.shape[0] - 1 == steps
            t_hist = [ts[0]]
            m_hist = [self.model_fn(x, ts[0])]
            x = after(x, ts[0], 0)
            for step in range(1, steps + 1):
                t = ts[step]
                if step < order:
                    k, corr = step, True
                else:
                    k = min(order, steps + 1 - step) if lower_order_final else order
                    corr = step != steps                      # no corrector (and no NFE) at the last step
                x, m_t = self.multistep_uni_pc_update(x, m_hist, t_hist, t, k, use_corrector=corr)
                if step < order:
                    if m_t is None:
                        m_t = self.model_fn(x, t)
                    x = after(x, t, step)
                    t_hist.append(t)
                    m_hist.append(m_t)
                else:
                    x = after(x, t, step)
                    t_hist = t_hist[1:] + [t]
                    if step < steps:
                        if m_t is None:
                            m_t = self.model_fn(x, t)
                        m_hist = m_hist[1:] + [m_t]
            if denoise_to_zero:
                t = torch.ones((1,)).to(device) * t_0
                x = self.denoise_to_zero_fn(x, t)
                x = after(x, t, steps + 1)
        return (x, track) if return_intermediate else x
