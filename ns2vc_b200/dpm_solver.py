"""DPM-Solver / DPM-Solver++ sampler — drop-in for the reference's ``sampler/dpm_solver.py``.

Public surface (constructor / method names, argument meaning, defaults) follows the reference
(``DPM_Solver`` :337-1245) so ``model.py:621-652`` works unchanged when ``sampler.dpm_solver``
resolves to this module.  The implementation is organised differently: all exponential-
integrator updates go through one coefficient helper (:class:`_Integrator`) that covers both
parameterisations (data prediction "dpmsolver++" and noise prediction "dpmsolver"), evaluated in
the reference's arithmetic order so results agree to fp32 rounding (bit-equal on the
multistep order-2 path, which the fused CUDA sampler in ``fused.py`` reproduces).

Fast path: ``sample(..., method='multistep')`` on a CUDA tensor whose model is the B200
denoiser (see ``fused.try_fused_dpm``) runs the whole loop as UNet-forward + one fused
sampler-step kernel per step with host-precomputed coefficients.
"""
from __future__ import annotations

import torch

from .schedule import NoiseScheduleVP, model_wrapper, interpolate_fn, expand_dims  # noqa: F401 (API re-export)


class _Integrator:
    """Scalars for one exponential-integrator step s -> t in half-logSNR.

    data prediction ("++"):  x_t = (sigma_t/sigma_s) x - alpha_t * expm1(-h) * m + ...
    noise prediction:        x_t = exp(log a_t - log a_s) x - sigma_t * expm1(h) * m - ...
    """

    def __init__(self, ns: NoiseScheduleVP, plus: bool, s, t):
        self.plus = plus
        self.lam_s, self.lam_t = ns.marginal_lambda(s), ns.marginal_lambda(t)
        self.h = self.lam_t - self.lam_s
        la_s, la_t = ns.marginal_log_mean_coeff(s), ns.marginal_log_mean_coeff(t)
        sg_s, sg_t = ns.marginal_std(s), ns.marginal_std(t)
        if plus:
            self.lin = sg_t / sg_s
            self.w = torch.exp(la_t)
        else:
            self.lin = torch.exp(la_t - la_s)
            self.w = sg_t
        self.ns = ns
        self.s = s
        self._la_s, self._sg_s = la_s, sg_s

    def e(self, r=None):
        """expm1(-+ r h)."""
        rh = self.h if r is None else r * self.h
        return torch.expm1(-rh) if self.plus else torch.expm1(rh)

    def phi2(self, phi1, h=None):
        h = self.h if h is None else h
        return phi1 / h + 1.0 if self.plus else phi1 / h - 1.0

    def at(self, r):
        """(time, linear coefficient, model weight) of the intermediate point lambda_s + r h."""
        u = self.ns.inverse_lambda(self.lam_s + r * self.h)
        la_u = self.ns.marginal_log_mean_coeff(u)
        sg_u = self.ns.marginal_std(u)
        if self.plus:
            return u, sg_u / self._sg_s, torch.exp(la_u)
        return u, torch.exp(la_u - self._la_s), sg_u


class DPM_Solver:
    def __init__(self, model_fn, noise_schedule, algorithm_type="dpmsolver++", correcting_x0_fn=None,
                 correcting_xt_fn=None, thresholding_max_val=1.0, dynamic_thresholding_ratio=0.995):
        assert algorithm_type in ["dpmsolver", "dpmsolver++"]
        self._wrapped = model_fn
        self.model = lambda x, t: model_fn(x, t.expand((x.shape[0])))
        self.noise_schedule = noise_schedule
        self.algorithm_type = algorithm_type
        self.correcting_x0_fn = self.dynamic_thresholding_fn if correcting_x0_fn == "dynamic_thresholding" else correcting_x0_fn
        self.correcting_xt_fn = correcting_xt_fn
        self.dynamic_thresholding_ratio = dynamic_thresholding_ratio
        self.thresholding_max_val = thresholding_max_val

    # ------------------------------------------------------------------ model plumbing
    @property
    def _plus(self):
        return self.algorithm_type == "dpmsolver++"

    def dynamic_thresholding_fn(self, x0, t):
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
            x0 = self.correcting_x0_fn(x0, t)
        return x0

    def model_fn(self, x, t):
        return self.data_prediction_fn(x, t) if self._plus else self.noise_prediction_fn(x, t)

    # ------------------------------------------------------------------ time grids
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

    def get_orders_and_timesteps_for_singlestep_solver(self, steps, order, skip_type, t_T, t_0, device):
        """Split ``steps`` NFE into single-step solvers of order <= ``order`` (reference :482-545)."""
        if order == 3:
            K = steps // 3 + 1
            tail = {0: [2, 1], 1: [1], 2: [2]}[steps % 3]
            orders = [3] * (K - len(tail)) + tail
        elif order == 2:
            K = (steps + 1) // 2
            orders = [2] * (steps // 2) + ([1] if steps % 2 else [])
        elif order == 1:
            K = 1
            orders = [1] * steps
        else:
            raise ValueError("'order' must be '1' or '2' or '3'.")
        if skip_type == "logSNR":
            outer = self.get_time_steps(skip_type, t_T, t_0, K, device)
        else:
            sel = torch.cumsum(torch.tensor([0] + orders), 0).to(device)
            outer = self.get_time_steps(skip_type, t_T, t_0, steps, device)[sel]
        return outer, orders

    def denoise_to_zero_fn(self, x, s):
        return self.data_prediction_fn(x, s)

    # ------------------------------------------------------------------ single-step updates
    def dpm_solver_first_update(self, x, s, t, model_s=None, return_intermediate=False):
        g = _Integrator(self.noise_schedule, self._plus, s, t)
        phi_1 = g.e()
        if model_s is None:
            model_s = self.model_fn(x, s)
        if self._plus:
            x_t = g.lin * x - g.w * phi_1 * model_s
        else:
            x_t = g.lin * x - (g.w * phi_1) * model_s
        return (x_t, {"model_s": model_s}) if return_intermediate else x_t

    def singlestep_dpm_solver_second_update(self, x, s, t, r1=0.5, model_s=None, return_intermediate=False,
                                            solver_type="dpmsolver"):
        if solver_type not in ["dpmsolver", "taylor"]:
            raise ValueError("'solver_type' must be either 'dpmsolver' or 'taylor', got {}".format(solver_type))
        r1 = 0.5 if r1 is None else r1
        g = _Integrator(self.noise_schedule, self._plus, s, t)
        s1, lin1, w1 = g.at(r1)
        phi_11, phi_1 = g.e(r1), g.e()
        if model_s is None:
            model_s = self.model_fn(x, s)
        x_s1 = lin1 * x - (w1 * phi_11) * model_s
        model_s1 = self.model_fn(x_s1, s1)
        base = g.lin * x - (g.w * phi_1) * model_s
        d = model_s1 - model_s
        if solver_type == "dpmsolver":
            x_t = base - (0.5 / r1) * (g.w * phi_1) * d
        elif self._plus:
            x_t = base + (1.0 / r1) * (g.w * g.phi2(phi_1)) * d
        else:
            x_t = base - (1.0 / r1) * (g.w * g.phi2(phi_1)) * d
        if return_intermediate:
            return x_t, {"model_s": model_s, "model_s1": model_s1}
        return x_t

    def singlestep_dpm_solver_third_update(self, x, s, t, r1=1.0 / 3.0, r2=2.0 / 3.0, model_s=None, model_s1=None,
                                           return_intermediate=False, solver_type="dpmsolver"):
        if solver_type not in ["dpmsolver", "taylor"]:
            raise ValueError("'solver_type' must be either 'dpmsolver' or 'taylor', got {}".format(solver_type))
        r1 = 1.0 / 3.0 if r1 is None else r1
        r2 = 2.0 / 3.0 if r2 is None else r2
        plus = self._plus
        g = _Integrator(self.noise_schedule, plus, s, t)
        s1, lin1, w1 = g.at(r1)
        s2, lin2, w2 = g.at(r2)
        phi_11, phi_12, phi_1 = g.e(r1), g.e(r2), g.e()
        phi_22 = g.phi2(g.e(r2), r2 * g.h)
        phi_2 = g.phi2(phi_1)
        phi_3 = phi_2 / g.h - 0.5
        if model_s is None:
            model_s = self.model_fn(x, s)
        if model_s1 is None:
            x_s1 = lin1 * x - (w1 * phi_11) * model_s
            model_s1 = self.model_fn(x_s1, s1)
        x_s2 = lin2 * x - (w2 * phi_12) * model_s
        c2 = r2 / r1 * (w2 * phi_22) * (model_s1 - model_s)
        x_s2 = x_s2 + c2 if plus else x_s2 - c2
        model_s2 = self.model_fn(x_s2, s2)
        base = g.lin * x - (g.w * phi_1) * model_s
        if solver_type == "dpmsolver":
            c = (1.0 / r2) * (g.w * phi_2) * (model_s2 - model_s)
            x_t = base + c if plus else base - c
        else:
            D1_0 = (1.0 / r1) * (model_s1 - model_s)
            D1_1 = (1.0 / r2) * (model_s2 - model_s)
            D1 = (r2 * D1_0 - r1 * D1_1) / (r2 - r1)
            D2 = 2.0 * (D1_1 - D1_0) / (r2 - r1)
            a = (g.w * phi_2) * D1
            x_t = (base + a if plus else base - a) - (g.w * phi_3) * D2
        if return_intermediate:
            return x_t, {"model_s": model_s, "model_s1": model_s1, "model_s2": model_s2}
        return x_t

    # ------------------------------------------------------------------ multistep updates
    def multistep_dpm_solver_second_update(self, x, model_prev_list, t_prev_list, t, solver_type="dpmsolver"):
        if solver_type not in ["dpmsolver", "taylor"]:
            raise ValueError("'solver_type' must be either 'dpmsolver' or 'taylor', got {}".format(solver_type))
        ns = self.noise_schedule
        m1, m0 = model_prev_list[-2], model_prev_list[-1]
        g = _Integrator(ns, self._plus, t_prev_list[-1], t)
        h_0 = g.lam_s - ns.marginal_lambda(t_prev_list[-2])
        r0 = h_0 / g.h
        D1_0 = (1.0 / r0) * (m0 - m1)
        phi_1 = g.e()
        base = g.lin * x - (g.w * phi_1) * m0
        if solver_type == "dpmsolver":
            return base - 0.5 * (g.w * phi_1) * D1_0
        c = (g.w * g.phi2(phi_1)) * D1_0
        return base + c if self._plus else base - c

    def multistep_dpm_solver_third_update(self, x, model_prev_list, t_prev_list, t, solver_type="dpmsolver"):
        ns = self.noise_schedule
        m2, m1, m0 = model_prev_list
        t2, t1, t0 = t_prev_list
        g = _Integrator(ns, self._plus, t0, t)
        lam2, lam1 = ns.marginal_lambda(t2), ns.marginal_lambda(t1)
        h_1 = lam1 - lam2
        h_0 = g.lam_s - lam1
        r0, r1 = h_0 / g.h, h_1 / g.h
        D1_0 = (1.0 / r0) * (m0 - m1)
        D1_1 = (1.0 / r1) * (m1 - m2)
        D1 = D1_0 + (r0 / (r0 + r1)) * (D1_0 - D1_1)
        D2 = (1.0 / (r0 + r1)) * (D1_0 - D1_1)
        phi_1 = g.e()
        phi_2 = g.phi2(phi_1)
        phi_3 = phi_2 / g.h - 0.5
        base = g.lin * x - (g.w * phi_1) * m0
        a = (g.w * phi_2) * D1
        return (base + a if self._plus else base - a) - (g.w * phi_3) * D2

    def singlestep_dpm_solver_update(self, x, s, t, order, return_intermediate=False, solver_type="dpmsolver",
                                     r1=None, r2=None):
        if order == 1:
            return self.dpm_solver_first_update(x, s, t, return_intermediate=return_intermediate)
        if order == 2:
            return self.singlestep_dpm_solver_second_update(x, s, t, return_intermediate=return_intermediate,
                                                            solver_type=solver_type, r1=r1)
        if order == 3:
            return self.singlestep_dpm_solver_third_update(x, s, t, return_intermediate=return_intermediate,
                                                           solver_type=solver_type, r1=r1, r2=r2)
        raise ValueError("Solver order must be 1 or 2 or 3, got {}".format(order))

    def multistep_dpm_solver_update(self, x, model_prev_list, t_prev_list, t, order, solver_type="dpmsolver"):
        if order == 1:
            return self.dpm_solver_first_update(x, t_prev_list[-1], t, model_s=model_prev_list[-1])
        if order == 2:
            return self.multistep_dpm_solver_second_update(x, model_prev_list, t_prev_list, t, solver_type=solver_type)
        if order == 3:
            return self.multistep_dpm_solver_third_update(x, model_prev_list, t_prev_list, t, solver_type=solver_type)
        raise ValueError("Solver order must be 1 or 2 or 3, got {}".format(order))

    # ------------------------------------------------------------------ adaptive step size
    def dpm_solver_adaptive(self, x, order, t_T, t_0, h_init=0.05, atol=0.0078, rtol=0.05, theta=0.9, t_err=1e-5,
                            solver_type="dpmsolver"):
        """Embedded-pair adaptive solver ("DPM-Solver-12"/"-23", reference :917-984)."""
        ns = self.noise_schedule
        s = t_T * torch.ones((1,)).to(x)
        lambda_s = ns.marginal_lambda(s)
        lambda_0 = ns.marginal_lambda(t_0 * torch.ones_like(s).to(x))
        h = h_init * torch.ones_like(s).to(x)
        x_prev = x
        nfe = 0
        if order == 2:
            def lower(x, s, t):
                return self.dpm_solver_first_update(x, s, t, return_intermediate=True)

            def higher(x, s, t, **kw):
                return self.singlestep_dpm_solver_second_update(x, s, t, r1=0.5, solver_type=solver_type, **kw)
        elif order == 3:
            def lower(x, s, t):
                return self.singlestep_dpm_solver_second_update(x, s, t, r1=1.0 / 3.0, return_intermediate=True,
                                                                solver_type=solver_type)

            def higher(x, s, t, **kw):
                return self.singlestep_dpm_solver_third_update(x, s, t, r1=1.0 / 3.0, r2=2.0 / 3.0,
                                                               solver_type=solver_type, **kw)
        else:
            raise ValueError("For adaptive step size solver, order must be 2 or 3, got {}".format(order))
        while torch.abs((s - t_0)).mean() > t_err:
            t = ns.inverse_lambda(lambda_s + h)
            x_lower, kw = lower(x, s, t)
            x_higher = higher(x, s, t, **kw)
            delta = torch.max(torch.ones_like(x).to(x) * atol, rtol * torch.max(torch.abs(x_lower), torch.abs(x_prev)))
            err = (x_higher - x_lower) / delta
            E = torch.sqrt(torch.square(err.reshape((err.shape[0], -1))).mean(dim=-1, keepdim=True)).max()
            if torch.all(E <= 1.0):
                x, s, x_prev = x_higher, t, x_lower
                lambda_s = ns.marginal_lambda(s)
            h = torch.min(theta * h * torch.float_power(E, -1.0 / order).float(), lambda_0 - lambda_s)
            nfe += order
        print("adaptive solver nfe", nfe)
        return x

    def add_noise(self, x, t, noise=None):
        alpha_t, sigma_t = self.noise_schedule.marginal_alpha(t), self.noise_schedule.marginal_std(t)
        if noise is None:
            noise = torch.randn((t.shape[0], *x.shape), device=x.device)
        x = x.reshape((-1, *x.shape))
        xt = expand_dims(alpha_t, x.dim()) * x + expand_dims(sigma_t, x.dim()) * noise
        return xt.squeeze(0) if t.shape[0] == 1 else xt

    def inverse(self, x, steps=20, t_start=None, t_end=None, order=2, skip_type="time_uniform", method="multistep",
                lower_order_final=True, denoise_to_zero=False, solver_type="dpmsolver", atol=0.0078, rtol=0.05,
                return_intermediate=False):
        t_0 = 1.0 / self.noise_schedule.total_N if t_start is None else t_start
        t_T = self.noise_schedule.T if t_end is None else t_end
        assert t_0 > 0 and t_T > 0, "Time range needs to be greater than 0. For discrete-time DPMs, it needs to be in [1 / N, 1], where N is the length of betas array"
        return self.sample(x, steps=steps, t_start=t_0, t_end=t_T, order=order, skip_type=skip_type, method=method,
                           lower_order_final=lower_order_final, denoise_to_zero=denoise_to_zero,
                           solver_type=solver_type, atol=atol, rtol=rtol, return_intermediate=return_intermediate)

    # ------------------------------------------------------------------ driver
    def sample(self, x, steps=20, t_start=None, t_end=None, order=2, skip_type="time_uniform", method="multistep",
               lower_order_final=True, denoise_to_zero=False, solver_type="dpmsolver", atol=0.0078, rtol=0.05,
               return_intermediate=False):
        t_0 = 1.0 / self.noise_schedule.total_N if t_end is None else t_end
        t_T = self.noise_schedule.T if t_start is None else t_start
        assert t_0 > 0 and t_T > 0, "Time range needs to be greater than 0. For discrete-time DPMs, it needs to be in [1 / N, 1], where N is the length of betas array"
        if return_intermediate:
            assert method in ["multistep", "singlestep", "singlestep_fixed"], "Cannot use adaptive solver when saving intermediate values"
        if self.correcting_xt_fn is not None:
            assert method in ["multistep", "singlestep", "singlestep_fixed"], "Cannot use adaptive solver when correcting_xt_fn is not None"
        device = x.device

        if (method == "multistep" and order == 2 and self._plus and solver_type == "dpmsolver"
                and not return_intermediate and not denoise_to_zero and self.correcting_x0_fn is None
                and self.correcting_xt_fn is None and not (lower_order_final and steps < 10) and steps >= 2):
            from . import fused
            out = fused.try_fused_dpm(self, x, steps, skip_type, t_T, t_0)
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
            if method == "adaptive":
                x = self.dpm_solver_adaptive(x, order=order, t_T=t_T, t_0=t_0, atol=atol, rtol=rtol, solver_type=solver_type)
            elif method == "multistep":
                assert steps >= order
                ts = self.get_time_steps(skip_type=skip_type, t_T=t_T, t_0=t_0, N=steps, device=device)
                assert ts.shape[0] - 1 == steps
                t_hist = [ts[0]]
                m_hist = [self.model_fn(x, ts[0])]
                x = after(x, ts[0], 0)
                for step in range(1, steps + 1):
                    t = ts[step]
                    if step < order:
                        k = step                                   # warm-up with lower orders
                    elif lower_order_final and steps < 10:
                        k = min(order, steps + 1 - step)
                    else:
                        k = order
                    x = self.multistep_dpm_solver_update(x, m_hist, t_hist, t, k, solver_type=solver_type)
                    x = after(x, t, step)
                    if step < order:
                        t_hist.append(t)
                        m_hist.append(self.model_fn(x, t))
                    else:
                        t_hist = t_hist[1:] + [t]
                        # the model is not evaluated after the final update (NFE == steps)
                        m_hist = m_hist[1:] + [self.model_fn(x, t) if step < steps else m_hist[-1]]
            elif method in ["singlestep", "singlestep_fixed"]:
                if method == "singlestep":
                    outer, orders = self.get_orders_and_timesteps_for_singlestep_solver(
                        steps=steps, order=order, skip_type=skip_type, t_T=t_T, t_0=t_0, device=device)
                else:
                    K = steps // order
                    orders = [order] * K
                    outer = self.get_time_steps(skip_type=skip_type, t_T=t_T, t_0=t_0, N=K, device=device)
                for step, k in enumerate(orders):
                    s, t = outer[step], outer[step + 1]
                    inner = self.get_time_steps(skip_type=skip_type, t_T=s.item(), t_0=t.item(), N=k, device=device)
                    lam = self.noise_schedule.marginal_lambda(inner)
                    h = lam[-1] - lam[0]
                    r1 = None if k <= 1 else (lam[1] - lam[0]) / h
                    r2 = None if k <= 2 else (lam[2] - lam[0]) / h
                    x = self.singlestep_dpm_solver_update(x, s, t, k, solver_type=solver_type, r1=r1, r2=r2)
                    x = after(x, t, step)
                step = len(orders) - 1
            else:
                raise ValueError("Got wrong method {}".format(method))
            if denoise_to_zero:
                t = torch.ones((1,)).to(device) * t_0
                x = self.denoise_to_zero_fn(x, t)
                x = after(x, t, steps + 1 if method == "multistep" else step + 1)
        return (x, track) if return_intermediate else x
