"""DPM-Solver++ sampler — drop-in for the part of the reference's ``sampler/dpm_solver.py`` that NS2VC uses.

``NaturalSpeech2.sample(method='dpmsolver')`` (reference ``model.py:621-652``) builds
``NoiseScheduleVP('discrete', betas)``, ``model_wrapper(..., model_type='x_start')`` and
``DPM_Solver(model_fn, ns, algorithm_type='dpmsolver++').sample(x, steps, order=2,
skip_type='time_uniform', method='multistep')``.  This module keeps exactly that surface (same
constructor and ``sample`` signature, so the call site is unchanged when ``sampler.dpm_solver``
resolves here): data-prediction multistep solver of order 1 / 2 on the three time grids.  Every other
mode of the reference class (noise-prediction 'dpmsolver', singlestep / adaptive solvers, order 3,
'taylor', thresholding / x_t correction hooks, intermediate outputs) raises ``NotImplementedError`` —
they are not on NS2VC's path (unsupported modes are rejected loudly).

There is ONE implementation of the step arithmetic: the per-step scalars of ``coefs.dpmpp_2m_table``
(host fp32, the reference's op order, ``dpm_solver.py:547-592, 796-852``).  The fused CUDA path
(``fused.try_fused_dpm``: UNet forward + ``dpm_step_kernel`` per step, whole loop as one CUDA graph)
and the generic loop below (any model closure, any device) consume the same table.
"""
from __future__ import annotations

import torch

from . import coefs
from .schedule import NoiseScheduleVP, model_wrapper, interpolate_fn, expand_dims  # noqa: F401 (API re-export)


def _unsupported(what: str):
    raise NotImplementedError(f"ns2vc_b200.dpm_solver: {what} is not implemented — supported: algorithm_type='dpmsolver++', "
                              "method='multistep', order 1|2, solver_type='dpmsolver', skip_type time_uniform|logSNR|time_quadratic")


def time_grid(ns: NoiseScheduleVP, skip_type: str, t_T: float, t_0: float, N: int, device) -> torch.Tensor:
    """The N+1 time points from t_T down to t_0 (reference :453-480)."""
    if skip_type == "logSNR":
        lambda_T = ns.marginal_lambda(torch.tensor(t_T).to(device))
        lambda_0 = ns.marginal_lambda(torch.tensor(t_0).to(device))
        return ns.inverse_lambda(torch.linspace(lambda_T.cpu().item(), lambda_0.cpu().item(), N + 1).to(device))
    if skip_type == "time_uniform":
        return torch.linspace(t_T, t_0, N + 1).to(device)
    if skip_type == "time_quadratic":
        return torch.linspace(t_T ** 0.5, t_0 ** 0.5, N + 1).pow(2).to(device)
    raise ValueError("Unsupported skip_type {}, need to be 'logSNR' or 'time_uniform' or 'time_quadratic'".format(skip_type))


class DPM_Solver:
    def __init__(self, model_fn, noise_schedule, algorithm_type="dpmsolver++", correcting_x0_fn=None,
                 correcting_xt_fn=None, thresholding_max_val=1.0, dynamic_thresholding_ratio=0.995):
        assert algorithm_type in ["dpmsolver", "dpmsolver++"]
        if algorithm_type != "dpmsolver++":
            _unsupported("algorithm_type='dpmsolver' (noise prediction)")
        if correcting_x0_fn is not None or correcting_xt_fn is not None:
            _unsupported("correcting_x0_fn / correcting_xt_fn")
        self._wrapped = model_fn
        self.model = lambda x, t: model_fn(x, t.expand((x.shape[0])))
        self.noise_schedule = noise_schedule
        self.algorithm_type = algorithm_type

    def get_time_steps(self, skip_type, t_T, t_0, N, device):
        return time_grid(self.noise_schedule, skip_type, t_T, t_0, N, device)

    def data_prediction_fn(self, x, t):
        """x0 from the wrapped noise model (reference :433-442)."""
        noise = self.model(x, t)
        alpha_t, sigma_t = self.noise_schedule.marginal_alpha(t), self.noise_schedule.marginal_std(t)
        return (x - sigma_t * noise) / alpha_t

    def sample(self, x, steps=20, t_start=None, t_end=None, order=2, skip_type="time_uniform", method="multistep",
               lower_order_final=True, denoise_to_zero=False, solver_type="dpmsolver", atol=0.0078, rtol=0.05,
               return_intermediate=False):
        ns = self.noise_schedule
        t_0 = 1.0 / ns.total_N if t_end is None else t_end
        t_T = ns.T if t_start is None else t_start
        assert t_0 > 0 and t_T > 0, "Time range needs to be greater than 0. For discrete-time DPMs, it needs to be in [1 / N, 1], where N is the length of betas array"
        if method != "multistep":
            _unsupported(f"method={method!r}")
        if order not in (1, 2):
            _unsupported(f"order={order}")
        if solver_type != "dpmsolver":
            _unsupported(f"solver_type={solver_type!r}")
        if return_intermediate:
            _unsupported("return_intermediate")
        assert steps >= order
        first = None                                       # model output at ts[0] when the fused-path probe already evaluated it
        if order == 2 and not denoise_to_zero and not (lower_order_final and steps < 10):
            from . import fused
            out, first = fused.try_fused_dpm(self, x, steps, skip_type, t_T, t_0)
            if out is not None:
                return out
        # Generic loop: the same per-step scalars as the fused kernel (coefs.DpmStep), applied with torch ops.
        ts = self.get_time_steps(skip_type=skip_type, t_T=t_T, t_0=t_0, N=steps, device=x.device)
        table = coefs.dpmpp_2m_table(ns, ts, lower_order_final, order=order)
        f = lambda v: torch.tensor(v, dtype=torch.float32, device=x.device)
        m_prev = None
        with torch.no_grad():
            for k, st in enumerate(table):
                noise = first if (k == 0 and first is not None) else self.model(x, ts[k])
                m0 = (x - f(st.sigma_s) * noise) / f(st.alpha_s)                       # data_prediction_fn (:437-439)
                r = f(st.c_x) * x - f(st.c_m) * m0                                    # first-order update (:569-576)
                if st.order == 2:
                    r = r - f(st.c_d) * (f(st.inv_r0) * (m0 - m_prev))               # second-order term (:813-831)
                x, m_prev = r, m0
            if denoise_to_zero:
                x = self.data_prediction_fn(x, torch.ones((1,)).to(x.device) * t_0)
        return x
