"""UniPC sampler — drop-in for the part of the reference's ``sampler/uni_pc.py`` that NS2VC uses.

``NaturalSpeech2.sample(method='unipc')`` (reference ``model.py:655-686``) calls
``UniPC(model_fn, ns, variant='bh2').sample(x, steps, order=2, skip_type='time_uniform',
method='multistep')`` with an x_start model.  This module keeps that surface: the data-prediction
B(h) predictor-corrector of order 2 with ``lower_order_final`` (variants bh1 / bh2).  Everything
else of the reference class (noise prediction, 'vary_coeff', other orders, thresholding /
correction hooks, intermediate outputs) raises ``NotImplementedError``.

One implementation of the step arithmetic: ``coefs.unipc_bh2_table`` (host fp32 scalars in the
reference's op order, ``uni_pc.py:471-588``) feeds both the fused CUDA path
(``fused.try_fused_unipc`` + ``unipc_step_kernel``) and the generic torch loop below.

Deliberate difference: the reference's wrapper broadcasts ``alpha_t[B]`` against ``[B,C,T]`` without
``expand_dims`` (``uni_pc.py:191``) and so only works for B = 1; here every sample gets its own scalar
(identical for B = 1).  The 'cosine' schedule of the reference's copy of NoiseScheduleVP is not carried.
"""
from __future__ import annotations

import torch

from . import coefs
from .dpm_solver import time_grid
from .schedule import NoiseScheduleVP as _NoiseScheduleVP, model_wrapper, interpolate_fn, expand_dims  # noqa: F401


class NoiseScheduleVP(_NoiseScheduleVP):
    """The uni_pc copy of the schedule never clips the log-alpha tail (reference uni_pc.py:78-87)."""

    def __init__(self, schedule="discrete", betas=None, alphas_cumprod=None, continuous_beta_0=0.1,
                 continuous_beta_1=20.0, dtype=torch.float32):
        if schedule == "cosine":
            raise ValueError("the 'cosine' schedule is not supported by ns2vc_b200 (NS2VC uses 'discrete')")
        super().__init__(schedule, betas, alphas_cumprod, continuous_beta_0, continuous_beta_1, dtype, clip_alpha=False)


def _unsupported(what: str):
    raise NotImplementedError(f"ns2vc_b200.uni_pc: {what} is not implemented — supported: algorithm_type='data_prediction', "
                              "variant bh1|bh2, method='multistep', order=2, lower_order_final=True")


class UniPC:
    def __init__(self, model_fn, noise_schedule, algorithm_type="data_prediction", correcting_x0_fn=None,
                 correcting_xt_fn=None, thresholding_max_val=1.0, dynamic_thresholding_ratio=0.995, variant="bh1"):
        assert algorithm_type in ["data_prediction", "noise_prediction"]
        if algorithm_type != "data_prediction":
            _unsupported("algorithm_type='noise_prediction'")
        if variant not in ("bh1", "bh2"):
            _unsupported(f"variant={variant!r}")
        if correcting_x0_fn is not None or correcting_xt_fn is not None:
            _unsupported("correcting_x0_fn / correcting_xt_fn")
        self._wrapped = model_fn
        self.model = lambda x, t: model_fn(x, t.expand((x.shape[0])))
        self.noise_schedule = noise_schedule
        self.variant = variant
        self.predict_x0 = True

    def get_time_steps(self, skip_type, t_T, t_0, N, device):
        return time_grid(self.noise_schedule, skip_type, t_T, t_0, N, device)

    def sample(self, x, steps=20, t_start=None, t_end=None, order=2, skip_type="time_uniform", method="multistep",
               lower_order_final=True, denoise_to_zero=False, atol=0.0078, rtol=0.05, return_intermediate=False):
        ns = self.noise_schedule
        t_0 = 1.0 / ns.total_N if t_end is None else t_end
        t_T = ns.T if t_start is None else t_start
        assert t_0 > 0 and t_T > 0, "Time range needs to be greater than 0. For discrete-time DPMs, it needs to be in [1 / N, 1], where N is the length of betas array"
        if method != "multistep":
            raise ValueError("Got wrong method {}".format(method))
        if order != 2 or not lower_order_final:
            _unsupported(f"order={order}, lower_order_final={lower_order_final}")
        if return_intermediate or denoise_to_zero:
            _unsupported("return_intermediate / denoise_to_zero")
        assert steps >= order
        first = None                                       # model output at ts[0] when the fused-path probe already evaluated it
        if self.variant == "bh2" and steps >= 3:
            from . import fused
            out, first = fused.try_fused_unipc(self, x, steps, skip_type, t_T, t_0)
            if out is not None:
                return out
        # Generic loop: the per-step scalars of the fused kernel (coefs.UniPcStep) applied with torch ops.
        ts = self.get_time_steps(skip_type=skip_type, t_T=t_T, t_0=t_0, N=steps, device=x.device)
        table = coefs.unipc_bh2_table(ns, ts, self.variant)
        f = lambda v: torch.tensor(v, dtype=torch.float32, device=x.device)
        x_prev, x_eval, m0, m1 = x, x, None, None
        with torch.no_grad():
            for k, st in enumerate(table):
                noise = first if (k == 0 and first is not None) else self.model(x_eval, ts[k])
                mt = (x_eval - f(st.sigma_t) * noise) / f(st.alpha_t)                  # data_prediction_fn (:304-312)
                xt = x_eval
                if st.corr_order > 0:                                                  # corrector at ts[k] (:533-567)
                    xbar = f(st.c_x) * x_prev - f(st.c_m) * m0
                    d1t = mt - m0
                    inner = f(st.rho1) * d1t if st.corr_order == 1 else f(st.rho0) * ((m1 - m0) / f(st.rk)) + f(st.rho1) * d1t
                    xt = xbar - f(st.ab) * inner
                xpred = f(st.n_c_x) * xt - f(st.n_c_m) * mt                            # predictor to ts[k+1] (:540-559)
                if st.pred_order == 2:
                    xpred = xpred - f(st.nab) * (f(0.5) * ((m0 - mt) / f(st.nrk)))
                m1, m0 = m0, mt
                x_prev, x_eval = xt, xpred
        return x_eval
