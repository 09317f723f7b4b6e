"""Fused sampling loops: UNet forward (C-ABI) + one fused sampler-step kernel per step.

``DenoiserSession`` is the explicit fast API (device-resident x, content, prompt):
conditioning prepared once per utterance batch, per-step scalars precomputed on the host
(``coefs.py``), no host synchronisation inside the loop (the reference syncs every step on
``assert torch.isnan(x).any() == False``, model.py:404).

``try_fused_dpm`` / ``try_fused_unipc`` let the drop-in ``DPM_Solver`` / ``UniPC`` classes
take this path when the model closure they were given turns out to wrap our UNet (detected by
tracing the first model call), so ``model.py:621-686`` benefits unchanged.
"""
from __future__ import annotations

import collections
import ctypes as C
import hashlib
import os
from typing import Optional

import torch

from . import _lib, coefs
from .unet import UNet1DConditionModel, trace_calls


def schedule_signature(ns) -> tuple:
    """Content key of a noise schedule.  The reference builds a NEW NoiseScheduleVP inside every ``sample()``
    (model.py:621-622, 655-656), so object identity is useless as a cache key (and a recycled ``id()`` could
    alias a different beta table); two schedules with the same knots share tables and captured graphs."""
    if getattr(ns, "schedule", None) == "discrete":
        la = ns.log_alpha_array.detach().to("cpu", torch.float32).contiguous()
        return ("discrete", int(ns.total_N), hashlib.sha1(la.numpy().tobytes()).hexdigest())
    return (str(getattr(ns, "schedule", "?")), float(getattr(ns, "beta_0", 0.0)), float(getattr(ns, "beta_1", 0.0)), int(getattr(ns, "total_N", 0)))


_TABLES = collections.OrderedDict()     # per-step scalars do not depend on the shape: shared by every session (the CLI: a new shape per slice)


def _step_table(kind, ns, ts, extra, key):
    tab = _TABLES.get(key)
    if tab is None:
        tab = coefs.dpmpp_2m_table(ns, ts, extra) if kind == "dpm" else coefs.unipc_bh2_table(ns, ts, extra)
        _TABLES[key] = tab
        while len(_TABLES) > 16:
            _TABLES.popitem(last=False)
    else:
        _TABLES.move_to_end(key)
    return tab


class DenoiserSession:
    """One utterance-batch shape (B, T, S) on one GPU: UNet engine, static input buffers, prepared
    conditioning and the whole sampling loop captured as one CUDA graph.

    The C calls are stream-ordered and allocation-free, so the N-step loop (prepare_cond + the timestep
    table of all N evaluation times + N x (UNet forward + fused sampler step), every kernel linked to its
    predecessor by programmatic dependent launch) is captured once per (sampler, time grid, schedule) and
    replayed; new inputs are copied into the static buffers.  ``NS2VC_GRAPH=0`` disables the capture.

    The reference asserts ``not isnan(x)`` on the host before every denoiser call (model.py:404: one
    device->host sync per step).  Here the sampler-step kernel accumulates a device flag and the host checks
    it ONCE after the run, raising the same ``AssertionError``."""

    MAX_GRAPHS = 4                                           # LRU bound of captured loops per session
    CAPTURE_AFTER = 3                                        # the loop is captured on its third run, replayed from the fourth

    def __init__(self, unet: UNet1DConditionModel, content_BCT: Optional[torch.Tensor], prompt_BSC: torch.Tensor,
                 prompt_mask: Optional[torch.Tensor], T: Optional[int] = None):
        if not prompt_BSC.is_cuda:
            raise RuntimeError("DenoiserSession needs CUDA tensors (no CPU path)")
        self.unet = unet
        self.dev = prompt_BSC.device
        self.B, self.S = prompt_BSC.shape[0], prompt_BSC.shape[1]
        self.Cc = unet.cfg.in_channels - unet.latent_channels
        if self.Cc > 0:
            if content_BCT is None or content_BCT.shape[1] != self.Cc:
                raise ValueError(f"content must be [B, {self.Cc}, T]")
            self.T = content_BCT.shape[2]
        else:
            if T is None:
                raise ValueError("T is required when the model has no content channels")
            self.T = T
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.content = torch.empty((self.B, self.Cc, self.T), **f32) if self.Cc > 0 else None
        self.prompt = torch.empty((self.B, self.S, unet.cfg.cross_attention_dim), **f32)
        self.mask = torch.empty((self.B, self.S), dtype=torch.uint8, device=self.dev) if prompt_mask is not None else None
        self.L = _lib.lib()
        self.h = unet.engine(self.dev)
        self.Cl, self.Co = unet.latent_channels, unet.cfg.out_channels
        self.x_in = torch.empty((self.B, self.Cl, self.T), **f32)
        self.first_out = torch.empty((self.B, self.Co, self.T), **f32)   # step-0 model output handed in by the drop-in samplers
        self.nan_flag = torch.zeros((1,), dtype=torch.int32, device=self.dev)
        self.ws = unet.workspace(self.B, self.T, self.S, self.dev)     # the module's shared grow-only scratch buffer
        self._graphs = collections.OrderedDict()
        self._wsig = unet._wsig
        self.set_cond(content_BCT, prompt_BSC, prompt_mask)

    def set_cond(self, content_BCT, prompt_BSC, prompt_mask):
        """Copy a new utterance batch (same shapes) into the static buffers."""
        if self.content is not None:
            self.content.copy_(content_BCT, non_blocking=True)
        self.prompt.copy_(prompt_BSC, non_blocking=True)
        if (prompt_mask is None) != (self.mask is None):
            raise ValueError("mask presence must not change within a session")
        if self.mask is not None:
            self.mask.copy_(prompt_mask.to(torch.bool), non_blocking=True)
        self._prepared = False

    def _stream(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    def prepare(self):
        with torch.cuda.device(self.dev):
            _lib.check(self.L.ns2vc_unet_prepare_cond(
                self.h, self.content.data_ptr() if self.content is not None else None,
                (self.Cc * self.T) if self.content is not None else 0, self.prompt.data_ptr(),
                self.mask.data_ptr() if self.mask is not None else None, self.B, self.T, self.S, self.ws.data_ptr(), self._stream()))
        self._prepared = True
        self.unet.__dict__["_cond_owner"] = self                 # the shared workspace holds THIS session's conditioning now

    def forward(self, x: torch.Tensor, t: torch.Tensor, out: torch.Tensor, film_rows: Optional[torch.Tensor] = None):
        """x [B,Cl,T] fp32 contiguous, t [B] fp32 (or B precomputed FiLM rows), out [B,Co,T] fp32 — all on the session device."""
        if not self._prepared or self.unet.__dict__.get("_cond_owner") is not self:
            self.prepare()
        with torch.cuda.device(self.dev):
            if film_rows is not None:
                _lib.check(self.L.ns2vc_unet_forward_film(self.h, x.data_ptr(), self.Cl * self.T, film_rows.data_ptr(), out.data_ptr(),
                                                          self.B, self.T, self.S, self.ws.data_ptr(), self._stream()))
            else:
                _lib.check(self.L.ns2vc_unet_forward(self.h, x.data_ptr(), self.Cl * self.T, t.data_ptr(), out.data_ptr(),
                                                     self.B, self.T, self.S, self.ws.data_ptr(), self._stream()))

    def time_table(self, tvals: torch.Tensor, table: torch.Tensor):
        """FiLM rows of every evaluation time of a run (tvals [steps, B] fp32) into ``table`` (needs prepare())."""
        with torch.cuda.device(self.dev):
            _lib.check(self.L.ns2vc_unet_time_table(self.h, tvals.data_ptr(), tvals.numel(), table.data_ptr(), self.B, self.T, self.S,
                                                    self.ws.data_ptr(), self._stream()))

    # ------------------------------------------------------------------ loop bodies (eager or under capture)
    def _film(self, ent, k):
        fw = ent["film_width"]
        return ent["table"][k * self.B * fw:(k + 1) * self.B * fw]

    def _body_dpm(self, ent, use_first):
        x = self.x_in.clone()
        n = x.numel()
        x_next, out = torch.empty_like(x), torch.empty_like(x)
        m_a, m_b = torch.empty_like(x), torch.empty_like(x)
        stream = self._stream()
        for k, st in enumerate(ent["steps"]):
            if k == 0 and use_first:
                out.copy_(self.first_out)
            else:
                self.forward(x, ent["tvals"][k], out, film_rows=self._film(ent, k))
            c = _lib.DpmCoef(st.alpha_s, st.sigma_s, st.c_x, st.c_m, st.c_d, st.inv_r0, st.order)
            with torch.cuda.device(self.dev):
                _lib.check(self.L.ns2vc_dpm_step(x.data_ptr(), out.data_ptr(), m_b.data_ptr(), C.byref(c), m_a.data_ptr(),
                                                 x_next.data_ptr(), n, self.nan_flag.data_ptr(), stream))
            x, x_next = x_next, x
            m_a, m_b = m_b, m_a
        return x

    def _body_unipc(self, ent, use_first):
        x_prev = self.x_in                                 # x at the previous time point (corrector base); never written
        x_eval = x_prev                                    # where the model is evaluated
        n = x_prev.numel()
        out = torch.empty_like(x_prev)
        m0 = m1 = None
        stream = self._stream()
        for k, st in enumerate(ent["steps"]):
            if k == 0 and use_first:
                out.copy_(self.first_out)
            else:
                self.forward(x_eval, ent["tvals"][k], out, film_rows=self._film(ent, k))
            m_t = torch.empty_like(x_prev)
            x_t = torch.empty_like(x_prev) if st.corr_order > 0 else None
            x_pred = torch.empty_like(x_prev)
            c = _lib.UniPcCoef(st.alpha_t, st.sigma_t, st.c_x, st.c_m, st.ab, st.rk, st.rho0, st.rho1, st.corr_order,
                               st.n_c_x, st.n_c_m, st.nab, st.nrk, st.pred_order)
            with torch.cuda.device(self.dev):
                _lib.check(self.L.ns2vc_unipc_step(
                    x_prev.data_ptr(), x_eval.data_ptr(), out.data_ptr(), m0.data_ptr() if m0 is not None else None,
                    m1.data_ptr() if m1 is not None else None, C.byref(c), m_t.data_ptr(),
                    x_t.data_ptr() if x_t is not None else None, x_pred.data_ptr(), n, self.nan_flag.data_ptr(), stream))
            # history m1 <- m0 <- m_t ; corrector base <- corrected x_t (x_eval itself at k = 0)
            m1, m0 = m0, m_t
            x_prev = x_t if x_t is not None else x_eval
            x_eval = x_pred
        return x_eval

    def _loop(self, kind, ent, use_first):
        """prepare_cond + timestep table + the N-step loop; returns the final latents (a fresh tensor)."""
        self.nan_flag.zero_()
        self.prepare()
        self.time_table(ent["tvals"], ent["table"])
        res = self._body_dpm(ent, use_first) if kind == "dpm" else self._body_unipc(ent, use_first)
        return res.clone()

    def _check_nan(self):
        if int(self.nan_flag.item()) != 0:
            # same exception type as the reference's per-call guard (model.py:404)
            raise AssertionError("NaN in the denoiser input during the fused sampling run (reference model.py:404)")

    def _run(self, kind, x_T, ns, ts, first_out, extra):
        assert self.Cl == self.Co, "x_start parameterisation needs out_channels == latent channels"
        if self.unet._wsig != self._wsig:                  # weights were re-packed: captured graphs are stale
            self._graphs.clear()
            self.h = self.unet.engine(self.dev)
            self._wsig = self.unet._wsig
            self._prepared = False
        self.x_in.copy_(x_T, non_blocking=True)
        use_first = first_out is not None
        if use_first:
            self.first_out.copy_(first_out, non_blocking=True)
        key = (kind, tuple(float(v) for v in ts), extra, schedule_signature(ns), use_first)
        use_graph = os.environ.get("NS2VC_GRAPH", "1") != "0"
        ent = self._graphs.get(key)
        if ent is None:
            steps = _step_table(kind, ns, ts, extra, key[:4])
            tvals = torch.tensor([[st.t_input] * self.B for st in steps], dtype=torch.float32).to(self.dev)
            nrows = tvals.numel()
            table = torch.empty(int(self.L.ns2vc_unet_time_table_floats(self.h, nrows)), dtype=torch.float32, device=self.dev)
            ent = {"steps": steps, "tvals": tvals, "table": table, "film_width": int(self.L.ns2vc_unet_film_width(self.h)),
                   "graph": None, "out": None, "runs": 0}
            self._graphs[key] = ent
            while len(self._graphs) > self.MAX_GRAPHS:     # LRU: the oldest captured loop (graph + tables) is dropped
                self._graphs.popitem(last=False)
        else:
            self._graphs.move_to_end(key)
        ent["runs"] += 1
        if not use_graph:
            res = self._loop(kind, ent, use_first)
        elif ent["graph"] is None and ent["runs"] < self.CAPTURE_AFTER:
            # eager: the first run builds the launch program and sets kernel attributes; a shape / schedule seen only once or
            # twice (the CLI: every slice a new length) never pays for a capture (~0.1-1 s for 6 000+ kernel nodes)
            res = self._loop(kind, ent, use_first)
        else:
            if ent["graph"] is None:
                torch.cuda.synchronize(self.dev)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    ent["out"] = self._loop(kind, ent, use_first)
                ent["graph"] = g
            ent["graph"].replay()
            self._prepared = True
            self.unet.__dict__["_cond_owner"] = self
            res = ent["out"].clone()
        self._check_nan()
        return res

    # ------------------------------------------------------------------ public samplers
    def sample_dpmpp_2m(self, x_T: torch.Tensor, ns, ts: torch.Tensor, lower_order_final: bool = True,
                        first_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """DPM-Solver++ multistep order 2 over time points ``ts`` (N+1 values), x_start model.
        Equivalent to reference DPM_Solver.sample(method='multistep', order=2) (dpm_solver.py:1171-1213).
        ``first_out``: the model output at ts[0] if the caller already evaluated it (saves one forward)."""
        return self._run("dpm", x_T, ns, ts, first_out, bool(lower_order_final))

    def sample_unipc(self, x_T: torch.Tensor, ns, ts: torch.Tensor, variant: str = "bh2",
                     first_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """UniPC multistep order 2, data prediction, lower_order_final (uni_pc.py:606-658)."""
        return self._run("unipc", x_T, ns, ts, first_out, variant)


def get_session(unet: UNet1DConditionModel, content_BCT, prompt_BSC, prompt_mask, T=None) -> DenoiserSession:
    """Session cache per (B, T, S, mask?) on the module: captured graphs and static buffers are reused
    across utterance batches of the same shape."""
    B, S = prompt_BSC.shape[0], prompt_BSC.shape[1]
    Tn = content_BCT.shape[2] if content_BCT is not None else T
    key = (B, Tn, S, prompt_mask is not None, str(prompt_BSC.device))
    cache = unet.__dict__.setdefault("_sessions", {})
    sess = cache.get(key)
    if sess is None or sess.h != unet.engine(prompt_BSC.device):
        while len(cache) >= 8:                               # bounded: oldest shape first (dicts keep insertion order)
            cache.pop(next(iter(cache)))
        sess = DenoiserSession(unet, content_BCT, prompt_BSC, prompt_mask, T=T)
        cache = unet.__dict__.setdefault("_sessions", {})     # (the workspace may have grown and dropped the old sessions)
        cache[key] = sess
    else:
        sess.set_cond(content_BCT, prompt_BSC, prompt_mask)
    return sess


def _fast_path_enabled() -> bool:
    return os.environ.get("NS2VC_B200_FUSED", "1") != "0"


def _trace_first_call(solver, x: torch.Tensor, t0: torch.Tensor):
    """Run the solver's first model evaluation through the user's closure while recording which
    denoiser calls it makes.  Returns (record, noise) if the closure is exactly one call of our
    UNet whose output is returned unchanged (x_start model, no guidance), else (None, noise)."""
    wrapped = getattr(solver, "_wrapped", None)
    from .schedule import WrappedModel
    with trace_calls() as recs:
        noise = solver.model(x, t0)
    if not isinstance(wrapped, WrappedModel) or wrapped.model_type != "x_start" or wrapped.guidance_type != "uncond":
        return None, noise
    if len(recs) != 1:
        return None, noise
    r = recs[0]
    u = r.unet
    Cl = u.latent_channels
    if u.cfg.out_channels != Cl or r.sample.shape[1] != u.cfg.in_channels or tuple(r.sample.shape[::2]) != tuple(x.shape[::2]):
        return None, noise
    if x.shape[1] != Cl or x.dtype != torch.float32:
        return None, noise
    if not torch.equal(r.sample[:, :Cl], x):
        return None, noise
    # the closure must hand back the UNet output itself: noise == (x - alpha*out)/sigma
    ns = wrapped.noise_schedule
    tt = t0.expand(x.shape[0])
    a, s = ns.marginal_alpha(tt), ns.marginal_std(tt)
    expect = (x - a[:, None, None] * r.output) / s[:, None, None]
    if not torch.equal(expect, noise):
        return None, noise
    return r, noise


def _session_from_record(r) -> DenoiserSession:
    u = r.unet
    Cl = u.latent_channels
    content = r.sample[:, Cl:] if r.sample.shape[1] > Cl else None
    return get_session(u, content, r.ehs, r.mask, T=r.sample.shape[2])


def _try_fused(solver, x, steps, skip_type, t_T, t_0, kind):
    """(latents, None) when the fused path ran; (None, first_noise) when the closure is not ours — first_noise is the
    model output of the solver's first evaluation (already paid for by the trace), or None if nothing was evaluated."""
    if not (_fast_path_enabled() and x.is_cuda and skip_type in ("time_uniform", "time_quadratic", "logSNR")):
        return None, None
    ns = solver.noise_schedule
    ts = solver.get_time_steps(skip_type=skip_type, t_T=t_T, t_0=t_0, N=steps, device="cpu")
    with torch.no_grad():
        rec, noise = _trace_first_call(solver, x, ts[0].to(x.device))
        if rec is None:
            return None, noise
        sess = _session_from_record(rec)
        if kind == "dpm":
            return sess.sample_dpmpp_2m(x, ns, ts, lower_order_final=True, first_out=rec.output), None
        return sess.sample_unipc(x, ns, ts, variant=solver.variant, first_out=rec.output), None


def try_fused_dpm(solver, x, steps, skip_type, t_T, t_0):
    return _try_fused(solver, x, steps, skip_type, t_T, t_0, "dpm")


def try_fused_unipc(solver, x, steps, skip_type, t_T, t_0):
    return _try_fused(solver, x, steps, skip_type, t_T, t_0, "unipc")
