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

import ctypes as C
import os
from typing import Optional

import torch

from . import _lib, coefs
from .unet import UNet1DConditionModel, trace_calls


class DenoiserSession:
    """One utterance-batch shape (B, T, S) on one GPU: UNet engine, static input buffers, prepared
    conditioning and (optionally) the whole sampling loop captured as one CUDA graph.

    The C calls are stream-ordered and allocation-free, so the N-step loop (prepare_cond + N x (UNet
    forward + fused sampler step), ~335 kernels per step linked by programmatic dependent launch) is
    captured once per (sampler, steps) and replayed; new inputs are copied into the static buffers.
    ``NS2VC_GRAPH=0`` disables the capture (eager launches)."""

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
        # Lanes: the utterances of a batch are independent, and every kernel of the step is a short
        # dependent-latency chain that leaves most SMs idle, so the batch is cut into sub-batches whose
        # whole sampling loops run concurrently on separate streams (own workspace + launch program each,
        # shared packed weights).  Off by default: GEMM CTAs own a whole SM (198 KB smem), so lanes do not overlap yet.
        want = int(os.environ.get("NS2VC_LANES", "1"))   # measured r01 (cfg2): 1 lane 1678, 2 lanes 1608, 4 lanes 1595, 8 lanes 1540 steps/s
        n_lanes = max(1, min(want, self.B))
        while self.B % n_lanes:
            n_lanes -= 1
        per = self.B // n_lanes
        self.lanes = []
        for g in range(n_lanes):
            n = C.c_size_t()
            _lib.check(self.L.ns2vc_unet_workspace_bytes(self.h, per, self.T, self.S, C.byref(n)))
            self.lanes.append(dict(sl=slice(g * per, (g + 1) * per), B=per,
                                   ws=torch.empty(int(n.value), dtype=torch.uint8, device=self.dev),
                                   stream=torch.cuda.Stream(device=self.dev) if n_lanes > 1 else None))
        self.ws = self.lanes[0]["ws"]
        self._graphs = {}
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

    def _prepare_lane(self, lane):
        sl = lane["sl"]
        with torch.cuda.device(self.dev):
            _lib.check(self.L.ns2vc_unet_prepare_cond(
                self.h, self.content[sl].data_ptr() if self.content is not None else None,
                (self.Cc * self.T) if self.content is not None else 0, self.prompt[sl].data_ptr(),
                self.mask[sl].data_ptr() if self.mask is not None else None, lane["B"], self.T, self.S, lane["ws"].data_ptr(), self._stream()))

    def _forward_lane(self, lane, x, t, out):
        with torch.cuda.device(self.dev):
            _lib.check(self.L.ns2vc_unet_forward(self.h, x.data_ptr(), self.Cl * self.T, t.data_ptr(), out.data_ptr(),
                                                 lane["B"], self.T, self.S, lane["ws"].data_ptr(), self._stream()))

    def prepare(self):
        for lane in self.lanes:
            self._prepare_lane(lane)
        self._prepared = True

    def forward(self, x: torch.Tensor, t: torch.Tensor, out: torch.Tensor):
        """x [B,Cl,T] fp32 contiguous, t [B] fp32, out [B,Co,T] fp32 — all on the session device."""
        if not self._prepared:
            self.prepare()
        for lane in self.lanes:
            sl = lane["sl"]
            self._forward_lane(lane, x[sl], t[sl], out[sl])

    # ------------------------------------------------------------------ loop bodies (eager or under capture)
    def _lane_dpm(self, lane, table, tvals, first_out, result):
        sl = lane["sl"]
        x = self.x_in[sl].clone()
        n = x.numel()
        x_next, out = torch.empty_like(x), torch.empty_like(x)
        m_a, m_b = torch.empty_like(x), torch.empty_like(x)
        stream = self._stream()
        for k, st in enumerate(table):
            if k == 0 and first_out is not None:
                out.copy_(first_out[sl])
            else:
                self._forward_lane(lane, x, tvals[k][sl], out)
            c = _lib.DpmCoef(st.alpha_s, st.sigma_s, st.c_x, st.c_m, st.c_d, st.inv_r0, st.order)
            with torch.cuda.device(self.dev):
                _lib.check(self.L.ns2vc_dpm_step(x.data_ptr(), out.data_ptr(), m_b.data_ptr(), C.byref(c), m_a.data_ptr(),
                                                 x_next.data_ptr(), n, stream))
            x, x_next = x_next, x
            m_a, m_b = m_b, m_a
        result[sl].copy_(x)

    def _lane_unipc(self, lane, table, tvals, first_out, result):
        sl = lane["sl"]
        x_prev = self.x_in[sl]                             # x at the previous time point (corrector base); never written
        x_eval = x_prev                                    # where the model is evaluated
        n = x_prev.numel()
        out = torch.empty_like(x_prev)
        m0 = m1 = None
        stream = self._stream()
        for k, st in enumerate(table):
            if k == 0 and first_out is not None:
                out.copy_(first_out[sl])
            else:
                self._forward_lane(lane, x_eval, tvals[k][sl], out)
            m_t = torch.empty_like(x_prev)
            x_t = torch.empty_like(x_prev) if st.corr_order > 0 else None
            x_pred = torch.empty_like(x_prev)
            c = _lib.UniPcCoef(st.alpha_t, st.sigma_t, st.c_x, st.c_m, st.ab, st.rk, st.rho0, st.rho1, st.corr_order,
                               st.n_c_x, st.n_c_m, st.nab, st.nrk, st.pred_order)
            with torch.cuda.device(self.dev):
                _lib.check(self.L.ns2vc_unipc_step(
                    x_prev.data_ptr(), x_eval.data_ptr(), out.data_ptr(), m0.data_ptr() if m0 is not None else None,
                    m1.data_ptr() if m1 is not None else None, C.byref(c), m_t.data_ptr(),
                    x_t.data_ptr() if x_t is not None else None, x_pred.data_ptr(), n, stream))
            # history m1 <- m0 <- m_t ; corrector base <- corrected x_t (x_eval itself at k = 0)
            m1, m0 = m0, m_t
            x_prev = x_t if x_t is not None else x_eval
            x_eval = x_pred
        result[sl].copy_(x_eval)

    def _loop(self, kind, table, tvals, first_out=None):
        """prepare_cond + the N-step loop of every lane; lanes run concurrently on their own streams."""
        body = self._lane_dpm if kind == "dpm" else self._lane_unipc
        result = torch.empty_like(self.x_in)
        cur = torch.cuda.current_stream(self.dev)
        for lane in self.lanes:
            if lane["stream"] is None:
                self._prepare_lane(lane)
                body(lane, table, tvals, first_out, result)
            else:
                lane["stream"].wait_stream(cur)
                with torch.cuda.stream(lane["stream"]):
                    self._prepare_lane(lane)
                    body(lane, table, tvals, first_out, result)
        for lane in self.lanes:
            if lane["stream"] is not None:
                cur.wait_stream(lane["stream"])
        self._prepared = True
        return result

    def _run(self, kind, x_T, ns, ts, first_out, extra):
        assert self.Cl == self.Co, "x_start parameterisation needs out_channels == latent channels"
        if self.unet._wsig != self._wsig:                  # weights were re-packed: captured graphs are stale
            self._graphs.clear()
            self.h = self.unet.engine(self.dev)
            self._wsig = self.unet._wsig
            self._prepared = False
        self.x_in.copy_(x_T, non_blocking=True)
        key = (kind, tuple(float(v) for v in ts), extra, id(ns))
        use_graph = os.environ.get("NS2VC_GRAPH", "1") != "0"
        ent = self._graphs.get(key)
        if ent is None:
            table = coefs.dpmpp_2m_table(ns, ts, extra) if kind == "dpm" else coefs.unipc_bh2_table(ns, ts, extra)
            tvals = torch.tensor([[st.t_input] * self.B for st in table], dtype=torch.float32).to(self.dev)
            ent = {"table": table, "tvals": tvals, "graph": None, "out": None, "warm": False}
            self._graphs[key] = ent
        if not use_graph:
            return self._loop(kind, ent["table"], ent["tvals"], first_out)
        if ent["graph"] is None:
            if not ent["warm"]:
                # first run eagerly: builds the launch programs, sets kernel attributes, warms the allocator
                res = self._loop(kind, ent["table"], ent["tvals"], first_out)
                ent["warm"] = True
                return res
            torch.cuda.synchronize(self.dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                ent["out"] = self._loop(kind, ent["table"], ent["tvals"], None)
            ent["graph"] = g
        ent["graph"].replay()
        self._prepared = True
        return ent["out"].clone()

    # ------------------------------------------------------------------ public samplers
    def sample_dpmpp_2m(self, x_T: torch.Tensor, ns, ts: torch.Tensor, lower_order_final: bool = True,
                        first_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """DPM-Solver++ multistep order 2 over time points ``ts`` (N+1 values), x_start model.
        Equivalent to reference DPM_Solver.sample(method='multistep', order=2) (dpm_solver.py:1171-1213)."""
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
        if len(cache) > 4:
            cache.clear()
        sess = DenoiserSession(unet, content_BCT, prompt_BSC, prompt_mask, T=T)
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


def try_fused_dpm(solver, x, steps, skip_type, t_T, t_0):
    if not (_fast_path_enabled() and x.is_cuda and skip_type in ("time_uniform", "time_quadratic", "logSNR")):
        return None
    ns = solver.noise_schedule
    ts = solver.get_time_steps(skip_type=skip_type, t_T=t_T, t_0=t_0, N=steps, device="cpu")
    with torch.no_grad():
        rec, _ = _trace_first_call(solver, x, ts[0].to(x.device))
        if rec is None:
            return None
        sess = _session_from_record(rec)
        return sess.sample_dpmpp_2m(x, ns, ts, lower_order_final=True, first_out=rec.output)


def try_fused_unipc(solver, x, steps, skip_type, t_T, t_0):
    if not (_fast_path_enabled() and x.is_cuda and skip_type in ("time_uniform", "time_quadratic", "logSNR")):
        return None
    ns = solver.noise_schedule
    ts = solver.get_time_steps(skip_type=skip_type, t_T=t_T, t_0=t_0, N=steps, device="cpu")
    with torch.no_grad():
        rec, _ = _trace_first_call(solver, x, ts[0].to(x.device))
        if rec is None:
            return None
        sess = _session_from_record(rec)
        return sess.sample_unipc(x, ns, ts, variant=solver.variant, first_out=rec.output)
