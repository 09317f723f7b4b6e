"""The CLI's real call pattern (reference infer.py:99-140, inference/infer_tool.py:141-187): one B=1 slice after another, every
slice a different length T, each through the drop-in UniPC at the reference's default 30 steps (model.py:655-686) with a FRESH
NoiseScheduleVP per call.  A new (B, T, S) means a new launch program, tensor maps, workspace and session; nothing is replayed
from a graph the first time a shape is seen.  Reports, per slice: wall time of the first call (cold shape), of an eager
repeat, of the capturing call and of a graph replay, and what the first-call overhead amounts to."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ns2vc_b200 import uni_pc as upc
from ns2vc_b200.arch import ns2vc_denoiser_config
from ns2vc_b200.synth import linear_betas, make_inputs, make_state_dict
from ns2vc_b200.unet import UNet1DConditionModel
from oracle import unet_oracle   # sequence_mask only (test-side helper)

dev = torch.device("cuda", 0)
cfg = ns2vc_denoiser_config()
unet = UNet1DConditionModel(in_channels=356, out_channels=100, block_out_channels=(128, 256, 384, 512), norm_num_groups=8,
                            cross_attention_dim=256, attention_head_dim=8, addition_embed_type="text", resnet_time_scale_shift="scale_shift")
unet.load_state_dict(make_state_dict(cfg, 0)); unet = unet.to(dev).eval()
betas = linear_betas(1000).to(dev)
g = torch.Generator().manual_seed(0)
lengths = [int(v) for v in torch.randint(280, 1100, (20,), generator=g)]
S = 187


def closure(inp):
    content, prompt, plen = inp["content"].to(dev), inp["prompt"].to(dev), inp["refer_lengths"].to(dev)
    def fn(x, t, **kw):
        assert torch.isnan(x).any() == False  # noqa: E712  (model.py:404)
        p = prompt.permute(1, 0, 2); c = content.permute(1, 2, 0)
        mask = unet_oracle.sequence_mask(plen, p.size(1)).to(torch.bool)
        return unet(torch.cat([x, c], dim=1), t, p, encoder_attention_mask=mask).sample
    return fn


def sample(inp, x):
    ns = upc.NoiseScheduleVP("discrete", betas=betas)                      # fresh object per call, as model.py:655-656
    mf = upc.model_wrapper(closure(inp), ns, model_type="x_start", model_kwargs={})
    return upc.UniPC(mf, ns, variant="bh2").sample(x, steps=30, order=2, skip_type="time_uniform", method="multistep")


def timed(fn):
    torch.cuda.synchronize(dev); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize(dev); return out, (time.perf_counter() - t0) * 1e3

with torch.no_grad():
    warm = make_inputs(1, 256, S, seed=99)
    sample(warm, warm["x"].to(dev)); sample(warm, warm["x"].to(dev))       # library load, kernel attributes, allocator
    rows = []
    for i, T in enumerate(lengths):
        inp = make_inputs(1, T, S, seed=100 + i)
        x = inp["x"].to(dev)
        _, cold = timed(lambda: sample(inp, x))                            # new shape: session, launch program, tensor maps; eager loop
        _, eager = timed(lambda: sample(inp, x))                           # known shape, still eager (DenoiserSession.CAPTURE_AFTER = 3)
        _, capture = timed(lambda: sample(inp, x))                         # third run: captures the loop as one CUDA graph
        _, replay = timed(lambda: sample(inp, x))                          # replays it
        rows.append((T, cold, eager, capture, replay))
    print("slice  T   cold_ms  eager_ms  capture_ms  replay_ms")
    for i, (T, c, e, cp, r) in enumerate(rows):
        print(f"{i:4d} {T:5d} {c:8.1f} {e:9.1f} {cp:11.1f} {r:10.1f}")
    tc, te, tr = sum(r[1] for r in rows), sum(r[2] for r in rows), sum(r[4] for r in rows)
    print(f"20 slices, every T new: {tc:.0f} ms cold total; {te:.0f} ms for the same calls on known shapes (eager); {tr:.0f} ms replayed from graphs")
    print(f"first-call overhead (session + launch program + tensor maps + first-call trace): {(tc - te) / len(rows):.1f} ms per slice = {100 * (tc - te) / tc:.1f} % of a cold slice")
