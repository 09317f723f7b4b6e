"""Condition encoders under the CLI's call pattern (reference infer.py:99-140): one B=1 slice after another, every slice a new
length T (new launch program, tensor maps, workspace growth).  Per slice: wall time of the first Pre_model.infer of that shape and of a
repeat on the now-known shape."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ns2vc_b200.pre_model import Pre_model
from ns2vc_b200.synth import make_pre_inputs, make_pre_state_dict

dev = torch.device("cuda", 0)
cfg = {"phoneme_encoder": dict(in_channels=256, hidden_channels=256, out_channels=256, n_layers=6, p_dropout=0.2),
       "prompt_encoder": dict(in_channels=100, hidden_channels=256, out_channels=256, n_layers=6, p_dropout=0.2)}
pre = Pre_model(cfg); pre.load_state_dict(make_pre_state_dict(cfg, 0)); pre = pre.to(dev).eval()
g = torch.Generator().manual_seed(0)
lengths = [int(v) for v in torch.randint(280, 1100, (20,), generator=g)]
S = 187


def data_of(T, seed):
    p = make_pre_inputs(1, T, S, seed=seed)
    return (p["c"].to(dev), p["refer"].to(dev), None, None, None, p["lengths"].to(dev), p["refer_lengths"].to(dev), None)


def timed(fn):
    torch.cuda.synchronize(dev); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize(dev); return out, (time.perf_counter() - t0) * 1e3

w = data_of(256, 99)
pre.infer(w); pre.infer(w)                                                 # library load, weight packing, kernel attributes
rows = []
for i, T in enumerate(lengths):
    d = data_of(T, 100 + i)
    _, cold = timed(lambda: pre.infer(d))
    _, warm = timed(lambda: pre.infer(d))
    rows.append((T, cold, warm))
print("slice  T   cold_ms  warm_ms")
for i, (T, c, wm) in enumerate(rows):
    print(f"{i:4d} {T:5d} {c:8.2f} {wm:8.2f}")
tc, tw = sum(r[1] for r in rows), sum(r[2] for r in rows)
print(f"20 B=1 slices, every T new: {tc:.1f} ms cold total, {tw:.1f} ms on known shapes: first-call overhead {(tc - tw) / len(rows):.2f} ms per slice "
      f"(launch program + {pre.launch_count()} launches' tensor maps; no device allocation unless the workspace grows)")
