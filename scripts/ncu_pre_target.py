"""ncu target for the condition encoders: load + pack, one warm-up infer, then ONE Pre_model.infer at the cfg2 shape inside a
cudaProfilerStart/Stop window (run ncu with --profile-from-start off)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ns2vc_b200.pre_model import Pre_model
from ns2vc_b200.synth import make_pre_inputs, make_pre_state_dict

B, T, S = 8, 1024, 256
cfg = {"phoneme_encoder": dict(in_channels=256, hidden_channels=256, out_channels=256, n_layers=6, p_dropout=0.2),
       "prompt_encoder": dict(in_channels=100, hidden_channels=256, out_channels=256, n_layers=6, p_dropout=0.2)}
pre = Pre_model(cfg)
pre.load_state_dict(make_pre_state_dict(cfg, 0))
pre = pre.cuda().eval()
pin = make_pre_inputs(B, T, S, seed=5)
data = (pin["c"].cuda(), pin["refer"].cuda(), None, None, None, pin["lengths"].cuda(), pin["refer_lengths"].cuda(), None)
pre.infer(data)
torch.cuda.synchronize()
torch.cuda.profiler.start()
c, p = pre.infer(data)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", float(c.abs().mean()), float(p.abs().mean()), "launches", pre.launch_count())
