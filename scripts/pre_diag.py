"""Per-layer error of the GPU condition encoders against the oracle on the full fixture shape, for one backend / switch set
(run once per environment: NS2VC_GEMM_BACKEND=simt, NS2VC_ATTN_P=split, ...).  Prints worst err/tol per tap."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ns2vc_b200.pre_model import Pre_model  # noqa: E402
from oracle import pre_model_oracle as po  # noqa: E402

B, T, S = (int(v) for v in (sys.argv[1:4] if len(sys.argv) >= 4 else (2, 48, 32)))
cfg = {"phoneme_encoder": dict(in_channels=256, hidden_channels=256, out_channels=256, n_layers=6, p_dropout=0.2),
       "prompt_encoder": dict(in_channels=100, hidden_channels=256, out_channels=256, n_layers=6, p_dropout=0.2)}
m = Pre_model(cfg)
sd = po.synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=0)
m.load_state_dict(sd)
m = m.cuda().eval()
g = torch.Generator().manual_seed(7)
c = torch.randn((B, 256, T), generator=g)
refer = torch.randn((B, 100, S), generator=g)
lengths = torch.tensor([max(1, T - 13 * i) for i in range(B)])
refer_lengths = torch.tensor([max(1, S - 7 * i) for i in range(B)])
ref_taps = {}
with torch.no_grad():
    rc, rp = po.pre_model_infer(sd, c, refer, lengths, refer_lengths, 6, 6, ref_taps)
data = (c.cuda(), refer.cuda(), None, None, None, lengths.cuda(), refer_lengths.cuda(), None)
taps = m.taps(data)
gc, gp = m.infer(data)


def worst(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    e = (a - b).abs()
    return (e / (1e-4 + 1e-3 * b.abs())).max().item(), e.max().item(), b.pow(2).mean().sqrt().item()


print("env", {k: v for k, v in os.environ.items() if k.startswith("NS2VC")}, "launches", m.launch_count())
for k, ref in ref_taps.items():
    ref = ref.squeeze(-1).unsqueeze(1) if k == "ref_enc" else ref.transpose(0, 1)
    w, e, r = worst(taps[k], ref)
    print(f"  {k:32s} err/tol {w:6.2f}  max_abs {e:.2e}  rms {r:.2e}")
for name, a, b in (("content", gc, rc), ("prompt", gp, rp)):
    w, e, r = worst(a, b)
    print(f"  {name:32s} err/tol {w:6.2f}  max_abs {e:.2e}  rms {r:.2e}")
