"""In-situ timeline of one UNet forward replayed from a CUDA graph: [first CTA entry, last CTA exit] of every launch."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ns2vc_b200 import _lib, api
from ns2vc_b200.arch import ns2vc_denoiser_config
from ns2vc_b200.fused import DenoiserSession
from ns2vc_b200.synth import make_inputs, make_state_dict
from ns2vc_b200.unet import UNet1DConditionModel
B, T, S = 8, 1024, 256
cfg = ns2vc_denoiser_config()
unet = UNet1DConditionModel(in_channels=356, out_channels=100, block_out_channels=(128, 256, 384, 512), norm_num_groups=8,
                            cross_attention_dim=256, attention_head_dim=8, addition_embed_type="text", resnet_time_scale_shift="scale_shift")
unet.load_state_dict(make_state_dict(cfg, 0)); unet = unet.cuda().eval()
inp = make_inputs(B, T, S, seed=0)
sess = DenoiserSession(unet, inp["content"].permute(1, 2, 0).contiguous().cuda(), inp["prompt"].permute(1, 0, 2).contiguous().cuda(),
                       api.sequence_mask(inp["refer_lengths"].cuda(), S))
x = inp["x"].cuda(); t = torch.full((B,), 500.0, device="cuda"); o = torch.empty_like(x)
for _ in range(3): sess.forward(x, t, o)
torch.cuda.synchronize()
L = _lib.lib(); h = unet.engine(torch.device("cuda", 0))
n = L.ns2vc_unet_launch_count(h)
buf = torch.empty(n * 2, dtype=torch.int64, device="cuda")
def reset():
    v = buf.view(n, 2); v[:, 0] = 0x7fffffffffffffff; v[:, 1] = 0
reset()
_lib.check(L.ns2vc_unet_set_span_trace(h, buf.data_ptr(), n))
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    sess.forward(x, t, o)
_lib.check(L.ns2vc_unet_set_span_trace(h, None, 0))
g.replay(); torch.cuda.synchronize(); reset(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
print("graph replay of 1 forward: %.3f ms" % e0.elapsed_time(e1))
tr = buf.view(n, 2).cpu()
names = [L.ns2vc_profile_kind_name(L.ns2vc_unet_launch_kind(h, i)).decode() for i in range(n)]
rows = [(i, names[i], int(tr[i, 0]), int(tr[i, 1])) for i in range(n) if int(tr[i, 1]) > 0]
t0 = rows[0][2]
dur = collections.defaultdict(float); cnt = collections.Counter(); gap = collections.defaultdict(float)
prev_end = None
out = []
for (i, k, a, b) in rows:
    d = (b - a) / 1e3
    gp = (a - prev_end) / 1e3 if prev_end is not None else 0.0     # negative = overlapped its predecessor (PDL)
    dur[k] += d; cnt[k] += 1; gap[k] += max(gp, 0.0)
    out.append(f"{i:4d} {k:14s} start {(a - t0) / 1e3:9.1f} us  dur {d:7.2f}  gap_before {gp:6.2f}")
    prev_end = b if prev_end is None else max(prev_end, b)
print("\n".join(out[:120]))
span = (max(r[3] for r in rows) - t0) / 1e3
print("forward span %.1f us (spans of traced kernels only)" % span)
for k in dur: print(f"{k:14s} n={cnt[k]:4d} sum_dur={dur[k]:8.1f} us  mean={dur[k]/cnt[k]:6.2f}  idle_before_sum={gap[k]:8.1f}")
