"""In-kernel timeline of every GEMM launch of one forward (globaltimer stamps of CTA (0,0))."""
import sys, os
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
n = 200
buf = torch.zeros(n * 32, dtype=torch.int64, device="cuda")
_lib.check(L.ns2vc_unet_set_trace(h, buf.data_ptr(), n))
sess.forward(x, t, o); torch.cuda.synchronize()
_lib.check(L.ns2vc_unet_set_trace(h, None, 0))
full = buf.view(n, 32).cpu()
tr = full[:, :8]
keep = tr[:, 0] > 0
tr = tr[keep]; ep = full[keep][:, 8:16]; pp = full[keep][:, 16:21]
t00 = int(tr[0, 0])
names = ["entry", "prologue_done", "pdl_wait_done", "first_full", "mma_issued", "acc_ready", "epi_done", "exit"]
print("idx start_us " + " ".join(f"d_{n}" for n in names[1:]) + " gap_from_prev_exit")
prev_exit = None
tot = {k: 0 for k in names[1:]}; gaps = 0
for i in range(tr.shape[0]):
    r = [int(v) for v in tr[i]]
    d = [(r[j] - r[0]) / 1e3 for j in range(1, 8)]
    gap = (r[0] - prev_exit) / 1e3 if prev_exit else 0.0
    prev_exit = r[7]
    if i < 60: print(i, f"{(r[0]-t00)/1e3:8.1f}", " ".join(f"{v:6.2f}" for v in d), f"{gap:7.2f}")
    for k, v in zip(names[1:], d): tot[k] += v
    gaps += gap
print("mean per launch (us since entry):", {k: round(v / tr.shape[0], 2) for k, v in tot.items()}, "mean gap to next gemm entry", round(gaps / tr.shape[0], 2))
print("forward span us", (int(tr[-1, 7]) - t00) / 1e3, "n gemm", tr.shape[0])

enames = ["acc_ready", "tmem->reg", "bias+res", "stores", "staged", "fenced", "atomics/done", "cta_sync"]
print("epilogue sub-steps, SM cycles since acc_ready (warp 2 lane 0 of CTA (0,0)):")
print("idx " + " ".join(f"{n:>12s}" for n in enames[1:]))
acc = [0.0] * 8; cnt = [0] * 8
for i in range(ep.shape[0]):
    e = [int(v) for v in ep[i]]
    d = [(e[j] - e[0]) if e[j] else 0 for j in range(8)]
    if i < 40: print(f"{i:3d} " + " ".join(f"{v:12d}" for v in d[1:]))
    for j in range(8):
        if e[j]: acc[j] += d[j]; cnt[j] += 1
print("mean " + " ".join(f"{(acc[j] / cnt[j] if cnt[j] else 0):12.0f}" for j in range(1, 8)))

pn = ["wait_done", "affine_ready", "panel0_full", "panel0_done", "panels_done"]
print("panel mode (us since kernel entry; first transform thread of CTA 0), then first MMA / acc_ready / exit of the same launch:")
print("idx " + " ".join(f"{n:>13s}" for n in pn) + "    first_full    acc_ready         exit")
k = 0
for i in range(tr.shape[0]):
    if int(pp[i, 0]) == 0: continue
    r0 = int(tr[i, 0])
    if k < 70: print(f"{i:3d} " + " ".join(f"{(int(v) - r0) / 1e3:13.2f}" for v in pp[i]) + " ".join(f"{(int(tr[i, j]) - r0) / 1e3:13.2f}" for j in (3, 5, 7)))
    k += 1
