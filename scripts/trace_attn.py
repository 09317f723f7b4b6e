"""Per-key-tile SM-clock timeline of CTA (0,0,0) of the first attention launches of one forward (attention_v2.cu)."""
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
n = 32
buf = torch.zeros(n * 2048, dtype=torch.int64, device="cuda")
_lib.check(L.ns2vc_unet_set_attn_trace(h, buf.data_ptr(), n))
sess.forward(x, t, o); torch.cuda.synchronize()
_lib.check(L.ns2vc_unet_set_attn_trace(h, None, 0))
full = buf.view(n, 2048).cpu()
tr = full[:, :256].reshape(n, 16, 16)
names = ["top", "s_full", "S->reg", "max_xchg", "exp", "o_full", "P_done", "o_acc", "mma:S_go", "mma:S_iss", "mma:p_full", "mma:PV_iss", "tma:empty", "tma:iss"]
for li in (0, 1, 4, 5, 12, 13):
    a = tr[li]
    base = int(a[0, 0]) if int(a[0, 0]) else int(a[a > 0].min())
    print(f"--- attention launch {li}: cycles since the softmax loop top of tile 0 (SM clock, ~1.9 GHz)")
    print("tile " + " ".join(f"{nm:>10s}" for nm in names))
    for j in range(16):
        if int(a[j, 0]) == 0 and int(a[j, 10]) == 0: break
        print(f"{j:4d} " + " ".join(f"{(int(a[j, k]) - base) if int(a[j, k]) else 0:10d}" for k in range(14)))

import collections
for li in (0, 1, 4):
    c = full[li, 256:256 + 3 * 597].reshape(597, 3)
    c = c[c[:, 0] > 0]
    t0 = int(c[:, 0].min())
    per_sm = collections.Counter(int(v) for v in c[:, 2])
    starts = sorted((int(v) - t0) / 1e3 for v in c[:, 0]); ends = sorted((int(v) - t0) / 1e3 for v in c[:, 1])
    late = sum(1 for v in starts if v > 5.0)
    print(f"--- launch {li}: {c.shape[0]} CTAs on {len(per_sm)} SMs, CTAs per SM min/max {min(per_sm.values())}/{max(per_sm.values())}; "
          f"{late} CTAs started > 5 us after the first; start pct (us) 50/90/100: {starts[len(starts)//2]:.1f}/{starts[int(len(starts)*0.9)]:.1f}/{starts[-1]:.1f}; "
          f"end pct 10/50/90/100: {ends[len(ends)//10]:.1f}/{ends[len(ends)//2]:.1f}/{ends[int(len(ends)*0.9)]:.1f}/{ends[-1]:.1f}")
    conc = collections.Counter()
    for s_, e_, sm in c.tolist():
        conc[sm] = max(conc[sm], sum(1 for s2, e2, sm2 in c.tolist() if sm2 == sm and s2 <= s_ < e2))
    print("    max simultaneously resident CTAs on one SM:", max(conc.values()), " histogram:", sorted(collections.Counter(conc.values()).items()))
