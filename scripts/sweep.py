"""Throughput over the sequence-length sweep of SURVEY.md 8(d) (cfg2 / cfg3 shapes): denoiser-steps/s of the fused
50-step DPM-Solver++(2M) and UniPC-bh2 loops (CUDA-graph replay, inputs resident) and the time of one Pre_model.infer (condition
encoders) at the same shape, one line of JSON per point."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ns2vc_b200 import api
from ns2vc_b200.arch import ns2vc_denoiser_config
from ns2vc_b200.fused import get_session
from ns2vc_b200.pre_model import Pre_model
from ns2vc_b200.synth import make_inputs, make_pre_inputs, make_pre_state_dict, make_state_dict
from ns2vc_b200.unet import UNet1DConditionModel

cfg = ns2vc_denoiser_config()
unet = UNet1DConditionModel(in_channels=356, out_channels=100, block_out_channels=(128, 256, 384, 512), norm_num_groups=8,
                            cross_attention_dim=256, attention_head_dim=8, addition_embed_type="text", resnet_time_scale_shift="scale_shift")
unet.load_state_dict(make_state_dict(cfg, 0)); unet = unet.cuda().eval()
PRE_CFG = {"phoneme_encoder": dict(in_channels=256, hidden_channels=256, out_channels=256, n_layers=6, p_dropout=0.2),
           "prompt_encoder": dict(in_channels=100, hidden_channels=256, out_channels=256, n_layers=6, p_dropout=0.2)}
pre = Pre_model(PRE_CFG); pre.load_state_dict(make_pre_state_dict(PRE_CFG, 0)); pre = pre.cuda().eval()
ns = api.default_schedule()
NFE = 50
points = [(8, 256), (8, 512), (8, 1024), (4, 2048), (2, 4096), (1, 1024)]
for B, T in points:
    S = 256
    inp = make_inputs(B, T, S, seed=5)
    content = inp["content"].permute(1, 2, 0).contiguous().cuda(); prompt = inp["prompt"].permute(1, 0, 2).contiguous().cuda()
    mask = api.sequence_mask(inp["refer_lengths"].cuda(), S); x = inp["x"].cuda()
    ts = torch.linspace(1.0, 1e-3, NFE + 1)
    pin = make_pre_inputs(B, T, S, seed=5)
    data = (pin["c"].cuda(), pin["refer"].cuda(), None, None, None, pin["lengths"].cuda(), pin["refer_lengths"].cuda(), None)
    for _ in range(3): pre.infer(data)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): pre.infer(data)
    e1.record(); torch.cuda.synchronize()
    print(json.dumps({"B": B, "T": T, "S": S, "what": "Pre_model.infer", "ms_per_call": round(e0.elapsed_time(e1) / 10, 3), "launches": pre.launch_count()}), flush=True)
    for kind in ("dpmpp_2m", "unipc_bh2"):
        def run():
            sess = get_session(unet, content, prompt, mask)
            return sess.sample_dpmpp_2m(x, ns, ts) if kind == "dpmpp_2m" else sess.sample_unipc(x, ns, ts)
        for _ in range(3): out = run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2): out = run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 2
        assert torch.isfinite(out).all()
        print(json.dumps({"B": B, "T": T, "S": S, "sampler": kind, "nfe": NFE, "ms_per_run": round(ms, 2), "ms_per_forward": round(ms / NFE, 3),
                          "denoiser_steps_per_s": round(B * NFE / (ms / 1e3), 1)}), flush=True)
