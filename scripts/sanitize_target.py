"""Small full-architecture run for compute-sanitizer: one prepare_cond + 2 fused sampler steps at a ragged shape."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ns2vc_b200 import api
from ns2vc_b200.arch import ns2vc_denoiser_config
from ns2vc_b200.fused import DenoiserSession
from ns2vc_b200.synth import make_inputs, make_state_dict
from ns2vc_b200.unet import UNet1DConditionModel
os.environ.setdefault("NS2VC_GRAPH", "0")
B, T, S = 2, 200, 40
cfg = ns2vc_denoiser_config()
unet = UNet1DConditionModel(in_channels=356, out_channels=100, block_out_channels=(128, 256, 384, 512), norm_num_groups=8,
                            cross_attention_dim=256, attention_head_dim=8, addition_embed_type="text", resnet_time_scale_shift="scale_shift")
unet.load_state_dict(make_state_dict(cfg, 0)); unet = unet.cuda().eval()
inp = make_inputs(B, T, S, ragged=True, seed=3)
sess = DenoiserSession(unet, inp["content"].permute(1, 2, 0).contiguous().cuda(), inp["prompt"].permute(1, 0, 2).contiguous().cuda(),
                       api.sequence_mask(inp["refer_lengths"].cuda(), S))
out = sess.sample_dpmpp_2m(inp["x"].cuda(), api.default_schedule(), torch.linspace(1.0, 1e-3, 3))
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
