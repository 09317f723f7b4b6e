"""Drop-in check against the real reference tree (only where /root/reference exists: the build
container).  `ns2vc_b200.install()` + the reference's own model.py must construct, expose the same
state_dict contract, and strict-load a reference-shaped checkpoint."""
import json
import os
import sys
from unittest.mock import MagicMock

import pytest
import torch

REF = os.environ.get("NS2VC_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "unet1d")), reason="reference tree not present")


def test_reference_model_py_builds_on_our_unet():
    import ns2vc_b200
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("unet1d", "sampler", "model", "modules", "utils", "operations")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REF)
    try:
        for name in ("matplotlib", "matplotlib.pyplot", "vocos", "accelerate", "librosa", "soundfile", "tensorboardX"):
            sys.modules.setdefault(name, MagicMock())
        ns2vc_b200.install()
        import model as ref_model
        from ns2vc_b200.unet import UNet1DConditionModel
        cfg = json.load(open(os.path.join(REF, "config.json")))
        ref_pre_keys = {k: tuple(v.shape) for k, v in ref_model.Pre_model(cfg).state_dict().items()}   # the reference's own class
        ns2vc_b200.install_pre_model(ref_model)
        ns2 = ref_model.NaturalSpeech2(cfg)
        from ns2vc_b200.pre_model import Pre_model
        assert isinstance(ns2.pre_model, Pre_model)
        ours = {k: tuple(v.shape) for k, v in ns2.pre_model.state_dict().items()}
        assert list(ours.items()) == list(ref_pre_keys.items())          # same keys, order and shapes as the reference's Pre_model
        assert sum(p.numel() for p in ns2.pre_model.parameters()) == 34923404
        unet = ns2.diff_model.unet
        assert isinstance(unet, UNet1DConditionModel)
        assert sum(p.numel() for p in unet.parameters()) == 66076900
        assert unet.latent_channels == cfg["diffusion_encoder"]["in_channels"]
        # a checkpoint written by the reference has exactly these keys under diff_model.unet.
        from ns2vc_b200.arch import ns2vc_denoiser_config, param_shapes
        keys = [k for k in ns2.state_dict() if k.startswith("diff_model.unet.")]
        assert [k[len("diff_model.unet."):] for k in keys] == list(param_shapes(ns2vc_denoiser_config()).keys())
        ns2.load_state_dict(ns2.state_dict(), strict=True)
        from sampler.dpm_solver import DPM_Solver
        from sampler.uni_pc import UniPC
        from ns2vc_b200 import dpm_solver, uni_pc
        assert DPM_Solver is dpm_solver.DPM_Solver and UniPC is uni_pc.UniPC
    finally:
        sys.path.remove(REF)
        for k in list(sys.modules):
            if k.split(".")[0] in ("unet1d", "sampler", "model", "modules", "utils", "operations"):
                del sys.modules[k]
        sys.modules.update(saved)
