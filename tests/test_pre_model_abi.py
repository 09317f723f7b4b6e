"""Condition encoders (`Pre_model`), host side: the Python drop-in and the C registry expose the
reference's state_dict contract (fixtures written by the unmodified reference), errors are loud, no CPU path.  No GPU compute."""
import ctypes as C

import pytest
import torch

from ns2vc_b200 import _lib
from ns2vc_b200.pre_model import Pre_model, pre_param_shapes


def _registry(cfg):
    L = _lib.lib()
    m = Pre_model(cfg)
    h = C.c_void_p()
    ccfg = m._c_cfg()
    _lib.check(L.ns2vc_pre_create(C.byref(ccfg), C.byref(h)))
    try:
        got = {}
        for i in range(L.ns2vc_pre_num_weights(h)):
            name, shp, nd = C.c_char_p(), (C.c_int64 * 4)(), C.c_int()
            _lib.check(L.ns2vc_pre_weight_info(h, i, C.byref(name), shp, C.byref(nd)))
            got[name.value.decode()] = tuple(shp[k] for k in range(nd.value))
    finally:
        L.ns2vc_pre_destroy(h)
    return m, got


@pytest.mark.parametrize("name", ["tiny", "full"])
def test_state_dict_contract_matches_reference_fixture(gold, name):
    g = gold(f"pre_model_{name}.pt")
    m, reg = _registry(g["cfg"])
    shapes = pre_param_shapes(g["cfg"])
    assert list(shapes.items()) == [(k, tuple(v)) for k, v in g["shapes"].items()]      # same keys, same ORDER, same shapes
    assert reg == shapes
    sd = m.state_dict()
    assert list(sd.keys()) == list(shapes.keys()) and all(tuple(v.shape) == shapes[k] for k, v in sd.items())
    assert sum(p.numel() for p in m.parameters()) == g["n_params"]
    if name == "full":
        assert g["n_params"] == 34923404                      # demo.ipynb:447
    m.load_state_dict({k: torch.zeros(v) for k, v in shapes.items()}, strict=True)


def test_errors_are_loud():
    L = _lib.lib()
    cfg = {"phoneme_encoder": dict(in_channels=64, hidden_channels=64, out_channels=40, n_layers=1),
           "prompt_encoder": dict(in_channels=100, hidden_channels=64, out_channels=40, n_layers=1)}
    m = Pre_model(cfg)
    h = C.c_void_p()
    ccfg = m._c_cfg()
    _lib.check(L.ns2vc_pre_create(C.byref(ccfg), C.byref(h)))
    try:
        n = C.c_size_t()
        assert L.ns2vc_pre_workspace_bytes(h, 1, 16, 8, C.byref(n)) != 0 and b"finalize" in L.ns2vc_last_error()
        shape = (C.c_int64 * 1)(3)
        assert L.ns2vc_pre_load_weight(h, b"not.a.key", 1, shape, 1, None) != 0 and b"Unexpected key" in L.ns2vc_last_error()
    finally:
        L.ns2vc_pre_destroy(h)
    bad = _lib.PreCfg()
    for f, _ in _lib.PreCfg._fields_:
        setattr(bad, f, getattr(ccfg, f))
    bad.phone_in = 32                                          # content channels != hidden: spk_proj(g) could not be added (model.py:130)
    assert L.ns2vc_pre_create(C.byref(bad), C.byref(h)) != 0 and b"spk_proj" in L.ns2vc_last_error()
    with pytest.raises(NotImplementedError):
        Pre_model({"phoneme_encoder": dict(last_ln=False), "prompt_encoder": {}})


def test_no_cpu_path():
    cfg = {"phoneme_encoder": dict(in_channels=64, hidden_channels=64, out_channels=40, n_layers=1),
           "prompt_encoder": dict(in_channels=100, hidden_channels=64, out_channels=40, n_layers=1)}
    m = Pre_model(cfg).eval()
    data = (torch.zeros(1, 64, 9), torch.zeros(1, 100, 5), None, None, None, torch.tensor([9]), torch.tensor([5]), None)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.infer(data)


@pytest.mark.parametrize("hidden,out,lp,lr", [(64, 40, 0, 1), (128, 256, 3, 2), (256, 256, 6, 6), (512, 512, 1, 1)])
def test_registry_matches_python_over_configurations(hidden, out, lp, lr):
    cfg = {"phoneme_encoder": dict(in_channels=hidden, hidden_channels=hidden, out_channels=out, n_layers=lp),
           "prompt_encoder": dict(in_channels=100, hidden_channels=hidden, out_channels=out, n_layers=lr)}
    m, reg = _registry(cfg)
    assert reg == pre_param_shapes(cfg)
    assert sum(p.numel() for p in m.parameters()) == sum(int(torch.tensor(s).prod()) for s in reg.values())


def test_bad_configurations_rejected():
    L = _lib.lib()
    base = dict(phone_in=64, phone_hidden=64, phone_out=40, phone_layers=1, prompt_in=100, prompt_hidden=64, prompt_out=40, prompt_layers=1,
                ref_dim=100, ref_heads=1, n_heads=8, ffn_kernel=9)
    for bad, needle in ((dict(ffn_kernel=8), b"ffn_kernel"), (dict(ffn_kernel=11), b"ffn_kernel"), (dict(phone_hidden=60, phone_in=60), b"hidden width"),
                        (dict(prompt_in=80), b"ref_dim"), (dict(n_heads=0), b"head")):
        c = _lib.PreCfg(**{**base, **bad})
        h = C.c_void_p()
        assert L.ns2vc_pre_create(C.byref(c), C.byref(h)) != 0 and needle in L.ns2vc_last_error(), (bad, L.ns2vc_last_error())
