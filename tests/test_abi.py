"""CPU: the C-ABI library loads, exports every symbol include/dppo_b200.h declares, and its
host-only entry points agree with the oracle.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import diffusionpolicyoptimization_b200 as dp
from diffusionpolicyoptimization_b200 import _lib as L
from oracle import dppo_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "dppo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dppo_[a-z0-9_]+)\s*\(", src)))


def test_exports_match_header(lib):
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/dppo_b200.h but not exported"
    assert sorted(L.EXPORTS) == syms, "python binding list and header disagree"


def test_abi_version_and_cfg_layout(lib):
    assert lib.dppo_abi_version() == 1
    assert lib.dppo_cfg_size() == C.sizeof(L.DppoCfg)
    cfg = dp.default_cfg()
    assert (cfg.obs_dim, cfg.action_dim, cfg.horizon_steps, cfg.denoising_steps, cfg.ft_denoising_steps) == (11, 3, 4, 20, 10)
    assert cfg.actor_act == L.ACT_RELU and cfg.critic_act == L.ACT_MISH
    assert abs(cfg.pretrain_weight_decay - 1e-6) < 1e-12 and abs(cfg.weight_decay - 0.004) < 1e-9


@pytest.mark.parametrize("task", ["hopper", "walker2d"])
def test_num_params(lib, task):
    d = O.Dims(**O.TASKS[task])
    cfg = dp.default_cfg()
    cfg.obs_dim, cfg.action_dim = d.obs_dim, d.action_dim
    assert lib.dppo_num_params(C.byref(cfg), L.NET_ACTOR) == d.n_actor()
    assert lib.dppo_num_params(C.byref(cfg), L.NET_ACTOR_FT) == d.n_actor()
    assert lib.dppo_num_params(C.byref(cfg), L.NET_CRITIC) == d.n_critic()


@pytest.mark.parametrize("T", [1, 5, 20, 100])
def test_schedule_matches_oracle(lib, T):
    from diffusionpolicyoptimization_b200.model.diffusion.sampling import ddpm_schedule, SCHEDULE_ROWS
    if T == 1:
        pytest.skip("T=1 divides by zero in the reference's own formula")
    s = ddpm_schedule(T)
    o = O.schedule_table(T)
    assert SCHEDULE_ROWS == O.SCHEDULE_ROWS
    for i, k in enumerate(SCHEDULE_ROWS):
        # identical formula; libm vs numpy transcendental rounding may differ by an ulp or two
        np.testing.assert_allclose(s[k], o[i], rtol=3e-7, atol=1e-37, err_msg=k)


def test_invalid_config_and_no_gpu_fail_loudly(lib):
    import torch
    cfg = dp.default_cfg()
    cfg.ft_denoising_steps = 99
    h = C.c_void_p()
    assert lib.dppo_create(C.byref(cfg), 0, C.byref(h)) != 0
    assert b"invalid configuration" in lib.dppo_last_error()
    if not torch.cuda.is_available():
        cfg = dp.default_cfg()
        assert lib.dppo_create(C.byref(cfg), 0, C.byref(h)) != 0
        assert b"no CPU fallback" in lib.dppo_last_error()
        with pytest.raises(dp.DppoError):
            dp.Engine(cfg)


def test_network_containers_roundtrip():
    net = dp.DiffusionMLP(action_dim=3, horizon_steps=4, cond_dim=11, mlp_dims=[512, 512, 512], activation_type="ReLU",
                          residual_style=True, seed=0)
    assert net.num_params() == 553020
    ws = net.get_weights()
    assert [w.shape for w in ws] == [tuple(s) for s in O.Dims().actor_shapes()]
    flat = net.get_flat_weights()
    net.set_flat_weights(flat * 2)
    np.testing.assert_array_equal(net.get_flat_weights(), flat * 2)
    cr = dp.CriticObs(cond_dim=11, mlp_dims=[256, 256, 256], residual_style=True, seed=0)
    assert cr.num_params() == 134913
    with pytest.raises(ValueError):
        dp.DiffusionMLP(action_dim=3, horizon_steps=4, cond_dim=11, residual_style=False)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "diffusionpolicyoptimization_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("dppo_oracle is test infrastructure", ""), f"{f} mentions the oracle"
