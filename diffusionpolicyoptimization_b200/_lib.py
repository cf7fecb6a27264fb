"""ctypes binding of libdppo_b200.so (C ABI declared in include/dppo_b200.h).

The library is the product: there is no Python / CPU fallback.  If the shared object is missing
`load()` raises; if no sm_100 GPU is visible `dppo_create` fails and `check()` raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdppo_b200.so")

NET_ACTOR, NET_ACTOR_FT, NET_CRITIC, NET_ACTOR_EMA = 0, 1, 2, 3
ACT_RELU, ACT_MISH = 0, 1
PREC_FP32, PREC_BF16, PREC_BF16X3 = 0, 1, 2
OPT_PRETRAIN, OPT_FINETUNE = 0, 1

# every symbol include/dppo_b200.h declares (tests check the .so exports all of them)
EXPORTS = (
    "dppo_abi_version", "dppo_last_error", "dppo_cfg_default", "dppo_cfg_size", "dppo_ddpm_schedule", "dppo_num_params",
    "dppo_create", "dppo_destroy", "dppo_set_weights", "dppo_get_weights", "dppo_set_opt_state",
    "dppo_get_opt_state", "dppo_set_ft_denoising_steps", "dppo_set_grad_clip_norm", "dppo_actor_forward", "dppo_value",
    "dppo_sample", "dppo_sample_host", "dppo_set_env_normalization", "dppo_rollout_step", "dppo_logprobs", "dppo_logprobs_subsample", "dppo_ppo_step",
    "dppo_ppo_step_host", "dppo_ppo_step_indexed", "dppo_ppo_step_indexed_host", "dppo_gae", "dppo_pretrain_step", "dppo_ema_update", "dppo_comm_unique_id",
    "dppo_comm_init", "dppo_comm_ipc_export", "dppo_comm_ipc_attach", "dppo_comm_status", "dppo_launch_count", "dppo_tc_launch_count", "dppo_fused_launch_count", "dppo_last_path", "dppo_force_path", "dppo_profile_enable", "dppo_debug_tc_gemm",
    "dppo_profile_read", "dppo_debug_chain_timing", "dppo_debug_mma_probe", "dppo_debug_ffma_peak", "dppo_profile_read_class", "dppo_profile_read_exec", "dppo_debug_split_gemm", "dppo_debug_pair_gemm",
)


class DppoCfg(C.Structure):
    """Mirror of `struct dppo_cfg` (field order and types must match the header)."""
    _fields_ = [
        ("obs_dim", C.c_int32), ("action_dim", C.c_int32), ("horizon_steps", C.c_int32), ("cond_steps", C.c_int32),
        ("denoising_steps", C.c_int32), ("ft_denoising_steps", C.c_int32),
        ("time_dim", C.c_int32),
        ("actor_hidden", C.c_int32), ("critic_hidden", C.c_int32),
        ("actor_act", C.c_int32), ("critic_act", C.c_int32),
        ("precision", C.c_int32),
        ("denoised_clip_value", C.c_float), ("randn_clip_value", C.c_float), ("final_action_clip_value", C.c_float),
        ("min_sampling_denoising_std", C.c_float), ("min_logprob_denoising_std", C.c_float),
        ("gamma_denoising", C.c_float),
        ("clip_ploss_coef", C.c_float), ("clip_ploss_coef_base", C.c_float), ("clip_ploss_coef_rate", C.c_float),
        ("clip_vloss_coef", C.c_float),
        ("norm_adv", C.c_int32), ("reward_horizon", C.c_int32),
        ("vf_coef", C.c_float),
        ("logprob_clip_lo", C.c_float), ("logprob_clip_hi", C.c_float),
        ("adam_beta1", C.c_float), ("adam_beta2", C.c_float), ("adam_eps", C.c_float),
        ("weight_decay", C.c_float), ("pretrain_weight_decay", C.c_float),
    ]


_lib = None


def load():
    """Load the shared library (built in-tree by `__graft_entry__.build()` / `build.build()`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m diffusionpolicyoptimization_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, i32, i64, u64, f32, sz = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float, C.c_size_t
    cfgp = C.POINTER(DppoCfg)
    sig = {
        "dppo_abi_version": (C.c_int, []),
        "dppo_last_error": (C.c_char_p, []),
        "dppo_cfg_default": (None, [cfgp]),
        "dppo_cfg_size": (sz, []),
        "dppo_ddpm_schedule": (C.c_int, [i32, vp]),
        "dppo_num_params": (sz, [cfgp, i32]),
        "dppo_create": (C.c_int, [cfgp, i32, C.POINTER(vp)]),
        "dppo_destroy": (None, [vp]),
        "dppo_set_weights": (C.c_int, [vp, i32, vp, sz, i32, vp]),
        "dppo_get_weights": (C.c_int, [vp, i32, vp, sz, i32, vp]),
        "dppo_set_opt_state": (C.c_int, [vp, i32, vp, vp, sz, i64, i32, vp]),
        "dppo_get_opt_state": (C.c_int, [vp, i32, vp, vp, sz, C.POINTER(i64), i32, vp]),
        "dppo_set_ft_denoising_steps": (C.c_int, [vp, i32]),
        "dppo_set_grad_clip_norm": (C.c_int, [vp, f32]),
        "dppo_actor_forward": (C.c_int, [vp, i32, vp, vp, vp, i32, vp, vp]),
        "dppo_value": (C.c_int, [vp, vp, i32, vp, vp]),
        "dppo_sample": (C.c_int, [vp, vp, i32, i32, i32, f32, u64, u64, i64, vp, vp, vp, vp, vp]),
        "dppo_sample_host": (C.c_int, [vp, vp, i32, i32, i32, f32, u64, u64, i64, vp, vp, vp, vp, vp]),
        "dppo_set_env_normalization": (C.c_int, [vp, vp, vp, vp, vp]),
        "dppo_rollout_step": (C.c_int, [vp, vp, i32, i32, i32, f32, u64, u64, i64, vp, vp, vp, vp, vp, vp, i32, vp]),
        "dppo_logprobs": (C.c_int, [vp, vp, vp, i32, i32, vp, vp]),
        "dppo_logprobs_subsample": (C.c_int, [vp, vp, vp, vp, vp, i32, i32, vp, vp]),
        "dppo_ppo_step": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i64, f32, f32, f32, i32, vp, vp, vp]),
        "dppo_ppo_step_host": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i64, f32, f32, f32, i32, vp, vp]),
        "dppo_ppo_step_indexed": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, i64, vp, i32, i64, f32, f32, f32, i32, vp, vp, vp]),
        "dppo_ppo_step_indexed_host": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, i64, vp, i32, i64, f32, f32, f32, i32, vp, vp]),
        "dppo_gae": (C.c_int, [vp, vp, vp, vp, vp, i32, i32, C.c_double, C.c_double, C.c_double, vp, vp, vp]),
        "dppo_pretrain_step": (C.c_int, [vp, vp, vp, i32, i64, i64, vp, vp, u64, u64, f32, i32, vp, vp, vp]),
        "dppo_ema_update": (C.c_int, [vp, f32, vp]),
        "dppo_comm_unique_id": (C.c_int, [vp]),
        "dppo_comm_init": (C.c_int, [vp, vp, i32, i32]),
        "dppo_comm_ipc_export": (C.c_int, [vp, vp]),
        "dppo_comm_ipc_attach": (C.c_int, [vp, vp, i32, i32]),
        "dppo_comm_status": (C.c_int, [vp]),
        "dppo_launch_count": (i64, [vp]),
        "dppo_tc_launch_count": (i64, [vp]),
        "dppo_fused_launch_count": (i64, [vp]),
        "dppo_last_path": (C.c_int, [vp]),
        "dppo_force_path": (C.c_int, [vp, i32]),
        "dppo_profile_enable": (C.c_int, [vp, i32]),
        "dppo_debug_tc_gemm": (C.c_int, [vp, vp, i32, i64, vp, i64, i32, vp, i32, i64, i32, i32, i32, i32, vp, i32, vp, vp, vp]),
        "dppo_debug_split_gemm": (C.c_int, [vp, vp, i32, i64, vp, i32, i64, i32, i32, i32, i32, i32, vp, vp]),
        "dppo_debug_pair_gemm": (C.c_int, [vp, vp, i64, vp, i32, i64, i32, i32, i32, i32, vp, i32, vp, vp, vp]),
        "dppo_debug_chain_timing": (C.c_int, [vp, i32, vp, C.POINTER(C.c_int)]),
        "dppo_debug_mma_probe": (C.c_int, [vp, i32, i32, i32, i32, i32, vp]),
        "dppo_debug_ffma_peak": (C.c_int, [vp, C.POINTER(C.c_double)]),
        "dppo_profile_read_class": (C.c_int, [vp, i32, C.POINTER(C.c_double), C.POINTER(i64), C.POINTER(C.c_double)]),
        "dppo_profile_read_exec": (C.c_int, [vp, i32, C.POINTER(C.c_double)]),
        "dppo_profile_read": (C.c_int, [vp, C.POINTER(C.c_double), C.POINTER(i64), C.POINTER(C.c_double)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class DppoError(RuntimeError):
    pass


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().dppo_last_error()
        raise DppoError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


def default_cfg() -> DppoCfg:
    cfg = DppoCfg()
    load().dppo_cfg_default(C.byref(cfg))
    return cfg
