"""Generates tests/golden/ref_pretrain_loop.npz: three epochs of the reference's diffusion pre-training loop, produced by the
reference's own code: the `run` loop of agent/pretrain/train_diffusion_agent.py (:56-120), `class EMA`, `reset_parameters` and
`step_ema` of agent/pretrain/train_agent.py (both modules import hydra / wandb / omegaconf at the top, so the pieces are sliced
out of the files and exec'd verbatim), and the model classes (DiffusionModel, DiffusionMLP) imported unmodified - over the TF shim.
Third-party, hence stubbed with restated rules: the Keras AdamW object and `CosineDecayRestarts` (train_agent.py:113-123).

    python tests/golden/make_ref_pretrain_loop.py
"""
import logging
import math
import os
import sys
import textwrap
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("DPPO_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "tf_shim"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

import tensorflow as tf  # noqa: E402  (the shim)
from model.diffusion.diffusion import DiffusionModel  # noqa: E402  (reference)
from model.diffusion.mlp_diffusion import DiffusionMLP  # noqa: E402

from oracle import dppo_oracle as O  # noqa: E402

torch.set_grad_enabled(False)
tf.Module = type("Module", (), {"__init__": lambda self, *a, **k: None})
torch.Tensor.assign = lambda self, value: self.data.copy_(value)       # tf.Variable.assign on the shim's variables


def lines_of(path):
    return open(os.path.join(REF, path)).read().splitlines()


def slice_block(src, first, last_pred):
    a = next(i for i, l in enumerate(src) if l.strip().startswith(first))
    b = next(i for i in range(a + 1, len(src)) if last_pred(src, i))
    return textwrap.dedent("\n".join(src[a:b + 1])), (a + 1, b + 1)


ta = lines_of("agent/pretrain/train_agent.py")
tda = lines_of("agent/pretrain/train_diffusion_agent.py")
EMA_SRC, l1 = slice_block(ta, "class EMA(tf.Module):", lambda s, i: s[i].strip().startswith("return old * self.decay"))
RESET_SRC, l2 = slice_block(ta, "def reset_parameters(self):", lambda s, i: s[i].strip().startswith("self.ema_model.set_weights"))
STEP_SRC, l3 = slice_block(ta, "def step_ema(self):", lambda s, i: s[i].strip().startswith("self.ema.update_model_average"))
LOOP_SRC, l4 = slice_block(tda, "for _ in range(self.n_epochs):", lambda s, i: s[i].strip() == "self.epoch += 1")
print("exec'd reference lines: train_agent.py", l1, l2, l3, "train_diffusion_agent.py", l4)


class AdamWStub:
    def __init__(self, lr_fn, h, wd):
        self.lr_fn, self.h, self.wd, self.iterations, self.m, self.v = lr_fn, h, wd, 0, None, None

    def apply_gradients(self, grads_and_vars):
        grads, vs = zip(*list(grads_and_vars))
        if self.m is None:
            self.m = [torch.zeros_like(p) for p in vs]; self.v = [torch.zeros_like(p) for p in vs]
        lr = self.lr_fn(self.iterations); self.iterations += 1
        O.adamw_keras([p.data for p in vs], [g.detach() for g in grads], self.m, self.v, self.iterations, lr, self.h.beta1, self.h.beta2,
                      self.h.adam_eps, self.wd)


def cosine_decay_restarts(lr0, first, alpha):
    def f(step):
        frac = step / first; frac -= math.floor(frac)
        return lr0 * ((1 - alpha) * 0.5 * (1 + math.cos(math.pi * frac)) + alpha)
    return f


def np_(x):
    return x.detach().numpy().copy()


def main():
    seed, M, B, EPOCHS = 31, 300, 128, 3
    LR0, FIRST, ALPHA, WD, DECAY, START, FREQ = 1e-3, 5, 0.1, 1e-6, 0.9, 2, 1
    o = O.make_oracle("hopper", seed=seed)
    rng = np.random.default_rng(7)
    actions = rng.uniform(-1, 1, size=(M, 4, 3)).astype(np.float32)
    states = rng.uniform(-1, 1, size=(M, 1, 11)).astype(np.float32)
    drawlog = []

    def source(kind, shape, minval=0, maxval=None, dtype=None):
        r = np.random.default_rng(900 + len(drawlog))
        if kind == "uniform":
            v = torch.from_numpy(r.integers(minval, maxval, size=tuple(shape))).to(dtype)
        else:
            v = torch.from_numpy(r.standard_normal(size=tuple(shape)).astype(np.float32))
        drawlog.append((kind, v.clone()))
        return v
    tf.random.source = source

    def make_model():
        net = DiffusionMLP(action_dim=3, horizon_steps=4, cond_dim=11, time_dim=16, mlp_dims=[512] * 3, activation_type="ReLU", residual_style=True)
        mdl = DiffusionModel(network=net, horizon_steps=4, obs_dim=11, action_dim=3, denoising_steps=20, device="cpu")
        mdl.c_loss(actions=torch.zeros(2, 4, 3), conditions={"state": torch.zeros(2, 1, 11)})      # builds the layers (as :46-49 does)
        return mdl
    model, ema_model = make_model(), make_model()
    drawlog.clear()
    model.network.set_weights([np_(p) for p in o.actor])
    ns = dict(tf=tf, np=np, log=logging.getLogger("ref"))
    exec(EMA_SRC, ns)
    losses = []
    wandb = types.SimpleNamespace(log=lambda d, step=None, commit=None: losses.append(float(d["loss - train"])) if "loss - train" in d else None)
    self = types.SimpleNamespace(
        n_epochs=EPOCHS, epoch=1, model=model, ema_model=ema_model, ema=ns["EMA"](DECAY), optimizer=AdamWStub(cosine_decay_restarts(LR0, FIRST, ALPHA), o.h, WD),
        dataloader_train=[{"actions": torch.from_numpy(actions[r0:r0 + B]), "conditions": {"state": torch.from_numpy(states[r0:r0 + B])}}
                          for r0 in range(0, M, B)],
        dataloader_val=None, val_freq=1, update_ema_freq=FREQ, epoch_start_ema=START, save_model_freq=10 ** 9, save_model=lambda epoch: None,
        log_freq=1, use_wandb=True)
    exec(RESET_SRC, ns); exec(STEP_SRC, ns)
    self.reset_parameters = types.MethodType(ns["reset_parameters"], self); self.step_ema = types.MethodType(ns["step_ema"], self)
    self.reset_parameters()                                       # PreTrainAgent.__init__ ends with it (:125)
    ns.update(self=self, batch_to_device=lambda b, device=None: b, timer=lambda: 0.0, wandb=wandb)
    exec(LOOP_SRC, ns)
    assert len(drawlog) == 2 * 3 * EPOCHS and self.epoch == EPOCHS + 1
    out = dict(seed=np.array([seed]), cfg=np.array([M, B, EPOCHS, FIRST, START, FREQ]), hyper=np.array([LR0, ALPHA, WD, DECAY]),
               actions=actions, states=states, losses=np.array(losses),
               t=np.concatenate([np_(v).astype(np.int32) for k, v in drawlog if k == "uniform"]),
               noise=np.concatenate([np_(v).reshape(len(v), -1) for k, v in drawlog if k == "normal"]),
               net_fp=np.concatenate([np_(v).reshape(-1) for v in model.network.variables])[::97].copy(),
               ema_fp=np.concatenate([np_(v).reshape(-1) for v in ema_model.network.variables])[::97].copy(),
               opt_iterations=np.array([self.optimizer.iterations]))
    np.savez_compressed(os.path.join(HERE, "ref_pretrain_loop.npz"), **out)
    print("losses", losses, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
