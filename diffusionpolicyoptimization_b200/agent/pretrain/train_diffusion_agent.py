"""Diffusion pre-training loop over the B200 library with the dataset resident in HBM.

Mirrors `TrainDiffusionAgent.run` (agent/pretrain/train_diffusion_agent.py:46-110) and `PreTrainAgent`
(agent/pretrain/train_agent.py:40-162): sequential batches of the (un-shuffled, cached) dataset, eps-MSE loss with a
fresh t ~ U{0..T-1} and eps ~ N(0,1) per sample (diffusion.py:179-194), Keras-3 AdamW under
`CosineDecayRestarts(t_mul=1, m_mul=1)`, EMA of the network every `update_ema_freq` epochs (a plain copy before
`epoch_start_ema`).  Hydra, the tf.data pipeline, WandB and .h5 checkpoint files stay on the reference side.

The whole dataset (actions [M,Ta,Da], states [M,To,Do]) is uploaded once; a batch is a contiguous slice of the device
buffers, so an epoch moves no data over the host link and the only host sync is the epoch's mean loss.
"""
import logging
import math

import numpy as np
import torch

from ... import _lib as L

log = logging.getLogger(__name__)


class CosineDecayRestarts:
    """keras.optimizers.schedules.CosineDecayRestarts restated for t_mul = m_mul = 1 (train_agent.py:113-119):
    lr(step) = lr0 * ((1 - alpha) * 0.5 * (1 + cos(pi * frac(step / first_decay_steps))) + alpha)."""

    def __init__(self, initial_learning_rate, first_decay_steps, alpha=0.0):
        self.initial_learning_rate, self.first_decay_steps, self.alpha = initial_learning_rate, first_decay_steps, alpha

    def __call__(self, step):
        frac = step / self.first_decay_steps
        frac -= math.floor(frac)
        return self.initial_learning_rate * ((1 - self.alpha) * 0.5 * (1 + math.cos(math.pi * frac)) + self.alpha)


class TrainDiffusionAgent:
    def __init__(self, model, actions, states, *, n_epochs, batch_size, learning_rate=1e-3, lr_first_cycle_steps=3000,
                 lr_min=1e-4, ema_decay=0.995, epoch_start_ema=10, update_ema_freq=5, log_freq=1, draws_fn=None):
        """`model`: DiffusionModel of this package (its cfg carries pretrain_weight_decay).  `draws_fn(epoch, batch, n) ->
        (t int32 [n], noise [n,A])` replaces the library's Philox stream (tests)."""
        self.model, self.engine = model, model.engine
        dev = self.engine.dev
        M = actions.shape[0]
        self.actions = torch.as_tensor(np.ascontiguousarray(actions, np.float32)).reshape(M, -1).to(dev)
        self.states = torch.as_tensor(np.ascontiguousarray(states, np.float32)).reshape(M, -1).to(dev)
        self.n_epochs, self.batch_size = int(n_epochs), int(batch_size)
        self.lr_scheduler = CosineDecayRestarts(learning_rate, lr_first_cycle_steps, alpha=lr_min / learning_rate)
        self.ema_decay, self.epoch_start_ema, self.update_ema_freq = ema_decay, epoch_start_ema, update_ema_freq
        self.log_freq, self.draws_fn = log_freq, draws_fn
        self.epoch, self.opt_iterations = 1, 0
        self.loss_history = []
        self.reset_parameters()                                      # train_agent.py:125

    def reset_parameters(self):
        """EMA <- model (train_agent.py:130-132), device to device."""
        w = torch.empty(self.engine.n_actor, device=self.engine.dev)
        st = self.engine._stream()
        L.check(self.engine.lib.dppo_get_weights(self.engine.h, L.NET_ACTOR, w.data_ptr(), w.numel(), 1, st), "dppo_get_weights")
        L.check(self.engine.lib.dppo_set_weights(self.engine.h, L.NET_ACTOR_EMA, w.data_ptr(), w.numel(), 1, st), "dppo_set_weights")

    def step_ema(self):
        """train_agent.py:134-139."""
        if self.epoch < self.epoch_start_ema:
            self.reset_parameters()
            return
        self.engine.ema_update(self.ema_decay)

    def run_epoch(self):
        M, B = self.actions.shape[0], self.batch_size
        losses = []
        for n_batch, r0 in enumerate(range(0, M, B)):                # tf.data .batch(): the last batch may be short
            a, s = self.actions[r0: r0 + B], self.states[r0: r0 + B]
            kw = {}
            if self.draws_fn is not None:
                kw["t"], kw["noise"] = self.draws_fn(self.epoch, n_batch, a.shape[0])
            losses.append(self.engine.pretrain_step(a, s, lr=self.lr_scheduler(self.opt_iterations), apply=True,
                                                    seed=self.model.seed, offset=self.model._next_offset(), **kw))
            self.opt_iterations += 1
        loss_train = float(torch.cat(losses).mean())                 # np.mean(loss_train_epoch), train_diffusion_agent.py:76
        if self.epoch % self.update_ema_freq == 0:                   # :92-93
            self.step_ema()
        if self.epoch % self.log_freq == 0:
            log.info("%d: train loss %8.4f", self.epoch, loss_train)
        self.loss_history.append(loss_train)
        self.epoch += 1
        return loss_train

    def run(self):
        for _ in range(self.n_epochs):
            self.run_epoch()
        return self.loss_history

    # ---------------------------------------------------------------- checkpoints (agent/pretrain/train_agent.py:150-162)
    def _ema_weights(self):
        flat = self.engine.get_weights(L.NET_ACTOR_EMA)
        out, off = [], 0
        for shp in self.model.network.shapes:
            n = int(np.prod(shp)); out.append(flat[off:off + n].reshape(shp)); off += n
        return out

    def save_model(self, checkpoint_dir, epoch=None):
        """`state_{epoch}.weights.h5` (network) and `ema_state_{epoch}.weights.h5` (EMA copy) in the reference's Keras layout, plus
        `state_{epoch}.opt.npz` with the AdamW moments / step / schedule position the reference does not save."""
        import os
        from ...util.keras_h5 import save_keras_weights_h5
        epoch = self.epoch if epoch is None else epoch
        os.makedirs(checkpoint_dir, exist_ok=True)
        path = os.path.join(checkpoint_dir, f"state_{epoch}.weights.h5")
        net = self.model.network
        net.save_weights(path)
        save_keras_weights_h5(path.replace("state_", "ema_state_"), self._ema_weights(), net.keras_variable_paths())
        m, v, step = self.engine.get_opt_state(L.OPT_PRETRAIN)
        np.savez(os.path.join(checkpoint_dir, f"state_{epoch}.opt.npz"), m=m, v=v, step=step, epoch=self.epoch, opt_iterations=self.opt_iterations)
        log.info("Saved model to %s", path)
        return path

    def load_model(self, checkpoint_dir, epoch, with_optimizer=True):
        import os
        from ...util.keras_h5 import load_keras_weights_h5
        path = os.path.join(checkpoint_dir, f"state_{epoch}.weights.h5")
        net = self.model.network
        net.load_weights(path)
        ema = load_keras_weights_h5(path.replace("state_", "ema_state_"), net.keras_variable_paths(), net.shapes)
        self.engine.set_weights(L.NET_ACTOR_EMA, np.concatenate([w.reshape(-1) for w in ema]))
        side = os.path.join(checkpoint_dir, f"state_{epoch}.opt.npz")
        if with_optimizer and os.path.exists(side):
            z = np.load(side)
            self.engine.set_opt_state(L.OPT_PRETRAIN, z["m"], z["v"], int(z["step"]))
            self.epoch, self.opt_iterations = int(z["epoch"]), int(z["opt_iterations"])
