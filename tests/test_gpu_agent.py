"""GPU: the DPPO fine-tuning loop with the rollout resident in HBM (agent/finetune/train_ppo_diffusion_agent.py of this
package) against the CPU restatement of the reference's loop (oracle/dppo_loop.py) on a deterministic toy environment,
with the Gaussian draws and the minibatch permutations replayed on both sides."""
import numpy as np
import pytest
import torch

import diffusionpolicyoptimization_b200 as dp
from diffusionpolicyoptimization_b200 import _lib as L
from diffusionpolicyoptimization_b200.agent.finetune.train_ppo_diffusion_agent import TrainPPODiffusionAgent
from helpers import rel_err, max_abs
from oracle import dppo_loop as OL
from oracle import dppo_oracle as O
from toy_env import NormalizingVecWrapper, RawToyVecEnv, ToyVecEnv

pytestmark = pytest.mark.gpu

E, S, ACT_STEPS, LR = 8, 6, 4, 1e-4


def make_model(o, precision="fp32", device="cuda:0", **kw):
    d = o.d
    actor = dp.DiffusionMLP(action_dim=d.action_dim, horizon_steps=d.horizon_steps, cond_dim=d.obs_dim, time_dim=16,
                            mlp_dims=[512, 512, 512], activation_type="ReLU", residual_style=True)
    critic = dp.CriticObs(cond_dim=d.obs_dim, mlp_dims=[256, 256, 256], activation_type="Mish", residual_style=True)
    model = dp.PPODiffusion(gamma_denoising=0.99, clip_ploss_coef=0.01, clip_ploss_coef_base=0.01, clip_ploss_coef_rate=3,
                            randn_clip_value=3, min_sampling_denoising_std=0.1, min_logprob_denoising_std=0.1,
                            actor=actor, critic=critic, ft_denoising_steps=d.ft_denoising_steps, horizon_steps=d.horizon_steps,
                            obs_dim=d.obs_dim, action_dim=d.action_dim, denoising_steps=d.denoising_steps, device=device,
                            precision=precision, **kw)
    model.actor.set_flat_weights(O.flatten_params(o.actor))
    model.actor_ft.set_flat_weights(O.flatten_params(o.actor_ft))
    model.critic.set_flat_weights(O.flatten_params(o.critic))
    return model


def noise_fn(itr, step, B, A=12, T=20):
    rng = np.random.default_rng(1000 + 97 * itr + step)
    return rng.standard_normal((B, A)).astype(np.float32), rng.standard_normal((T, B, A)).astype(np.float32)


def shuffle_fn(itr, epoch, total):
    return np.random.default_rng(5000 + 13 * itr + epoch).permutation(total)


def test_two_iterations_match_the_reference_loop():
    o = O.make_oracle("hopper", seed=21)
    model = make_model(o)
    kw = dict(n_steps=S, act_steps=ACT_STEPS, batch_size=160, update_epochs=2, gamma=0.99, gae_lambda=0.95, target_kl=1)
    agent = TrainPPODiffusionAgent(model, ToyVecEnv(E, 11, 3, seed=3), n_envs=E, n_train_itr=2, actor_lr=LR, force_train=True,
                                   reset_at_iteration=False, reward_scale_running=True, noise_fn=noise_fn,
                                   shuffle_fn=shuffle_fn, **kw)
    venv = ToyVecEnv(E, 11, 3, seed=3)
    scaler = OL.RewardScaler(E)
    params = o.actor_ft + o.critic
    opt = dict(m=[torch.zeros_like(p) for p in params], v=[torch.zeros_like(p) for p in params], step=0)
    prev_obs, firsts0 = venv.reset_arg(), 1
    for itr in range(2):
        want, prev_obs, done = OL.ppo_iteration(o, opt, venv, itr, prev_obs, lr=LR, reward_scaler=scaler, reward_scale_const=1.0,
                                                noise_fn=noise_fn, shuffle_fn=shuffle_fn, firsts0=firsts0, **kw)
        firsts0 = done
        got = agent.run_iteration()
        K = o.d.ft_denoising_steps
        assert rel_err(agent.chains_trajs.reshape(S * E, K + 1, -1), want["chains_k"].reshape(S * E, K + 1, -1)) < 5e-4
        assert got["n_updates"] == want["n_updates"] == 6
        for k in ("pg_loss", "v_loss", "approx_kl", "ratio", "clipfrac", "explained_var"):
            assert abs(got[k] - want[k]) < 5e-3 * max(1.0, abs(want[k])) + 2e-5, (itr, k, got[k], want[k])
    assert agent.opt_iterations == opt["step"] == 12
    got_w = np.concatenate([model.engine.get_weights(L.NET_ACTOR_FT), model.engine.get_weights(L.NET_CRITIC)])
    want_w = O.flatten_params(params)
    moved = np.abs(want_w - np.concatenate([O.flatten_params(O.make_oracle("hopper", seed=21).actor_ft),
                                            O.flatten_params(O.make_oracle("hopper", seed=21).critic)]))
    assert moved.mean() > 2 * LR                      # 12 Adam steps really happened
    frac_bad = np.mean(np.abs(got_w - want_w) > 0.25 * LR)
    assert frac_bad < 1e-2, frac_bad
    np.testing.assert_array_equal(model.engine.get_weights(L.NET_ACTOR), O.flatten_params(o.actor))   # base net frozen


def test_fused_env_normalisation_equals_the_reference_wrapper():
    """the agent around a RAW env with `normalization=` (normalize_obs / slice / unnormalize_action on the device,
    pinned exchange) must reproduce, bit for bit, the agent around the same env behind the reference's normalising wrapper
    (env/gym_utils/wrapper/mujoco_locomotion_lowdim.py:57-62 restated in tests/toy_env.py)."""
    rng = np.random.default_rng(12)
    norm = {"obs_min": rng.uniform(-3, -1, 11).astype(np.float32), "obs_max": rng.uniform(1, 4, 11).astype(np.float32),
            "action_min": rng.uniform(-1.2, -0.8, 3).astype(np.float32), "action_max": rng.uniform(0.8, 1.2, 3).astype(np.float32)}
    res = []
    for fused in (False, True):
        o = O.make_oracle("hopper", seed=23)
        model = make_model(o)
        raw = RawToyVecEnv(E, 11, 3, norm, seed=3)
        venv = raw if fused else NormalizingVecWrapper(raw, norm)
        agent = TrainPPODiffusionAgent(model, venv, n_envs=E, n_steps=S, act_steps=ACT_STEPS, n_train_itr=1, batch_size=160, update_epochs=1,
                                       actor_lr=LR, force_train=True, noise_fn=noise_fn, shuffle_fn=shuffle_fn,
                                       normalization=norm if fused else None)
        n0 = model.engine.launch_count()
        r = agent.run_iteration()
        res.append((agent.obs_trajs.cpu().numpy().copy(), agent.chains_trajs.cpu().numpy().copy(), model.engine.get_weights(L.NET_ACTOR_FT).copy(),
                    r["pg_loss"], r["avg_episode_reward"]))
        model.engine.close()
    for a, b in zip(res[0], res[1]):
        np.testing.assert_array_equal(a, b)
    assert np.abs(res[0][0]).max() <= 1.0 + 1e-5                     # the rollout buffer holds normalised observations


def test_rollout_step_is_two_launches_on_the_cluster_path():
    o = O.make_oracle("hopper", seed=24)
    model = make_model(o)
    e = model.engine
    rng = np.random.default_rng(1)
    norm = {"obs_min": -np.ones(11, np.float32) * 2, "obs_max": np.ones(11, np.float32) * 3, "action_min": -np.ones(3, np.float32), "action_max": np.ones(3, np.float32) * 2}
    e.set_env_normalization(**norm)
    B, act_steps = 40, 4
    raw_obs = torch.from_numpy(rng.uniform(-2, 3, (B, 11))).pin_memory()
    raw_act = torch.zeros(B, act_steps * 3).pin_memory()
    obs_out = torch.zeros(B, 11, device="cuda"); act = torch.zeros(B, 12, device="cuda"); ch = torch.zeros(B, 11, 12, device="cuda")
    e.rollout_step(raw_obs, obs_out, act, ch, raw_act, act_steps, seed=5, offset=9)     # first call after new weights: + the two folded-output-layer tables
    n0 = e.launch_count()
    e.rollout_step(raw_obs, obs_out, act, ch, raw_act, act_steps, seed=5, offset=9)
    torch.cuda.synchronize()
    assert e.launch_count() - n0 == 2 and e.last_path() == 1         # normalise kernel + the persistent sampler (un-normalising epilogue)
    want_obs = (2 * ((raw_obs.numpy() - norm["obs_min"]) / (norm["obs_max"] - norm["obs_min"] + 1e-6) - 0.5)).astype(np.float32)
    np.testing.assert_array_equal(obs_out.cpu().numpy(), want_obs)
    a = act.cpu().numpy().reshape(B, 4, 3)[:, :act_steps]
    want_act = ((a + 1) / 2) * (norm["action_max"] - norm["action_min"]) + norm["action_min"]
    np.testing.assert_array_equal(raw_act.numpy().reshape(B, act_steps, 3), want_act.astype(np.float32))
    a2, _ = e.sample(obs_out, seed=5, offset=9)                      # same chain as the plain sampler on the normalised observations
    assert torch.equal(a2, act)
    e.close()


def test_checkpoints_in_the_reference_layout_round_trip(tmp_path):
    """agent/finetune/train_agent.py:127-142: `state_{itr}.weights.h5` (Keras layout, util/keras_h5.py) + the optimizer side-car."""
    from diffusionpolicyoptimization_b200.util.keras_h5 import H5Reader
    o = O.make_oracle("hopper", seed=25)
    model = make_model(o)
    agent = TrainPPODiffusionAgent(model, ToyVecEnv(E, 11, 3, seed=3), n_envs=E, n_steps=S, act_steps=ACT_STEPS, n_train_itr=1, batch_size=160,
                                   update_epochs=1, actor_lr=LR, force_train=True)
    agent.run_iteration()
    path = agent.save_model(str(tmp_path))
    names = set(H5Reader(path).datasets)
    assert "actor_ft/mlp_mean/input_layer/vars/0" in names and "critic/Q1/output_layer/vars/1" in names and len(names) == 12 + 12 + 8
    w = [model.engine.get_weights(n).copy() for n in (L.NET_ACTOR, L.NET_ACTOR_FT, L.NET_CRITIC)]
    mo, vo, so = model.engine.get_opt_state(L.OPT_FINETUNE)
    model2 = make_model(O.make_oracle("hopper", seed=26))
    agent2 = TrainPPODiffusionAgent(model2, ToyVecEnv(E, 11, 3, seed=3), n_envs=E, n_steps=S, act_steps=ACT_STEPS, n_train_itr=1, batch_size=160,
                                    update_epochs=1, actor_lr=LR, force_train=True)
    agent2.load(str(tmp_path), 1)
    for n, ww in zip((L.NET_ACTOR, L.NET_ACTOR_FT, L.NET_CRITIC), w):
        np.testing.assert_array_equal(model2.engine.get_weights(n), ww)
    m2, v2, s2 = model2.engine.get_opt_state(L.OPT_FINETUNE)
    assert s2 == so and np.array_equal(m2, mo) and np.array_equal(v2, vo) and agent2.itr == 1 and agent2.opt_iterations == agent.opt_iterations
    # the restored model behaves identically
    obs = torch.rand(64, 11, device="cuda") * 2 - 1
    a1, _ = model.engine.sample(obs, seed=1, offset=3); a2, _ = model2.engine.sample(obs, seed=1, offset=3)
    assert torch.equal(a1, a2)


def test_eval_iteration_does_not_train_and_tensor_mode_runs():
    """itr 0 is an eval iteration (val_freq, :70): deterministic sampling, no update.  Then a bf16 tensor-mode training
    iteration on the library's own Philox stream (>= 2048 rows per minibatch)."""
    o = O.make_oracle("hopper", seed=22)
    model = make_model(o, precision="bf16")
    n_envs, n_steps = 64, 8
    agent = TrainPPODiffusionAgent(model, ToyVecEnv(n_envs, 11, 3, max_episode_steps=3, seed=4), n_envs=n_envs, n_steps=n_steps,
                                   act_steps=ACT_STEPS, n_train_itr=3, batch_size=2560, update_epochs=3, actor_lr=3e-4, val_freq=2)
    w0 = model.engine.get_weights(L.NET_ACTOR_FT).copy()
    r0 = agent.run_iteration()
    assert r0["eval_mode"] and "pg_loss" not in r0 and r0["num_episode_finished"] > 0 and r0["step"] == 0
    np.testing.assert_array_equal(model.engine.get_weights(L.NET_ACTOR_FT), w0)
    n0 = model.engine.tc_launch_count() + model.engine.fused_launch_count()
    r1 = agent.run_iteration()
    assert not r1["eval_mode"] and r1["n_updates"] == 6 and np.isfinite(r1["loss"]) and r1["step"] == n_envs * ACT_STEPS * n_steps
    assert model.engine.tc_launch_count() + model.engine.fused_launch_count() > n0       # the tensor path ran
    assert np.abs(model.engine.get_weights(L.NET_ACTOR_FT) - w0).max() > 1e-4
    assert 0.0 <= r1["clipfrac"] <= 1.0 and abs(r1["ratio"] - 1.0) < 0.2


def test_pretrain_loop_matches_the_reference_loop():
    """agent/pretrain/train_diffusion_agent.py:56-93 + train_agent.py:113-139: sequential batches (short tail), AdamW under
    CosineDecayRestarts, EMA copy before epoch_start_ema and decay after — replayed with the same (t, eps) draws."""
    from diffusionpolicyoptimization_b200.agent.pretrain.train_diffusion_agent import CosineDecayRestarts, TrainDiffusionAgent
    o = O.make_oracle("hopper", seed=31)
    d = o.d
    actor = dp.DiffusionMLP(action_dim=3, horizon_steps=4, cond_dim=11, time_dim=16, mlp_dims=[512, 512, 512],
                            activation_type="ReLU", residual_style=True)
    model = dp.DiffusionModel(network=actor, horizon_steps=4, obs_dim=11, action_dim=3, denoising_steps=20, device="cuda:0",
                              precision="fp32")
    model.network.set_flat_weights(O.flatten_params(o.actor))
    rng = np.random.default_rng(7)
    M, B, EPOCHS = 300, 128, 3
    actions = rng.uniform(-1, 1, size=(M, 4, 3)).astype(np.float32)
    states = rng.uniform(-1, 1, size=(M, 1, 11)).astype(np.float32)

    def draws_fn(epoch, batch, n):
        r = np.random.default_rng(900 + 31 * epoch + batch)
        return r.integers(0, 20, size=n).astype(np.int32), r.standard_normal((n, 12)).astype(np.float32)

    agent = TrainDiffusionAgent(model, actions, states, n_epochs=EPOCHS, batch_size=B, learning_rate=1e-3, lr_first_cycle_steps=5,
                                lr_min=1e-4, ema_decay=0.9, epoch_start_ema=2, update_ema_freq=1, draws_fn=draws_fn)
    got_losses = agent.run()

    sched = CosineDecayRestarts(1e-3, 5, alpha=0.1)
    assert abs(sched(0) - 1e-3) < 1e-12 and abs(sched(5) - 1e-3) < 1e-12 and sched(4) < sched(1)     # restart every 5 steps
    net = [p.clone() for p in o.actor]
    ema = [p.clone() for p in net]
    m = [torch.zeros_like(p) for p in net]; v = [torch.zeros_like(p) for p in net]
    it, want_losses = 0, []
    for epoch in range(1, EPOCHS + 1):
        ep = []
        for nb, r0 in enumerate(range(0, M, B)):
            t, eps = draws_fn(epoch, nb, min(B, M - r0))
            oo = O.Oracle(o.d, o.h, net, o.actor_ft, o.critic)
            loss, g = oo.pretrain_grads(torch.from_numpy(actions[r0:r0 + B]), torch.from_numpy(states[r0:r0 + B]),
                                        torch.from_numpy(t).long(), torch.from_numpy(eps).reshape(-1, 4, 3))
            O.adamw_keras(net, g, m, v, it + 1, sched(it), o.h.beta1, o.h.beta2, o.h.adam_eps, 1e-6)
            it += 1
            ep.append(float(loss))
        want_losses.append(float(np.mean(ep)))
        if epoch < 2:
            ema = [p.clone() for p in net]
        else:
            O.ema_update(ema, net, 0.9)
    np.testing.assert_allclose(got_losses, want_losses, rtol=2e-3)
    assert agent.opt_iterations == it == 9
    got_w, want_w = model.engine.get_weights(L.NET_ACTOR), O.flatten_params(net)
    assert np.mean(np.abs(got_w - want_w) > 0.25 * 1e-3) < 1e-2
    got_e, want_e = model.engine.get_weights(L.NET_ACTOR_EMA), O.flatten_params(ema)
    assert np.mean(np.abs(got_e - want_e) > 0.25 * 1e-3) < 1e-2
    assert np.abs(got_e - got_w).max() > 1e-5                      # EMA really lags the model


def test_two_iterations_match_the_reference_fixture():
    """The CUDA agent against tests/golden/ref_loop.npz - two training iterations produced by the reference's own rollout and
    update blocks, model classes and reward scaler (exec'd / imported verbatim over the TF shim, tests/golden/make_ref_loop.py),
    on the same toy env with the same Gaussian draws and permutations."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_loop.npz"))
    n_envs, n_steps, act, batch, epochs = (int(v) for v in z["cfg"]); lr = float(z["lr"][0])
    o = O.make_oracle("hopper", seed=int(z["seed"][0]))
    model = make_model(o)
    agent = TrainPPODiffusionAgent(
        model, ToyVecEnv(n_envs, 11, 3, seed=3), n_envs=n_envs, n_steps=n_steps, act_steps=act, n_train_itr=2, batch_size=batch,
        update_epochs=epochs, gamma=0.99, gae_lambda=0.95, target_kl=1, actor_lr=lr, force_train=True, reset_at_iteration=False,
        reward_scale_running=True,
        noise_fn=lambda itr, s, B: (z[f"it{itr}_x_T"][s].reshape(B, -1), z[f"it{itr}_noise"][s].reshape(20, B, -1)),
        shuffle_fn=lambda itr, ep, total: z[f"it{itr}_perms"][ep])
    K = o.d.ft_denoising_steps
    for itr in range(2):
        k = f"it{itr}_"
        got = agent.run_iteration()
        assert rel_err(agent.chains_trajs.reshape(n_steps * n_envs, K + 1, -1), z[k + "chains"].reshape(n_steps * n_envs, K + 1, -1)) < 5e-4
        assert got["n_updates"] == int(z[k + "n_updates"][0])
        m = z[k + "metrics"]
        for name, idx in (("pg_loss", 0), ("v_loss", 2), ("approx_kl", 4), ("ratio", 5)):
            assert abs(got[name] - m[idx]) < 5e-3 * max(1.0, abs(m[idx])) + 2e-5, (itr, name, got[name], m[idx])
        assert abs(got["clipfrac"] - float(z[k + "clipfrac_mean"][0])) < 0.03
        assert abs(got["explained_var"] - float(z[k + "explained_var"][0])) < 5e-3
        fp = np.concatenate([model.engine.get_weights(L.NET_ACTOR_FT), model.engine.get_weights(L.NET_CRITIC)])[::97]
        assert np.mean(np.abs(fp - z[k + "weights_fp"]) > 0.25 * lr) < 1e-2
    assert agent.opt_iterations == 12


def test_pretrain_loop_matches_the_reference_fixture():
    """The CUDA pre-training loop against tests/golden/ref_pretrain_loop.npz (the reference's own run loop, EMA class and step_ema
    exec'd verbatim over the shim): per-epoch losses, final network and EMA weights."""
    import os
    from diffusionpolicyoptimization_b200.agent.pretrain.train_diffusion_agent import TrainDiffusionAgent
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_pretrain_loop.npz"))
    M, B, EPOCHS, FIRST, START, FREQ = (int(v) for v in z["cfg"]); lr0, alpha, wd, decay = (float(v) for v in z["hyper"])
    o = O.make_oracle("hopper", seed=int(z["seed"][0]))
    actor = dp.DiffusionMLP(action_dim=3, horizon_steps=4, cond_dim=11, time_dim=16, mlp_dims=[512, 512, 512],
                            activation_type="ReLU", residual_style=True)

    def hook(cfg):
        cfg.pretrain_weight_decay = wd
    model = dp.DiffusionModel(network=actor, horizon_steps=4, obs_dim=11, action_dim=3, denoising_steps=20, device="cuda:0",
                              precision="fp32", _cfg_hook=hook)
    model.network.set_flat_weights(O.flatten_params(o.actor))

    def draws_fn(epoch, nb, n):
        off = (epoch - 1) * M + nb * B
        return z["t"][off:off + n], z["noise"][off:off + n]
    agent = TrainDiffusionAgent(model, z["actions"], z["states"], n_epochs=EPOCHS, batch_size=B, learning_rate=lr0,
                                lr_first_cycle_steps=FIRST, lr_min=alpha * lr0, ema_decay=decay, epoch_start_ema=START,
                                update_ema_freq=FREQ, draws_fn=draws_fn)
    losses = agent.run()
    np.testing.assert_allclose(losses, z["losses"], rtol=2e-3)
    assert agent.opt_iterations == int(z["opt_iterations"][0])
    assert np.mean(np.abs(model.engine.get_weights(L.NET_ACTOR)[::97] - z["net_fp"]) > 0.25 * lr0) < 1e-2
    assert np.mean(np.abs(model.engine.get_weights(L.NET_ACTOR_EMA)[::97] - z["ema_fp"]) > 0.25 * lr0) < 1e-2
