"""GPU (-m gpu): the CUDA path through the C ABI against the oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star, fp32 mode): actions / chains within 1e-4 relative
(norm-wise: max|a-b| / max|b|), per-step log-probs within 1e-3 absolute.  Gradients and
post-AdamW weights: 1e-3 relative to the largest gradient entry / 1e-6 absolute on weights
(fp32 summation order differs between a tiled FFMA GEMM and the CPU BLAS).
"""
import glob
import os

import numpy as np
import pytest
import torch

from diffusionpolicyoptimization_b200 import _lib as L
from oracle import dppo_oracle as O
from helpers import make_engine, max_abs, rel_err

pytestmark = pytest.mark.gpu

ACT_RTOL = 1e-4
LOGP_ATOL = 1e-3
GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if not os.path.basename(p).startswith("ref_"))   # ref_*: tests/test_ref_golden.py


@pytest.fixture(scope="module", params=["hopper", "walker2d"])
def pair(request):
    o = O.make_oracle(request.param, seed=0)
    e = make_engine(o)
    yield o, e
    e.close()


def _flat_obs(obs):
    return obs.reshape(obs.shape[0], -1)


@pytest.mark.parametrize("B", [1, 5, 40, 67])
@pytest.mark.parametrize("path", [1, 2])     # 1 = persistent cluster sampler, 2 = layer-by-layer fp32
def test_sample_matches_oracle(pair, B, path):
    o, e = pair
    obs, x_T, noise = O.make_rollout_inputs(o, B, seed=B)
    want = o.sample(obs, x_T, noise)
    e.force_path(path)
    actions, chains = e.sample(_flat_obs(obs), x_T=x_T.reshape(B, -1), noise=noise.reshape(o.d.denoising_steps, B, -1))
    torch.cuda.synchronize()
    assert e.last_path() == path
    e.force_path(0)
    assert rel_err(actions, want.trajectories.reshape(B, -1)) < ACT_RTOL
    assert rel_err(chains, want.chains.reshape(B, o.d.ft_denoising_steps + 1, -1)) < ACT_RTOL


def test_sample_is_one_launch(pair):
    o, e = pair
    obs, x_T, noise = O.make_rollout_inputs(o, 40, seed=9)
    e.sample(_flat_obs(obs), x_T=x_T.reshape(40, -1), noise=noise.reshape(o.d.denoising_steps, 40, -1))
    n0 = e.launch_count()
    e.sample(_flat_obs(obs), x_T=x_T.reshape(40, -1), noise=noise.reshape(o.d.denoising_steps, 40, -1))
    assert e.last_path() == 1 and e.launch_count() - n0 == 1


@pytest.mark.parametrize("kw", [dict(deterministic=True), dict(use_base_policy=True), dict(min_sampling_std=0.25)])
def test_sample_modes(pair, kw):
    o, e = pair
    B = 12
    obs, x_T, noise = O.make_rollout_inputs(o, B, seed=21)
    okw = dict(kw)
    if "min_sampling_std" in okw:
        okw["min_sampling_denoising_std"] = okw.pop("min_sampling_std")
    want = o.sample(obs, x_T, noise, **okw)
    for path in (1, 2):
        e.force_path(path)
        actions, chains = e.sample(_flat_obs(obs), x_T=x_T.reshape(B, -1), noise=noise.reshape(o.d.denoising_steps, B, -1), **kw)
        e.force_path(0)
        assert rel_err(actions, want.trajectories.reshape(B, -1)) < ACT_RTOL
        assert rel_err(chains, want.chains.reshape(B, o.d.ft_denoising_steps + 1, -1)) < ACT_RTOL


def test_sample_all_steps_finetuned():
    """K == T: the initial x_T is chain[0] (diffusion_vpg.py:286-287)."""
    o = O.make_oracle("hopper", seed=4, ft_denoising_steps=20)
    e = make_engine(o)
    obs, x_T, noise = O.make_rollout_inputs(o, 7, seed=2)
    want = o.sample(obs, x_T, noise)
    for path in (1, 2):
        e.force_path(path)
        actions, chains = e.sample(_flat_obs(obs), x_T=x_T.reshape(7, -1), noise=noise.reshape(20, 7, -1))
        assert chains.shape == (7, 21, 12)
        assert rel_err(chains, want.chains.reshape(7, 21, -1)) < ACT_RTOL
    e.close()


def test_sample_host_entry_point(pair):
    o, e = pair
    B = 40
    obs, x_T, noise = O.make_rollout_inputs(o, B, seed=33)
    want = o.sample(obs, x_T, noise)
    A = o.d.A
    acts = np.empty((B, A), np.float32); ch = np.empty((B, o.d.ft_denoising_steps + 1, A), np.float32)
    e.sample_host(np.ascontiguousarray(_flat_obs(obs).numpy()), acts, ch,
                  x_T=np.ascontiguousarray(x_T.reshape(B, -1).numpy()), noise=np.ascontiguousarray(noise.reshape(-1, B, A).numpy()))
    assert rel_err(acts, want.trajectories.reshape(B, -1)) < ACT_RTOL
    assert rel_err(ch, want.chains.reshape(B, -1, A)) < ACT_RTOL


@pytest.mark.parametrize("N", [1, 129, 1000])
def test_actor_forward_and_value(pair, N):
    o, e = pair
    rng = np.random.default_rng(N)
    d = o.d
    x = torch.from_numpy(rng.standard_normal((N, d.horizon_steps, d.action_dim)).astype(np.float32))
    t = torch.from_numpy(rng.integers(0, d.denoising_steps, N))
    obs = torch.from_numpy(rng.uniform(-1, 1, (N, 1, d.obs_dim)).astype(np.float32))
    with torch.no_grad():
        want = O.diffusion_mlp(o.actor_ft, x, t, obs, d, o.h.actor_act)
        wantv = O.critic_obs(o.critic, obs, o.h.critic_act).reshape(-1)
    got = e.actor_forward(L.NET_ACTOR_FT, x.reshape(N, -1), t, _flat_obs(obs))
    assert rel_err(got, want.reshape(N, -1)) < 2e-5
    assert max_abs(e.value(_flat_obs(obs)), wantv) < 2e-5


@pytest.mark.parametrize("B", [1, 13, 200])
def test_logprobs_match_oracle(pair, B):
    o, e = pair
    obs, x_T, noise = O.make_rollout_inputs(o, B, seed=100 + B)
    chains = o.sample(obs, x_T, noise).chains
    want = o.get_logprobs(obs, chains)
    got = e.logprobs(_flat_obs(obs), chains.reshape(B, o.d.ft_denoising_steps + 1, -1))
    assert max_abs(got, want.reshape(B * o.d.ft_denoising_steps, -1)) < LOGP_ATOL
    want_b = o.get_logprobs(obs, chains, use_base_policy=True)
    got_b = e.logprobs(_flat_obs(obs), chains.reshape(B, o.d.ft_denoising_steps + 1, -1), use_base_policy=True)
    assert max_abs(got_b, want_b.reshape(B * o.d.ft_denoising_steps, -1)) < LOGP_ATOL


def _check_ppo(o, e, N, seed, clip_v=None):
    batch = O.make_ppo_batch(o, N, pool=max(16, N // 8), seed=seed)
    metrics, ga, gc = o.ppo_grads(*batch)
    want_g = np.concatenate([O.flatten_params(ga), O.flatten_params(gc)])
    lp = e.logprobs_subsample(_flat_obs(batch[0]), batch[1].reshape(N, -1), batch[2].reshape(N, -1), batch[3])
    with torch.no_grad():
        want_lp, _ = o.get_logprobs_subsample(batch[0], batch[1], batch[2], batch[3])
    assert max_abs(lp, want_lp.reshape(N, -1)) < LOGP_ATOL
    got_m, got_g = e.ppo_step(_flat_obs(batch[0]), batch[1].reshape(N, -1), batch[2].reshape(N, -1), batch[3], batch[4],
                              batch[5], batch[6], batch[7].reshape(N, -1), lr=0.0, apply=False, want_grads=True)
    torch.cuda.synchronize()
    got_m = got_m.cpu().numpy(); got_g = got_g.cpu().numpy()
    np.testing.assert_allclose(got_m, [float(m) for m in metrics], rtol=2e-3, atol=2e-6)
    nA = o.d.n_actor()
    for name, sl in (("actor_ft", slice(0, nA)), ("critic", slice(nA, None))):
        scale = np.abs(want_g[sl]).max()
        assert np.abs(got_g[sl] - want_g[sl]).max() < 1e-3 * scale, name
    return batch, want_g


@pytest.mark.parametrize("N", [64, 1000, 4099])
def test_ppo_loss_and_gradients(pair, N):
    o, e = pair
    _check_ppo(o, e, N, seed=N)


def test_ppo_value_clipping():
    o = O.make_oracle("hopper", seed=5, hyper=O.Hyper(clip_vloss_coef=0.05))
    e = make_engine(o)
    _check_ppo(o, e, 300, seed=8)
    e.close()


def test_ppo_adamw_step_matches_keras_semantics():
    o = O.make_oracle("hopper", seed=6)
    e = make_engine(o)
    N = 512
    batch = O.make_ppo_batch(o, N, pool=64, seed=12)
    params = [p.clone() for p in o.actor_ft] + [p.clone() for p in o.critic]
    m = [torch.zeros_like(p) for p in params]; v = [torch.zeros_like(p) for p in params]
    lr = 3e-4
    for step in (1, 2):
        oo = O.Oracle(o.d, o.h, o.actor, params[:12], params[12:])
        _, ga, gc = oo.ppo_grads(*batch)
        O.adamw_keras(params, ga + gc, m, v, step, lr, o.h.beta1, o.h.beta2, o.h.adam_eps, o.h.weight_decay)
        e.ppo_step(_flat_obs(batch[0]), batch[1].reshape(N, -1), batch[2].reshape(N, -1), batch[3], batch[4], batch[5],
                   batch[6], batch[7].reshape(N, -1), lr=lr, apply=True)
    got = np.concatenate([e.get_weights(L.NET_ACTOR_FT), e.get_weights(L.NET_CRITIC)])
    want = O.flatten_params(params)
    # Adam's first steps move every weight by ~lr regardless of gradient scale; entries whose
    # gradient is ~0 are sign-ambiguous, so compare in units of lr
    frac_bad = np.mean(np.abs(got - want) > 0.05 * lr)
    assert frac_bad < 2e-3, frac_bad
    mm, vv, st = e.get_opt_state(L.OPT_FINETUNE)
    assert st == 2
    np.testing.assert_allclose(mm, O.flatten_params(m), rtol=0, atol=1e-3 * np.abs(O.flatten_params(m)).max())
    e.close()


@pytest.mark.parametrize("clip", [0.03, 1e-3])     # 0.03: about half of the 20 variables exceed it
def test_ppo_per_variable_clip_by_norm(clip):
    """train_ppo_diffusion_agent.py:349-354: every variable's gradient is clipped to `clip` on its own before AdamW;
    the gradient handed back stays the unclipped one.  Adam's update is scale-free, so the first moments carry the check."""
    o = O.make_oracle("hopper", seed=16)
    e = make_engine(o)
    N = 384
    batch = O.make_ppo_batch(o, N, pool=64, seed=17)
    _, ga, gc = o.ppo_grads(*batch)
    raw = ga + gc
    norms = [float(g.norm()) for g in raw]
    assert any(n > clip for n in norms) and (clip < 0.01 or any(n <= clip for n in norms))   # both branches of the max()
    clipped = O.clip_by_norm(raw, clip)
    e.set_grad_clip_norm(clip)
    _, g = e.ppo_step(_flat_obs(batch[0]), batch[1].reshape(N, -1), batch[2].reshape(N, -1), batch[3], batch[4], batch[5],
                      batch[6], batch[7].reshape(N, -1), lr=0.0, apply=True, want_grads=True)
    want_raw = O.flatten_params(raw)
    assert np.abs(g.cpu().numpy() - want_raw).max() < 1e-3 * np.abs(want_raw).max()
    mm, _, st = e.get_opt_state(L.OPT_FINETUNE)
    assert st == 1
    off = 0
    for c, n in zip(clipped, norms):
        want = (c * (1 - o.h.beta1)).numpy().reshape(-1)
        got = mm[off:off + want.size]; off += want.size
        assert np.abs(got - want).max() < 2e-3 * np.abs(want).max(), (n, clip)
    assert off == mm.size
    # switching it off again restores the plain update
    e.set_grad_clip_norm(None)
    e.set_opt_state(L.OPT_FINETUNE, np.zeros_like(mm), np.zeros_like(mm), 0)
    e.ppo_step(_flat_obs(batch[0]), batch[1].reshape(N, -1), batch[2].reshape(N, -1), batch[3], batch[4], batch[5],
               batch[6], batch[7].reshape(N, -1), lr=0.0, apply=True)
    mm2, _, _ = e.get_opt_state(L.OPT_FINETUNE)
    assert np.abs(mm2 - want_raw * (1 - o.h.beta1)).max() < 2e-3 * np.abs(want_raw).max() * (1 - o.h.beta1)
    e.close()


@pytest.mark.parametrize("N", [64, 1500])
def test_pretrain_loss_and_gradients(pair, N):
    o, e = pair
    d = o.d
    rng = np.random.default_rng(N)
    x0 = torch.from_numpy(rng.uniform(-1, 1, (N, d.horizon_steps, d.action_dim)).astype(np.float32))
    obs = torch.from_numpy(rng.uniform(-1, 1, (N, 1, d.obs_dim)).astype(np.float32))
    t = torch.from_numpy(rng.integers(0, d.denoising_steps, N))
    nz = torch.from_numpy(rng.standard_normal(tuple(x0.shape)).astype(np.float32))
    want_loss, want_g = o.pretrain_grads(x0, obs, t, nz)
    loss, g = e.pretrain_step(x0.reshape(N, -1), _flat_obs(obs), lr=0.0, apply=False, t=t, noise=nz.reshape(N, -1), want_grads=True)
    assert abs(float(loss) - float(want_loss)) < 1e-4 * max(1.0, float(want_loss))
    wg = O.flatten_params(want_g)
    assert np.abs(g.cpu().numpy() - wg).max() < 1e-3 * np.abs(wg).max()


def test_ema_update(pair):
    o, e = pair
    w = e.get_weights(L.NET_ACTOR)
    e.set_weights(L.NET_ACTOR_EMA, w * 0.5)
    e.ema_update(0.995)
    np.testing.assert_allclose(e.get_weights(L.NET_ACTOR_EMA), 0.995 * 0.5 * w + 0.005 * w, rtol=1e-6, atol=1e-9)
    e.set_weights(L.NET_ACTOR_EMA, w)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_against_committed_golden(path):
    z = np.load(path)
    task = os.path.basename(path).split("_")[0]
    o = O.make_oracle(task, seed={"hopper": 0, "walker2d": 7}[task])
    np.testing.assert_array_equal(O.flatten_params(o.actor_ft)[::97], z["actor_ft_fp"])
    e = make_engine(o)
    B = z["obs"].shape[0]; A = o.d.A; K = o.d.ft_denoising_steps
    for path_id in (1, 2):
        e.force_path(path_id)
        actions, chains = e.sample(z["obs"].reshape(B, -1), x_T=z["x_T"].reshape(B, A), noise=z["noise"].reshape(-1, B, A))
        assert rel_err(chains, z["chains"].reshape(B, K + 1, A)) < ACT_RTOL
        a_det, _ = e.sample(z["obs"].reshape(B, -1), x_T=z["x_T"].reshape(B, A), noise=z["noise"].reshape(-1, B, A), deterministic=True)
        assert rel_err(a_det, z["actions_det"].reshape(B, A)) < ACT_RTOL
        a_base, _ = e.sample(z["obs"].reshape(B, -1), x_T=z["x_T"].reshape(B, A), noise=z["noise"].reshape(-1, B, A), use_base_policy=True)
        assert rel_err(a_base, z["actions_base"].reshape(B, A)) < ACT_RTOL
    e.force_path(0)
    assert max_abs(e.logprobs(z["obs"].reshape(B, -1), z["chains"].reshape(B, K + 1, A)), z["logp"].reshape(B * K, A)) < LOGP_ATOL
    N = z["ppo_obs"].shape[0]
    assert max_abs(e.value(z["ppo_obs"].reshape(N, -1)), z["value"]) < 1e-4
    m, g = e.ppo_step(z["ppo_obs"].reshape(N, -1), z["ppo_prev"].reshape(N, A), z["ppo_next"].reshape(N, A), z["ppo_inds"],
                      z["ppo_returns"], z["ppo_oldvalues"], z["ppo_adv"], z["ppo_oldlogp"].reshape(N, A), lr=0.0, apply=False, want_grads=True)
    np.testing.assert_allclose(m.cpu().numpy(), z["ppo_metrics"], rtol=2e-3, atol=2e-6)
    gf = g.cpu().numpy()[::97]
    assert np.abs(gf - z["ppo_grads_fp"]).max() < 1e-3 * np.abs(z["ppo_grads_fp"]).max()
    loss, pg = e.pretrain_step(z["pre_x0"].reshape(N, A), z["ppo_obs"].reshape(N, -1), lr=0.0, apply=False, t=z["pre_t"],
                               noise=z["pre_noise"].reshape(N, A), want_grads=True)
    assert abs(float(loss) - float(z["pre_loss"][0])) < 1e-4 * max(1.0, float(z["pre_loss"][0]))
    assert np.abs(pg.cpu().numpy()[::97] - z["pre_grads_fp"]).max() < 1e-3 * np.abs(z["pre_grads_fp"]).max()
    e.close()


def test_reference_api_shims():
    """The reference-facing Python surface: PPODiffusion(...)(cond=...), get_logprobs, c_loss, critic."""
    import diffusionpolicyoptimization_b200 as dp
    o = O.make_oracle("hopper", seed=11)
    actor = dp.DiffusionMLP(action_dim=3, horizon_steps=4, cond_dim=11, time_dim=16, mlp_dims=[512, 512, 512],
                            activation_type="ReLU", residual_style=True)
    critic = dp.CriticObs(cond_dim=11, mlp_dims=[256, 256, 256], activation_type="Mish", residual_style=True)
    model = dp.PPODiffusion(gamma_denoising=0.99, clip_ploss_coef=0.01, clip_ploss_coef_base=0.01, clip_ploss_coef_rate=3,
                            randn_clip_value=3, min_sampling_denoising_std=0.1, min_logprob_denoising_std=0.1,
                            actor=actor, critic=critic, ft_denoising_steps=10, horizon_steps=4, obs_dim=11, action_dim=3,
                            denoising_steps=20, device="cuda:0")
    model.actor.set_flat_weights(O.flatten_params(o.actor))
    model.actor_ft.set_flat_weights(O.flatten_params(o.actor_ft))
    model.critic.set_flat_weights(O.flatten_params(o.critic))
    obs, x_T, noise = O.make_rollout_inputs(o, 10, seed=1)
    want = o.sample(obs, x_T, noise)
    s = model(cond={"state": obs.cuda()}, deterministic=False, return_chain=True, x_T=x_T, noise=noise)
    assert s.trajectories.shape == (10, 4, 3) and s.chains.shape == (10, 11, 4, 3)
    assert rel_err(s.chains, want.chains) < ACT_RTOL
    lp = model.get_logprobs({"state": obs.cuda()}, s.chains)
    assert lp.shape == (100, 4, 3) and max_abs(lp, o.get_logprobs(obs, want.chains)) < LOGP_ATOL
    with torch.no_grad():
        assert max_abs(model.critic({"state": obs.cuda()}), O.critic_obs(o.critic, obs, "Mish")) < 1e-4
    batch = O.make_ppo_batch(o, 200, pool=32, seed=2)
    out = model.c_loss({"state": batch[0]}, *batch[1:], use_bc_loss=False, reward_horizon=4)
    metrics, _, _ = o.ppo_grads(*batch)
    np.testing.assert_allclose([float(x) for x in out], [float(x) for x in metrics], rtol=2e-3, atol=2e-6)
    # host path returns NumPy, Philox noise: finite, bounded, chain[-1] == action
    s2 = model(cond={"state": obs.numpy()})
    assert isinstance(s2.trajectories, np.ndarray) and np.isfinite(s2.chains).all()
    np.testing.assert_array_equal(s2.chains[:, -1], s2.trajectories)
    mu, logvar, eta = model.p_mean_var(x_T.cuda(), torch.full((10,), 3), {"state": obs.cuda()})
    wmu, wlv, _ = o.p_mean_var(x_T, torch.full((10,), 3), obs)
    assert max_abs(mu, wmu) < 1e-4


def test_ppo_step_indexed_matches_host_gathered_minibatch(pair):
    """the minibatch assembled on the device from resident rollout buffers and a flat (b*K + k)
    index list (train_ppo_diffusion_agent.py:287-312) gives bit-identical results to the same rows gathered on the host."""
    import diffusionpolicyoptimization_b200 as dp
    o, e = pair
    d = o.d
    K, A = d.ft_denoising_steps, d.A
    P, N = 300, 1000
    rng = np.random.default_rng(77)
    obs, x_T, noise = O.make_rollout_inputs(o, P, seed=5)
    chains = o.sample(obs, x_T, noise).chains.reshape(P, K + 1, A).numpy()
    obs = obs.reshape(P, -1).numpy()
    oldlogp = e.logprobs(obs, chains, use_base_policy=True).reshape(P, K, A).cpu().numpy()
    ret = rng.standard_normal(P).astype(np.float32); val = rng.standard_normal(P).astype(np.float32)
    adv = rng.standard_normal(P).astype(np.float32)
    inds = rng.integers(0, P * K, N).astype(np.int32)
    b, k = inds // K, inds % K
    m1, g1 = e.ppo_step(obs[b], chains[b, k], chains[b, k + 1], k.astype(np.int32), ret[b], val[b], adv[b], oldlogp[b, k],
                        lr=0.0, apply=False, want_grads=True)
    m2, g2 = e.ppo_step_indexed(obs, chains, oldlogp, ret, val, adv, inds, lr=0.0, apply=False, want_grads=True)
    torch.cuda.synchronize()
    assert torch.equal(m1, m2) and torch.equal(g1, g2)
    mh = np.zeros(8, np.float32)
    e.ppo_step_indexed(obs, chains, oldlogp, ret, val, adv, inds, lr=0.0, apply=False, metrics_host=mh)
    np.testing.assert_array_equal(mh, m1.cpu().numpy())
    bad = inds.copy(); bad[17] = P * K
    with pytest.raises(dp.DppoError):
        e.ppo_step_indexed(obs, chains, oldlogp, ret, val, adv, bad, lr=0.0, apply=False, metrics_host=mh)


def test_gae_on_device_is_bit_identical_to_the_numpy_scan(pair):
    """train_ppo_diffusion_agent.py:242-263 (float64 NumPy) vs dppo_gae; outputs compared after the fp32 cast."""
    o, e = pair
    rng = np.random.default_rng(3)
    S, E = 500, 40
    rewards = rng.standard_normal((S, E)) * 3.0                      # float64, like env rewards
    terminated = (rng.random((S, E)) < 0.02).astype(np.float32)
    values = rng.standard_normal((S, E)).astype(np.float32)
    next_values = rng.standard_normal(E).astype(np.float32)
    want_a, want_r = O.gae(rewards, terminated, values, next_values, reward_scale_const=0.1, gamma=0.999, gae_lambda=0.95)
    adv, ret = e.gae(rewards, terminated, values, next_values, reward_scale_const=0.1, gamma=0.999, gae_lambda=0.95)
    np.testing.assert_array_equal(adv.cpu().numpy(), want_a.astype(np.float32))
    np.testing.assert_array_equal(ret.cpu().numpy(), want_r.astype(np.float32))


def test_ffma_peak_probe_is_plausible(pair):
    """bench.py's roofline denominator of the strict-fp32 mode: the measured CUDA-core FFMA rate of this GPU (derived peak of a B200:
    148 SM x 128 lanes x 2 x 1.965 GHz = 74.5 TFLOP/s)."""
    o, e = pair
    v = e.ffma_peak_tflops()
    print(f"measured FFMA rate {v:.1f} TFLOP/s")
    assert 40.0 < v < 80.0, v
