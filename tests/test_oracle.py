"""CPU: pins the oracle against (a) the known answers SURVEY.md §8(a) derives from the reference's
formulas and (b) the committed golden fixtures (tests/golden/make_golden.py)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import dppo_oracle as O

GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if not os.path.basename(p).startswith("ref_"))   # ref_*: tests/test_ref_golden.py


def test_schedule_known_answers():
    c = O.ddpm_constants(20)
    sigma = np.exp(0.5 * c["ddpm_logvar_clipped"])
    want = [1e-10, 0.0758, 0.1232, 0.1617, 0.1952, 0.2257, 0.2546, 0.2827, 0.3107, 0.3392, 0.3690]
    np.testing.assert_allclose(sigma[:11], want, rtol=2e-3)
    np.testing.assert_allclose(sigma[-2:], [0.8573, 0.9965], rtol=2e-3)
    assert abs(c["sqrt_recipm1_alphas_cumprod"][19] - 406.2355) < 1e-2
    assert abs(c["ddpm_mu_coef1"][0] - 0.999997) < 1e-5 and c["ddpm_mu_coef2"][0] == 0.0
    assert abs(c["betas"][0] - 0.007993) < 1e-6 and abs(c["betas"][19] - 0.999) < 1e-7
    for v in c.values():
        assert v.dtype == np.float32


def test_param_counts():
    assert O.Dims(obs_dim=11, action_dim=3).n_actor() == 553020
    assert O.Dims(obs_dim=11, action_dim=3).n_critic() == 134913
    assert O.Dims(obs_dim=17, action_dim=6).n_actor() == 568392
    assert O.Dims(obs_dim=17, action_dim=6).n_critic() == 136449


def test_sinusoidal_and_mish():
    e = O.sinusoidal_pos_emb(torch.tensor([0, 3]), 16)
    assert e.shape == (2, 16)
    np.testing.assert_allclose(e[0].numpy(), [0] * 8 + [1] * 8, atol=1e-7)
    f = np.exp(-np.arange(8) * np.log(1e4) / 7)
    np.testing.assert_allclose(e[1].numpy(), np.concatenate([np.sin(3 * f), np.cos(3 * f)]), rtol=1e-5, atol=1e-6)
    x = torch.tensor([-2.0, 0.0, 1.5])
    np.testing.assert_allclose(O.mish(x).numpy(), (x * torch.tanh(torch.log1p(torch.exp(x)))).numpy(), rtol=1e-6)


def test_chain_layout_and_switch():
    """chain has K+1 entries: chain[0] = output of step t=K (base net), chain[-1] = action."""
    o = O.make_oracle("hopper", seed=3)
    obs, x_T, noise = O.make_rollout_inputs(o, 4, seed=5)
    s = o.sample(obs, x_T, noise)
    assert s.chains.shape == (4, 11, 4, 3)
    assert torch.equal(s.chains[:, -1], s.trajectories)
    # first K+... steps only use the base net: identical chain[0] when the ft net is replaced
    o2 = O.Oracle(o.d, o.h, o.actor, [p * 0 for p in o.actor_ft], o.critic)
    s2 = o2.sample(obs, x_T, noise)
    assert torch.equal(s.chains[:, 0], s2.chains[:, 0]) and not torch.equal(s.chains[:, 1], s2.chains[:, 1])
    # K == T records x_T first
    o3 = O.make_oracle("hopper", seed=3, ft_denoising_steps=20)
    s3 = o3.sample(obs, x_T, noise)
    assert s3.chains.shape == (4, 21, 4, 3) and torch.equal(s3.chains[:, 0], x_T)
    # deterministic: no noise at t=0
    sd = o.sample(obs, x_T, noise, deterministic=True)
    noise2 = noise.clone(); noise2[-1] += 1.0
    assert torch.equal(sd.trajectories, o.sample(obs, x_T, noise2, deterministic=True).trajectories)


def test_logprob_matches_torch_normal():
    o = O.make_oracle("hopper", seed=1)
    obs, x_T, noise = O.make_rollout_inputs(o, 6, seed=2)
    chains = o.sample(obs, x_T, noise).chains
    lp = o.get_logprobs(obs, chains)
    K = o.d.ft_denoising_steps
    k = 3
    t = torch.full((6,), K - 1 - k)
    mu, logvar, _ = o.p_mean_var(chains[:, k], t, obs)
    std = torch.clamp(torch.exp(0.5 * logvar), 0.1, 1e6).expand_as(mu)
    ref = torch.distributions.Normal(mu, std).log_prob(chains[:, k + 1])
    np.testing.assert_allclose(lp.reshape(6, K, 4, 3)[:, k].numpy(), ref.numpy(), rtol=1e-4, atol=1e-4)


def test_adamw_keras_first_step():
    p = [torch.tensor([1.0, -2.0])]; g = [torch.tensor([0.5, 0.25])]
    m = [torch.zeros(2)]; v = [torch.zeros(2)]
    O.adamw_keras(p, g, m, v, step=1, lr=1e-2, weight_decay=0.004)
    # first step: m/(sqrt(v)) -> sign(g) after bias correction (up to eps), plus decoupled decay
    want = torch.tensor([1.0, -2.0]) * (1 - 0.004 * 1e-2) - 1e-2 * torch.tensor([1.0, 1.0])
    np.testing.assert_allclose(p[0].numpy(), want.numpy(), rtol=1e-5)


def test_ppo_gradient_finite_difference():
    o = O.make_oracle("hopper", seed=2, actor_hidden=32, critic_hidden=16)
    batch = O.make_ppo_batch(o, 64, pool=16, seed=4)
    metrics, ga, gc = o.ppo_grads(*batch)
    idx, j = 11, 5            # output bias of the actor
    eps = 1e-3
    def loss_with(delta):
        ft = [p.clone() for p in o.actor_ft]; ft[idx][j] += delta
        out = o.ppo_loss(*batch, actor_ft=ft)
        return float(out[0] + o.h.vf_coef * out[2])
    fd = (loss_with(eps) - loss_with(-eps)) / (2 * eps)
    assert abs(fd - float(ga[idx][j])) < 5e-3 * max(1.0, abs(fd)) + 2e-4


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_reproduces_golden(path):
    z = np.load(path)
    task = os.path.basename(path).split("_")[0]
    seed = {"hopper": 0, "walker2d": 7}[task]
    o = O.make_oracle(task, seed=seed)
    np.testing.assert_array_equal(O.flatten_params(o.actor)[::97], z["actor_fp"])
    np.testing.assert_array_equal(O.flatten_params(o.actor_ft)[::97], z["actor_ft_fp"])
    s = o.sample(torch.from_numpy(z["obs"]), torch.from_numpy(z["x_T"]), torch.from_numpy(z["noise"]))
    np.testing.assert_allclose(s.chains.numpy(), z["chains"], rtol=1e-5, atol=1e-6)
    lp = o.get_logprobs(torch.from_numpy(z["obs"]), torch.from_numpy(z["chains"]))
    np.testing.assert_allclose(lp.numpy(), z["logp"], rtol=1e-4, atol=1e-5)
    batch = tuple(torch.from_numpy(z[k]) for k in ("ppo_obs", "ppo_prev", "ppo_next", "ppo_inds", "ppo_returns",
                                                   "ppo_oldvalues", "ppo_adv", "ppo_oldlogp"))
    metrics, ga, gc = o.ppo_grads(*batch)
    np.testing.assert_allclose([float(m) for m in metrics], z["ppo_metrics"], rtol=1e-4, atol=1e-6)
    g = np.concatenate([O.flatten_params(ga), O.flatten_params(gc)])[::97]
    np.testing.assert_allclose(g, z["ppo_grads_fp"], rtol=1e-3, atol=1e-7)


def test_gae_known_answer():
    """Hand-computed 3-step, 2-env case of train_ppo_diffusion_agent.py:242-263 (env 1 terminates at t=1)."""
    r = np.array([[1.0, 1.0], [2.0, 0.5], [0.0, 4.0]]); term = np.array([[0, 0], [0, 1], [0, 0]], np.float32)
    v = np.array([[0.5, 0.5], [1.0, 1.0], [2.0, 0.0]], np.float32); nv = np.array([3.0, 1.0], np.float32)
    g, lam = 0.9, 0.5
    adv, ret = O.gae(r, term, v, nv, 1.0, g, lam)
    gnv = (np.float32(g) * nv).astype(np.float64)       # the bootstrap term is an fp32 product in the reference (Python float x fp32 array)
    d2 = np.array([0.0 + gnv[0] - 2.0, 4.0 + gnv[1] - 0.0]); a2 = d2
    d1 = np.array([2.0 + g * 2.0 - 1.0, 0.5 + 0.0 - 1.0]); a1 = d1 + g * lam * np.array([1.0, 0.0]) * a2
    d0 = np.array([1.0 + g * 1.0 - 0.5, 1.0 + g * 1.0 - 0.5]); a0 = d0 + g * lam * a1
    np.testing.assert_allclose(adv, np.stack([a0, a1, a2]), rtol=1e-15)
    np.testing.assert_allclose(ret, adv + v, rtol=1e-15)
