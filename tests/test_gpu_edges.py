"""GPU: edge cases of the C ABI — the fp32 / tensor-mode switch-over sizes, ragged last row tiles, single rows, the cluster
sampler's size limit, and argument errors that must come back as error codes (DppoError), never as a crash or a silent no-op."""
import numpy as np
import pytest
import torch

import diffusionpolicyoptimization_b200 as dp
from diffusionpolicyoptimization_b200 import _lib as L
from helpers import make_engine, max_abs, rel_err
from oracle import dppo_oracle as O

pytestmark = pytest.mark.gpu


def _flat(x):
    return x.reshape(x.shape[0], -1)


def _ppo(e, batch, **kw):
    N = batch[0].shape[0]
    return e.ppo_step(_flat(batch[0]), batch[1].reshape(N, -1), batch[2].reshape(N, -1), batch[3], batch[4], batch[5], batch[6],
                      batch[7].reshape(N, -1), lr=0.0, apply=False, want_grads=True, **kw)


@pytest.fixture(scope="module")
def tensor_pair():
    o = O.make_oracle("walker2d", seed=3)
    e = make_engine(o, precision=L.PREC_BF16)
    yield o, e
    e.close()


@pytest.mark.parametrize("N", [2047, 2048, 2049, 2176 + 1])
def test_switch_over_sizes_and_ragged_tiles(tensor_pair, N):
    """2047 rows stay on the strict fp32 path, 2048 is the first tensor-mode size (16 full tiles), 2049 / 2177 leave a last
    tile with a single row."""
    o, e = tensor_pair
    batch = O.make_ppo_batch(o, N, pool=256, seed=N)
    metrics, ga, gc = o.ppo_grads(*batch)
    want_g = np.concatenate([O.flatten_params(ga), O.flatten_params(gc)])
    n0 = e.tc_launch_count() + e.fused_launch_count()
    m, g = _ppo(e, batch)
    torch.cuda.synchronize()
    used_tensor = (e.tc_launch_count() + e.fused_launch_count()) > n0
    assert used_tensor == (N >= 2048)
    # bf16 tensor mode: stated bound 0.15 of the largest gradient entry (0.056 - 0.097 measured on these batches)
    tol_m, tol_g = ((5e-2, 0.15) if used_tensor else (2e-3, 1e-3))
    np.testing.assert_allclose(m.cpu().numpy(), [float(x) for x in metrics], rtol=tol_m, atol=2e-3 if used_tensor else 2e-6)
    nA = o.d.n_actor()
    for sl in (slice(0, nA), slice(nA, None)):
        assert np.abs(g.cpu().numpy()[sl] - want_g[sl]).max() < tol_g * np.abs(want_g[sl]).max()
    lp = e.logprobs_subsample(_flat(batch[0]), batch[1].reshape(N, -1), batch[2].reshape(N, -1), batch[3])
    with torch.no_grad():
        want_lp, _ = o.get_logprobs_subsample(batch[0], batch[1], batch[2], batch[3])
    assert max_abs(lp, want_lp.reshape(N, -1)) < (0.15 if used_tensor else 1e-3)
    assert np.isfinite(lp.cpu().numpy()).all()


def test_single_row_calls():
    """One row everywhere: a one-element advantage batch normalises to 0 (population std 0 -> (a - a) / 1e-8 = 0)."""
    o = O.make_oracle("hopper", seed=4)
    e = make_engine(o)
    batch = O.make_ppo_batch(o, 1, pool=4, seed=1)
    metrics, ga, gc = o.ppo_grads(*batch)
    m, g = _ppo(e, batch)
    np.testing.assert_allclose(m.cpu().numpy(), [float(x) for x in metrics], rtol=2e-3, atol=2e-6)
    want_g = np.concatenate([O.flatten_params(ga), O.flatten_params(gc)])
    assert np.abs(g.cpu().numpy() - want_g).max() < 1e-3 * np.abs(want_g).max()
    obs, x_T, noise = O.make_rollout_inputs(o, 1, seed=2)
    want = o.sample(obs, x_T, noise)
    for path in (1, 2):
        e.force_path(path)
        a, c = e.sample(_flat(obs), x_T=x_T.reshape(1, -1), noise=noise.reshape(o.d.denoising_steps, 1, -1))
        assert rel_err(c, want.chains.reshape(1, o.d.ft_denoising_steps + 1, -1)) < 1e-4
    e.force_path(0)
    assert max_abs(e.logprobs(_flat(obs), want.chains.reshape(1, o.d.ft_denoising_steps + 1, -1)),
                   o.get_logprobs(obs, want.chains).reshape(o.d.ft_denoising_steps, -1)) < 1e-3
    assert max_abs(e.value(_flat(obs)), O.critic_obs(o.critic, obs).reshape(-1)) < 1e-4
    e.close()


def test_cluster_sampler_size_limit_falls_over_to_the_layered_path():
    """The persistent cluster sampler takes B <= 4096; one more row must switch paths and keep the numbers."""
    o = O.make_oracle("hopper", seed=5)
    e = make_engine(o)
    for B, path in ((4096, 1), (4097, 2)):
        obs, x_T, noise = O.make_rollout_inputs(o, B, seed=B)
        a, c = e.sample(_flat(obs), x_T=x_T.reshape(B, -1), noise=noise.reshape(o.d.denoising_steps, B, -1))
        torch.cuda.synchronize()
        assert e.last_path() == path
        want = o.sample(obs[-64:], x_T[-64:], noise[:, -64:])         # the last rows: the ragged end of the grid
        assert rel_err(c[-64:], want.chains.reshape(64, o.d.ft_denoising_steps + 1, -1)) < 1e-4
    e.close()


def test_argument_errors_are_reported():
    o = O.make_oracle("hopper", seed=6)
    e = make_engine(o)
    batch = O.make_ppo_batch(o, 8, pool=4, seed=1)
    with pytest.raises(dp.DppoError):                              # sharded rows without global advantage statistics
        _ppo(e, batch, n_global=16)
    with pytest.raises(dp.DppoError):                              # N_global < N
        _ppo(e, batch, n_global=4, adv_mean=0.0, adv_std=1.0)
    with pytest.raises(dp.DppoError):
        e.set_ft_denoising_steps(o.d.denoising_steps + 1)
    with pytest.raises(dp.DppoError):                              # flat (b, k) index outside the resident rollout
        P, K, A = 4, o.d.ft_denoising_steps, o.d.A
        e.ppo_step_indexed(torch.zeros(P, o.d.Do), torch.zeros(P, K + 1, A), torch.zeros(P, K, A), torch.zeros(P), torch.zeros(P),
                           torch.zeros(P), np.array([0, P * K], np.int32), lr=0.0, apply=False, metrics_host=np.zeros(8, np.float32))
    with pytest.raises(ValueError):                                # an output view of the wrong size
        e.sample(torch.zeros(3, o.d.Do), chains_out=torch.zeros(2, o.d.ft_denoising_steps + 1, o.d.A, device="cuda"))
    # the handle is still usable after the errors
    m, _ = _ppo(e, batch)
    assert np.isfinite(m.cpu().numpy()).all()
    e.close()
