"""GPU (-m gpu): the tcgen05 bf16 mode (DPPO_PREC_BF16) against the fp32 oracle.

BASELINE.json north_star allows "looser stated bounds for any bf16 mode".  The bounds, stated here:
operands (activations and weight copies) are rounded to bf16 (2^-9 relative), products are exact
and accumulated in fp32 in TMEM, epilogues / losses / optimizer run in fp32 on fp32 master weights.
  eps (actor forward), value ........ 2e-2 norm-wise relative
  per-step log-probs ................ 0.15 absolute (the 1/sigma^2 <= 100 factor amplifies eps error)
  PPO / pre-train gradients ......... 0.15 of the largest gradient entry in general; on this file's fixed batches 8e-2 is asserted; losses 5e-2 relative
  sampled actions (20-step chain) ... 0.1 norm-wise relative
The tests also assert that the tensor path (not the FFMA path) produced the numbers.
"""
import os

import numpy as np
import pytest
import torch

from diffusionpolicyoptimization_b200 import _lib as L
from oracle import dppo_oracle as O
from helpers import make_engine, max_abs, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["hopper-fused", "walker2d-fused", "walker2d-layered"])
def pair(request):
    """fused: whole-MLP layer-chain kernel; layered: one tcgen05 GEMM launch per layer (force_path 3)."""
    task, mode = request.param.split("-")
    o = O.make_oracle(task, seed=0)
    e = make_engine(o, precision=L.PREC_BF16)
    e.fused = mode == "fused"
    if not e.fused:
        e.force_path(3)
    yield o, e
    e.close()


DETERMINISTIC = os.environ.get("DPPO_DETERMINISTIC") == "1"     # fixed-order split-K reductions: one GEMM launch per product


def _counts(e, fused, layered):
    return fused if e.fused else layered


def _flat(obs):
    return obs.reshape(obs.shape[0], -1)


def test_bf16_forward_and_value(pair):
    o, e = pair
    N = 3000
    rng = np.random.default_rng(N)
    d = o.d
    x = torch.from_numpy(rng.standard_normal((N, d.horizon_steps, d.action_dim)).astype(np.float32))
    t = torch.from_numpy(rng.integers(0, d.denoising_steps, N))
    obs = torch.from_numpy(rng.uniform(-1, 1, (N, 1, d.obs_dim)).astype(np.float32))
    with torch.no_grad():
        want = O.diffusion_mlp(o.actor_ft, x, t, obs, d, o.h.actor_act)
        wantv = O.critic_obs(o.critic, obs, o.h.critic_act).reshape(-1)
    n0, f0 = e.tc_launch_count(), e.fused_launch_count()
    got = e.actor_forward(L.NET_ACTOR_FT, x.reshape(N, -1), t, _flat(obs))
    gv = e.value(_flat(obs))
    torch.cuda.synchronize()
    assert e.tc_launch_count() - n0 == _counts(e, 2, 8) and e.fused_launch_count() - f0 == _counts(e, 2, 0)
    err, errv = rel_err(got, want.reshape(N, -1)), rel_err(gv, wantv)
    print(f"bf16 eps rel err {err:.3e}, value rel err {errv:.3e}")
    assert err < 2e-2 and errv < 2e-2


def test_bf16_small_batches_stay_fp32(pair):
    o, e = pair
    obs, x_T, noise = O.make_rollout_inputs(o, 40, seed=3)
    want = o.sample(obs, x_T, noise)
    n0 = e.tc_launch_count()
    actions, chains = e.sample(_flat(obs), x_T=x_T.reshape(40, -1), noise=noise.reshape(o.d.denoising_steps, 40, -1))
    torch.cuda.synchronize()
    assert e.tc_launch_count() == n0 and e.last_path() == 1
    assert rel_err(actions, want.trajectories.reshape(40, -1)) < 1e-4


def test_bf16_logprobs(pair):
    o, e = pair
    B = 300
    obs, x_T, noise = O.make_rollout_inputs(o, B, seed=5)
    chains = o.sample(obs, x_T, noise).chains
    want = o.get_logprobs(obs, chains)
    n0 = e.tc_launch_count()
    got = e.logprobs(_flat(obs), chains.reshape(B, o.d.ft_denoising_steps + 1, -1))
    torch.cuda.synchronize()
    assert e.tc_launch_count() - n0 == _counts(e, 1, 4)
    err = max_abs(got, want.reshape(B * o.d.ft_denoising_steps, -1))
    print(f"bf16 logp abs err {err:.3e}")
    assert err < 0.15


def test_bf16_ppo_loss_and_gradients(pair):
    o, e = pair
    N = 4099
    batch = O.make_ppo_batch(o, N, pool=512, seed=N)
    metrics, ga, gc = o.ppo_grads(*batch)
    want_g = np.concatenate([O.flatten_params(ga), O.flatten_params(gc)])
    n0 = e.tc_launch_count()
    got_m, got_g = e.ppo_step(_flat(batch[0]), batch[1].reshape(N, -1), batch[2].reshape(N, -1), batch[3], batch[4],
                              batch[5], batch[6], batch[7].reshape(N, -1), lr=0.0, apply=False, want_grads=True)
    torch.cuda.synchronize()
    # fused: per net fwd 1 + bwd 1, one grouped dW launch for both; layered: 8 forward + 6 dX + 8 dW GEMMs
    assert DETERMINISTIC or e.tc_launch_count() - n0 == _counts(e, 5, 22)
    got_m = got_m.cpu().numpy(); got_g = got_g.cpu().numpy()
    print("bf16 ppo metrics", got_m, [float(m) for m in metrics])
    np.testing.assert_allclose(got_m, [float(m) for m in metrics], rtol=5e-2, atol=2e-3)
    nA = o.d.n_actor()
    ao = o.d  # noqa
    for name, sl in (("actor_ft", slice(0, nA)), ("critic", slice(nA, None))):
        scale = np.abs(want_g[sl]).max()
        err = np.abs(got_g[sl] - want_g[sl]).max() / scale
        print(f"bf16 ppo grad rel err {name}: {err:.3e}")
        assert err < 8e-2, name
    # every parameter tensor individually (catches a mis-scattered bias / time-MLP gradient)
    off = 0
    for i, p in enumerate(list(ga) + list(gc)):
        n = p.numel(); w = p.detach().numpy().reshape(-1); g = got_g[off:off + n]; off += n
        s = max(np.abs(w).max(), 1e-12)
        assert np.abs(g - w).max() < 0.1 * s + 1e-3 * np.abs(want_g).max(), (i, np.abs(g - w).max(), s)


def test_bf16_pretrain(pair):
    o, e = pair
    d = o.d
    N = 2500
    rng = np.random.default_rng(N)
    x0 = torch.from_numpy(rng.uniform(-1, 1, (N, d.horizon_steps, d.action_dim)).astype(np.float32))
    obs = torch.from_numpy(rng.uniform(-1, 1, (N, 1, d.obs_dim)).astype(np.float32))
    t = torch.from_numpy(rng.integers(0, d.denoising_steps, N))
    nz = torch.from_numpy(rng.standard_normal(tuple(x0.shape)).astype(np.float32))
    want_loss, want_g = o.pretrain_grads(x0, obs, t, nz)
    n0 = e.tc_launch_count()
    loss, g = e.pretrain_step(x0.reshape(N, -1), _flat(obs), lr=0.0, apply=False, t=t, noise=nz.reshape(N, -1), want_grads=True)
    torch.cuda.synchronize()
    assert DETERMINISTIC or e.tc_launch_count() - n0 == _counts(e, 3, 11)
    wg = O.flatten_params(want_g)
    err = np.abs(g.cpu().numpy() - wg).max() / np.abs(wg).max()
    print(f"bf16 pretrain loss {float(loss):.5f} vs {float(want_loss):.5f}, grad rel err {err:.3e}")
    assert abs(float(loss) - float(want_loss)) < 5e-2 * max(1.0, float(want_loss))
    assert err < 5e-2


def test_bf16_large_batch_sampling(pair):
    o, e = pair
    B = 2048
    obs, x_T, noise = O.make_rollout_inputs(o, B, seed=17)
    want = o.sample(obs, x_T, noise)
    n0 = e.tc_launch_count()
    actions, chains = e.sample(_flat(obs), x_T=x_T.reshape(B, -1), noise=noise.reshape(o.d.denoising_steps, B, -1))
    torch.cuda.synchronize()
    assert e.last_path() == _counts(e, 4, 3) and e.tc_launch_count() - n0 == _counts(e, 1, 4 * o.d.denoising_steps)
    err = rel_err(actions, want.trajectories.reshape(B, -1))
    # mean error is the meaningful figure for a 20-step stochastic chain; the max is dominated by
    # rows where the bf16 perturbation flips a clip decision
    mean_err = float((actions.cpu() - want.trajectories.reshape(B, -1)).abs().mean())
    print(f"bf16 sampled-action rel err {err:.3e}, mean abs err {mean_err:.3e}")
    assert mean_err < 2e-2 and err < 0.25


def test_bf16_adamw_step_runs_and_refreshes_operands(pair):
    """After an applied step the bf16 operand copies must follow the fp32 masters."""
    o, e = pair
    N = 2304
    batch = O.make_ppo_batch(o, N, pool=256, seed=1)
    args = [_flat(batch[0]), batch[1].reshape(N, -1), batch[2].reshape(N, -1), batch[3], batch[4], batch[5], batch[6], batch[7].reshape(N, -1)]
    w0 = e.get_weights(L.NET_ACTOR_FT).copy()
    lp0 = e.logprobs_subsample(*args[:4]).clone()
    e.ppo_step(*args, lr=1e-2, apply=True)
    w1 = e.get_weights(L.NET_ACTOR_FT)
    assert np.abs(w1 - w0).max() > 1e-3
    lp1 = e.logprobs_subsample(*args[:4])
    assert float((lp1 - lp0).abs().max()) > 1e-3      # the forward sees the updated weights
    e.set_weights(L.NET_ACTOR_FT, w0)
    lp2 = e.logprobs_subsample(*args[:4])
    assert float((lp2 - lp0).abs().max()) == 0.0       # and is deterministic


def test_bf16_host_entry_chunked_pipeline_matches_device_call(pair):
    """dppo_ppo_step_host overlaps the H2D copy of chunk c+1 with the compute of chunk c (chunks of whole 148 x 128-row waves);
    its metrics and post-step weights must agree with the single-chunk device-resident call on the same rows."""
    o, e = pair
    N = 40000        # > 2 waves of 148 x 128 rows -> 3 pipeline chunks
    rng = np.random.default_rng(5)
    d = o.d
    K = d.ft_denoising_steps
    obs = rng.uniform(-1, 1, (N, d.Do)).astype(np.float32)
    prev = rng.standard_normal((N, d.A)).astype(np.float32)
    nxt = (prev + 0.1 * rng.standard_normal((N, d.A))).astype(np.float32)
    inds = rng.integers(0, K, N).astype(np.int32)
    ret = rng.standard_normal(N).astype(np.float32); val = rng.standard_normal(N).astype(np.float32)
    adv = rng.standard_normal(N).astype(np.float32)
    olp = e.logprobs_subsample(obs, prev, nxt, inds).cpu().numpy() + 0.01 * rng.standard_normal((N, d.A)).astype(np.float32)
    olp = np.ascontiguousarray(olp, np.float32)
    w_a, w_c = e.get_weights(L.NET_ACTOR_FT).copy(), e.get_weights(L.NET_CRITIC).copy()
    n = e.n_actor + e.n_critic
    e.set_opt_state(L.OPT_FINETUNE, np.zeros(n, np.float32), np.zeros(n, np.float32), 0)     # earlier tests stepped the optimizer
    m_dev, g_dev = e.ppo_step(obs, prev, nxt, inds, ret, val, adv, olp, lr=1e-3, apply=True, want_grads=True)
    m_dev = m_dev.cpu().numpy(); g_dev = g_dev.cpu().numpy()
    wa1 = e.get_weights(L.NET_ACTOR_FT).copy(); wc1 = e.get_weights(L.NET_CRITIC).copy()
    # rewind weights and optimizer state, then take the same step through the host entry point
    e.set_weights(L.NET_ACTOR_FT, w_a); e.set_weights(L.NET_CRITIC, w_c)
    e.set_opt_state(L.OPT_FINETUNE, np.zeros(n, np.float32), np.zeros(n, np.float32), 0)
    m_host = np.zeros(8, np.float32)
    e.ppo_step_host(obs, prev, nxt, inds, ret, val, adv, olp, m_host, lr=1e-3, apply=True)
    np.testing.assert_allclose(m_host, m_dev, rtol=1e-4, atol=1e-6)
    wa2 = e.get_weights(L.NET_ACTOR_FT); wc2 = e.get_weights(L.NET_CRITIC)
    # first Adam step moves every weight by ~lr * sign(g): compare in units of lr, allowing sign flips where g ~ 0
    assert np.mean(np.abs(wa2 - wa1) > 0.05e-3) < 5e-3 and np.mean(np.abs(wc2 - wc1) > 0.05e-3) < 5e-3
    e.set_weights(L.NET_ACTOR_FT, w_a); e.set_weights(L.NET_CRITIC, w_c)
    e.set_opt_state(L.OPT_FINETUNE, np.zeros(n, np.float32), np.zeros(n, np.float32), 0)


def test_bf16_logprobs_row_chunking_is_seamless(pair):
    """dppo_logprobs walks the rows in chunks of 2^18 (whole chains per chunk): results around the chunk boundary and
    in the ragged last tile must equal those of small independent calls on the same chains."""
    o, e = pair
    K, A = o.d.ft_denoising_steps, o.d.A
    B = 27001                                     # 270 010 rows: two chunks, last tile ragged
    g = torch.Generator(device="cuda"); g.manual_seed(11)
    obs = torch.rand(B, o.d.Do, device="cuda", generator=g) * 2 - 1
    _, chains = e.sample(obs, seed=3, offset=4)
    lp = e.logprobs(obs, chains).reshape(B, K, A)
    cut = (1 << 18) // K                          # first chain of the second chunk
    for lo, hi in ((cut - 300, cut + 300), (B - 257, B), (0, 300)):
        part = e.logprobs(obs[lo:hi].contiguous(), chains[lo:hi].contiguous()).reshape(hi - lo, K, A)
        assert torch.equal(part, lp[lo:hi]), (lo, hi)
    assert torch.isfinite(lp).all()


def test_bf16_sampler_is_shard_invariant(pair):
    """Philox counters are keyed by the global row: sampling rows [lo, hi) with row_offset = lo reproduces the same rows
    of a single large call bit for bit (what makes the 8-GPU rollout independent of the shard layout)."""
    o, e = pair
    B = 4096
    g = torch.Generator(device="cuda"); g.manual_seed(12)
    obs = torch.rand(B, o.d.Do, device="cuda", generator=g) * 2 - 1
    a_all, c_all = e.sample(obs, seed=9, offset=1)
    lo, hi = 1024, 3200
    a_part, c_part = e.sample(obs[lo:hi].contiguous(), seed=9, offset=1, row_offset=lo)
    assert e.last_path() == _counts(e, 4, 3)
    assert torch.equal(a_part, a_all[lo:hi]) and torch.equal(c_part, c_all[lo:hi])


def test_bf16_index_driven_update_reads_the_rollout_buffers_directly(pair):
    """Tensor mode: dppo_ppo_step_indexed does not materialise the minibatch (the h0 pack and loss kernels address the rollout
    buffers through the flat indices); it must agree with dppo_ppo_step on the same rows gathered on the host."""
    o, e = pair
    if not e.fused:
        pytest.skip("the index-driven loads belong to the fused path")
    d = o.d
    P, K, A, N = 600, d.ft_denoising_steps, d.A, 4096
    obs, x_T, noise = O.make_rollout_inputs(o, P, seed=77)
    chains = o.sample(obs, x_T, noise).chains.reshape(P, K + 1, A)
    g = torch.Generator().manual_seed(5)
    olp = torch.randn(P, K, A, generator=g) * 0.3 - 1.0
    ret, val, adv = torch.randn(P, generator=g), torch.randn(P, generator=g), torch.randn(P, generator=g)
    flat = torch.randint(0, P * K, (N,), generator=g, dtype=torch.int64)
    b, k = flat // K, flat % K
    n0 = e.launch_count()
    m1, g1 = e.ppo_step_indexed(_flat(obs), chains, olp, ret, val, adv, flat.to(torch.int32).cuda(), lr=0.0, apply=False, want_grads=True)
    launches_indexed = e.launch_count() - n0
    m2, g2 = e.ppo_step(_flat(obs)[b], chains[b, k], chains[b, k + 1], k.to(torch.int32), ret[b], val[b], adv[b], olp[b, k],
                        lr=0.0, apply=False, want_grads=True)
    torch.cuda.synchronize()
    n1 = e.launch_count()
    e.ppo_step(_flat(obs)[b], chains[b, k], chains[b, k + 1], k.to(torch.int32), ret[b], val[b], adv[b], olp[b, k], lr=0.0, apply=False)
    # no gather kernel in front of the update (the deterministic mode keeps the gathered path: one more launch); the device variant
    # ends with the one-warp kernel that turns the metrics into NaN when an index was out of range
    assert launches_indexed == e.launch_count() - n1 + 1 + (1 if DETERMINISTIC else 0)
    assert torch.equal(m1, m2)                                        # the loss partial sums are reduced in a fixed order
    scale = float(g2.abs().max())
    assert float((g1 - g2).abs().max()) < 2e-3 * scale                # dW accumulates with atomics: order varies run to run
    with pytest.raises(L.DppoError):
        e.ppo_step_indexed(_flat(obs), chains, olp, ret, val, adv, np.full(N, P * K, np.int32), lr=0.0, apply=False,
                           metrics_host=np.zeros(8, np.float32))
    # an out-of-range index must never reach the optimizer: weights and moments untouched, NaN metrics on the device variant
    w0 = e.get_weights(L.NET_ACTOR_FT).copy(); m0, v0, st0 = e.get_opt_state(L.OPT_FINETUNE)
    bad = flat.to(torch.int32).clone(); bad[17] = P * K + 3
    mb = e.ppo_step_indexed(_flat(obs), chains, olp, ret, val, adv, bad.cuda(), lr=1e-3, apply=True)
    assert torch.isnan(mb).all()
    with pytest.raises(L.DppoError):
        e.ppo_step_indexed(_flat(obs), chains, olp, ret, val, adv, bad.numpy(), lr=1e-3, apply=True, metrics_host=np.zeros(8, np.float32))
    m1_, v1_, _ = e.get_opt_state(L.OPT_FINETUNE)
    assert np.array_equal(e.get_weights(L.NET_ACTOR_FT), w0) and np.array_equal(m1_, m0) and np.array_equal(v1_, v0)
    e.set_opt_state(L.OPT_FINETUNE, m0, v0, st0)                      # the step counter did advance: restore it for the shared fixture
