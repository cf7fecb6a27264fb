"""GPU (-m gpu): the split-precision tcgen05 mode (DPPO_PREC_BF16X3) against the fp32 oracle at north_star's fp32 tolerance.

Every operand is carried as 16-bit planes (two fp16 planes = 22 mantissa bits in the actor's forward GEMMs, two bf16 planes = 16
bits in the smooth Mish critic, the backward and the weight-gradient GEMMs) and every product is the sum of the exact plane products
with fp32 accumulation in tensor memory (csrc/ts_path.cuh), so the bounds here are the fp32 mode's (tests/test_gpu_parity.py), not the
bf16 mode's.  Measured on a B200 in brackets:
  eps ............................... 5e-6 norm-wise relative [1.0e-6]; value 5e-5 [1.2e-5]
  per-step log-probs ................ 1e-3 absolute (north_star) [5.7e-6]
  PPO / pre-train gradients ......... 1e-3 of the largest gradient entry [1.2e-4 / 2.0e-6]
  sampled actions (20-step chain) ... 1e-4 norm-wise relative (north_star) [2.1e-5 max abs]
The tests also assert that tcgen05 GEMMs (not the FFMA path) produced the numbers.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from diffusionpolicyoptimization_b200 import _lib as L
from oracle import dppo_oracle as O
from helpers import make_engine, max_abs, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["hopper", "walker2d"])
def pair(request):
    o = O.make_oracle(request.param, seed=0)
    e = make_engine(o, precision=L.PREC_BF16X3)
    yield o, e
    e.close()


def _flat(obs):
    return obs.reshape(obs.shape[0], -1)


@pytest.mark.parametrize("a_mn,b_mn,M,N,K,splits,planes", [
    (0, 1, 300, 512, 576, 1, 2),   # forward layer: K-major activations x MN-major weights, ragged M, K = H + 64
    (0, 1, 300, 512, 576, 1, 3),   # ... with three planes (the actor's forward GEMMs)
    (0, 1, 1000, 256, 320, 1, 3),
    (0, 0, 1000, 24, 512, 1, 2),   # output layer: narrow N
    (0, 0, 1000, 24, 512, 1, 3),
    (0, 1, 300, 512, 576, 1, 4),   # planes = 4: two fp16 planes (B scaled by 2^10), two accumulators - the forward format
    (0, 0, 1000, 24, 512, 1, 4),   # the actor's output layer
    (0, 0, 260, 512, 64, 1, 2),    # dv = dout W3^T: one k-block
    (0, 0, 515, 256, 256, 1, 2),   # critic backward
    (1, 1, 512, 512, 5000, 7, 2),  # weight gradient: both MN-major, K = rows (ragged), split-K partials
    (1, 1, 64, 512, 3000, 5, 2),   # dW0 = h0^T du: M = 64
    (1, 1, 512, 64, 3000, 3, 2),   # dW3 = v^T dout: N = 64
])
def test_split_gemm_kernel_matches_float64(pair, a_mn, b_mn, M, N, K, splits, planes):
    """One ts::split_gemm_kernel launch on random fp32 operands vs a float64 product.  Two planes: the error is that of 16-bit
    operands (2^-17 relative per operand), far below one bf16 ulp (2^-9); three planes: fp32 operands exactly, what is left is
    the fp32 accumulation in tensor memory."""
    o, e = pair
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A = rng.standard_normal((M, K)).astype(np.float32)
    B = rng.standard_normal((K, N)).astype(np.float32)
    a_dev = torch.from_numpy(np.ascontiguousarray(A.T if a_mn else A)).cuda()
    lda = M if a_mn else K
    # B K-major = stored [N][K]; MN-major = stored [K][N].  Leading dimensions must keep rows 16-byte aligned in bf16: pad to 8
    if b_mn:
        ldb = (N + 7) // 8 * 8
        b_host = np.zeros((K, ldb), np.float32); b_host[:, :N] = B
    else:
        ldb = K
        rows = max(N, 32)
        b_host = np.zeros((rows, K), np.float32); b_host[:N] = B.T
    b_dev = torch.from_numpy(b_host).cuda()
    out = torch.zeros(splits, M, N, device="cuda")
    n0 = e.tc_launch_count()
    L.check(e.lib.dppo_debug_split_gemm(e.h, C.c_void_p(a_dev.data_ptr()), a_mn, lda, C.c_void_p(b_dev.data_ptr()), b_mn, ldb,
                                        M, N, K, splits, planes, C.c_void_p(out.data_ptr()), e._stream()), "dppo_debug_split_gemm")
    torch.cuda.synchronize()
    assert e.tc_launch_count() - n0 == 1
    got = out.sum(0).cpu().numpy().astype(np.float64)
    want = A.astype(np.float64) @ B.astype(np.float64)
    err = np.abs(got - want).max() / np.abs(want).max()
    print(f"split gemm planes={planes} a_mn={a_mn} b_mn={b_mn} {M}x{N}x{K}/{splits}: rel err {err:.2e}")
    assert err < (2e-5 if planes == 2 else 3e-6), err


@pytest.mark.parametrize("b_mn,M,N,K,planes,act", [
    (1, 300, 512, 64, 3, 1),       # forward L0 (actor): one k-block, ReLU + bit masks, ragged M
    (1, 1000, 512, 576, 3, 1),     # forward L1 / L2 (actor): three planes, two accumulators
    (1, 515, 256, 320, 2, 2),      # critic forward: two planes, Mish
    (0, 700, 512, 512, 2, 0),      # backward: K-major weights
    (0, 260, 512, 64, 2, 0),       # dv = dout W3^T
    (1, 40000, 512, 512, 3, 1),    # many tiles per pair (persistent loop, both accumulator buffers)
    (1, 300, 512, 64, 4, 1),       # planes = 4: the forward layers as they run - fp16 planes in and out, two accumulators
    (1, 1000, 512, 576, 4, 1),
    (1, 40000, 512, 512, 4, 1),
])
def test_pair_gemm_kernel_matches_float64(pair, b_mn, M, N, K, planes, act):
    """tsp::pair_gemm_kernel (cta_group::2, TMA-store epilogue) on random fp32 operands vs float64."""
    o, e = pair
    rng = np.random.default_rng(M + 3 * N + 5 * K)
    A = rng.standard_normal((M, K)).astype(np.float32)
    B = (rng.standard_normal((K, N)) / np.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    a_dev = torch.from_numpy(A).cuda()
    b_dev = torch.from_numpy(np.ascontiguousarray(B if b_mn else B.T)).cuda()
    bias_dev = torch.from_numpy(bias).cuda()
    out = torch.zeros(M, N, device="cuda")
    mask = torch.zeros(M, N // 32, device="cuda", dtype=torch.int32)
    n0 = e.tc_launch_count()
    L.check(e.lib.dppo_debug_pair_gemm(e.h, C.c_void_p(a_dev.data_ptr()), K, C.c_void_p(b_dev.data_ptr()), b_mn, N if b_mn else K,
                                       M, N, K, planes, C.c_void_p(bias_dev.data_ptr()), act, C.c_void_p(out.data_ptr()),
                                       C.c_void_p(mask.data_ptr()) if act == 1 else None, e._stream()), "dppo_debug_pair_gemm")
    torch.cuda.synchronize()
    assert e.tc_launch_count() - n0 == 1
    pre = A.astype(np.float64) @ B.astype(np.float64) + bias
    want = np.maximum(pre, 0) if act == 1 else (pre * np.tanh(np.logaddexp(0, pre)) if act == 2 else pre)
    err = np.abs(out.cpu().numpy() - want).max() / np.abs(want).max()
    print(f"pair gemm planes={planes} b_mn={b_mn} act={act} {M}x{N}x{K}: rel err {err:.2e}")
    assert err < (2e-5 if planes == 2 else 3e-6), err
    if act == 1:
        bits = mask.cpu().numpy().view(np.uint32)
        got = ((bits[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(M, N).astype(bool)
        sure = np.abs(pre) > 1e-4                     # away from the kink both must agree
        assert (got == (pre > 0))[sure].all()


def test_split_forward_value_logprobs(pair):
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
    n0 = e.tc_launch_count()
    got = e.actor_forward(L.NET_ACTOR_FT, x.reshape(N, -1), t, _flat(obs))
    gv = e.value(_flat(obs))
    torch.cuda.synchronize()
    assert e.tc_launch_count() - n0 == 6                       # three plane GEMMs per network: L0, L1, the folded [block.l2 ; output] layer
    err, errv = rel_err(got, want.reshape(N, -1)), rel_err(gv, wantv)
    print(f"bf16x3 eps rel err {err:.3e}, value rel err {errv:.3e}")
    assert err < 5e-6 and errv < 5e-5
    # get_logprobs over whole chains (diffusion_vpg.py:343-425)
    B = 512
    obs_b, x_T, noise = O.make_rollout_inputs(o, B, seed=3)
    chains = o.sample(obs_b, x_T, noise).chains
    with torch.no_grad():
        want_lp = o.get_logprobs(obs_b, chains).reshape(B * d.ft_denoising_steps, -1)
    lp = e.logprobs(_flat(obs_b), chains.reshape(B, d.ft_denoising_steps + 1, -1))
    err_lp = max_abs(lp, want_lp)
    print(f"bf16x3 log-prob abs err {err_lp:.3e}")
    assert err_lp < 1e-3


def test_split_ppo_step_matches_oracle(pair):
    o, e = pair
    N = 4099                                                    # ragged: not a multiple of the 128-row tile
    batch = O.make_ppo_batch(o, N, pool=512, seed=3)
    metrics, ga, gc = o.ppo_grads(*batch)
    n0 = e.tc_launch_count()
    m, g = e.ppo_step(_flat(batch[0]), batch[1].reshape(N, -1), batch[2].reshape(N, -1), batch[3], batch[4], batch[5], batch[6],
                      batch[7].reshape(N, -1), lr=0.0, apply=False, want_grads=True)
    torch.cuda.synchronize()
    assert e.tc_launch_count() - n0 == 2 * (3 + 2) + 1          # per net: 3 forward + 2 backward plane GEMMs (dv = dout W3^T is never formed); one grouped weight-gradient launch
    wg = np.concatenate([O.flatten_params(ga), O.flatten_params(gc)])
    g = g.cpu().numpy()
    nA = e.n_actor
    err_a = float(np.abs(g[:nA] - wg[:nA]).max() / np.abs(wg[:nA]).max())
    err_c = float(np.abs(g[nA:] - wg[nA:]).max() / np.abs(wg[nA:]).max())
    print(f"bf16x3 PPO grads: actor {err_a:.3e} critic {err_c:.3e} of max")
    assert err_a < 1e-3 and err_c < 1e-3
    np.testing.assert_allclose(m.cpu().numpy(), [float(x) for x in metrics], rtol=2e-3, atol=2e-6)
    # bit-reproducible (fixed-order split-K reduction)
    m2, g2 = e.ppo_step(_flat(batch[0]), batch[1].reshape(N, -1), batch[2].reshape(N, -1), batch[3], batch[4], batch[5], batch[6],
                        batch[7].reshape(N, -1), lr=0.0, apply=False, want_grads=True)
    assert torch.equal(g2.cpu(), torch.from_numpy(g)) and torch.equal(m2, m)


def test_split_ppo_step_applies_adamw_and_refreshes_operands(pair):
    """After the update the split operand copies must be rebuilt from the new master weights: the next forward sees them."""
    o, e = pair
    N = 2560
    batch = O.make_ppo_batch(o, N, pool=256, seed=4)
    fb = (_flat(batch[0]), batch[1].reshape(N, -1), batch[2].reshape(N, -1), batch[3], batch[4], batch[5], batch[6], batch[7].reshape(N, -1))
    w_ft0, w_c0 = e.get_weights(L.NET_ACTOR_FT), e.get_weights(L.NET_CRITIC)
    try:
        lp0 = e.logprobs_subsample(fb[0], fb[1], fb[2], fb[3]).clone()
        e.ppo_step(*fb, lr=1e-3, apply=True)
        lp1 = e.logprobs_subsample(fb[0], fb[1], fb[2], fb[3])
        assert float((lp1 - lp0).abs().max()) > 1e-4
        # same weights in a fresh fp32 engine give the same log-probs within the parity tolerance
        e32 = make_engine(o, precision=L.PREC_FP32)
        e32.set_weights(L.NET_ACTOR_FT, e.get_weights(L.NET_ACTOR_FT)); e32.set_weights(L.NET_CRITIC, e.get_weights(L.NET_CRITIC))
        lp32 = e32.logprobs_subsample(fb[0], fb[1], fb[2], fb[3])
        err_lp, err_v = max_abs(lp1, lp32), rel_err(e.value(fb[0]), e32.value(fb[0]))
        e32.close()
        assert err_lp < 1e-4 and err_v < 5e-5, (err_lp, err_v)
    finally:
        e.set_weights(L.NET_ACTOR_FT, w_ft0); e.set_weights(L.NET_CRITIC, w_c0)     # the fixture is shared
        e.set_opt_state(L.OPT_FINETUNE, np.zeros(e.n_actor + e.n_critic, np.float32), np.zeros(e.n_actor + e.n_critic, np.float32), 0)


def test_split_pretrain_and_large_batch_sampler(pair):
    o, e = pair
    d = o.d
    N = 2304
    rng = np.random.default_rng(5)
    acts = torch.from_numpy(rng.uniform(-1, 1, (N, d.horizon_steps, d.action_dim)).astype(np.float32))
    st = torch.from_numpy(rng.uniform(-1, 1, (N, 1, d.obs_dim)).astype(np.float32))
    tt = torch.from_numpy(rng.integers(0, d.denoising_steps, N))
    nz = torch.from_numpy(rng.standard_normal((N, d.horizon_steps, d.action_dim)).astype(np.float32))
    want_l, want_g = o.pretrain_grads(acts, st, tt, nz)
    n0 = e.tc_launch_count()
    loss, pg = e.pretrain_step(acts.reshape(N, -1), _flat(st), lr=1e-3, apply=False, t=tt, noise=nz.reshape(N, -1), want_grads=True)
    torch.cuda.synchronize()
    assert e.tc_launch_count() - n0 == 6                        # 3 forward + 2 backward + the grouped weight-gradient launch
    wg = O.flatten_params(want_g)
    err_g = float(np.abs(pg.cpu().numpy() - wg).max() / np.abs(wg).max())
    print(f"bf16x3 pre-train: loss {float(loss):.6f} vs {float(want_l):.6f}, grad {err_g:.3e} of max")
    assert abs(float(loss) - float(want_l)) < 1e-5 * abs(float(want_l)) and err_g < 1e-3
    # VPGDiffusion.call for a large batch: T x (4 split GEMMs + update), injected noise
    B = 2304
    obs, x_T, noise = O.make_rollout_inputs(o, B, seed=6)
    want = o.sample(obs, x_T, noise)
    n0 = e.tc_launch_count()
    a, ch = e.sample(_flat(obs), x_T=x_T.reshape(B, -1), noise=noise.reshape(d.denoising_steps, B, -1))
    torch.cuda.synchronize()
    assert e.last_path() == 3 and e.tc_launch_count() - n0 == 3 * d.denoising_steps     # L0, L1 and the folded [block.l2 ; output] layer per step
    diff = (a.cpu() - want.trajectories.reshape(B, -1)).abs()
    print(f"bf16x3 sampler: max {float(diff.max()):.2e} mean {float(diff.mean()):.2e}")
    assert rel_err(a, want.trajectories.reshape(B, -1)) < 1e-4 and float(diff.mean()) < 1e-6
    assert float((ch.cpu() - want.chains.reshape(B, d.ft_denoising_steps + 1, -1)).abs().mean()) < 1e-6


def test_split_index_driven_update_reads_the_rollout_buffers_directly(pair):
    """dppo_ppo_step_indexed in the bf16x3 mode: no materialised minibatch (h0 pack, advantage statistics and loss kernel address the
    resident rollout through the flat indices); bit-identical to dppo_ppo_step on the rows gathered on the host (train_ppo_diffusion_agent.py:292-312)."""
    o, e = pair
    d = o.d
    P, K, A, N = 600, d.ft_denoising_steps, d.A, 4096
    obs, x_T, noise = O.make_rollout_inputs(o, P, seed=77)
    chains = o.sample(obs, x_T, noise).chains.reshape(P, K + 1, A)
    g = torch.Generator().manual_seed(5)
    olp = torch.randn(P, K, A, generator=g) * 0.3 - 1.0
    ret, val, adv = torch.randn(P, generator=g), torch.randn(P, generator=g), torch.randn(P, generator=g)
    flat = torch.randint(0, P * K, (N,), generator=g, dtype=torch.int64)
    b, k = flat // K, flat % K
    e.ppo_step_indexed(_flat(obs), chains, olp, ret, val, adv, flat.to(torch.int32).cuda(), lr=0.0, apply=False)
    # (the folded output layers of the current weights are now built: they cost one small launch per net after a weight change)
    n0 = e.launch_count()
    m1, g1 = e.ppo_step_indexed(_flat(obs), chains, olp, ret, val, adv, flat.to(torch.int32).cuda(), lr=0.0, apply=False, want_grads=True)
    launches_indexed = e.launch_count() - n0
    n1 = e.launch_count()
    m2, g2 = e.ppo_step(_flat(obs)[b], chains[b, k], chains[b, k + 1], k.to(torch.int32), ret[b], val[b], adv[b], olp[b, k],
                        lr=0.0, apply=False, want_grads=True)
    torch.cuda.synchronize()
    assert launches_indexed == e.launch_count() - n1 + 1         # + the NaN-on-bad-index kernel of the device variant, no gather kernel
    assert torch.equal(m1, m2) and torch.equal(g1, g2)
    w0 = e.get_weights(L.NET_ACTOR_FT).copy()
    bad = flat.to(torch.int32).clone(); bad[3] = -1
    assert torch.isnan(e.ppo_step_indexed(_flat(obs), chains, olp, ret, val, adv, bad.cuda(), lr=1e-3, apply=True)).all()
    np.testing.assert_array_equal(e.get_weights(L.NET_ACTOR_FT), w0)
    m0, v0, s0 = e.get_opt_state(L.OPT_FINETUNE)
    e.set_opt_state(L.OPT_FINETUNE, m0, v0, s0 - 1)              # the skipped update still counted a step: undo for the shared fixture
