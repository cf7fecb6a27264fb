"""GPU (-m gpu): BASELINE.json's full sizes (walker2d shapes, 50 000-row PPO minibatch, 18 944-env rollout) through
size-independent properties - the oracle cannot run these sizes in seconds, so the checks are:
  * additivity over row shards: grads(A u B) == grads(A) + grads(B) when every call divides by the same N_global
    (the contract the multi-GPU path and the chunked host pipeline rely on);
  * the metrics of the union are the sums of the shards' metric partials;
  * determinism of the Philox sampler and of forward passes; idempotence of apply=False steps;
  * the chain recorded by the sampler reproduces itself: log-probs of a sampled chain under the sampling weights are finite,
    and re-sampling with injected noise equal to the normalised residuals returns the same chain (round trip)."""
import numpy as np
import pytest
import torch

from diffusionpolicyoptimization_b200 import _lib as L
from oracle import dppo_oracle as O
from helpers import make_engine

pytestmark = pytest.mark.gpu
N = 50_000


@pytest.fixture(scope="module", params=["bf16", "fp32"])
def eng(request):
    o = O.make_oracle("walker2d", seed=0)
    e = make_engine(o, precision=L.PREC_BF16 if request.param == "bf16" else L.PREC_FP32)
    e.mode = request.param
    yield o, e
    e.close()


def _batch(e, n, seed):
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    P, K = 4096, e.K
    obs = torch.rand(P, e.Do, device="cuda", generator=g) * 2 - 1
    _, chains = e.sample(obs, seed=seed, offset=1)
    olp = e.logprobs(obs, chains, use_base_policy=True).reshape(P, K, e.A)
    val = e.value(obs)
    flat = torch.randint(0, P * K, (n,), device="cuda", generator=g)
    b, k = flat // K, flat % K
    return [obs[b].contiguous(), chains[b, k].contiguous(), chains[b, k + 1].contiguous(), k.to(torch.int32).contiguous(),
            torch.randn(n, device="cuda", generator=g), (val[b] + 0.1 * torch.randn(n, device="cuda", generator=g)).contiguous(),
            torch.randn(n, device="cuda", generator=g), olp[b, k].contiguous()]


def test_ppo_gradients_are_additive_over_row_shards_at_full_size(eng):
    o, e = eng
    batch = _batch(e, N, seed=1)
    mean, std = float(batch[6].mean()), float(batch[6].std(unbiased=False))
    m_all, g_all = e.ppo_step(*batch, lr=0.0, apply=False, n_global=N, adv_mean=mean, adv_std=std, want_grads=True)
    cut = 23_041                                   # ragged split, neither part a multiple of 128
    parts = []
    for lo, hi in ((0, cut), (cut, N)):
        sl = [t[lo:hi].contiguous() for t in batch]
        parts.append(e.ppo_step(*sl, lr=0.0, apply=False, n_global=N, adv_mean=mean, adv_std=std, want_grads=True))
    g_sum = parts[0][1] + parts[1][1]
    m_sum = parts[0][0] + parts[1][0]
    scale = float(g_all.abs().max())
    # fp32 mode: only the summation order differs; bf16 mode: identical bf16 operands per row, fp32 accumulation order differs
    assert float((g_sum - g_all).abs().max()) < 2e-4 * scale
    np.testing.assert_allclose(m_sum.cpu().numpy(), m_all.cpu().numpy(), rtol=1e-4, atol=1e-6)
    # idempotence of apply=False
    m2, g2 = e.ppo_step(*batch, lr=0.0, apply=False, n_global=N, adv_mean=mean, adv_std=std, want_grads=True)
    assert float((g2 - g_all).abs().max()) < 2e-5 * scale and torch.allclose(m2, m_all, rtol=1e-5, atol=1e-7)


def test_forward_passes_are_deterministic_and_row_independent_at_full_size(eng):
    o, e = eng
    batch = _batch(e, N, seed=2)
    lp1 = e.logprobs_subsample(batch[0], batch[1], batch[2], batch[3])
    lp2 = e.logprobs_subsample(batch[0], batch[1], batch[2], batch[3])
    assert torch.equal(lp1, lp2) and torch.isfinite(lp1).all()
    perm = torch.randperm(N, device="cuda")
    lp3 = e.logprobs_subsample(batch[0][perm].contiguous(), batch[1][perm].contiguous(), batch[2][perm].contiguous(), batch[3][perm].contiguous())
    if e.mode == "bf16":
        assert torch.equal(lp3, lp1[perm])          # a row's result does not depend on its tile or position
    else:
        assert float((lp3 - lp1[perm]).abs().max()) < 1e-5
    v1 = e.value(batch[0]); v2 = e.value(batch[0][perm].contiguous())
    assert float((v2 - v1[perm]).abs().max()) < (0.0 if e.mode == "bf16" else 1e-5) + 1e-30


def test_sampler_round_trip_at_full_size(eng):
    """Sample with Philox, recover the clipped normalised noise of the fine-tuned steps from the recorded chain
    (x_{k+1} = mu_k + sigma_k * eps_k), and check |eps| <= randn_clip_value and that the chain's own log-probs equal
    -eps^2/2 - log(sigma sqrt(2 pi)) within the mode's tolerance."""
    o, e = eng
    B = 148 * 128
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    obs = torch.rand(B, e.Do, device="cuda", generator=g) * 2 - 1
    a1, c1 = e.sample(obs, seed=11, offset=5)
    a2, c2 = e.sample(obs, seed=11, offset=5)
    assert torch.equal(a1, a2) and torch.equal(c1, c2)                  # counter-based noise: bit-reproducible
    a3, _ = e.sample(obs, seed=11, offset=6)
    assert not torch.equal(a1, a3)
    assert torch.equal(c1[:, -1], a1) and torch.isfinite(c1).all()
    lp = e.logprobs(obs, c1).reshape(B, e.K, e.A)
    sched = O.schedule_table(o.d.denoising_steps)
    logvar = torch.tensor(sched[O.SCHEDULE_ROWS.index("ddpm_logvar_clipped")], device="cuda")
    t = torch.arange(e.K - 1, -1, -1, device="cuda")
    sd = torch.exp(0.5 * logvar[t]).clamp(min=o.h.min_logprob_denoising_std)            # [K]
    z2 = -2.0 * (lp + torch.log(sd)[None, :, None] + 0.9189385332046727)                 # = eps^2 when sampling std == log-prob std
    same_std = sd >= o.h.min_sampling_denoising_std - 1e-7
    tol = 0.5 if e.mode == "bf16" else 2e-3
    z2s = z2[:, same_std]
    assert float(z2s.min()) > -tol and float(z2s.max()) < o.h.randn_clip_value ** 2 + (3.0 if e.mode == "bf16" else 0.05)
    # the noise is a standard normal clamped to +-3: E[eps^2] = (2 Phi(3) - 1) - 6 phi(3) + 9 * 2 Q(3) = 0.9950
    assert abs(float(z2s.mean()) - 0.995) < (0.05 if e.mode == "bf16" else 0.01)
