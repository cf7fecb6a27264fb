"""GPU (-m gpu): the tcgen05 GEMM kernel in isolation against torch.matmul on the same bf16 operands
(fp32 reference of the op; products of bf16 values are exact in fp32, so only summation order differs)."""
import ctypes as C

import numpy as np
import pytest
import torch

import diffusionpolicyoptimization_b200 as dp
from diffusionpolicyoptimization_b200 import _lib as L

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = dp.Engine(dp.default_cfg(), 0)
    yield e
    e.close()


def run(e, A, a_mn, B, b_mn, M, N, K, splits=1, bias=None, act=0, A2=None, K2=0, want_bf16=False):
    out = torch.full((splits, M, N), float("nan"), device="cuda")
    ob = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16) if want_bf16 else None
    p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    rc = e.lib.dppo_debug_tc_gemm(e.h, p(A), int(a_mn), A.stride(0), p(A2), 0 if A2 is None else A2.stride(0), K2,
                                  p(B), int(b_mn), B.stride(0), M, N, K, splits, p(bias), act, p(out), p(ob), e._stream())
    L.check(rc, "dppo_debug_tc_gemm")
    torch.cuda.synchronize()
    return out, ob


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 512), (1000, 512, 576), (300, 64, 200), (4099, 24, 512)])
def test_gemm_layouts(eng, a_mn, b_mn, M, N, K):
    if N < 64 and b_mn:
        Npad = 64
    else:
        Npad = N
    g = torch.Generator(device="cuda"); g.manual_seed(M + N + K)
    Af = torch.randn(M, K, device="cuda", generator=g)
    Bf = torch.randn(K, Npad, device="cuda", generator=g)
    Ab, Bb = Af.to(torch.bfloat16), Bf.to(torch.bfloat16)
    want = Ab.float() @ Bb.float()
    A = Ab.t().contiguous() if a_mn else Ab.contiguous()            # [K][M] or [M][K]
    B = Bb.contiguous() if b_mn else Bb.t().contiguous()            # [K][N] or [N][K]
    # leading dimensions must be multiples of 8 elements (16 B): pad K/M/N strides where needed
    def pad_ld(t):
        ld = (t.shape[1] + 7) // 8 * 8
        buf = torch.zeros(t.shape[0], ld, device="cuda", dtype=torch.bfloat16)
        buf[:, :t.shape[1]] = t
        return buf[:, :t.shape[1]]
    A, B = pad_ld(A), pad_ld(B)
    out, _ = run(eng, A, a_mn, B, b_mn, M, Npad, K)
    err = (out[0] - want).abs().max().item() / max(want.abs().max().item(), 1e-6)
    assert err < 2e-5, err


def test_gemm_splitk_bias_act_concat(eng):
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    M, N, K, K2 = 700, 512, 512, 64
    A = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    A2 = torch.randn(M, K2, device="cuda", generator=g).to(torch.bfloat16)
    B = (torch.randn(K + K2, N, device="cuda", generator=g) * 0.05).to(torch.bfloat16)   # MN-major [K][N]
    bias = torch.randn(N, device="cuda", generator=g)
    want = torch.relu(torch.cat([A, A2], 1).float() @ B.float() + bias)
    out, ob = run(eng, A, 0, B, 1, M, N, K, bias=bias, act=1, A2=A2, K2=K2, want_bf16=True)
    assert (out[0] - want).abs().max().item() < 2e-4
    assert (ob.float() - want).abs().max().item() < 0.02 * want.abs().max().item()
    # split-K over the row dimension (the dW = X^T D shape): partials sum to the product
    R = 5000
    X = torch.randn(R, 512, device="cuda", generator=g).to(torch.bfloat16)
    D = torch.randn(R, 256, device="cuda", generator=g).to(torch.bfloat16)
    out, _ = run(eng, X, 1, D, 1, 512, 256, R, splits=7)
    want = X.float().t() @ D.float()
    assert not torch.isnan(out).any()
    assert (out.sum(0) - want).abs().max().item() < 1e-3 * want.abs().max().item()
