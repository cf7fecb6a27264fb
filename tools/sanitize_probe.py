"""Small run of every split-precision program in one process (PPO update with the folded forward, grouped dW, dW3 assembly and tail; log-prob forward;
large-batch sampler; 40-env cluster sampler; pre-train step): a quick smoke for memory checkers / debug builds.
    python tools/sanitize_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import dppo_oracle as O
from helpers import make_engine
from diffusionpolicyoptimization_b200 import _lib as L

o = O.make_oracle("hopper", seed=0)
e = make_engine(o, precision=L.PREC_BF16X3)
N = 2304
b = O.make_ppo_batch(o, N, pool=256, seed=1)
args = [x.cuda() for x in b]
for i in (0, 1, 2, 7): args[i] = args[i].reshape(N, -1).contiguous()
for _ in range(2):
    m = e.ppo_step(*args, lr=1e-4, apply=True)
lp = e.logprobs_subsample(args[0], args[1], args[2], args[3])
obs = torch.rand(N, o.d.Do, device="cuda") * 2 - 1
a, c = e.sample(obs, seed=1, offset=2)
a40, c40 = e.sample(obs[:40], seed=1, offset=2)
acts = torch.rand(N, o.d.A, device="cuda") * 2 - 1
loss = e.pretrain_step(acts, obs, lr=1e-4, apply=True)
torch.cuda.synchronize()
print("ok", float(m[0]), float(lp.mean()), float(a.abs().mean()), float(a40.abs().mean()), e.last_path())
e.close()
