"""Dev tool: a longer run of the fine-tuning loop (tensor mode, library Philox noise) on the toy env - finite losses, value loss falling."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import diffusionpolicyoptimization_b200 as dp
from diffusionpolicyoptimization_b200.agent.finetune.train_ppo_diffusion_agent import TrainPPODiffusionAgent
from diffusionpolicyoptimization_b200.util.scheduler import CosineAnnealingWarmupRestarts2
from oracle import dppo_oracle as O
from toy_env import ToyVecEnv
from test_gpu_agent import make_model
o = O.make_oracle("hopper", seed=5)
model = make_model(o, precision=(sys.argv[1] if len(sys.argv) > 1 else "bf16"))
E, S = 256, 10
sched = CosineAnnealingWarmupRestarts2(1e-4, 1000, 1.0, 1e-4, 1e-4, 10, 1.0)
agent = TrainPPODiffusionAgent(model, ToyVecEnv(E, 11, 3, max_episode_steps=5, seed=9), n_envs=E, n_steps=S, act_steps=4, n_train_itr=12,
                               batch_size=6400, update_epochs=3, actor_lr=sched, val_freq=4, max_grad_norm=1.0)
for r in agent.run():
    if r["eval_mode"]:
        print(f"itr {r['itr']:2d} eval  reward {r['avg_episode_reward']:.3f} episodes {r['num_episode_finished']}")
    else:
        print(f"itr {r['itr']:2d} train loss {r['loss']:.4f} pg {r['pg_loss']:+.5f} v {r['v_loss']:.4f} kl {r['approx_kl']:.2e} clipfrac {r['clipfrac']:.3f} "
              f"ev {r['explained_var']:+.3f} reward {r['avg_episode_reward']:.3f} updates {r['n_updates']} t {r['time']*1e3:.0f} ms")
        assert np.isfinite(r["loss"])
