"""Learning-rate schedule of the fine-tuning agents, host side: util/scheduler.py:57-176 (`CosineAnnealingWarmupRestarts2`, the
class `TrainPPOAgent` hands to the Keras optimizer, train_ppo_agent.py:35-43).  Called with the optimizer's iteration count.

Within a cycle of length L (restart c, position s): linear warm-up from `initial_learning_rate` to the cycle's peak over
`warmup_steps`, then half a cosine back down to `initial_learning_rate` (NOT to `min_lr`: the reference's formula never uses
it); the peak decays by `gamma` per cycle, cycles grow by `cycle_mult`.  With the shipped cfg (initial = max) it is constant.
"""
import math


class CosineAnnealingWarmupRestarts2:
    def __init__(self, initial_learning_rate, first_cycle_steps, cycle_mult=1.0, max_lr=0.1, min_lr=0.001, warmup_steps=0,
                 gamma=1.0, last_epoch=-1):
        assert warmup_steps < first_cycle_steps
        self.initial_learning_rate, self.first_cycle_steps, self.cycle_mult = initial_learning_rate, first_cycle_steps, cycle_mult
        self.base_max_lr, self.min_lr, self.warmup_steps, self.gamma = max_lr, min_lr, warmup_steps, gamma
        self.last_epoch = last_epoch

    def _locate(self, step):
        """-> (cycle index, step inside the cycle, cycle length)"""
        L0, m = self.first_cycle_steps, self.cycle_mult
        if step < L0:
            return 0, step, L0
        if m == 1.0:
            return step // L0, step % L0, L0
        n = int(math.log(step / L0 * (m - 1) + 1, m))
        return n, step - int(L0 * (m ** n - 1) / (m - 1)), L0 * m ** n

    def __call__(self, step):
        step = int(step)
        cycle, s, length = self._locate(step)
        peak = self.base_max_lr * self.gamma ** cycle
        self.last_epoch = step
        lo = self.initial_learning_rate
        if s < self.warmup_steps:
            return (peak - lo) * s / self.warmup_steps + lo
        return lo + (peak - lo) * (1 + math.cos(math.pi * (int(s) - self.warmup_steps) / (length - self.warmup_steps))) / 2
