"""Generates tests/golden/ref_scheduler.npz from the reference's util/scheduler.py (imported unmodified; its Keras base class comes
from the TF shim - the class is plain Python arithmetic).   python tests/golden/make_ref_scheduler.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("DPPO_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "tf_shim"))
sys.path.insert(0, REF)
from util.scheduler import CosineAnnealingWarmupRestarts2  # noqa: E402  (reference)

CASES = {
    "yaml": dict(initial_learning_rate=1e-4, first_cycle_steps=1000, cycle_mult=1.0, max_lr=1e-4, min_lr=1e-4, warmup_steps=10, gamma=1.0),
    "warm": dict(initial_learning_rate=1e-5, first_cycle_steps=50, cycle_mult=1.0, max_lr=1e-3, min_lr=1e-4, warmup_steps=5, gamma=0.8),
    "grow": dict(initial_learning_rate=2e-5, first_cycle_steps=20, cycle_mult=2.0, max_lr=5e-4, min_lr=1e-5, warmup_steps=3, gamma=0.5),
}
out = {}
steps = np.arange(0, 400)
for name, kw in CASES.items():
    sch = CosineAnnealingWarmupRestarts2(**kw)
    out[name + "_lr"] = np.array([sch(int(t)) for t in steps], np.float64)
    out[name + "_kw"] = np.array([kw[k] for k in ("initial_learning_rate", "first_cycle_steps", "cycle_mult", "max_lr", "min_lr", "warmup_steps", "gamma")], np.float64)
out["steps"] = steps
np.savez_compressed(os.path.join(HERE, "ref_scheduler.npz"), **out)
print({k: v.shape for k, v in out.items()})
