"""Network geometry containers.  The reference builds Keras layers (model/common/mlp.py:95-206);
here a network is only its flat fp32 parameter vector in Keras variable-creation order — the
arithmetic lives in libdppo_b200.so."""
import math
from typing import List, Sequence, Tuple

import numpy as np

activation_ids = {"ReLU": 0, "Mish": 1}   # model/common/mlp.py:6-14 (the two any cfg uses)


def residual_mlp_shapes(dim_list: Sequence[int]) -> List[Tuple[int, ...]]:
    """ResidualMLP(dim_list) variable shapes (mlp.py:102-139).  Only the one-block geometry
    [in, H, H, H, out] used by every reference cfg is supported by the kernels."""
    if len(dim_list) != 5 or not (dim_list[1] == dim_list[2] == dim_list[3]):
        raise ValueError("libdppo_b200 supports ResidualMLP with dim_list [in, H, H, H, out] (one residual block)")
    i, h, o = dim_list[0], dim_list[1], dim_list[4]
    return [(i, h), (h,), (h, h), (h,), (h, h), (h,), (h, o), (o,)]


def glorot_init(shapes, rng: np.random.Generator) -> List[np.ndarray]:
    """Keras Dense defaults: glorot_uniform kernels, zero biases."""
    out = []
    for s in shapes:
        if len(s) == 2:
            lim = math.sqrt(6.0 / (s[0] + s[1]))
            out.append(rng.uniform(-lim, lim, size=s).astype(np.float32))
        else:
            out.append(np.zeros(s, np.float32))
    return out


class _Net:
    """Weight list <-> flat vector; bound to an Engine net id once owned by a diffusion model."""

    def __init__(self, shapes, seed=None):
        self.shapes = [tuple(s) for s in shapes]
        self._host = glorot_init(self.shapes, np.random.default_rng(seed))
        self._engine, self._net = None, None

    def num_params(self):
        return int(sum(int(np.prod(s)) for s in self.shapes))

    def _bind(self, engine, net):
        self._engine, self._net = engine, net
        engine.set_weights(net, self._flat(self._host))
        self._host = None

    @staticmethod
    def _flat(ws):
        return np.concatenate([np.asarray(w, np.float32).reshape(-1) for w in ws])

    def get_weights(self):
        flat = self._flat(self._host) if self._engine is None else self._engine.get_weights(self._net)
        out, off = [], 0
        for s in self.shapes:
            n = int(np.prod(s))
            out.append(flat[off:off + n].reshape(s).copy())
            off += n
        return out

    def set_weights(self, weights):
        weights = [np.asarray(w, np.float32) for w in weights]
        assert [tuple(w.shape) for w in weights] == self.shapes, "weight shapes do not match the network"
        if self._engine is None:
            self._host = weights
        else:
            self._engine.set_weights(self._net, self._flat(weights))

    def get_flat_weights(self):
        return self._flat(self.get_weights())

    def set_flat_weights(self, flat):
        flat = np.asarray(flat, np.float32).reshape(-1)
        assert flat.size == self.num_params()
        if self._engine is None:
            out, off = [], 0
            for s in self.shapes:
                n = int(np.prod(s))
                out.append(flat[off:off + n].reshape(s).copy())
                off += n
            self._host = out
        else:
            self._engine.set_weights(self._net, flat)

    # reference checkpoints are Keras .weights.h5 (finetune/train_agent.py:127-142); h5py is not a
    # dependency here, so flat .npz files carry the same variable order.
    @staticmethod
    def _npz_path(path):
        path = str(path)
        return path if path.endswith(".npz") else path + ".npz"        # np.savez appends the suffix: save and load must agree

    def save_weights(self, path):
        path = str(path)
        if path.endswith(".h5"):
            from ...util.keras_h5 import save_keras_weights_h5
            return save_keras_weights_h5(path, self.get_weights(), self.keras_variable_paths())
        np.savez(self._npz_path(path), **{f"v{i}": w for i, w in enumerate(self.get_weights())})

    def load_weights(self, path):
        path = str(path)
        if path.endswith(".h5"):
            from ...util.keras_h5 import load_keras_weights_h5
            return self.set_weights(load_keras_weights_h5(path, self.keras_variable_paths(), self.shapes))
        z = np.load(self._npz_path(path))
        self.set_weights([z[f"v{i}"] for i in range(len(self.shapes))])
