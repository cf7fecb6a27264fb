"""CPU: the pure-Python `.weights.h5` reader / writer (diffusionpolicyoptimization_b200/util/keras_h5.py).

h5py and Keras are not installed in the build image, so what can be pinned here is (a) the HDF5 structures against the file
format specification by an independent walk of the bytes (signatures, sizes, alignment, B-tree keys) and (b) the round trip
through the package's own reader, including groups with more links than one symbol-table node holds.  The Keras variable
paths are the documented ASSUMPTION of util/keras_h5.py."""
import os
import struct

import numpy as np
import pytest

from diffusionpolicyoptimization_b200.util import keras_h5 as KH


def test_round_trip_and_structure(tmp_path):
    rng = np.random.default_rng(0)
    data = {p: rng.standard_normal(shp).astype(np.float32)
            for p, shp in zip(KH.keras_paths_diffusion_mlp("actor"),
                              [(16, 32), (32,), (32, 16), (16,), (57, 512), (512,), (512, 512), (512,), (512, 512), (512,), (512, 24), (24,)])}
    data["critic/Q1/output_layer/vars/0"] = rng.standard_normal((256, 1)).astype(np.float32)
    for i in range(19):                                           # a group with more links than one symbol-table node holds (8)
        data[f"many/d{i:02d}"] = np.full((3,), i, np.float64)
    data["scalar_like/empty"] = np.zeros((0,), np.float32)
    path = str(tmp_path / "x.weights.h5")
    KH.write_h5(path, data)
    raw = open(path, "rb").read()
    # superblock version 0: signature, 8-byte offsets / lengths, group K values, end-of-file address = file size
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and raw[8] == 0 and raw[13] == 8 and raw[14] == 8
    assert struct.unpack_from("<HH", raw, 16) == (4, 16)
    assert struct.unpack_from("<Q", raw, 40)[0] == len(raw)
    root_hdr = struct.unpack_from("<Q", raw, 64)[0]
    assert root_hdr % 8 == 0 and raw[root_hdr] == 1               # object header version 1, 8-byte aligned
    assert raw.count(b"TREE") >= 1 and raw.count(b"SNOD") >= 3 and raw.count(b"HEAP") >= 1
    r = KH.H5Reader(path)
    assert set(r.datasets) == set(data)
    for k, v in data.items():
        assert r.datasets[k].dtype == v.dtype and r.datasets[k].shape == v.shape
        np.testing.assert_array_equal(r.datasets[k], v)
    assert [k for k in r.datasets if k.startswith("many/")] == sorted(k for k in data if k.startswith("many/"))   # name order


def test_keras_paths_and_loader_fallback(tmp_path):
    paths = KH.keras_paths_diffusion_mlp()
    assert paths[0] == "time_embedding/layers/dense/vars/0" and paths[3] == "time_embedding/layers/dense_1/vars/1"
    assert paths[6] == "mlp_mean/residual_blocks/two_layer_pre_activation_res_net_linear/l1/vars/0" and paths[-1] == "mlp_mean/output_layer/vars/1"
    assert KH.keras_paths_critic_obs("critic")[0] == "critic/Q1/input_layer/vars/0"
    shapes = [(16, 32), (32,), (32, 16), (16,), (39, 64), (64,), (64, 64), (64,), (64, 64), (64,), (64, 12), (12,)]
    rng = np.random.default_rng(1)
    ws = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    p1 = str(tmp_path / "a.weights.h5")
    KH.save_keras_weights_h5(p1, ws, paths)
    for a, b in zip(KH.load_keras_weights_h5(p1, paths, shapes), ws):
        np.testing.assert_array_equal(a, b)
    # a file with other group names (another Keras version): matched by shape in file order, distinct shapes only
    p2 = str(tmp_path / "b.weights.h5")
    KH.write_h5(p2, {f"layers/x{i:02d}/vars/0": w for i, w in enumerate(ws)})
    got = KH.load_keras_weights_h5(p2, paths, shapes)
    for i in (0, 2, 4, 10, 11):                                   # shapes that occur once are recovered exactly
        np.testing.assert_array_equal(got[i], ws[i])
    with pytest.raises(KeyError):
        KH.load_keras_weights_h5(p2, paths, shapes[:-1] + [(7,)])


def test_network_save_load_paths(tmp_path):
    """_Net.save_weights / load_weights: `.npz` gets its suffix on both sides; `.h5` goes through the Keras layout."""
    from diffusionpolicyoptimization_b200.model.diffusion.mlp_diffusion import DiffusionMLP
    a = DiffusionMLP(action_dim=3, horizon_steps=4, cond_dim=11, mlp_dims=[64, 64, 64], activation_type="ReLU", residual_style=True, seed=3)
    b = DiffusionMLP(action_dim=3, horizon_steps=4, cond_dim=11, mlp_dims=[64, 64, 64], activation_type="ReLU", residual_style=True, seed=4)
    for name in ("ckpt", "ckpt2.npz", "state_7.weights.h5"):
        a.save_weights(str(tmp_path / name))
        b.load_weights(str(tmp_path / name))
        for x, y in zip(a.get_weights(), b.get_weights()):
            np.testing.assert_array_equal(x, y)
        b.set_flat_weights(np.zeros(b.num_params(), np.float32))
    assert os.path.exists(str(tmp_path / "ckpt.npz"))
