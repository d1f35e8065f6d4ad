"""Algebraic model (BASELINE config #5) on CPU: oracle and torch-backend mirror vs vectors from the unmodified reference."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_err
from oracle import vol_oracle as O
import lt_b200
from lt_b200 import testing


@pytest.fixture(scope="module")
def alg_case():
    B, V, S = 2, 4, 128
    model = lt_b200.AlgebraicTriangulationNet(testing.make_alg_config(num_layers=50), device="cpu", backend="torch")
    testing.randomize_backbone_weights(model, seed=5, calib_size=S)
    images, batch = testing.make_batch(B, V, image_size=S, seed=9)
    return model, images, batch


def _check(kp3d, kp2d, heat, conf, g):
    assert np.abs(np.asarray(conf) - g["confidences"]).max() < 1e-5
    assert rel_err(np.asarray(heat)[:, :, :, ::2, ::2], g["heatmaps_sub"]) < 1e-3
    assert np.array_equal(np.asarray(heat).reshape(2, 4, 17, -1).argmax(-1), g["heatmaps_argmax"])
    assert np.abs(np.asarray(kp2d) - g["keypoints_2d"]).max() < 0.05          # pixels (128-px images)
    assert np.abs(np.asarray(kp3d) - g["keypoints_3d"]).max() < 1.0           # mm (the reference's SVD runs in float32)


def test_oracle_matches_reference_vectors(alg_case):
    model, images, batch = alg_case
    g = np.load(os.path.join(GOLDEN, "algebraic_r50.npz"))
    sd = model.state_dict()
    assert abs(float(sum(v.double().abs().sum() for v in sd.values())) - float(g["sd_checksum"][0])) / float(g["sd_checksum"][0]) < 1e-4
    assert np.array_equal(testing.image_projections(batch), g["proj"])
    _check(*O.algebraic_forward(sd, images, g["proj"]), g)


def test_torch_backend_module_matches_reference_vectors(alg_case):
    model, images, batch = alg_case
    g = np.load(os.path.join(GOLDEN, "algebraic_r50.npz"))
    with torch.no_grad():
        kp3d, kp2d, heat, conf = model(images, torch.from_numpy(g["proj"]), batch)
    _check(kp3d.numpy(), kp2d.numpy(), heat.numpy(), conf.numpy(), g)


def test_dlt_oracle_recovers_known_points():
    rng = np.random.RandomState(0)
    cams = testing.make_cameras(4, image_size=384)
    P = np.stack([c.projection for c in cams]).astype(np.float32)[None]           # (1, 4, 3, 4)
    X = rng.randn(1, 17, 3) * 300 + [0, 0, 900]
    Xh = np.concatenate([X, np.ones((1, 17, 1))], -1)
    uvw = np.einsum("bvij,bkj->bvki", P.astype(np.float64), Xh)
    kp2d = (uvw[..., :2] / uvw[..., 2:3]).astype(np.float32)
    out = O.triangulate_batch_of_points(P, kp2d, rng.rand(1, 4, 17).astype(np.float32) + 0.1)
    assert np.abs(out - X).max() < 0.05
