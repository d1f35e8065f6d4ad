"""Pin the CPU oracle (oracle/vol_oracle.py) against the golden vectors generated from the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_err
from oracle import vol_oracle as O
from lt_b200 import testing
import lt_b200


@pytest.mark.parametrize("tag", ["a", "b", "c"])
@pytest.mark.parametrize("agg", ["sum", "max", "softmax", "conf"])
def test_unproject_matches_reference_vectors(tag, agg):
    g = np.load(os.path.join(GOLDEN, "unproject.npz"))
    out = O.unproject_heatmaps(g[tag + "_heat"], g[tag + "_proj"], g[tag + "_coord"], agg, g[tag + "_conf"])
    ref = g["%s_out_%s" % (tag, agg)]
    assert out.shape == ref.shape
    assert rel_err(out, ref) < 2e-5


def test_unproject_unknown_aggregation_raises():
    g = np.load(os.path.join(GOLDEN, "unproject.npz"))
    with pytest.raises(ValueError):
        O.unproject_heatmaps(g["c_heat"], g["c_proj"], g["c_coord"], "median")


@pytest.mark.parametrize("softmax", [True, False])
def test_softargmax_matches_reference_vectors(softmax):
    g = np.load(os.path.join(GOLDEN, "softargmax.npz"))
    kp, v = O.integrate_tensor_3d_with_coordinates(g["vols"], g["coord"], softmax)
    assert rel_err(kp, g["kp_%d" % softmax]) < 1e-5
    assert rel_err(v, g["v_%d" % softmax]) < 1e-5


@pytest.fixture(scope="module")
def r50_case():
    """Same seeds / recipe as make_golden.gen_forward()."""
    B, V, S, n = 2, 2, 128, 32
    cfg = testing.make_config(num_layers=50, volume_size=n)
    model = lt_b200.VolumetricTriangulationNet(cfg, device="cpu", backend="torch")
    testing.randomize_weights(model, seed=1, calib_size=S)
    images, batch = testing.make_batch(B, V, image_size=S, seed=3)
    return model, images, batch, n


def test_full_forward_oracle_matches_reference_vectors(r50_case):
    model, images, batch, n = r50_case
    g = np.load(os.path.join(GOLDEN, "forward_r50.npz"))
    sd = model.state_dict()
    drift = abs(float(sum(v.double().abs().sum() for v in sd.values())) - float(g["sd_checksum"][0])) / float(g["sd_checksum"][0])
    assert drift < 1e-4, "weight recipe drifted from the one the fixtures were generated with"
    base = np.stack([k[6, :3] for k in batch["keypoints_3d"]])
    kp, feats, vols, coords, inter = O.volumetric_forward(sd, images, batch["cameras"], base, volume_size=n,
                                                          return_intermediates=True)
    assert rel_err(feats[:, :, ::4, ::3, ::3], g["features_sub"]) < 1e-3
    assert np.abs(coords[:, ::5, ::5, ::5].numpy() - g["coord_sub"]).max() < 1e-3          # mm
    assert rel_err(inter["unprojected"][:, ::4, ::3, ::3, ::3], g["vol_in_sub"]) < 1e-3
    assert rel_err(inter["logits"][:, :, ::3, ::3, ::3], g["logits_sub"]) < 1e-3
    assert rel_err(vols[:, :, ::3, ::3, ::3], g["volumes_sub"]) < 2e-3
    assert np.abs(kp.numpy() - g["keypoints"]).max() < 0.5                                  # mm
    flat = vols.reshape(vols.shape[0], vols.shape[1], -1)
    assert np.array_equal(flat.argmax(-1).numpy(), g["volumes_argmax"])


def test_torch_backend_module_matches_reference_vectors(r50_case):
    """The nn.Module boundary (torch formulation) against the same vectors: state_dict layout, host geometry, outputs."""
    model, images, batch, n = r50_case
    g = np.load(os.path.join(GOLDEN, "forward_r50.npz"))
    model.eval()
    with torch.no_grad():
        kp, feats, vols, conf, cuboids, coords, base = model(images, None, batch)
    assert conf is None and len(cuboids) == images.shape[0]
    assert np.allclose(np.stack([c.position for c in cuboids]), g["cuboid_position"])
    assert np.allclose(base.numpy(), g["base_points"])
    assert rel_err(feats[:, :, ::4, ::3, ::3], g["features_sub"]) < 1e-3
    assert np.abs(coords[:, ::5, ::5, ::5].numpy() - g["coord_sub"]).max() < 1e-3
    assert rel_err(vols[:, :, ::3, ::3, ::3], g["volumes_sub"]) < 2e-3
    assert np.abs(kp.numpy() - g["keypoints"]).max() < 0.5
    assert np.array_equal(vols.reshape(2, 17, -1).argmax(-1).numpy(), g["volumes_argmax"])
