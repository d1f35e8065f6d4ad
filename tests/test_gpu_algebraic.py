"""GPU parity of the 'next' rows: algebraic model (config #5), confidence heads, conf/conf_norm aggregation, DLT kernel."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_err
from oracle import vol_oracle as O
import lt_b200
from lt_b200 import capi, multiview, op, testing

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_dlt_kernel_vs_oracle_and_truth():
    rng = np.random.RandomState(1)
    cams = testing.make_cameras(5, image_size=384)
    P = np.repeat(np.stack([c.projection for c in cams]).astype(np.float32)[None], 3, axis=0)      # (3, 5, 3, 4)
    X = rng.randn(3, 17, 3) * 300 + [0, 0, 900]
    Xh = np.concatenate([X, np.ones((3, 17, 1))], -1)
    uvw = np.einsum("bvij,bkj->bvki", P.astype(np.float64), Xh)
    kp2d = (uvw[..., :2] / uvw[..., 2:3] + rng.randn(3, 5, 17, 2) * 2.0).astype(np.float32)        # 2 px noise
    conf = (rng.rand(3, 5, 17) + 0.1).astype(np.float32)
    want = O.triangulate_batch_of_points(P, kp2d, conf)
    got = multiview.triangulate_batch_of_points(torch.from_numpy(P).to(DEV), torch.from_numpy(kp2d).to(DEV), torch.from_numpy(conf).to(DEV))
    assert np.abs(got.cpu().numpy() - want).max() < 1e-2                                            # mm
    got1 = multiview.triangulate_batch_of_points(torch.from_numpy(P).to(DEV), torch.from_numpy(kp2d).to(DEV))
    assert np.abs(got1.cpu().numpy() - O.triangulate_batch_of_points(P, kp2d)).max() < 1e-2


def test_integrate_tensor_2d_vs_oracle():
    rng = np.random.RandomState(2)
    hm = (rng.randn(3, 17, 24, 20) * 3).astype(np.float32)
    kp_w, hm_w = O.integrate_tensor_2d(hm, True)
    kp, hmn = op.integrate_tensor_2d(torch.from_numpy(hm).to(DEV), True)
    assert np.abs(kp.cpu().numpy() - kp_w).max() < 1e-3 and rel_err(hmn.cpu().numpy(), hm_w) < 3e-5


@pytest.mark.parametrize("mode", ["simt", "tc"])
def test_algebraic_forward_vs_reference_vectors(mode):
    B, V, S = 2, 4, 128
    g = np.load(os.path.join(GOLDEN, "algebraic_r50.npz"))
    holder = lt_b200.AlgebraicTriangulationNet(testing.make_alg_config(num_layers=50), device="cpu", backend="torch")
    testing.randomize_backbone_weights(holder, seed=5, calib_size=S)
    model = lt_b200.AlgebraicTriangulationNet(testing.make_alg_config(num_layers=50), device=DEV, backend="native", conv_mode=mode)
    model.load_state_dict(holder.state_dict(), strict=True)
    model = model.to(DEV).eval()
    images, batch = testing.make_batch(B, V, image_size=S, seed=9)
    with torch.no_grad():
        kp3d, kp2d, heat, conf = model(images.to(DEV), torch.from_numpy(g["proj"]).to(DEV), batch)
    torch.cuda.synchronize()
    kp3d_o, kp2d_o, heat_o, conf_o = O.algebraic_forward(holder.state_dict(), images, g["proj"])
    e = dict(conf=float(np.abs(conf.cpu().numpy() - conf_o).max()), heat=rel_err(heat.cpu().numpy(), heat_o),
             kp2d=float(np.abs(kp2d.cpu().numpy() - kp2d_o).max()), kp3d=float(np.abs(kp3d.cpu().numpy() - kp3d_o).max()))
    print("algebraic[%s] vs oracle: %s" % (mode, e))
    assert e["conf"] < 1e-4 and e["heat"] < 1e-3 and e["kp2d"] < 0.02 and e["kp3d"] < 0.5
    assert np.array_equal(heat.reshape(B, V, 17, -1).argmax(-1).cpu().numpy(), g["heatmaps_argmax"])
    assert np.abs(kp3d.cpu().numpy() - g["keypoints_3d"]).max() < 1.0        # vs the reference (float32 SVD there)
    assert np.abs(kp2d.cpu().numpy() - g["keypoints_2d"]).max() < 0.05


@pytest.mark.parametrize("agg", ["conf", "conf_norm", "sum", "max"])
def test_volumetric_other_aggregations_vs_oracle(agg):
    """conf / conf_norm use the vol_confidences head (triangulation.py:230-231, :268-269); sum / max are template variants."""
    B, V, S, n = 1, 3, 128, 32
    cfg = testing.make_config(num_layers=50, volume_size=n, aggregation=agg)
    holder = lt_b200.VolumetricTriangulationNet(cfg, device="cpu", backend="torch")
    testing.randomize_weights(holder, seed=3, calib_size=S)
    if agg.startswith("conf"):
        g = torch.Generator().manual_seed(0)
        for lin in (holder.backbone.vol_confidences.head[i] for i in (0, 2, 4)):
            lin.weight.data = torch.randn(lin.weight.shape, generator=g) * (1.0 / lin.weight.shape[1]) ** 0.5
    sd = holder.state_dict()
    model = lt_b200.VolumetricTriangulationNet(testing.make_config(num_layers=50, volume_size=n, aggregation=agg), device=DEV,
                                               backend="native", conv_mode="tc", use_cuda_graph=False)
    model.load_state_dict(sd, strict=True)
    model = model.to(DEV).eval()
    images, batch = testing.make_batch(B, V, image_size=S, seed=7)
    with torch.no_grad():
        kp, feats, vols, conf, cuboids, coords, base = model(images.to(DEV), None, batch)
    torch.cuda.synchronize()
    bp = np.stack([k[6, :3] for k in batch["keypoints_3d"]])
    kp_o, feats_o, vols_o, coords_o, inter = O.volumetric_forward(sd, images, batch["cameras"], bp, volume_size=n, aggregation=agg,
                                                                  return_intermediates=True)
    if agg.startswith("conf"):
        assert conf is not None and np.abs(conf.cpu().numpy() - inter["vol_confidences"]).max() < 1e-4
    else:
        assert conf is None
    e_v, e_k = rel_err(vols.cpu().numpy(), vols_o.numpy()), float((kp.cpu() - kp_o).abs().max())
    print("aggregation %s: volumes %.2e keypoints %.4f mm" % (agg, e_v, e_k))
    assert e_v < 1e-3 and e_k < 0.5
