"""GPU parity of the configuration variants reachable from the reference's yaml files, native path vs the CPU oracle:
BasicBlock backbones (ResNet-18/34, pose_resnet.py:25-54), `style: caffe` bottlenecks (:98-137), the CMU axis transfer
(triangulation.py:336-339), `volume_softmax: false` (op.py:90-91), the non-softmax 2-D integration (op.py:25-41) at op level and
through AlgebraicTriangulationNet(heatmap_softmax=False), and inputs whose side is not a multiple of 32."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import vol_oracle as O
import lt_b200
from lt_b200 import op, testing

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _case(cfg, B=1, V=2, S=128, seed=1):
    holder = lt_b200.VolumetricTriangulationNet(cfg, device="cpu", backend="torch")
    testing.randomize_weights(holder, seed=seed, calib_size=S)
    images, batch = testing.make_batch(B, V, image_size=S, seed=seed + 2)
    return holder.state_dict(), images, batch


def _native(cfg, sd, mode="tc"):
    m = lt_b200.VolumetricTriangulationNet(cfg, device=DEV, backend="native", conv_mode=mode, use_cuda_graph=False)
    m.load_state_dict(sd, strict=True)
    return m.to(DEV).eval()


def _check(out, ref, B, tol_feat=3e-4, tol_vol=1e-3, tol_kp=0.5):
    kp, feats, vols, _, _, coords, _ = out
    kp_o, feats_o, vols_o, coords_o = ref[:4]
    assert torch.equal(coords.cpu(), coords_o)
    e_f, e_v = rel_err(feats.cpu().numpy(), feats_o.numpy()), rel_err(vols.cpu().numpy(), vols_o.numpy())
    e_k = float((kp.cpu() - kp_o).abs().max())
    print("features %.2e volumes %.2e keypoints %.4f mm" % (e_f, e_v, e_k))
    assert e_f < tol_feat and e_v < tol_vol and e_k < tol_kp
    return e_f, e_v, e_k


@pytest.mark.parametrize("layers,style", [(18, "simple"), (34, "simple"), (50, "caffe")])
def test_backbone_variants_match_oracle(layers, style):
    n = 32
    cfg = testing.make_config(num_layers=layers, volume_size=n, style=style)
    sd, images, batch = _case(cfg)
    base = np.stack([k[6, :3] for k in batch["keypoints_3d"]])
    ref = O.volumetric_forward(sd, images, batch["cameras"], base, volume_size=n, style=style)
    model = _native(testing.make_config(num_layers=layers, volume_size=n, style=style), sd)
    with torch.no_grad():
        out = model(images.to(DEV), None, batch)
    _check(out, ref, 1)
    assert torch.equal(out[2].reshape(1, 17, -1).argmax(-1).cpu(), ref[2].reshape(1, 17, -1).argmax(-1))


def test_cmu_axis_transfer_coco_matches_oracle():
    """kind=coco (hip midpoint base point, y-up rotation axis) with transfer_cmu_to_human36m (triangulation.py:286-296, 336-339)."""
    n = 32
    cfg = testing.make_config(num_layers=18, volume_size=n, kind="coco")
    cfg.model.transfer_cmu_to_human36m = True
    sd, images, batch = _case(cfg, seed=5)
    base = np.stack([(k[11, :3] + k[12, :3]) / 2 for k in batch["keypoints_3d"]])
    ref = O.volumetric_forward(sd, images, batch["cameras"], base, volume_size=n, kind="coco", transfer_cmu=True)
    cfg2 = testing.make_config(num_layers=18, volume_size=n, kind="coco")
    cfg2.model.transfer_cmu_to_human36m = True
    model = _native(cfg2, sd)
    with torch.no_grad():
        out = model(images.to(DEV), None, batch)
    _check(out, ref, 1)


def test_volume_softmax_false_matches_oracle():
    """volume_softmax: false -> ReLU volumes, un-normalised expectation (op.py:90-91); volume_multiplier applied first."""
    n = 32
    cfg = testing.make_config(num_layers=18, volume_size=n, volume_softmax=False, volume_multiplier=0.01)
    sd, images, batch = _case(cfg, seed=7)
    base = np.stack([k[6, :3] for k in batch["keypoints_3d"]])
    ref = O.volumetric_forward(sd, images, batch["cameras"], base, volume_size=n, volume_softmax=False, volume_multiplier=0.01)
    model = _native(testing.make_config(num_layers=18, volume_size=n, volume_softmax=False, volume_multiplier=0.01), sd)
    with torch.no_grad():
        kp, feats, vols, _, _, coords, _ = model(images.to(DEV), None, batch)
    assert rel_err(feats.cpu().numpy(), ref[1].numpy()) < 3e-4
    assert rel_err(vols.cpu().numpy(), ref[2].numpy()) < 1e-3
    # the ReLU expectation is a plain weighted sum of coordinates (thousands of mm x mass): relative bar
    assert rel_err(kp.cpu().numpy(), ref[0].numpy()) < 1e-3


@pytest.mark.parametrize("softmax", [True, False])
def test_integrate_tensor_2d_native_matches_oracle(softmax):
    g = torch.Generator().manual_seed(3)
    hm = torch.randn(3, 17, 24, 20, generator=g) * 2.0
    kp, out = op.integrate_tensor_2d(hm.to(DEV), softmax=softmax)
    kp_o, out_o = O.integrate_tensor_2d(hm.numpy(), softmax)
    assert rel_err(out.cpu().numpy(), out_o) < 1e-5
    assert float(np.abs(kp.cpu().numpy() - kp_o).max()) < 1e-3      # pixels


def test_algebraic_heatmap_softmax_false_native_vs_torch_backend():
    """AlgebraicTriangulationNet with heatmap_softmax=False (ReLU mass normalisation, op.py:25-41) on the native path."""
    cfg = testing.make_config(num_layers=18)
    cfg.model.use_confidences = False
    cfg.model.heatmap_softmax = False
    cfg.model.heatmap_multiplier = 1.0
    holder = lt_b200.AlgebraicTriangulationNet(cfg, device="cpu", backend="torch")
    testing.randomize_backbone_weights(holder, seed=2, calib_size=128)
    sd = holder.state_dict()
    images, batch = testing.make_batch(2, 2, image_size=128, seed=4)
    from lt_b200 import multiview
    proj = torch.from_numpy(multiview.stack_projections(batch["cameras"], (128, 128), (128, 128))).float()
    with torch.no_grad():
        want = holder.eval()(images, proj, batch)
    cfg2 = testing.make_config(num_layers=18)
    cfg2.model.use_confidences = False
    cfg2.model.heatmap_softmax = False
    cfg2.model.heatmap_multiplier = 1.0
    m = lt_b200.AlgebraicTriangulationNet(cfg2, device=DEV, backend="native", conv_mode="tc")
    m.load_state_dict(sd)
    m = m.to(DEV).eval()
    with torch.no_grad():
        got = m(images.to(DEV), proj.to(DEV), batch)
    assert rel_err(got[2].cpu().numpy(), want[2].numpy()) < 1e-3            # ReLU heat-maps
    assert float((got[1].cpu() - want[1]).abs().max()) < 0.05               # 2-D key points, pixels
    # triangulated key points, mm.  Two views only: the DLT turns a 0.05 px difference of a 2-D key point into millimetres, so the
    # 3-D bar against the torch backend is loose, and the DLT itself is checked tightly on the native 2-D key points.
    assert float((got[0].cpu() - want[0]).abs().max()) < 5.0
    same_2d = multiview.triangulate_batch_of_points(proj, got[1].cpu(), confidences_batch=got[3].cpu(), backend="torch")
    assert float((got[0].cpu() - same_2d).abs().max()) < 0.5          # float32 SVD in the torch op, fp64 Jacobi in the kernel


def test_input_side_not_multiple_of_32():
    """160 is a multiple of 32, 176 is not: 176 -> 88/44/22/11/6 -> 48 (not 44 = 176 // 4): the intrinsics must be rescaled by
    the feature-map size the kernels really produce (reference: heatmaps.shape, triangulation.py:264-265)."""
    n, S = 32, 176
    from lt_b200.triangulation import backbone_map_size
    assert backbone_map_size(160) == 40 and backbone_map_size(S) == 48
    cfg = testing.make_config(num_layers=18, volume_size=n)
    sd, images, batch = _case(cfg, S=S, seed=9)
    base = np.stack([k[6, :3] for k in batch["keypoints_3d"]])
    ref = O.volumetric_forward(sd, images, batch["cameras"], base, volume_size=n)
    assert tuple(ref[1].shape[3:]) == (48, 48)
    model = _native(testing.make_config(num_layers=18, volume_size=n), sd)
    with torch.no_grad():
        out = model(images.to(DEV), None, batch)
    _check(out, ref, 1)
