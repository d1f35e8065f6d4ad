"""Oracle and host mirror vs the reference itself (only where /root/reference is mounted: the authoring
container).  On the GPU box these skip; the committed golden vectors carry the pin there."""
import sys

import numpy as np
import pytest
import torch

from conftest import REFERENCE, has_reference, rel_err

pytestmark = pytest.mark.skipif(not has_reference(), reason="reference checkout not mounted")

from oracle import vol_oracle as O  # noqa: E402
from lt_b200 import testing, torch_ops  # noqa: E402
import lt_b200  # noqa: E402


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, REFERENCE)
    from mvn.utils import op, multiview, volumetric
    from mvn.models import v2v, pose_resnet
    return dict(op=op, multiview=multiview, volumetric=volumetric, v2v=v2v, pose_resnet=pose_resnet)


@pytest.mark.parametrize("shape", [(1, 2, 3, 9, 13, 5), (2, 8, 4, 12, 12, 6), (1, 9, 2, 8, 6, 4)])
@pytest.mark.parametrize("agg", ["sum", "max", "softmax", "conf"])
def test_unproject_random(ref, shape, agg):
    B, V, C, h, w, n = shape
    rng = np.random.RandomState(B * 100 + V)
    heat = rng.randn(B, V, C, h, w).astype(np.float32)
    cams = testing.make_cameras(V, image_size=48, radius=3000.0)
    proj = np.stack([np.stack([O.projection_after_resize(c.K, c.R, c.t, (48, 48), (h, w)) for c in cams])] * B)
    coord = np.stack([O.coord_volume(rng.randn(3) * 100 + [0, 0, 900], 2600.0, n) for _ in range(B)])
    conf = rng.rand(B, V, C).astype(np.float32)
    want = ref["op"].unproject_heatmaps(torch.from_numpy(heat), torch.from_numpy(proj), torch.from_numpy(coord), agg,
                                        torch.from_numpy(conf)).numpy()
    assert rel_err(O.unproject_heatmaps(heat, proj, coord, agg, conf), want) < 2e-5
    got_t = torch_ops.unproject_heatmaps(torch.from_numpy(heat), torch.from_numpy(proj), torch.from_numpy(coord), agg,
                                         torch.from_numpy(conf)).numpy()
    assert rel_err(got_t, want) < 2e-5


def test_coord_volume_and_projection_bit_exact(ref):
    base = np.array([123.4, -56.7, 910.1])
    sys.path.insert(0, REFERENCE)
    from mvn.models.triangulation import VolumetricTriangulationNet as RefNet
    net = RefNet(testing.make_config(num_layers=18, volume_size=32), device="cpu").eval()
    images, batch = testing.make_batch(1, 2, image_size=64, seed=5, camera_cls=ref["multiview"].Camera)
    batch["keypoints_3d"][0][6, :3] = base
    with torch.no_grad():
        out = net(images, None, batch)
    coords = out[5].numpy()[0]
    mine = O.coord_volume(base, 2500.0, 32)
    assert np.array_equal(mine, coords)
    cam = batch["cameras"][1][0]
    c2 = ref["multiview"].Camera(cam.R, cam.t, cam.K)
    c2.update_after_resize((64, 64), (16, 16))
    assert np.array_equal(O.projection_after_resize(cam.K, cam.R, cam.t, (64, 64), (16, 16)), c2.projection.astype(np.float32))


def test_v2v_and_backbone_functional_restatement(ref):
    torch.manual_seed(0)
    v = ref["v2v"].V2VModel(32, 17).eval()
    for m in v.modules():
        if isinstance(m, torch.nn.BatchNorm3d):
            m.running_mean.normal_(0, 0.1)
            m.running_var.uniform_(0.5, 1.5)
    x = torch.randn(1, 32, 32, 32, 32)
    with torch.no_grad():
        want = v(x)
    assert rel_err(O.v2v_forward(v.state_dict(), x), want) < 1e-5
    cfg = testing.AttrDict(num_layers=34, style="simple", num_joints=17, alg_confidences=False, vol_confidences=False,
                           init_weights=False)
    bb = ref["pose_resnet"].get_pose_net(cfg, device="cpu").eval()
    xi = torch.randn(2, 3, 64, 64)
    with torch.no_grad():
        hm, feats, _, _ = bb(xi)
    ohm, ofe = O.pose_resnet_forward(bb.state_dict(), xi)
    assert rel_err(ofe, feats) < 1e-5 and rel_err(ohm, hm) < 1e-5
