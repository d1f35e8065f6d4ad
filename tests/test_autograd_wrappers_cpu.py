"""CPU check of the `backend="hybrid"` autograd plumbing (autograd_ops.py): the C-ABI calls are replaced by torch stand-ins
with the same contracts (layouts, accumulate-into semantics), so that the wrappers' permutes / reshapes / saved tensors /
returned gradient slots are exercised without a GPU.  The CUDA kernels themselves are covered by tests/test_gpu_hybrid.py."""
import numpy as np
import pytest
import torch

from lt_b200 import autograd_ops, capi, torch_ops

AGG_NAME = {v: k for k, v in capi.AGG.items()}


def _fake_unproject(feats_cl, proj, coord, conf, out_cl, out_format, agg):
    B, V, h, w, C = feats_cl.shape
    n = round(coord.shape[1] ** (1 / 3))
    vol = torch_ops.unproject_heatmaps(feats_cl.permute(0, 1, 4, 2, 3), proj, coord.view(B, n, n, n, 3), AGG_NAME[agg], conf)
    out_cl.copy_(vol.reshape(B, C, -1).permute(0, 2, 1))


def _fake_unproject_bwd(feats_cl, proj, coord, conf, g_cl, grad_feats, grad_conf, agg):
    B, V, h, w, C = feats_cl.shape
    n = round(coord.shape[1] ** (1 / 3))
    f = feats_cl.clone().requires_grad_(True)
    c = None if conf is None else conf.clone().requires_grad_(True)
    with torch.enable_grad():
        vol = torch_ops.unproject_heatmaps(f.permute(0, 1, 4, 2, 3), proj, coord.view(B, n, n, n, 3), AGG_NAME[agg], c)
        vol.backward(g_cl.permute(0, 2, 1).reshape(vol.shape))
    grad_feats.add_(f.grad)                      # accumulate-into contract of lt_unproject_aggregate_bwd
    if grad_conf is not None:
        grad_conf.add_(c.grad)


def _fake_softargmax(logits, bs, vs, cs, coord, out, kp, ws, B, J, nvox, mult, softmax):
    n = round(nvox ** (1 / 3))
    k, v = torch_ops.integrate_tensor_3d_with_coordinates(logits.view(B, J, n, n, n) * mult, coord.view(B, n, n, n, 3), softmax)
    kp.copy_(k)
    out.copy_(v.view_as(out))


def _fake_softargmax_bwd(probs, coord, g_kp, g_vol, grad_logits, scratch, B, J, nvox, mult, softmax):
    p = probs.reshape(B, J, nvox)
    t = torch.einsum("bjc,bnc->bjn", g_kp, coord)
    if g_vol is not None:
        t = t + g_vol.reshape(B, J, nvox)
    if softmax:
        g = mult * p * (t - (p * t).sum(-1, keepdim=True))
    else:
        g = mult * t * (p > 0)
    grad_logits.copy_(g.view_as(grad_logits))


@pytest.fixture
def fake_capi(monkeypatch):
    monkeypatch.setattr(capi, "unproject_aggregate", _fake_unproject)
    monkeypatch.setattr(capi, "unproject_aggregate_bwd", _fake_unproject_bwd)
    monkeypatch.setattr(capi, "softargmax3d", _fake_softargmax)
    monkeypatch.setattr(capi, "softargmax3d_bwd", _fake_softargmax_bwd)
    monkeypatch.setattr(capi, "softargmax3d_workspace_bytes", lambda B, J, nvox: 64)


@pytest.mark.parametrize("agg", ["sum", "softmax", "conf", "max"])
def test_unproject_wrapper_gradients(fake_capi, agg):
    torch.manual_seed(1)
    B, V, C, h, w, n = 2, 3, 8, 5, 7, 4
    heat = torch.randn(B, V, C, h, w)
    proj = torch.randn(B, V, 3, 4)
    proj[:, :, 2, 3] += 6.0                      # positive depths
    coord = torch.randn(B, n, n, n, 3)
    conf = torch.rand(B, V, C)
    g = torch.randn(B, C, n, n, n)
    res = []
    for fn in (lambda a, c: torch_ops.unproject_heatmaps(a, proj, coord, agg, c), lambda a, c: autograd_ops.unproject_heatmaps(a, proj, coord, agg, c)):
        a, c = heat.clone().requires_grad_(True), conf.clone().requires_grad_(True)
        out = fn(a, c)
        out.backward(g)
        res.append((out.detach(), a.grad, c.grad))
    assert torch.allclose(res[0][0], res[1][0], atol=1e-5)
    assert torch.allclose(res[0][1], res[1][1], atol=1e-5)
    if agg == "conf":
        assert torch.allclose(res[0][2], res[1][2], atol=1e-4)
    else:
        assert res[1][2] is None


@pytest.mark.parametrize("softmax", [True, False])
def test_softargmax_wrapper_gradients(fake_capi, softmax):
    torch.manual_seed(2)
    B, J, n = 2, 5, 4
    vols = torch.randn(B, J, n, n, n)
    coord = torch.randn(B, n, n, n, 3) * 10
    g_kp, g_vol = torch.randn(B, J, 3), torch.randn(B, J, n, n, n)
    res = []
    for fn in (torch_ops.integrate_tensor_3d_with_coordinates, autograd_ops.integrate_tensor_3d_with_coordinates):
        v = vols.clone().requires_grad_(True)
        kp, p = fn(v, coord, softmax)
        ((kp * g_kp).sum() + (p * g_vol).sum()).backward()
        res.append((kp.detach(), p.detach(), v.grad))
    for a, b in zip(res[0], res[1]):
        assert torch.allclose(a, b, atol=1e-4, rtol=1e-4)
    # keypoints-only loss: the volumes gradient slot arrives as None
    v = vols.clone().requires_grad_(True)
    kp, _ = autograd_ops.integrate_tensor_3d_with_coordinates(v, coord, softmax)
    (kp * g_kp).sum().backward()
    v2 = vols.clone().requires_grad_(True)
    kp2, _ = torch_ops.integrate_tensor_3d_with_coordinates(v2, coord, softmax)
    (kp2 * g_kp).sum().backward()
    assert torch.allclose(v.grad, v2.grad, atol=1e-4, rtol=1e-4)
