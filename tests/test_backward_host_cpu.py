"""Gradient arithmetic of csrc/backward.cu checked on the CPU: `lt_test_*_bwd_host` run the kernels' own per-item code
(`__host__ __device__`) with host pointers; the reference gradients come from torch autograd of the torch formulation
(`torch_ops`, pinned to the reference).  What this cannot cover -- launch geometry, atomics, shared-memory accumulation of
the confidence gradient -- is left to tests/test_gpu_hybrid.py."""
import numpy as np
import pytest
import torch

from lt_b200 import capi, testing, torch_ops
from oracle import vol_oracle as O


def _p(t):
    assert not t.is_cuda and t.is_contiguous() and t.dtype == torch.float32
    return t.data_ptr()


def _scene(B, V, C, h, w, n, seed):
    rng = np.random.RandomState(seed)
    heat = rng.randn(B, V, C, h, w).astype(np.float32)
    cams = testing.make_cameras(V, image_size=48, radius=3000.0)
    proj = np.stack([np.stack([O.projection_after_resize(c.K, c.R, c.t, (48, 48), (h, w)) for c in cams])] * B)
    coord = np.stack([O.coord_volume(rng.randn(3) * 100 + [0, 0, 900], 2800.0, n) for _ in range(B)])
    conf = rng.rand(B, V, C).astype(np.float32)
    return [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)) for a in (heat, proj, coord, conf)]


@pytest.mark.parametrize("shape", [(2, 4, 8, 12, 12, 6), (1, 3, 4, 9, 13, 5), (1, 5, 8, 8, 8, 4)])
@pytest.mark.parametrize("agg", ["sum", "softmax", "conf", "max"])
def test_unproject_backward_item_code_vs_torch_autograd(shape, agg):
    B, V, C, h, w, n = shape
    heat, proj, coord, conf = _scene(*shape, seed=sum(shape))
    nvox = n ** 3
    g = torch.randn(B, C, n, n, n)
    a = heat.clone().requires_grad_(True)
    c = conf.clone().requires_grad_(True)
    out = torch_ops.unproject_heatmaps(a, proj, coord, agg, c)
    out.backward(g)
    feats_cl = heat.permute(0, 1, 3, 4, 2).contiguous()
    g_cl = g.reshape(B, C, nvox).permute(0, 2, 1).contiguous()
    grad_feats = torch.zeros_like(feats_cl)
    grad_conf = torch.zeros(B, V, C) if agg == "conf" else None
    rc = capi.lib().lt_test_unproject_aggregate_bwd_host(_p(feats_cl), _p(proj.contiguous()), _p(coord.reshape(B, nvox, 3).contiguous()),
                                                         _p(conf) if agg == "conf" else None, _p(g_cl), _p(grad_feats),
                                                         _p(grad_conf) if grad_conf is not None else None, B, V, C, h, w, nvox, capi.AGG[agg])
    assert rc == 0, capi.lib().lt_last_error_string()
    want = a.grad.permute(0, 1, 3, 4, 2)
    scale = float(max(want.abs().max(), want.std()))
    assert float((grad_feats - want).abs().max()) <= 2e-5 * scale
    if agg == "conf":
        assert float((grad_conf - c.grad).abs().max()) <= 1e-4 * float(c.grad.abs().max())


@pytest.mark.parametrize("softmax", [True, False])
@pytest.mark.parametrize("with_gvol", [True, False])
def test_softargmax_backward_item_code_vs_torch_autograd(softmax, with_gvol):
    torch.manual_seed(5)
    B, J, n = 2, 5, 6
    nvox = n ** 3
    vols = torch.randn(B, J, n, n, n) * 2
    coord = torch.randn(B, n, n, n, 3) * 50
    g_kp, g_vol = torch.randn(B, J, 3), torch.randn(B, J, n, n, n)
    v = vols.clone().requires_grad_(True)
    kp, p = torch_ops.integrate_tensor_3d_with_coordinates(v, coord, softmax)
    loss = (kp * g_kp).sum() + ((p * g_vol).sum() if with_gvol else 0.0)
    loss.backward()
    grad = torch.empty(B, J, nvox)
    rc = capi.lib().lt_test_softargmax3d_bwd_host(_p(p.detach().reshape(B, J, nvox).contiguous()), _p(coord.reshape(B, nvox, 3).contiguous()),
                                                  _p(g_kp), _p(g_vol.reshape(B, J, nvox).contiguous()) if with_gvol else None, _p(grad),
                                                  B, J, nvox, 1.0, int(softmax))
    assert rc == 0
    want = v.grad.reshape(B, J, nvox)
    assert float((grad - want).abs().max()) <= 2e-5 * float(max(want.abs().max(), want.std()))
