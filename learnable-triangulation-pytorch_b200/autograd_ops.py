"""Differentiable wrappers of the two native custom ops (`backend="hybrid"`): the forward is the same hand-written kernel
the inference path uses, the backward is `csrc/backward.cu` (lt_unproject_aggregate_bwd, lt_softargmax3d_bwd).

This is the first stage of SURVEY section 8f row 1: with it the reference training loop (`train.py:159-243`,
`total_loss.backward()` at :236) runs the unprojection + aggregation and the soft-argmax -- the ops the reference implements
as Python loops over (sample, view) pairs with ~10 passes over a (V, C, N^3) staging tensor -- on the native kernels while
the convolutions stay on torch/cuDNN autograd.  Gradients: feature maps, `conf` confidences, V2V logits; projection matrices
and coordinate volumes carry none (they do not in the reference either: they come from numpy camera data).
"""
import torch

from . import capi


class UnprojectHeatmapsFn(torch.autograd.Function):
    """op.unproject_heatmaps (reference op.py:99-166), NCHW features in, NCDHW volume out."""

    @staticmethod
    def forward(ctx, heatmaps, proj_matricies, coord_volumes, vol_confidences, agg):
        B, V, C, h, w = heatmaps.shape
        vol_shape = tuple(coord_volumes.shape[1:4])
        nvox = vol_shape[0] * vol_shape[1] * vol_shape[2]
        feats_cl = heatmaps.detach().float().permute(0, 1, 3, 4, 2).contiguous()          # (B, V, h, w, C)
        proj = proj_matricies.detach().float().contiguous()
        coord = coord_volumes.detach().float().reshape(B, nvox, 3).contiguous()
        conf = None
        if agg == capi.AGG["conf"]:
            conf = vol_confidences.detach().float().reshape(B, V, C).contiguous()
        out_cl = torch.empty((B, nvox, C), dtype=torch.float32, device=heatmaps.device)
        capi.unproject_aggregate(feats_cl, proj, coord, conf, out_cl, capi.FMT_F32, agg)
        ctx.save_for_backward(feats_cl, proj, coord, conf if conf is not None else torch.empty(0, device=heatmaps.device))
        ctx.agg, ctx.vol_shape, ctx.has_conf = agg, vol_shape, conf is not None
        ctx.conf_shape = None if vol_confidences is None else tuple(vol_confidences.shape)
        return out_cl.permute(0, 2, 1).reshape(B, C, *vol_shape).contiguous()

    @staticmethod
    def backward(ctx, grad_out):
        feats_cl, proj, coord, conf = ctx.saved_tensors
        conf = conf if ctx.has_conf else None
        B, V, h, w, C = feats_cl.shape
        nvox = coord.shape[1]
        g_cl = grad_out.float().reshape(B, C, nvox).permute(0, 2, 1).contiguous()          # (B, nvox, C)
        grad_feats = torch.zeros_like(feats_cl)
        need_conf = ctx.has_conf and ctx.needs_input_grad[3]
        grad_conf = torch.zeros((B, V, C), dtype=torch.float32, device=feats_cl.device) if need_conf else None
        capi.unproject_aggregate_bwd(feats_cl, proj, coord, conf, g_cl, grad_feats, grad_conf, ctx.agg)
        grad_heat = grad_feats.permute(0, 1, 4, 2, 3)                                       # (B, V, C, h, w) view
        return grad_heat, None, None, (grad_conf.reshape(ctx.conf_shape) if need_conf else None), None


class IntegrateTensor3dFn(torch.autograd.Function):
    """op.integrate_tensor_3d_with_coordinates (reference op.py:84-96): (keypoints, normalised volumes)."""

    @staticmethod
    def forward(ctx, volumes, coord_volumes, softmax):
        B, J = volumes.shape[:2]
        nvox = volumes[0, 0].numel()
        logits = volumes.detach().float().contiguous()
        coord = coord_volumes.detach().float().reshape(B, nvox, 3).contiguous()
        out = torch.empty_like(logits)
        keypoints = torch.empty((B, J, 3), dtype=torch.float32, device=volumes.device)
        ws = torch.empty(capi.softargmax3d_workspace_bytes(B, J, nvox) // 4 + 1, dtype=torch.float32, device=volumes.device)
        capi.softargmax3d(logits, J * nvox, 1, nvox, coord, out, keypoints, ws, B, J, nvox, 1.0, softmax)
        ctx.save_for_backward(out, coord)
        ctx.softmax = bool(softmax)
        return keypoints, out

    @staticmethod
    def backward(ctx, grad_kp, grad_vol):
        probs, coord = ctx.saved_tensors
        B, J = probs.shape[:2]
        nvox = coord.shape[1]
        dev = probs.device
        g_kp = (grad_kp if grad_kp is not None else torch.zeros((B, J, 3), device=dev)).float().contiguous()
        g_vol = None if grad_vol is None else grad_vol.float().contiguous()
        grad_logits = torch.empty_like(probs)
        scratch = torch.empty(B * J, dtype=torch.float32, device=dev)
        capi.softargmax3d_bwd(probs, coord, g_kp, g_vol, grad_logits, scratch, B, J, nvox, 1.0, ctx.softmax)
        return grad_logits, None, None


def unproject_heatmaps(heatmaps, proj_matricies, coord_volumes, volume_aggregation_method="sum", vol_confidences=None):
    agg = capi.AGG["conf" if volume_aggregation_method.startswith("conf") else volume_aggregation_method]
    return UnprojectHeatmapsFn.apply(heatmaps, proj_matricies, coord_volumes, vol_confidences, agg)


def integrate_tensor_3d_with_coordinates(volumes, coord_volumes, softmax=True):
    return IntegrateTensor3dFn.apply(volumes, coord_volumes, softmax)
