"""Checker glue around the CPU oracle -- TEST INFRASTRUCTURE (same rules as vol_oracle.py: only tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline leg import it; the product path never does).

`compare_forward` runs the oracle's eval forward (reference triangulation.py:245-355 restated in vol_oracle.py) on the
given inputs, times it, and compares the outputs of a native forward of the SAME inputs with it, using the parity
definition of SURVEY.md 8(d): max|a-b| / max(|b|, std b); keypoints in millimetres; arg-max voxel indices bit-exact.
"""
import time

import numpy as np
import torch

from . import vol_oracle as O

# contract of BASELINE.json north_star: 1e-3 relative fp32, arg-max joint indices bit-exact
CONTRACT_REL = 1e-3
CUBOID_MM = 2500.0


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    denom = max(float(np.abs(b).max()), float(b.std()), 1e-30)
    return float(np.abs(a - b).max()) / denom


def oracle_forward(sd, images, batch, volume_size, **kw):
    """-> (outputs tuple incl. intermediates, seconds)."""
    base = np.stack([np.asarray(k)[6, :3] for k in batch["keypoints_3d"]])
    t0 = time.perf_counter()
    out = O.volumetric_forward(sd, images, batch["cameras"], base, volume_size=volume_size, return_intermediates=True, **kw)
    return out, time.perf_counter() - t0


def compare_outputs(native_out, oracle_out):
    """native_out: the 7-tuple of VolumetricTriangulationNet.forward; oracle_out: oracle_forward()[0]."""
    kp, feats, vols, _, _, coords, _ = native_out
    kp_o, feats_o, vols_o, coords_o = oracle_out[:4]
    B, J = vols_o.shape[:2]
    am = vols.reshape(B, J, -1).argmax(-1).cpu()
    am_o = vols_o.reshape(B, J, -1).argmax(-1)
    res = {
        "features_rel": rel_err(feats.cpu().numpy(), feats_o.numpy()),
        "volumes_rel": rel_err(vols.cpu().numpy(), vols_o.numpy()),
        "keypoints_mm": float((kp.cpu() - kp_o).abs().max()),
        "keypoints_rel": float((kp.cpu() - kp_o).abs().max()) / CUBOID_MM,
        "argmax_equal": bool(torch.equal(am, am_o)),
        "coord_volumes_bit_exact": bool(torch.equal(coords.cpu(), coords_o)),
        "samples": int(B),
        "tolerance_rel": CONTRACT_REL,
    }
    res["ok"] = bool(res["features_rel"] < CONTRACT_REL and res["volumes_rel"] < CONTRACT_REL and
                     res["keypoints_rel"] < CONTRACT_REL and res["argmax_equal"] and res["coord_volumes_bit_exact"])
    return res


def compare_stages(model, images, batch, oracle_out, dev):
    """Intermediates of the native engine (unprojected volume, V2V logits) against the oracle's, same inputs."""
    from lt_b200 import capi
    inter = oracle_out[4]
    coords_o = oracle_out[3]
    eng = model.engine()
    eng.prepare()
    B, V = images.shape[:2]
    with torch.no_grad():
        feats = eng.backbone_features(images.to(dev).reshape(B * V, *images.shape[2:]))
        proj = torch.from_numpy(np.ascontiguousarray(inter["proj"], dtype=np.float32)).to(dev)
        coord = coords_o.to(dev).contiguous()
        vol = eng.unproject(feats, B, V, proj, coord, capi.AGG[model.volume_aggregation_method])
        got_vol = eng._as_f32(vol).data.permute(0, 4, 1, 2, 3).cpu().numpy()
        logits = eng.v2v(vol)
        got_logits = logits.data[..., :inter["logits"].shape[1]].permute(0, 4, 1, 2, 3).cpu().numpy()
    return {"unprojected_rel": rel_err(got_vol, inter["unprojected"]), "logits_rel": rel_err(got_logits, inter["logits"])}
