"""Host-side logic and the C-ABI surface, no GPU needed."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT
import lt_b200
from lt_b200 import capi, multiview, testing, volumetric, op


def test_library_loads_and_exports_every_declared_symbol():
    lib = capi.lib()                     # builds with nvcc if needed; raises if missing
    header = open(os.path.join(ROOT, "include", "lt_b200.h")).read()
    declared = set(re.findall(r"\b(lt_[a-z0-9_]+)\s*\(", header))
    declared -= {"lt_conv_desc"}
    assert declared, "no declarations parsed"
    assert declared == set(capi.SIGNATURES), (declared ^ set(capi.SIGNATURES))
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.lt_version() >= 100


def test_conv_desc_layout_matches_header():
    header = open(os.path.join(ROOT, "include", "lt_b200.h")).read()
    body = header[header.index("typedef struct lt_conv_desc {"):header.index("} lt_conv_desc;")]
    fields = []
    for line in body.splitlines()[1:]:
        line = line.split("/*")[0].strip()
        for ctype, c in (("int ", ctypes.c_int), ("void* ", ctypes.c_void_p), ("size_t ", ctypes.c_size_t)):
            if line.startswith(ctype):
                fields += [(f.strip(), c) for f in line[len(ctype):].rstrip(";").split(",")]
    assert fields == [(f[0], f[1]) for f in capi.ConvDesc._fields_]
    assert ctypes.sizeof(capi.ConvDesc) == sum(ctypes.sizeof(c) for _, c in fields)   # 32 ints, then 8-byte aligned


def test_options_layout_matches_header():
    """capi.Options mirrors `struct lt_options` field by field (all ints), and the defaults come from the library."""
    header = open(os.path.join(ROOT, "include", "lt_b200.h")).read()
    body = header[header.index("typedef struct lt_options {"):header.index("} lt_options;")]
    fields = []
    for line in body.splitlines()[1:]:
        line = line.split("/*")[0].strip()
        if line.startswith("int "):
            fields += [f.strip() for f in line[4:].rstrip(";").split(",")]
    assert fields == [f[0] for f in capi.Options._fields_]
    o = capi.get_options()
    assert o["tc_splitk"] == 1 and o["unproject_cpl"] in (4, 8) and o["pair_prof"] == 0


def test_stack_projections_equals_per_camera_path():
    cams = testing.make_cameras(3, image_size=384)
    cameras = [[testing.Camera(c.R, c.t, c.K) for _ in range(2)] for c in cams]
    before = [c[0].K.copy() for c in cameras]
    P = multiview.stack_projections(cameras, (384, 384), (96, 96))
    assert P.shape == (2, 3, 3, 4) and P.dtype == np.float32
    for v in range(3):
        c = testing.Camera(cams[v].R, cams[v].t, cams[v].K)
        c.update_after_resize((384, 384), (96, 96))
        assert np.array_equal(P[1, v], c.projection.astype(np.float32))
        assert np.array_equal(cameras[v][0].K, before[v]), "caller's cameras must not be mutated"


def test_rotation_matrix_identity_and_orthonormal():
    assert np.allclose(volumetric.get_rotation_matrix([0, 0, 1], 0.0), np.eye(3))
    R = volumetric.get_rotation_matrix([0, 0, 1], 0.7)
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-12)


def test_attrdict_hasattr_semantics():
    cfg = testing.make_config()
    assert not hasattr(cfg.model, "transfer_cmu_to_human36m")
    assert cfg.model.backbone.num_layers == 152


def test_constructor_mutates_config_like_reference_and_freezes_head():
    cfg = testing.make_config(num_layers=18, volume_size=32, aggregation="conf")
    m = lt_b200.VolumetricTriangulationNet(cfg, device="cpu", backend="torch")
    assert cfg.model.backbone.vol_confidences is True and cfg.model.backbone.alg_confidences is False
    assert hasattr(m.backbone, "vol_confidences")
    assert all(not p.requires_grad for p in m.backbone.final_layer.parameters())
    assert all(p.requires_grad for p in m.process_features.parameters())


def test_state_dict_keys_match_reference_key_set():
    want = json.load(open(os.path.join(GOLDEN, "state_dict_r152.json")))
    m = lt_b200.VolumetricTriangulationNet(testing.make_config(num_layers=152), device="cpu", backend="torch")
    have = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert have == want
    assert len(have) == 1311


def test_native_backend_refuses_cpu_tensors_and_training():
    m = lt_b200.VolumetricTriangulationNet(testing.make_config(num_layers=18, volume_size=32), device="cpu")
    assert m.backend == "native"
    images, batch = testing.make_batch(1, 2, image_size=64)
    with pytest.raises(RuntimeError):
        m.eval()(images, None, batch)
    with pytest.raises(RuntimeError):
        op.unproject_heatmaps(torch.zeros(1, 1, 4, 4, 4), torch.zeros(1, 1, 3, 4), torch.zeros(1, 2, 2, 2, 3))
    with pytest.raises(ValueError):
        op.unproject_heatmaps(torch.zeros(1, 1, 4, 4, 4), torch.zeros(1, 1, 3, 4), torch.zeros(1, 2, 2, 2, 3), "median")


def test_header_is_plain_c_and_the_library_is_callable_from_c(tmp_path):
    """The drop-in boundary is a C ABI: include/lt_b200.h must compile as C99 (no C++-isms, no torch types) and a plain C program must be
    able to dlopen the library and call it (no GPU work here: version, default options, error string of a rejected call)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "use_lt.c"
    src.write_text(r'''
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include "lt_b200.h"
int main(int argc, char** argv) {
  void* h = dlopen(argv[1], RTLD_NOW);
  if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
  int (*version)(void) = (int (*)(void))dlsym(h, "lt_version");
  void (*defaults)(lt_options*) = (void (*)(lt_options*))dlsym(h, "lt_default_options");
  const char* (*last_error)(void) = (const char* (*)(void))dlsym(h, "lt_last_error_string");
  int (*coord)(const float*, const float*, const float*, const float*, float*, int, int, int, void*) =
      (int (*)(const float*, const float*, const float*, const float*, float*, int, int, int, void*))dlsym(h, "lt_coord_volume_fwd");
  if (!version || !defaults || !last_error || !coord) return 3;
  lt_options o;
  memset(&o, 0xff, sizeof o);
  defaults(&o);
  int rc = coord(0, 0, 0, 0, 0, 1, 64, 0, 0);          /* null pointers: must be rejected with a message, not crash */
  printf("%d %d %d %d %s\n", version(), o.tc_splitk, o.pair_two_acc, rc, last_error());
  return 0;
}
''')
    exe = tmp_path / "use_lt"
    inc = os.path.join(ROOT, "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", inc, str(src), "-o", str(exe), "-ldl"], check=True)
    capi.lib()
    out = subprocess.run([str(exe), capi.LIB_PATH], capture_output=True, text=True, check=True).stdout.split(None, 4)
    assert int(out[0]) >= 100 and int(out[1]) == 1 and int(out[2]) == 1
    assert int(out[3]) < 0 and "null pointer" in out[4]
