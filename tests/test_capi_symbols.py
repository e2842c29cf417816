"""CPU-side checks of the drop-in boundary: libcubemap_b200.so loads without a GPU and exports every function that
include/cubemap_b200.h declares; creation without a CUDA device fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    txt = open(os.path.join(ROOT, "include", "cubemap_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cslam_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from cubemapslam_b200 import _capi
    lib = _capi.lib()
    names = declared_functions()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.cslam_version() >= 100


def test_no_device_is_a_loud_error():
    import torch
    if torch.cuda.is_available():
        pytest.skip("runs on the CPU-only container")
    from cubemapslam_b200 import _capi, config
    from cubemapslam_b200.frontend import FrontEnd
    cfg = config.lafida_450()
    with pytest.raises(_capi.CslamError) as ei:
        FrontEnd(cfg, np.full((1350, 1350), 255, np.uint8), max_batch=1)
    assert "no CUDA device" in str(ei.value) or "error -1" in str(ei.value)
    from cubemapslam_b200.matcher import ORBMatcher
    from cubemapslam_b200.optimizer import Optimizer
    with pytest.raises(_capi.CslamError):
        ORBMatcher()
    with pytest.raises(_capi.CslamError):
        Optimizer()


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under cubemapslam_b200/ may import it."""
    pkg = os.path.join(ROOT, "cubemapslam_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(import oracle|from oracle)|#include\s+\"[./]*oracle", txt, flags=re.M):
                    offenders.append(f)
    assert not offenders, offenders


def test_settings_fixture_and_parser(tmp_path):
    from cubemapslam_b200 import config
    cfg = config.camera("lafida_cam0_params")
    assert cfg["Camera.Iw"] == 754 and cfg["CubeFace.w"] == 650 and cfg["ORBextractor.nFeatures"] == 2000
    assert abs(cfg["Camera.pol11"] - 0.810799620714366) < 1e-15 and cfg["Camera.nrinvpol"] == 12
    f = config.front_1024()
    assert f["Camera.Ih"] == 1024 and f["Camera.v0"] == 512.0 and f["Camera.nrinvpol"] == 10
    # the parser for the reference's "key: value" settings dialect (what a deployment would feed it)
    p = tmp_path / "s.yaml"
    p.write_text("%YAML:1.0\n# comment\nCamera.Iw: 754\nCamera.a2: 1.5e-03   # trailing\nViewer.Name: x\n")
    y = config.load_settings(str(p), Camera_Iw=800)
    assert y["Camera.Iw"] == 800 and y["Camera.a2"] == 1.5e-3 and y["Viewer.Name"] == "x"
    m = config.load_mask("gray_lafida_cubemap_mask_450")
    assert m.shape == (1350, 1350) and set(np.unique(m)) == {0, 255} and 0.29 < (m > 0).mean() < 0.31
