"""The C++ oracle extractor against an independent Python/cv2 transcription of the reference control flow, plus
BASELINE config 1 (lafida yaml, 450-px faces) sanity and golden-vector checks."""
import os
import numpy as np
import pytest

from cubemapslam_b200 import config, synth

cv2 = pytest.importorskip("cv2")
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _canvas(oracle, cfg, idx):
    cp = oracle.cam_params(cfg)
    m1, m2 = oracle.build_maps(cp)
    return cp, oracle.warp(cp, synth.fisheye_frame(cfg, idx), m1, m2), (m1, m2)


def test_stages_match_cv2_transcription(oracle):
    from tests.ref_cv2_extractor import extract_stages
    cfg = config.camera("lafida_cam0_params", CubeFace_w=150, CubeFace_h=150)
    cp, canvas, _ = _canvas(oracle, cfg, 3)
    assert canvas.shape == (450, 450)
    ex = oracle.ORBextractor(500, 1.2, 8, 20, 7, 150, 150)
    mask = np.full(canvas.shape, 255, np.uint8)
    kps, desc = ex(canvas, mask)
    pyr, cands, dist, blurred, per = extract_stages(canvas, 500, 1.2, 8, 20, 7)
    assert list(ex.features_per_level) == per
    total = 0
    for l in range(8):
        assert np.array_equal(ex.level_image(l), pyr[l]), "pyramid level %d" % l
        c = ex.stage(l, 0)
        ref = np.array(cands[l], np.float32).reshape(-1, 3)
        assert len(c) == len(ref)
        assert np.array_equal(np.stack([c["x"], c["y"], c["response"]], 1), ref), "candidates level %d" % l
        d = ex.stage(l, 1)
        refd = np.array(dist[l], np.float32).reshape(-1, 4)
        assert len(d) == len(refd)
        assert np.array_equal(np.stack([d["x"], d["y"], d["response"], d["angle"]], 1), refd), "distributed level %d" % l
        assert np.all(d["octave"] == l)
        if len(d):
            assert np.array_equal(ex.level_image(l, blurred=True), blurred[l]), "blur level %d" % l
        total += len(d)
    assert 300 < len(kps) <= total                # full-255 mask; only corner-tile / bounds culling removes points


def test_cull_and_scaling(oracle):
    cfg = config.lafida_450()
    cp, canvas, _ = _canvas(oracle, cfg, 0)
    mask = config.load_mask("gray_lafida_cubemap_mask_450")
    assert mask.shape == canvas.shape == (1350, 1350)
    ex = oracle.ORBextractor(2000, 1.2, 8, 20, 7, 450, 450)
    kps, desc = ex(canvas, mask)
    assert 400 < len(kps) <= 2000 + 24 and desc.shape == (len(kps), 32)
    xi = (kps["x"] + np.float32(0.5)).astype(np.int32); yi = (kps["y"] + np.float32(0.5)).astype(np.int32)
    assert np.all(mask[yi, xi] != 0)
    assert np.all(np.diff(kps["octave"]) >= 0)                       # level-major order
    sizes = {o: float(int(np.float32(31) * ex.scale[o])) for o in range(8)}
    assert all(k["size"] == sizes[int(k["octave"])] for k in kps)
    # every descriptor row must be reproducible from the blurred level + det_sincos
    assert desc.any(axis=1).all()


def test_golden_config1(oracle):
    """Golden vector for BASELINE config 1 (tests/golden/make_golden.py wrote it from this oracle after the cv2
    transcription test above passed); guards the oracle itself against silent drift."""
    path = os.path.join(GOLD, "extract_lafida450_frame0.npz")
    g = np.load(path)
    cfg = config.lafida_450()
    cp, canvas, _ = _canvas(oracle, cfg, 0)
    mask = config.load_mask("gray_lafida_cubemap_mask_450")
    kps, desc = oracle.ORBextractor(2000, 1.2, 8, 20, 7, 450, 450)(canvas, mask)
    assert np.array_equal(kps.view(np.uint8), g["kps"].view(np.uint8)) and np.array_equal(desc, g["desc"])
    assert int(canvas.astype(np.uint64).sum()) == int(g["canvas_sum"])
