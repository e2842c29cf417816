"""GPU parity: CUDA warp + ORB extraction vs the CPU oracle, bit-exact, through the C ABI (tests marked gpu)."""
import numpy as np
import pytest

from cubemapslam_b200 import config, synth

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


def _fe(cfg, mask, **kw):
    from cubemapslam_b200.frontend import FrontEnd
    return FrontEnd(cfg, mask, **kw)


@pytest.fixture(scope="module")
def lafida(oracle):
    cfg = config.lafida_450()
    mask = config.load_mask("gray_lafida_cubemap_mask_450")
    cp = oracle.cam_params(cfg)
    m1, m2 = oracle.build_maps(cp)
    fe = _fe(cfg, mask, max_batch=4)
    yield cfg, mask, cp, (m1, m2), fe
    fe.close()


def test_tables_and_maps(oracle, lafida):
    cfg, mask, cp, (m1, m2), fe = lafida
    ex = oracle.ORBextractor(2000, 1.2, 8, 20, 7, 450, 450)
    t = fe.tables()
    assert np.array_equal(t["scale"], ex.scale) and np.array_equal(t["inv_sigma2"], ex.inv_sigma2)
    assert np.array_equal(t["features_per_level"], ex.features_per_level) and np.array_equal(t["umax"], ex.umax)
    g1, g2 = fe.maps()
    assert np.array_equal(g1, m1) and np.array_equal(g2, m2)


def test_warp_bit_exact(oracle, lafida):
    cfg, mask, cp, (m1, m2), fe = lafida
    frames = np.stack([synth.fisheye_frame(cfg, i) for i in range(3)])
    got = fe.warp(frames)
    for i in range(3):
        assert np.array_equal(got[i], oracle.warp(cp, frames[i], m1, m2)), "frame %d" % i
    # corner tiles of a caller-provided canvas stay untouched
    canvas = np.full((1350, 1350), 7, np.uint8)
    out = fe.warp(frames[0], canvas[None].copy())
    assert np.all(out[:450, :450] == 7) and np.all(out[900:, 900:] == 7) and np.array_equal(out[450:900], got[0][450:900])


def test_stages_and_extract_config1(oracle, lafida):
    cfg, mask, cp, (m1, m2), fe = lafida
    frames = np.stack([synth.fisheye_frame(cfg, i) for i in range(2)])
    res = fe.run(frames)
    ex = oracle.ORBextractor(2000, 1.2, 8, 20, 7, 450, 450)
    for f in range(2):
        canvas = oracle.warp(cp, frames[f], m1, m2)
        kps, desc = ex(canvas, mask)
        for l in range(8):
            assert np.array_equal(fe.level_image(f, l), ex.level_image(l)), "pyramid f%d l%d" % (f, l)
            c = ex.stage(l, 0)
            ref = np.stack([c["x"], c["y"], c["response"]], 1).astype(np.int32)
            got = fe.candidates(f, l)
            assert len(got) == len(ref), "candidate count f%d l%d: %d vs %d" % (f, l, len(got), len(ref))
            key = lambda a: a[np.lexsort((a[:, 0], a[:, 1]))]
            assert np.array_equal(key(got), key(ref)), "candidates f%d l%d" % (f, l)
        gk, gd = res[f]
        assert len(gk) == len(kps), "keypoint count f%d: %d vs %d" % (f, len(gk), len(kps))
        assert np.array_equal(gk.view(np.uint8), kps.view(np.uint8)), "keypoints f%d" % f
        assert np.array_equal(gd, desc), "descriptors f%d" % f


def test_golden_config1(lafida):
    import os
    cfg, mask, cp, maps, fe = lafida
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "extract_lafida450_frame0.npz"))
    kps, desc = fe.run(synth.fisheye_frame(cfg, 0))
    assert np.array_equal(kps.view(np.uint8), g["kps"].view(np.uint8)) and np.array_equal(desc, g["desc"])


def test_extract_host_canvas_and_full_mask(oracle, lafida):
    cfg, mask, cp, (m1, m2), fe = lafida
    canvas = oracle.warp(cp, synth.fisheye_frame(cfg, 5), m1, m2)
    rng = np.random.default_rng(0)
    canvas[:450, :450] = rng.integers(0, 256, (450, 450), dtype=np.uint8)      # textured corner tile: culled as UNKNOWN_FACE
    kps, desc = oracle.ORBextractor(2000, 1.2, 8, 20, 7, 450, 450)(canvas, mask)
    gk, gd = fe.extract(canvas)
    assert np.array_equal(gk.view(np.uint8), kps.view(np.uint8)) and np.array_equal(gd, desc)
    # and the warp path afterwards must see zeroed corners again
    k2, d2 = fe.run(synth.fisheye_frame(cfg, 0))
    r2 = oracle.ORBextractor(2000, 1.2, 8, 20, 7, 450, 450)(oracle.warp(cp, synth.fisheye_frame(cfg, 0), m1, m2), mask)
    assert np.array_equal(k2.view(np.uint8), r2[0].view(np.uint8)) and np.array_equal(d2, r2[1])


def test_config2_frame_650(oracle):
    cfg = config.front_1024()
    mask = config.load_mask("gray_cubemap_front_mask_650")
    fe = _fe(cfg, mask, max_batch=2)
    cp = oracle.cam_params(cfg)
    m1, m2 = oracle.build_maps(cp)
    frames = np.stack([synth.fisheye_frame(cfg, i) for i in (0, 7)])
    res = fe.run(frames)
    ex = oracle.ORBextractor(3000, 1.2, 8, 20, 7, 650, 650)
    for f in range(2):
        kps, desc = ex(oracle.warp(cp, frames[f], m1, m2), mask)
        assert len(res[f][0]) == len(kps) and len(kps) > 1500
        assert np.array_equal(res[f][0].view(np.uint8), kps.view(np.uint8)) and np.array_equal(res[f][1], desc)
    fe.close()


def test_degenerate_frames(oracle, lafida):
    cfg, mask, cp, (m1, m2), fe = lafida
    black = np.zeros((cp.Ih, cp.Iw), np.uint8)
    kps, desc = fe.run(black)
    assert len(kps) == 0 and desc.shape == (0, 32)
    flat = np.full((cp.Ih, cp.Iw), 200, np.uint8)       # only the image-circle / face seams produce corners
    kps, desc = fe.run(flat)
    rk, rd = oracle.ORBextractor(2000, 1.2, 8, 20, 7, 450, 450)(oracle.warp(cp, flat, m1, m2), mask)
    assert np.array_equal(kps.view(np.uint8), rk.view(np.uint8)) and np.array_equal(desc, rd)


@pytest.mark.parametrize("face", [560, 570, 610, 333])
def test_face_size_sweep(oracle, face):
    """Face sizes whose last FAST cell row / column is shorter than 7 px (k_fast indexed shared memory at ny < 0);
    560 -> level 1, 570 -> level 0, 610 -> level 3 have a 4-row last cell row at scale 1.2."""
    cfg = config.camera("lafida_cam0_params", CubeFace_w=face, CubeFace_h=face)
    mask = np.full((3 * face, 3 * face), 255, np.uint8)
    cp = oracle.cam_params(cfg)
    m1, m2 = oracle.build_maps(cp)
    fe = _fe(cfg, mask, max_batch=1)
    frame = synth.fisheye_frame(cfg, 3)
    gk, gd = fe.run(frame)
    rk, rd = oracle.ORBextractor(2000, 1.2, 8, 20, 7, face, face)(oracle.warp(cp, frame, m1, m2), mask)
    assert len(gk) == len(rk) and len(rk) > 500
    assert np.array_equal(gk.view(np.uint8), rk.view(np.uint8)) and np.array_equal(gd, rd)
    fe.close()


def test_reference_generated_goldens(lafida):
    """Golden vectors produced by the reference's own ORBExtractor.cpp / CamModelGeneral.cpp compiled unmodified (oracle/_ref,
    tests/golden/make_golden_ref.py): config 1 (lafida, 450-px faces) and a config-2 frame (front camera, 650-px faces, 3000 features)."""
    import os
    gold = os.path.join(os.path.dirname(__file__), "golden")
    cfg, mask, cp, maps, fe = lafida
    g = np.load(os.path.join(gold, "ref_extract_lafida450_frame0.npz"))
    kps, desc = fe.run(synth.fisheye_frame(cfg, int(g["frame_idx"])))
    assert np.array_equal(kps.view(np.uint8), g["kps"].view(np.uint8)) and np.array_equal(desc, g["desc"])
    cfg2 = config.front_1024()
    fe2 = _fe(cfg2, config.load_mask("gray_cubemap_front_mask_650"), max_batch=1)
    g = np.load(os.path.join(gold, "ref_extract_front650_frame7.npz"))
    kps, desc = fe2.run(synth.fisheye_frame(cfg2, int(g["frame_idx"])))
    assert len(kps) == len(g["kps"]) > 1500
    assert np.array_equal(kps.view(np.uint8), g["kps"].view(np.uint8)) and np.array_equal(desc, g["desc"])
    fe2.close()
