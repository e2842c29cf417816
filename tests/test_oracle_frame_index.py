"""oracle/frame_index.h (rays, 5x50x50 grid, GetFeaturesInArea with its cube-face wrap-around cases) == the reference's own Frame code
compiled in oracle/_ref/libref.so (src/Frame.cpp:158-176,251-716,728-760, include/CamModelGeneral.h:494-513)."""
import os

import numpy as np
import pytest

from cubemapslam_b200 import config, synth

ref = pytest.importorskip("oracle.ref")
if not (ref.available() or os.path.isdir("/root/reference")):
    pytest.skip("oracle/_ref/libref.so not built and no reference tree to build it from", allow_module_level=True)


@pytest.fixture(scope="module", params=[("lafida", 450), ("front", 650)])
def frame(request, oracle):
    name, W = request.param
    if name == "lafida":
        cfg = config.lafida_450(); mask = config.load_mask("gray_lafida_cubemap_mask_450"); nf = 2000
    else:
        cfg = config.front_1024(); mask = np.full((1950, 1950), 255, np.uint8); nf = 3000      # full mask: features near every face seam
    cp = oracle.cam_params(cfg)
    r = ref.Ref(cp)
    m1, m2 = r.build_maps()
    canvas = r.warp(synth.fisheye_frame(cfg, 2), m1, m2)
    rex = r.extractor(nf, 1.2, 8, 20, 7)
    F = ref.RefFrame(rex, canvas, mask)
    return W, F


def test_rays_and_grid(oracle, frame):
    W, F = frame
    rays, faces = oracle.key_point_rays(F.kps, W, W)
    assert np.array_equal(rays, F.rays)                       # bit-exact unit bearing vectors
    g = oracle.FrameGrid(F.kps, W, W)
    start, idx = g.csr()
    assert len(idx) == F.N == F.grid_count.sum()
    assert np.array_equal(np.diff(start).reshape(5, 50, 50), F.grid_count)
    rng = np.random.default_rng(0)
    for _ in range(300):
        f, c, r_ = rng.integers(0, 5), rng.integers(0, 50), rng.integers(0, 50)
        cell = (f * 50 + c) * 50 + r_
        assert np.array_equal(idx[start[cell]:start[cell + 1]], F.grid_cell(f, c, r_))


def test_features_in_area_all_wraparound_cases(oracle, frame):
    W, F = frame
    g = oracle.FrameGrid(F.kps, W, W)
    rng = np.random.default_rng(1)
    tiles = {0: (1, 1), 1: (0, 1), 2: (2, 1), 3: (1, 0), 4: (1, 2)}
    n_checked = n_nonempty = 0
    for face, (tc, tr) in tiles.items():
        for k in range(1500):
            # uniform in the face, with half of the samples pushed against an edge / a corner so that every overflow case is exercised
            u, v = rng.uniform(0, W, 2)
            mode = k % 4
            if mode >= 1:
                u = rng.choice([rng.uniform(0, 40), rng.uniform(W - 40, W - 0.01)])
            if mode >= 2:
                v = rng.choice([rng.uniform(0, 40), rng.uniform(W - 40, W - 0.01)])
            if mode == 3:
                u, v = v, u
            x = np.float32(tc * W + u); y = np.float32(tr * W + v)
            r = np.float32(rng.choice([7.0, 15.0, 15.0 * 1.728, 40.0, 75.0]))
            lv = int(rng.integers(0, 8)); lo, hi = [(-1, -1), (lv - 1, lv), (lv - 1, lv + 1), (0, 3)][k % 4]
            a = g.features_in_area(x, y, r, lo, hi); b = F.features_in_area(x, y, r, lo, hi)
            assert np.array_equal(a, b), (face, float(x), float(y), float(r), lo, hi, a, b)
            n_checked += 1; n_nonempty += len(b) > 0
    # centres outside the five faces return nothing
    assert len(g.features_in_area(10.0, 10.0, 30.0)) == 0 == len(F.features_in_area(10.0, 10.0, 30.0))
    assert n_nonempty > 0.2 * n_checked
