"""CPU check of the product's wrap-around TABLE (cubemapslam_b200/csrc/area_table.cuh, exported as the host utility cslam_area_rects) against the
oracle's branch-by-branch restatement of Frame::GetFeaturesInArea (itself pinned to the compiled reference in tests/test_oracle_frame_index.py)."""
import numpy as np

from cubemapslam_b200 import tracker


def test_area_table_equals_oracle_cases(oracle):
    rng = np.random.default_rng(5)
    tiles = {0: (1, 1), 1: (0, 1), 2: (2, 1), 3: (1, 0), 4: (1, 2)}
    seen = set()
    for W in (450, 650):
        for face, (tc, tr) in tiles.items():
            for k in range(4000):
                u, v = rng.uniform(0, W, 2)
                if k % 4 >= 1:
                    u = rng.choice([rng.uniform(0, 60), rng.uniform(W - 60, W - 0.01)])
                if k % 4 >= 2:
                    v = rng.choice([rng.uniform(0, 60), rng.uniform(W - 60, W - 0.01)])
                if k % 4 == 3:
                    u, v = v, u
                x = np.float32(tc * W + u); y = np.float32(tr * W + v); r = np.float32(rng.choice([4.0, 15.0, 26.0, 54.0, 90.0]))
                a = tracker.area_rects(x, y, r, W, W); b = oracle.area_rects(x, y, r, W, W)
                assert a.shape == b.shape and np.array_equal(a, b), (face, float(x), float(y), float(r), a, b)
                seen.add((face, len(b), tuple(b[:, 0])))
    assert len(seen) >= 30                                  # every face met in-face, edge and corner cases
    assert len(tracker.area_rects(5.0, 5.0, 20.0, 450, 450)) == 0
