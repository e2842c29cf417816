"""Pin the oracle's OpenCV-primitive models bit-exactly to opencv-python 4.13."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


def _img(rng, h, w, smooth=True):
    if not smooth:
        return rng.integers(0, 256, (h, w), dtype=np.uint8)
    base = cv2.resize(rng.integers(0, 256, (h // 8 + 1, w // 8 + 1), dtype=np.uint8), (w, h), interpolation=cv2.INTER_LINEAR).astype(np.int32)
    for _ in range(40):
        x, y = int(rng.integers(0, w - 20)), int(rng.integers(0, h - 20))
        base[y:y + int(rng.integers(4, 20)), x:x + int(rng.integers(4, 20))] = int(rng.integers(0, 256))
    base += rng.integers(-3, 4, (h, w))
    return np.clip(base, 0, 255).astype(np.uint8)


def test_remap_bilinear(oracle):
    rng = np.random.default_rng(1)
    src = _img(rng, 120, 160)
    mx = rng.uniform(-3, 163, (90, 100)).astype(np.float32)
    my = rng.uniform(-3, 123, (90, 100)).astype(np.float32)
    mx[0, :10] = 0; my[0, :10] = 0                      # the reference's "invalid" (0,0) entries
    mx[1, :10] = np.arange(10); my[1, :10] = 5.5
    ref = cv2.remap(src, mx, my, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
    assert np.array_equal(oracle.remap_bilinear(src, mx, my), ref)


@pytest.mark.parametrize("sw,sh,dw,dh", [(390, 390, 325, 325), (325, 271, 271, 226), (1350, 300, 1125, 250), (37, 53, 31, 44)])
def test_resize_linear(oracle, sw, sh, dw, dh):
    rng = np.random.default_rng(2)
    src = _img(rng, sh, sw, smooth=False)
    assert np.array_equal(oracle.resize_linear(src, dw, dh), cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR))


@pytest.mark.parametrize("thr", [20, 7])
@pytest.mark.parametrize("shape", [(37, 37), (37, 29), (120, 97), (12, 40)])
def test_fast_nms(oracle, thr, shape):
    rng = np.random.default_rng(3)
    img = _img(rng, shape[0] + 10, shape[1] + 10)
    roi = img[5:5 + shape[0], 5:5 + shape[1]]          # non-contiguous view, like the reference's cell ROI
    det = cv2.FastFeatureDetector_create(thr, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    kps = det.detect(roi)
    ref = np.array([[int(k.pt[0]), int(k.pt[1]), int(k.response)] for k in kps], np.int32).reshape(-1, 3)
    got = oracle.fast(np.ascontiguousarray(roi), thr)
    assert got.shape == ref.shape and np.array_equal(got, ref)
    assert all(k.size == 7 and k.angle == -1 and k.octave == 0 for k in kps)


@pytest.mark.parametrize("shape", [(64, 64), (200, 231), (9, 300), (5, 5)])
def test_gaussian7(oracle, shape):
    rng = np.random.default_rng(4)
    img = _img(rng, max(shape[0], 16), max(shape[1], 16), smooth=False)[:shape[0], :shape[1]].copy()
    ref = cv2.GaussianBlur(img, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
    assert np.array_equal(oracle.gaussian7(img), ref)


def test_fast_atan2(oracle):
    rng = np.random.default_rng(5)
    y = rng.integers(-200000, 200000, 100000).astype(np.float32)
    x = rng.integers(-200000, 200000, 100000).astype(np.float32)
    y[:4] = [0, 0, 1, -1]; x[:4] = [0, 5, 0, 0]
    got = oracle.fast_atan2(y, x)
    ref = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in zip(y, x)], np.float32)
    assert np.array_equal(got, ref)


def test_det_sincos_accuracy(oracle):
    """det_sincosf is DEFINED as fp64-accurate sin/cos rounded once to fp32 (the reference's glibc cosf/sinf is
    not reproducible across CPUs). Check it against numpy's fp64 sin/cos."""
    rng = np.random.default_rng(6)
    deg = rng.uniform(0, 360, 400000).astype(np.float32)
    a = (deg * np.float32(np.pi / 180.0)).astype(np.float32)
    s, c = oracle.sincos(a)
    rs = np.sin(a.astype(np.float64)).astype(np.float32); rc = np.cos(a.astype(np.float64)).astype(np.float32)
    assert (s != rs).mean() < 1e-5 and (c != rc).mean() < 1e-5
    assert np.max(np.abs(s.astype(np.float64) - np.sin(a.astype(np.float64)))) < 6.1e-8
