"""Settings files in the reference's YAML dialect (OpenCV FileStorage "key: value" lines).

Mirrors what System / Tracking read (reference src/System.cpp:63-91, src/Tracking.cpp:61-96): the same keys,
`poly` always 5 coefficients and `invpoly` 12, zero padded (src/System.cpp:67-72).
"""
import os

_FIXTURES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")   # camera parameters + bit-packed cube masks of the reference settings


def load_settings(path, **overrides):
    """Parse a reference settings YAML into a flat dict {key: float}; `overrides` use '_' for '.' in key names
    (e.g. CubeFace_w=450) or are given through the `overrides` mapping semantics of dict.update."""
    cfg = {}
    with open(path) as f:
        for line in f:
            line = line.split("#", 1)[0].strip()
            if not line or line.startswith("%") or ":" not in line:
                continue
            k, v = line.split(":", 1)
            v = v.strip()
            try:
                cfg[k.strip()] = float(v)
            except ValueError:
                cfg[k.strip()] = v
    for k, v in overrides.items():
        cfg[k.replace("_", ".", 1)] = float(v)
    return cfg


def fixture(name):
    return os.path.join(_FIXTURES, name)


def camera(name, **overrides):
    """Parameters of one of the reference's settings files (cubemapslam_b200/data/cameras.json holds the values of
    Config/{lafida_cam0,front_cam,left_cam}_params.yaml under the reference's own key names)."""
    import json
    cfg = dict(json.load(open(fixture("cameras.json")))["cameras"][name])
    for k, v in overrides.items():
        cfg[k.replace("_", ".", 1)] = float(v)
    return cfg


def load_mask(name):
    """Binary cubemap mask (0/255, uint8) shipped by the reference under Masks/<name>.png, stored bit-packed in
    cubemapslam_b200/data/cubemap_masks.npz."""
    import numpy as np
    z = np.load(fixture("cubemap_masks.npz"))
    h, w = (int(v) for v in z[name + "_shape"])
    return (np.unpackbits(z[name + "_bits"])[:h * w].reshape(h, w) * 255).astype(np.uint8)


def lafida_450():
    """BASELINE config 1: lafida_cam0_params.yaml with 450-px faces (mask gray_lafida_cubemap_mask_450.png)."""
    return camera("lafida_cam0_params", CubeFace_w=450, CubeFace_h=450)


def front_1024():
    """BASELINE configs 2/5: front_cam_params.yaml with a 1280x1024 synthetic sensor (Ih 1024, v0 512), 650-px faces."""
    return camera("front_cam_params", Camera_Ih=1024, Camera_v0=512.0)
