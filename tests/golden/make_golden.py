"""Writes the committed golden vectors under tests/golden/ from the C++ oracle.

Run ONLY after tests/test_oracle_cv2.py and test_oracle_extractor.py::test_stages_match_cv2_transcription pass:
those pin the oracle to OpenCV 4.13 and to the independent transcription of the reference control flow; the golden
files then freeze that state (the reference itself ships no golden vectors: SURVEY.md §4).
  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import cv2  # noqa: E402
import oracle as orc  # noqa: E402
from cubemapslam_b200 import config, synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def extract_golden():
    cfg = config.lafida_450()
    cp = orc.cam_params(cfg)
    m1, m2 = orc.build_maps(cp)
    canvas = orc.warp(cp, synth.fisheye_frame(cfg, 0), m1, m2)
    mask = config.load_mask("gray_lafida_cubemap_mask_450")
    kps, desc = orc.ORBextractor(2000, 1.2, 8, 20, 7, 450, 450)(canvas, mask)
    np.savez_compressed(os.path.join(HERE, "extract_lafida450_frame0.npz"), kps=kps, desc=desc, canvas_sum=np.uint64(canvas.astype(np.uint64).sum()),
                        map_probe=np.stack([m1[::97, ::89], m2[::97, ::89]]))
    print("extract golden:", len(kps), "keypoints")


def match_golden():
    A, angA, B, angB, perm = synth.descriptor_pair(0, n=500)
    n, m, d, s = orc.match_bruteforce(A, angA, B, angB, 0.6, 50, True)
    rng = np.random.default_rng(9)
    nodeA = rng.integers(0, 40, 500).astype(np.int32); nodeB = nodeA[perm].copy()
    nodeB[rng.random(500) < 0.1] = 41
    valid = (rng.random(500) < 0.8).astype(np.uint8)
    nb, mf = orc.search_by_bow(A, angA, valid, nodeA, B, angB, nodeB, 0.7, True)
    np.savez_compressed(os.path.join(HERE, "match_pair0_n500.npz"), bf_n=n, bf_match=m, bf_dist=d, bf_second=s, nodeA=nodeA, nodeB=nodeB, valid=valid,
                        bow_n=nb, bow_match=mf)
    print("match golden: bf", n, "bow", nb)


def ba_golden():
    p = synth.ba_problem(nKF=8, nMP=300, kmin=2, kmax=6, faceW=450, seed=11, radius=1.5)
    r = orc.local_ba(p["Tcw"], p["kf_fixed"], p["pts"], p["eMP"], p["eKF"], p["kpxy"], p["inv_sigma2"], 450, 450)
    q = synth.pose_problem(n=200, faceW=450, seed=5)
    o = orc.pose_opt(q["Tcw"], q["Xw"], q["kpxy"], q["inv_sigma2"], 450, 450)
    np.savez_compressed(os.path.join(HERE, "ba_small.npz"), pose64=r["pose64"], pts64=r["pts64"], outlier=r["outlier"], log=r["log"],
                        po_pose64=o["pose64"], po_outlier=o["outlier"], po_inliers=o["inliers"], po_log=o["log"])
    print("ba golden: iters", r["iters"], "outliers", int(r["outlier"].sum()), "| pose inliers", o["inliers"])


if __name__ == "__main__":
    extract_golden()
    match_golden()
    ba_golden()
