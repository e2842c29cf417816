"""Runs LocalBA config 4 (used under ncu / for timing)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cubemapslam_b200 import synth
from cubemapslam_b200.optimizer import Optimizer
p = synth.ba_problem()
o = Optimizer()
a = (p["Tcw"], p["kf_fixed"], p["pts"], p["eMP"], p["eKF"], p["kpxy"], p["inv_sigma2"], 650, 650)
o.LocalBundleAdjustment(*a, its1=1, its2=0)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    t0 = time.perf_counter(); g = o.LocalBundleAdjustment(*a); dt = time.perf_counter() - t0
    print("iters", g["iters"], "trials", g["trials"], "ms", round(dt * 1e3, 2), "launches", o.launches)
