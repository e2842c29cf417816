"""Runs LocalBA config 4 (timing, per-kernel event timing, used under ncu)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cubemapslam_b200 import synth
from cubemapslam_b200.optimizer import Optimizer
dense = len(sys.argv) > 2 and sys.argv[2] == "dense"
p = synth.ba_problem(nKF=50, nMP=4000, kmin=50, kmax=50, radius=9.0) if dense else synth.ba_problem()
o = Optimizer()
a = (p["Tcw"], p["kf_fixed"], p["pts"], p["eMP"], p["eKF"], p["kpxy"], p["inv_sigma2"], 650, 650)
o.LocalBundleAdjustment(*a, its1=1, its2=0)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    t0 = time.perf_counter(); g = o.LocalBundleAdjustment(*a); dt = time.perf_counter() - t0
    print("edges", len(p["eMP"]), "iters", g["iters"], "trials", g["trials"], "ms", round(dt * 1e3, 2), "launches", o.launches)
o.set_timing(True)
g = o.LocalBundleAdjustment(*a)
tm = o.timing(); o.set_timing(False)
tot = sum(v[0] for v in tm.values())
for k, (ms, n) in sorted(tm.items(), key=lambda kv: -kv[1][0]):
    print("  %-16s %7.1f us avg over %3d  (%4.1f%%)" % (k, 1e3 * ms / n, n, 100 * ms / tot))
print("  GPU time per LM trial (events): %.1f us" % (1e3 * tot / max(g["trials"], 1)))
