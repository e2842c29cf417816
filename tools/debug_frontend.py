import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
import oracle as orc
from cubemapslam_b200 import config, synth
from cubemapslam_b200.frontend import FrontEnd
cfg = config.lafida_450(); mask = config.load_mask("gray_lafida_cubemap_mask_450")
cp = orc.cam_params(cfg); m1, m2 = orc.build_maps(cp)
fe = FrontEnd(cfg, mask, max_batch=2)
frame = synth.fisheye_frame(cfg, 0)
gk, gd = fe.run(frame)
ex = orc.ORBextractor(2000, 1.2, 8, 20, 7, 450, 450)
canvas = orc.warp(cp, frame, m1, m2)
kps, desc = ex(canvas, mask)
print("n gpu", len(gk), "n ref", len(kps))
for l in range(8):
    pyr_ok = np.array_equal(fe.level_image(0, l), ex.level_image(l))
    c = ex.stage(l, 0); ref = np.stack([c["x"], c["y"], c["response"]], 1).astype(np.int32); got = fe.candidates(0, l)
    key = lambda a: a[np.lexsort((a[:, 0], a[:, 1]))]
    same = len(got) == len(ref) and np.array_equal(key(got), key(ref))
    d = ex.stage(l, 1)
    g_l = gk[gk["octave"] == l]; r_l = kps[kps["octave"] == l]
    print("L%d pyr %s cand gpu %d ref %d same %s | dist ref %d | final gpu %d ref %d" % (l, pyr_ok, len(got), len(ref), same, len(d), len(g_l), len(r_l)))
    if not same:
        sg = set(map(tuple, got.tolist())); sr = set(map(tuple, ref.tolist()))
        print("   only gpu:", sorted(sg - sr)[:8], " only ref:", sorted(sr - sg)[:8])
    # compare final keypoints as sets of level coords
    if len(g_l) != len(r_l) or not np.array_equal(g_l.view(np.uint8), r_l.view(np.uint8)):
        sg = set((float(k["x"]), float(k["y"])) for k in g_l); sr = set((float(k["x"]), float(k["y"])) for k in r_l)
        print("   kp set equal:", sg == sr, "common", len(sg & sr), " first gpu", g_l[:2], " first ref", r_l[:2])
        if sg == sr:
            order_same = np.array_equal(g_l["x"], r_l["x"]) and np.array_equal(g_l["y"], r_l["y"])
            print("   order same:", order_same, "angle same:", np.array_equal(np.sort(g_l["angle"]), np.sort(r_l["angle"])))
print("---- score probe level 0")
img = ex.level_image(0)
got = fe.candidates(0, 0)
c = ex.stage(0, 0); ref = np.stack([c["x"], c["y"], c["response"]], 1).astype(np.int32)
sr = set(map(tuple, ref.tolist()))
n = 0
for (x, y, r) in got.tolist():
    if (x, y, r) in sr: continue
    X, Y = x + 16, y + 16
    k = orc.fast(np.ascontiguousarray(img[Y - 3:Y + 4, X - 3:X + 4]), 0)
    print((x, y, r), "oracle score at same pixel:", k.tolist(), "pixels:", img[Y, X - 3:X + 4].tolist())
    n += 1
    if n > 12: break
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/dbg_l0.npz", img=img, got=got, ref=ref)
