"""Runs a few front-end batches (used under ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
from cubemapslam_b200 import config, synth
from cubemapslam_b200.frontend import FrontEnd
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = config.front_1024()
mask = config.load_mask("gray_cubemap_front_mask_650")
fe = FrontEnd(cfg, mask, max_batch=B)
frames = np.stack([synth.fisheye_frame(cfg, i % 4) for i in range(B)])
for _ in range(reps):
    out = fe.run(frames)
print("kps", [len(o[0]) for o in out[:4]])
