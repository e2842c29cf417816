"""Runs the secondary bench legs (matching, PoseOptimization, tracking) once at reduced sizes - the command profiled under ncu for
profiles/r02_*: `ncu --set full -k regex:... python tools/run_legs.py`. Prints the legs' own JSON (numbers under a profiler are not bench values)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench_extra

ap = argparse.ArgumentParser()
ap.add_argument("--match-pairs", type=int, default=512)
ap.add_argument("--match-steps", type=int, default=1)
ap.add_argument("--pose-frames", type=int, default=512)
ap.add_argument("--track-frames", type=int, default=128)
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--legs", default="match,pose,tracking,mapping")
args = ap.parse_args()
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
out = {}
legs = args.legs.split(",")
if "match" in legs:
    out["match"] = bench_extra._match_leg(torch, dev, args, 0)
if "pose" in legs:
    out["pose_optimization"] = bench_extra._pose_leg(args, 0)
if "tracking" in legs:
    out["tracking"] = bench_extra._tracking_leg(torch, dev, args, 0)
if "latency" in legs:
    out["single_frame_latency"] = bench_extra._latency_leg(args, 0)
if "mapping" in legs:
    out["local_mapping"] = bench_extra._mapping_leg(args, 0)
print(json.dumps(out))
