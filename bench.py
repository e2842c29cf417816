#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on the host cores

Workload at N=1 = BASELINE.json configs[1]: batched fisheye->cubemap warp + ORB extraction of 4096 synthetic
1280x1024 frames, 650-px faces (front_cam_params.yaml with Ih=1024, nFeatures 3000). A "step" is one pass over all
4096 frames (launched in batches of --batch). Frames are sharded over ranks without any collective; per-rank work is
fixed ("weak" scaling), `value` is the whole-job frames/s = N * frames * K / max-over-ranks device time.
Extra legs (matching pairs/s, LocalBA LM iterations/s) are reported under "extra" in the same JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from cubemapslam_b200 import config, synth  # noqa: E402

N_BASE = 16          # distinct synthetic frames generated with the SURVEY §8d recipe; the rest are circular shifts of them


def load_mask():
    import cv2
    return config.load_mask("gray_cubemap_front_mask_650")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampling during the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.idx = gpu_index; self.rows = []; self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True); self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = []; mx = None; reasons = set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def make_frames_device(torch, cfg, frames, device):
    """frames x Ih x Iw uint8 on the device: N_BASE recipe frames, the others are horizontal rolls re-masked by the image circle."""
    Iw, Ih = int(cfg["Camera.Iw"]), int(cfg["Camera.Ih"])
    base = np.stack([synth.fisheye_frame(cfg, i) for i in range(min(N_BASE, frames))])
    tb = torch.from_numpy(base).to(device)
    circle = torch.from_numpy((synth.fisheye_frame(cfg, 0) > 0) | (base[0] > 0)).to(device)   # support of the image circle
    yy, xx = np.mgrid[0:Ih, 0:Iw]
    invp = [cfg.get("Camera.pol%d" % i, 0.0) for i in range(int(cfg["Camera.nrinvpol"]))]
    import math
    fov = cfg["Camera.fov"] / 2.0 * math.pi / 180.0
    rho = synth._horner(invp, math.atan(-math.cos(fov) / math.sin(fov)))
    circle = torch.from_numpy(((xx - cfg["Camera.u0"]) ** 2 + (yy - cfg["Camera.v0"]) ** 2 <= rho ** 2)).to(device)
    out = torch.empty((frames, Ih, Iw), dtype=torch.uint8, device=device)
    for i in range(frames):
        b = tb[i % len(base)]
        shift = (i // len(base)) * 7
        out[i] = torch.roll(b, shifts=shift, dims=1) * circle if shift else b
    return out, base


def geometry_bytes(cfg, nlevels=8, scale=1.2):
    """Algorithmic bytes per frame of each kernel: every input byte read once, every output written once."""
    Iw, Ih, W = int(cfg["Camera.Iw"]), int(cfg["Camera.Ih"]), int(cfg["CubeFace.w"])
    sizes = []; s = np.float32(1.0)
    for l in range(nlevels):
        inv = np.float32(1.0) / s
        sizes.append(int(np.rint(np.float32(3 * W) * inv)) ** 2)
        s = np.float32(s * np.float32(scale))
    return {"k_warp": Iw * Ih + 5 * W * W, "k_pyramid": sum(sizes[:-1]) + sum(sizes[1:]), "k_fast": sum(sizes), "levels_px": sizes}


def run_ours(args):
    import torch
    from cubemapslam_b200.frontend import FrontEnd
    from cubemapslam_b200 import _capi
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    cfg = config.front_1024()
    mask = load_mask()
    B = args.batch
    fe = FrontEnd(cfg, mask, max_batch=B, device=local)
    frames = args.frames
    fish, base_host = make_frames_device(torch, cfg, frames, dev)
    fsz = fish.shape[1] * fish.shape[2]
    kp_cap = fe.kp_cap
    kps = torch.empty((frames, kp_cap, 28), dtype=torch.uint8, device=dev)
    desc = torch.empty((frames, kp_cap, 32), dtype=torch.uint8, device=dev)
    nout = torch.empty((frames,), dtype=torch.int32, device=dev)
    stream = torch.cuda.ExternalStream(fe.stream, device=dev)
    torch.cuda.synchronize()

    def one_step():
        for c in range(0, frames, B):
            b = min(B, frames - c)
            fe.run_dev(fish.data_ptr() + c * fsz, b, kps.data_ptr() + c * kp_cap * 28, desc.data_ptr() + c * kp_cap * 32, nout.data_ptr() + c * 4)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_step()
    fe.sync()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = fe.launches
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        one_step()
    e1.record(stream)
    fe.sync()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = fe.launches - l0
    clocks = sampler.stop() if rank == 0 else None
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    value = world * frames * args.steps / (ms_max / 1e3)
    nkp = float(nout.float().mean().item())

    # ---- end-to-end through the reference-facing host call (pinned host frames in, host keypoints out)
    ef = min(args.e2e_frames, frames)
    h_in = torch.empty((ef, fish.shape[1], fish.shape[2]), dtype=torch.uint8).pin_memory()
    h_in.copy_(fish[:ef].cpu())
    # page-locked result buffers: the library DMAs straight into caller memory when it is pinned
    t_kps = torch.empty((B, kp_cap, 28), dtype=torch.uint8).pin_memory(); t_desc = torch.empty((B, kp_cap, 32), dtype=torch.uint8).pin_memory()
    t_n = torch.empty((B,), dtype=torch.int32).pin_memory()
    h_kps = t_kps.numpy(); h_desc = t_desc.numpy(); h_n = t_n.numpy()
    h_np = h_in.numpy()

    def e2e_pass():
        tot = 0
        for c in range(0, ef, B):
            b = min(B, ef - c)
            fe.run_raw(h_np[c:c + b], h_kps, h_desc, h_n)
            tot += int(h_n[:b].sum())
        return tot
    e2e_pass()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        tot = e2e_pass()
    barrier()
    t_e2e = time.perf_counter() - t0
    te = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * ef * args.e2e_steps / float(te.item())
    h2d = ef * fsz; d2h = ef * (kp_cap * 60 + 4)

    sharded = None
    if world > 1 and not args.no_extra:
        import bench_extra
        sharded = bench_extra.run_sharded_ba(local, rank, world, dist, torch)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel (separate short pass with events between launches)
    gb = geometry_bytes(cfg)
    fe.set_timing(True)
    for c in range(0, min(frames, 4 * B), B):
        fe.run_dev(fish.data_ptr() + c * fsz, B, kps.data_ptr() + c * kp_cap * 28, desc.data_ptr() + c * kp_cap * 32, nout.data_ptr() + c * 4)
        fe.sync()
    tm = fe.timing(); fe.set_timing(False)
    total_ms = sum(v[0] for v in tm.values()) or 1.0
    shares = {k: round(v[0] / total_ms, 4) for k, v in tm.items()}
    algo = {"k_warp": gb["k_warp"], "k_pyramid": gb["k_pyramid"], "k_fast": gb["k_fast"],
            "k_describe": int(nkp * (43 * 43 + 31 * 31 + 60)), "k_distribute": None}
    dom = max(tm, key=lambda k: tm[k][0])
    hbm, tf, how = peaks()
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        if dom in tj.get("dram_bytes_per_frame", {}):
            traffic = int(tj["dram_bytes_per_frame"][dom] * B)       # per launch group of one batch, like `achieved`
    roof = {"kernel": dom, "bound": "hbm", "peak": hbm, "peak_source": how, "unit": "GB/s", "traffic": traffic, "share_of_step": shares[dom], "kernel_shares": shares}
    if dom == "k_fast":
        roof["note"] = ("k_fast is ALU-pipe bound, not HBM bound: the exact FAST arc score costs ~60 integer instructions per pixel "
                        "(ncu: ALU pipe 68 % active, DRAM 2-3 %); `frac` is reported against HBM as the contract asks")
    nbatches = min(frames, 4 * B) // B
    if algo.get(dom):
        # per launch: bytes of one batch; k_pyramid/k_fast are nlevels(-1) launches per batch -> use the per-batch total
        per_batch_ms = tm[dom][0] / nbatches
        roof["achieved"] = round(algo[dom] * B / (per_batch_ms / 1e3) / 1e9, 2)
        roof["frac"] = round(roof["achieved"] / hbm, 4)
        roof["algorithmic_bytes_per_frame"] = algo[dom]
    else:
        roof["achieved"] = None; roof["frac"] = None; roof["note"] = "latency-bound quadtree kernel; no HBM roofline applies"
    whole = sum(v for v in [gb["k_warp"], gb["k_pyramid"], gb["k_fast"], 2 * gb["k_fast"], int(nkp * (43 * 43 + 31 * 31))])
    roof["pipeline_algorithmic_bytes_per_frame"] = whole          # SURVEY §8d accounting (incl. the reference's whole-level blur)
    roof["pipeline_frac"] = round(whole * (value / world) / 1e9 / hbm, 4)

    # ---- CPU baseline: the oracle port on the host cores, bounded sample of the same workload
    cpu = cpu_baseline(cfg, mask, base_host, args)
    extra = {}
    if sharded is not None:
        extra = {"local_ba_sharded": sharded}
    elif not args.no_extra:
        import bench_extra
        fe.close()
        del fish, kps, desc
        torch.cuda.empty_cache()
        extra = bench_extra.run(args, local)
    out = {"metric": "warp+ORB-extract frames/sec", "value": round(value, 1), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": round(ms_max / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "u8", "data": "synthetic (SURVEY §8d recipe: %d distinct frames, others are re-masked circular shifts)" % N_BASE,
           "config": {"workload": "configs[1]: warp + ORB extract, %d synthetic 1280x1024 frames per GPU, 650-px faces, nFeatures 3000" % frames,
                      "frames_per_gpu": frames, "batch": B, "l2": "inputs (%.1f GB) larger than L2" % (frames * fsz / 1e9), "parallelism": "frame-shard x%d" % world,
                      "mean_keypoints_per_frame": round(nkp, 1)},
           "clocks": clocks, "gpu_launches": int(launches),
           "e2e": {"value": round(e2e_value, 1), "unit": "frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                   "frames_per_step": ef, "steps": args.e2e_steps},
           "roofline": roof, "cpu_baseline": cpu, "extra": extra}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def host_cores():
    """CPUs this process may really use: the cgroup CPU quota if there is one (the GPU boxes expose 128 logical CPUs but a
    16-CPU quota), else os.cpu_count()."""
    n = os.cpu_count() or 1
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = max(1, min(n, int(round(int(q) / int(per)))))
    except Exception:
        pass
    return n


def cpu_baseline(cfg, mask, base_frames, args, threads=None, nframes=None):
    import oracle as orc
    quota = host_cores()
    cores = threads or min(os.cpu_count() or 1, 2 * quota)     # 2 worker threads per granted CPU measured fastest on the box
    cp = orc.cam_params(cfg)
    m1, m2 = orc.build_maps(cp)
    n = nframes or 20 * cores          # ~10 s of CPU work at the measured rate
    fr = np.stack([base_frames[i % len(base_frames)] for i in range(n)])
    orc.warp_extract_batch(cp, fr[:min(n, cores)], m1, m2, mask, 3000, 1.2, 8, 20, 7, cores)     # warm-up
    t0 = time.perf_counter()
    tot = orc.warp_extract_batch(cp, fr, m1, m2, mask, 3000, 1.2, 8, 20, 7, cores)
    dt = time.perf_counter() - t0
    return {"value": round(n / dt, 2), "unit": "frames/s", "cores": quota, "threads": cores, "kind": "port",
            "sample": "%d frames of the same workload, %d independent worker threads on %d granted CPUs (C++ oracle, -O3)" % (n, cores, quota), "seconds": round(dt, 2),
            "mean_keypoints_per_frame": round(tot / n, 1)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = config.front_1024()
    mask = load_mask()
    base = np.stack([synth.fisheye_frame(cfg, i) for i in range(N_BASE)])
    vals = []
    for i in range(args.warmup + args.steps):
        c = cpu_baseline(cfg, mask, base, args, nframes=256)
        if i >= args.warmup:
            vals.append(c)
    v = float(np.mean([c["value"] for c in vals]))
    c = vals[-1]; c["value"] = round(v, 2)
    out = {"impl": "reference", "metric": "warp+ORB-extract frames/sec", "value": round(v, 2), "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": round(1e3 * c["seconds"], 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
           "data": "synthetic", "config": {"workload": "configs[1]: warp + ORB extract, 1280x1024 frames, 650-px faces, nFeatures 3000 (bounded sample per step)"},
           "cpu_baseline": c, "e2e": {"value": round(v, 2), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=4096)
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--e2e-frames", type=int, default=1024)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--match-pairs", type=int, default=4096)
    ap.add_argument("--match-steps", type=int, default=2)
    ap.add_argument("--pose-frames", type=int, default=2048)
    ap.add_argument("--no-extra", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
