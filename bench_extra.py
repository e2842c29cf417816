"""Secondary legs of bench.py (same JSON line, key "extra"): Hamming matching pairs/s (BASELINE config 3) and LocalBA
LM iterations/s (config 4), each with its own roofline figure and CPU-oracle baseline. Rank 0, single GPU."""
import os
import time

import numpy as np

from cubemapslam_b200 import synth


def _match_leg(torch, dev, args, local):
    from cubemapslam_b200.matcher import ORBMatcher
    P, n = args.match_pairs, 2000
    nb = 16
    base = [synth.descriptor_pair(p, n=n) for p in range(nb)]
    A = torch.from_numpy(np.stack([b[0] for b in base])).to(dev); aA = torch.from_numpy(np.stack([b[1] for b in base])).to(dev)
    B = torch.from_numpy(np.stack([b[2] for b in base])).to(dev); aB = torch.from_numpy(np.stack([b[3] for b in base])).to(dev)
    reps = (P + nb - 1) // nb
    # P pairs resident in HBM: tiles of the 16 generated pairs with the B side rolled so every pair is a distinct problem
    dA = A.repeat(reps, 1, 1)[:P].contiguous(); daA = aA.repeat(reps, 1)[:P].contiguous()
    dB = torch.cat([torch.roll(B, shifts=r, dims=1) for r in range(reps)])[:P].contiguous()
    daB = torch.cat([torch.roll(aB, shifts=r, dims=1) for r in range(reps)])[:P].contiguous()
    m = ORBMatcher(0.6, True, max_pairs=P, max_features=2048, device=local)
    match = torch.empty((P, n), dtype=torch.int32, device=dev); nm = torch.empty((P,), dtype=torch.int32, device=dev)
    stream = torch.cuda.ExternalStream(m.stream, device=dev)
    out = {}

    def bf():
        m.match_bruteforce_dev(dA.data_ptr(), daA.data_ptr(), n, dB.data_ptr(), daB.data_ptr(), n, P, match.data_ptr(), nm.data_ptr())

    rng = np.random.default_rng(1)
    nodeA = torch.from_numpy(rng.integers(0, 100, (P, n)).astype(np.int32)).to(dev)
    nodeB = torch.from_numpy(rng.integers(0, 100, (P, n)).astype(np.int32)).to(dev)
    valid = torch.ones((P, n), dtype=torch.uint8, device=dev)
    bowm = ORBMatcher(0.7, True, max_pairs=P, max_features=2048, device=local)
    bstream = torch.cuda.ExternalStream(bowm.stream, device=dev)

    def bow():
        bowm.search_by_bow_dev(dA.data_ptr(), daA.data_ptr(), valid.data_ptr(), nodeA.data_ptr(), n, dB.data_ptr(), daB.data_ptr(), nodeB.data_ptr(), n, P,
                               match.data_ptr(), nm.data_ptr())
    for name, fn, st, mm in (("bruteforce", bf, stream, m), ("search_by_bow", bow, bstream, bowm)):
        for _ in range(2):
            fn()
        mm.sync()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(args.match_steps):
            fn()
        e1.record(st)
        mm.sync()
        ms = e0.elapsed_time(e1) / args.match_steps
        out[name] = {"pairs_per_s": round(P / (ms / 1e3), 1), "ms_per_%d_pairs" % P: round(ms, 3)}
    bfr = out["bruteforce"]["pairs_per_s"]
    out["bruteforce"]["popc32_per_s"] = round(bfr * 3.2e7, 3)
    out["bruteforce"]["hbm_GBs_algorithmic"] = round(bfr * 192e3 / 1e9, 2)
    out["bruteforce"]["mean_matches"] = float(nm.float().mean().item()) if False else None
    # CPU oracle: bounded sample, all cores
    import oracle as orc
    import bench as _b
    cores = min(os.cpu_count() or 1, 2 * _b.host_cores())
    s = nb
    hA = np.stack([b[0] for b in base[:s]]); haA = np.stack([b[1] for b in base[:s]]); hB = np.stack([b[2] for b in base[:s]]); haB = np.stack([b[3] for b in base[:s]])
    t0 = time.perf_counter()
    orc.match_bruteforce_batch(hA, haA, hB, haB, 0.6, 50, True, nthreads=min(cores, s))
    dt = time.perf_counter() - t0
    out["cpu_baseline"] = {"bruteforce_pairs_per_s": round(s / dt, 2), "cores": min(cores, s), "kind": "port", "sample": "%d pairs of 2000x2000" % s}
    out["config"] = "configs[2]: 2000x2000 descriptors x %d pairs resident in HBM" % P
    m.close(); bowm.close()
    return out


def _ba_leg(args, local):
    from cubemapslam_b200.optimizer import Optimizer
    import oracle as orc
    p = synth.ba_problem()   # config 4: 50 KF x 20k points, ~170k edges
    o = Optimizer(device=local)
    o.LocalBundleAdjustment(p["Tcw"], p["kf_fixed"], p["pts"], p["eMP"], p["eKF"], p["kpxy"], p["inv_sigma2"], 650, 650, its1=1, its2=0)   # warm-up
    t0 = time.perf_counter()
    g = o.LocalBundleAdjustment(p["Tcw"], p["kf_fixed"], p["pts"], p["eMP"], p["eKF"], p["kpxy"], p["inv_sigma2"], 650, 650)
    dt = time.perf_counter() - t0
    out = {"config": "configs[3]: LocalBA 50 KF x 20000 MapPoints, %d edges, optimize(5)+optimize(10) schedule" % len(p["eMP"]),
           "lm_iters": int(g["iters"]), "lm_trials": int(g["trials"]), "seconds_e2e": round(dt, 4), "lm_iters_per_s": round(g["iters"] / dt, 2),
           "lm_trials_per_s": round(g["trials"] / dt, 2), "note": "host call: upload, LM loop with one D2H per trial, write-back included"}
    t0 = time.perf_counter()
    r = orc.local_ba(p["Tcw"], p["kf_fixed"], p["pts"], p["eMP"], p["eKF"], p["kpxy"], p["inv_sigma2"], 650, 650)
    dtc = time.perf_counter() - t0
    out["cpu_baseline"] = {"lm_iters_per_s": round(r["iters"] / dtc, 3), "lm_iters": int(r["iters"]), "seconds": round(dtc, 3), "cores": 1, "kind": "port",
                           "sample": "the same problem, whole schedule (the reference runs LocalBA on one thread)"}
    pe = np.abs(g["pose64"] - r["pose64"]).max() / np.abs(r["pose64"]).max()
    out["parity_rel_pose"] = float(pe)
    o.close()
    return out


def _pose_leg(args, local):
    """Batched Optimizer::PoseOptimization (config 5's pose-only BA stage): F independent frames x ~600 correspondences, host call."""
    from cubemapslam_b200.optimizer import Optimizer
    import oracle as orc
    base = [synth.pose_problem(n=600, faceW=650, seed=100 + i, outlier_frac=0.1) for i in range(8)]
    F = args.pose_frames
    probs = [base[i % len(base)] for i in range(F)]
    off = np.cumsum([0] + [len(q["Xw"]) for q in probs]).astype(np.int32)
    T = np.stack([q["Tcw"] for q in probs]); Xw = np.concatenate([q["Xw"] for q in probs]); kp = np.concatenate([q["kpxy"] for q in probs])
    w = np.concatenate([q["inv_sigma2"] for q in probs])
    o = Optimizer(device=local)
    o.PoseOptimization(T[:8], Xw[:off[8]], kp[:off[8]], w[:off[8]], 650, 650, offset=off[:9])
    t0 = time.perf_counter()
    g = o.PoseOptimization(T, Xw, kp, w, 650, 650, offset=off)
    dt = time.perf_counter() - t0
    t0 = time.perf_counter()
    r = [orc.pose_opt(q["Tcw"], q["Xw"], q["kpxy"], q["inv_sigma2"], 650, 650) for q in base]
    dtc = time.perf_counter() - t0
    ok = all(int(g["inliers"][i]) == r[i]["inliers"] for i in range(len(base)))
    o.close()
    return {"config": "configs[4] stage: PoseOptimization, %d frames x ~%d correspondences per host call" % (F, len(base[0]["Xw"])), "frames_per_s": round(F / dt, 1),
            "seconds": round(dt, 4), "cpu_baseline": {"frames_per_s": round(len(base) / dtc, 1), "cores": 1, "kind": "port", "sample": "%d frames" % len(base)},
            "inliers_equal_to_oracle": bool(ok)}


def run(args, local):
    import torch
    dev = torch.device("cuda", local)
    out = {}
    try:
        out["match"] = _match_leg(torch, dev, args, local)
    except Exception as e:  # the headline metric must still be reported
        out["match"] = {"error": repr(e)}
    try:
        out["local_ba"] = _ba_leg(args, local)
    except Exception as e:
        out["local_ba"] = {"error": repr(e)}
    try:
        out["pose_optimization"] = _pose_leg(args, local)
    except Exception as e:
        out["pose_optimization"] = {"error": repr(e)}
    return out


def run_sharded_ba(local, rank, world, dist, torch):
    """LocalBA config 4 with landmarks sharded over all ranks (NCCL all-reduce of the reduced camera system); all ranks call."""
    from cubemapslam_b200.optimizer import Optimizer
    p = synth.ba_problem()
    a = (p["Tcw"], p["kf_fixed"], p["pts"], p["eMP"], p["eKF"], p["kpxy"], p["inv_sigma2"], 650, 650)
    o = Optimizer(device=local)
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.from_numpy(Optimizer.nccl_unique_id()))
    dist.broadcast(idt, 0)
    o.init_nccl(idt.cpu().numpy(), rank, world)
    o.LocalBundleAdjustment(*a, its1=1, its2=0)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    g = o.LocalBundleAdjustment(*a)
    torch.cuda.synchronize(); dist.barrier()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    o.close()
    return {"config": "configs[3] landmark-sharded over %d GPUs, NCCL all-reduce of [S|g] per LM trial" % world, "lm_iters": int(g["iters"]),
            "seconds": round(float(tt.item()), 4), "lm_iters_per_s": round(g["iters"] / float(tt.item()), 2)}
