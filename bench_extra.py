"""Secondary legs of bench.py (same JSON line, key "extra"). Every leg carries its own roofline statement, CPU baseline (same workload, bounded
sample, core count stated) and an end-to-end figure through the host entry points:
  match              BASELINE configs[2]: 2000 x 2000 all-pairs Hamming + SearchByBoW, 4096 pairs
  local_ba           configs[3]: LocalBundleAdjustment 50 KF x 20k MapPoints (the "LocalBA iters/sec" half of the metric) + the dense k=50 variant
  pose_optimization  PoseOptimization batches
  tracking           configs[4]: warp -> extract -> frame index -> SearchByProjection(last frame) -> pose-only BA, device-resident per frame batch
  local_mapping      SURVEY §8(f) rank 4: batched ComputeDistinctiveDescriptors, Fuse search, SearchForTriangulation (host calls)
Multi-GPU (run_multi): landmark-sharded LocalBA with parity against the 1-GPU run, pair-sharded matching, frame-sharded tracking."""
import json
import os
import time

import numpy as np

from cubemapslam_b200 import config, synth

ROOT = os.path.dirname(os.path.abspath(__file__))


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


def _cores():
    import bench as _b
    q = _b.host_cores()
    return q, min(os.cpu_count() or 1, 2 * q)


def _match_leg(torch, dev, args, local, rank=0, world=1):
    from cubemapslam_b200.matcher import ORBMatcher
    P, n = args.match_pairs, 2000
    nb = 16
    base = [synth.descriptor_pair(p + 16 * rank, n=n) for p in range(nb)]
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

    # SearchByBoW: both sides share the node of their true partner (FeatureVectors of two views of the same scene), ~100 nodes at levelsup 4
    rng = np.random.default_rng(1 + rank)
    nodeA_h = rng.integers(0, 100, (nb, n)).astype(np.int32)
    nodeB_h = np.stack([nodeA_h[i][base[i][4]] for i in range(nb)])
    flip = rng.random((nb, n)) < 0.1
    nodeB_h[flip] = rng.integers(0, 100, int(flip.sum()))
    nodeA = torch.from_numpy(nodeA_h).to(dev).repeat(reps, 1)[:P].contiguous()
    nodeB = torch.cat([torch.roll(torch.from_numpy(nodeB_h).to(dev), shifts=r, dims=1) for r in range(reps)])[:P].contiguous()
    valid = torch.ones((P, n), dtype=torch.uint8, device=dev)
    bowm = ORBMatcher(0.7, True, max_pairs=P, max_features=2048, device=local)
    bstream = torch.cuda.ExternalStream(bowm.stream, device=dev)

    def bow():
        bowm.search_by_bow_dev(dA.data_ptr(), daA.data_ptr(), valid.data_ptr(), nodeA.data_ptr(), n, dB.data_ptr(), daB.data_ptr(), nodeB.data_ptr(), n, P,
                               match.data_ptr(), nm.data_ptr())
    for name, fn, st, mm in (("bruteforce", bf, stream, m), ("search_by_bow", bow, bstream, bowm)):
        for _ in range(3):
            fn()
        mm.sync()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(args.match_steps):
            fn()
        e1.record(st)
        mm.sync()
        ms = e0.elapsed_time(e1) / args.match_steps
        out[name] = {"pairs_per_s": round(P / (ms / 1e3), 1), "ms_per_%d_pairs" % P: round(ms, 3), "mean_matches": round(float(nm.float().mean().item()), 1)}
    hbm, how = _peaks()
    popc_peak = m.ubench_popc()
    bfr = out["bruteforce"]["pairs_per_s"]
    out["bruteforce"]["roofline"] = {"bound": "popc issue (not HBM: 192 KB moved per 3.2e7 popc32)", "achieved_popc32_per_s": round(bfr * 3.2e7, 3),
                                     "measured_peak_popc32_per_s": round(popc_peak, 3), "frac": round(bfr * 3.2e7 / popc_peak, 4),
                                     "peak_source": "cslam_ubench_popc on this GPU (XOR+POPC+ADD chains on registers)", "hbm_GBs_algorithmic": round(bfr * 192e3 / 1e9, 2),
                                     "frac_of_hbm": round(bfr * 192e3 / 1e9 / hbm, 5)}
    bwr = out["search_by_bow"]["pairs_per_s"]
    out["search_by_bow"]["roofline"] = {"bound": "latency (in-CTA sort of both FeatureVectors + in-node sequential order); 192 KB per pair", "hbm_GBs_algorithmic": round(bwr * 192e3 / 1e9, 2),
                                        "frac_of_hbm": round(bwr * 192e3 / 1e9 / hbm, 4)}
    if rank == 0:
        # end to end through the host call (host descriptors in, host matches out)
        s = min(256, P)
        hA = dA[:s].cpu().numpy(); haA = daA[:s].cpu().numpy(); hB = dB[:s].cpu().numpy(); haB = daB[:s].cpu().numpy()
        m.match_bruteforce(hA[:8], haA[:8], hB[:8], haB[:8])
        t0 = time.perf_counter()
        m.match_bruteforce(hA, haA, hB, haB)
        dt = time.perf_counter() - t0
        out["bruteforce"]["e2e"] = {"pairs_per_s": round(s / dt, 1), "h2d_bytes": int(s * n * 72), "d2h_bytes": int(s * (n * 12 + 4)), "call": "cslam_match_bruteforce (host buffers)"}
        # CPU oracle: bounded sample, all cores
        import oracle as orc
        quota, cores = _cores()
        ss = nb
        cA = np.stack([b[0] for b in base[:ss]]); caA = np.stack([b[1] for b in base[:ss]]); cB = np.stack([b[2] for b in base[:ss]]); caB = np.stack([b[3] for b in base[:ss]])
        t0 = time.perf_counter()
        orc.match_bruteforce_batch(cA, caA, cB, caB, 0.6, 50, True, nthreads=min(cores, ss))
        dt = time.perf_counter() - t0
        t0 = time.perf_counter()
        orc.match_bruteforce(cA[0], caA[0], cB[0], caB[0], 0.6, 50, True)
        d1 = time.perf_counter() - t0
        out["cpu_baseline"] = {"bruteforce_pairs_per_s": round(ss / dt, 2), "cores": quota, "threads": min(cores, ss), "single_thread_pairs_per_s": round(1.0 / d1, 2), "kind": "port",
                               "sample": "%d pairs of 2000x2000 (the reference has no all-pairs function: DESIGN.md §2)" % ss}
    out["config"] = "configs[2]: 2000x2000 descriptors x %d pairs resident in HBM (per GPU)" % P
    m.close(); bowm.close()
    return out


def _ba_roofline(p, g, seconds, hbm):
    E = len(p["eMP"]); nP = int((1 - np.asarray(p["kf_fixed"])).sum()); nMP = len(p["pts"])
    k = np.bincount(p["eMP"], minlength=nMP)
    tuples = float((k * (k + 1) // 2).sum())
    # algorithmic bytes per LM trial: stream the edges (64 B) + pose / point state; flops: linearise ~450/edge, sparse Schur 27 + 108k + 216 k(k+1)/2 per landmark
    bytes_trial = 64.0 * E + 56.0 * nP + 24.0 * nMP
    flop_trial = 450.0 * E + float((27 + 108 * k + 216 * (k * (k + 1) // 2)).sum()) + (6 * nP) ** 3 / 3.0
    per_trial = seconds / max(g["trials"], 1)
    return {"bound": "latency / fp64 (a trial is ~12 dependent launches; not HBM-bound at this size)", "algorithmic_bytes_per_trial": int(bytes_trial), "algorithmic_fp64_flop_per_trial": int(flop_trial),
            "achieved_GBs": round(bytes_trial / per_trial / 1e9, 2), "frac_of_hbm": round(bytes_trial / per_trial / 1e9 / hbm, 5), "achieved_fp64_GFLOPs": round(flop_trial / per_trial / 1e9, 1),
            "co_observation_tuples": int(tuples), "seconds_per_trial_e2e": round(per_trial, 6)}


def _ba_leg(args, local, dense=False):
    from cubemapslam_b200.optimizer import Optimizer
    import oracle as orc
    hbm, _ = _peaks()
    p = synth.ba_problem(nKF=50, nMP=4000, kmin=50, kmax=50, radius=9.0) if dense else synth.ba_problem()   # config 4: 50 KF x 20k points, ~180k edges
    o = Optimizer(device=local)
    a = (p["Tcw"], p["kf_fixed"], p["pts"], p["eMP"], p["eKF"], p["kpxy"], p["inv_sigma2"], 650, 650)
    o.LocalBundleAdjustment(*a, its1=1, its2=0)   # warm-up
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        g = o.LocalBundleAdjustment(*a)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    dt = best
    out = {"config": ("dense variant: 50 KF x 4000 MapPoints each seen by every KF, %d edges" if dense else "configs[3]: LocalBA 50 KF x 20000 MapPoints, %d edges, optimize(5)+optimize(10) schedule") % len(p["eMP"]),
           "lm_iters": int(g["iters"]), "lm_trials": int(g["trials"]), "seconds_e2e": round(dt, 4), "lm_iters_per_s": round(g["iters"] / dt, 2),
           "lm_trials_per_s": round(g["trials"] / dt, 2), "launches_per_call": None,
           "e2e": {"lm_iters_per_s": round(g["iters"] / dt, 2), "note": "the figure above IS end to end: host float32 poses / points / edges in, upload, co-observation lists, LM loop with one 32-byte D2H per trial, float32 write-back"},
           "roofline": _ba_roofline(p, g, dt, hbm)}
    if not dense or len(p["eMP"]) < 400000:
        t0 = time.perf_counter()
        r = orc.local_ba(*a)
        dtc = time.perf_counter() - t0
        out["cpu_baseline"] = {"lm_iters_per_s": round(r["iters"] / dtc, 3), "lm_iters": int(r["iters"]), "seconds": round(dtc, 3), "cores": 1, "kind": "port",
                               "sample": "the same problem, whole schedule, one thread (the reference runs LocalBA on the LocalMapping thread; g2o's OpenMP is off)"}
        out["parity_rel_pose"] = float(np.abs(g["pose64"] - r["pose64"]).max() / np.abs(r["pose64"]).max())
        out["lm_log_equal"] = bool(np.array_equal(g["log"][:, 2:], r["log"][:, 2:])); out["outliers_equal"] = bool(np.array_equal(g["outlier"], r["outlier"]))
    o.close()
    return out


def _pose_leg(args, local):
    """Batched Optimizer::PoseOptimization: F independent frames x ~600 correspondences, host call."""
    from cubemapslam_b200.optimizer import Optimizer
    import oracle as orc
    F = args.pose_frames
    nd = 64
    base = [synth.pose_problem(n=600, faceW=650, seed=100 + i, outlier_frac=0.1) for i in range(nd)]
    probs = [base[i % nd] for i in range(F)]
    off = np.cumsum([0] + [len(q["Xw"]) for q in probs]).astype(np.int32)
    T = np.stack([q["Tcw"] for q in probs]); Xw = np.concatenate([q["Xw"] for q in probs]); kp = np.concatenate([q["kpxy"] for q in probs])
    w = np.concatenate([q["inv_sigma2"] for q in probs])
    o = Optimizer(device=local)
    o.PoseOptimization(T[:8], Xw[:off[8]], kp[:off[8]], w[:off[8]], 650, 650, offset=off[:9])
    t0 = time.perf_counter()
    g = o.PoseOptimization(T, Xw, kp, w, 650, 650, offset=off)
    dt = time.perf_counter() - t0
    t0 = time.perf_counter()
    r = [orc.pose_opt(q["Tcw"], q["Xw"], q["kpxy"], q["inv_sigma2"], 650, 650) for q in base[:16]]
    dtc = time.perf_counter() - t0
    ok = all(int(g["inliers"][i]) == r[i]["inliers"] for i in range(16))
    hbm, _ = _peaks()
    byt = float(off[-1]) * 24 * 4 * 10          # ~40 passes over the correspondences (4 rounds x <= 10 iterations), 24 B each
    o.close()
    return {"config": "PoseOptimization, %d frames (%d distinct problems) x ~%d correspondences per host call" % (F, nd, len(base[0]["Xw"])), "frames_per_s": round(F / dt, 1),
            "seconds": round(dt, 4), "e2e": {"frames_per_s": round(F / dt, 1), "note": "host call: upload, kernel, download included"},
            "roofline": {"bound": "latency (one CTA per frame runs the whole 4 x optimize(10) schedule)", "achieved_GBs_upper": round(byt / dt / 1e9, 2), "frac_of_hbm_upper": round(byt / dt / 1e9 / hbm, 5)},
            "cpu_baseline": {"frames_per_s": round(16 / dtc, 1), "cores": 1, "kind": "port", "sample": "16 frames, one thread (the reference runs it on the Tracking thread)"},
            "inliers_equal_to_oracle": bool(ok)}


def _tracking_leg(torch, dev, args, local, rank=0, world=1):
    """BASELINE configs[4]: the Tracking hot loop per frame - warp + extract (front end), rays + grid (frame index), SearchByProjection against the
    previous frame's MapPoints, pose-only BA - device-resident, frames sharded over GPUs as independent tasks (real tracking is
    sequential; the bench runs independent synthetic per-frame tasks that each carry their own prior pose and 3-D points)."""
    from cubemapslam_b200.frontend import FrontEnd
    from cubemapslam_b200.optimizer import Optimizer
    from cubemapslam_b200.tracker import Tracker
    import bench as _b
    cfg = config.front_1024(); mask = _b.load_mask()
    F = args.track_frames; B = min(args.batch, F)
    fe = FrontEnd(cfg, mask, max_batch=B, device=local)
    cap = fe.kp_cap
    tr = Tracker(max_frames=1, max_features=64, device=local)
    op = Optimizer(device=local)
    G = 16                                            # frame f+16 is frame f shifted by 7 px (bench.make_frames_device): the "same scene a moment later"
    NF = F + G
    fish, _ = _b.make_frames_device(torch, cfg, NF, dev)
    fsz = fish.shape[1] * fish.shape[2]
    u8 = torch.uint8
    kps = torch.zeros((NF, cap, 28), dtype=u8, device=dev); desc = torch.zeros((NF, cap, 32), dtype=u8, device=dev); nout = torch.zeros((NF,), dtype=torch.int32, device=dev)
    rays = torch.zeros((NF, cap, 3), dtype=torch.float32, device=dev); cs = torch.zeros((NF, 12501), dtype=torch.int16, device=dev); ci = torch.zeros((NF, cap), dtype=torch.int16, device=dev)
    match = torch.zeros((F, cap), dtype=torch.int32, device=dev); nmatch = torch.zeros((F,), dtype=torch.int32, device=dev)
    cXw = torch.zeros((F, cap, 3), dtype=torch.float32, device=dev); ckp = torch.zeros((F, cap, 2), dtype=torch.float32, device=dev); cw = torch.zeros((F, cap), dtype=torch.float32, device=dev)
    ccount = torch.zeros((F,), dtype=torch.int32, device=dev); outl = torch.zeros((F, cap), dtype=u8, device=dev); inl = torch.zeros((F,), dtype=torch.int32, device=dev)
    fs = torch.cuda.ExternalStream(fe.stream, device=dev); ts = torch.cuda.ExternalStream(tr.stream, device=dev); os_ = torch.cuda.ExternalStream(op.stream, device=dev)
    sc = np.ones(8, np.float32)
    for l in range(1, 8):
        sc[l] = sc[l - 1] * np.float32(1.2)
    invs2 = torch.from_numpy((np.float32(1.0) / (sc * sc)).astype(np.float32)).to(dev)
    cth = float(np.float32(np.cos(np.float64(np.float32(cfg["Camera.fov"]) / np.float32(2) * (np.float32(3.1415926535897932384626) / np.float32(180))))))
    # task f: "last" frame = extracted frame f (its key points get MapPoints at depth 4 along their rays, world = last camera), "current" = frame f+16, prior pose = identity + 1 cm
    def extract_all():
        for c in range(0, NF, B):
            b = min(B, NF - c)
            fe.run_dev(fish.data_ptr() + c * fsz, b, kps.data_ptr() + c * cap * 28, desc.data_ptr() + c * cap * 32, nout.data_ptr() + c * 4)
    extract_all(); fe.sync()
    ts.wait_stream(fs)
    tr.frame_index_dev(kps.data_ptr(), nout.data_ptr(), NF, cap, 650, 650, rays.data_ptr(), cs.data_ptr(), ci.data_ptr()); tr.sync()
    Xw = (rays * 4.0).contiguous()
    has = (torch.arange(cap, device=dev)[None, :] < nout[:, None]).to(u8).contiguous(); obs = torch.ones_like(has); taken = torch.zeros_like(has)
    T0 = torch.eye(4, dtype=torch.float32, device=dev).repeat(F, 1, 1).contiguous(); T0[:, 0, 3] = 0.01
    Tcw = T0.clone()

    def one_pass():
        fs.wait_stream(os_)
        extract_all()
        ts.wait_stream(fs)
        tr.frame_index_dev(kps.data_ptr(), nout.data_ptr(), NF, cap, 650, 650, rays.data_ptr(), cs.data_ptr(), ci.data_ptr())
        # current = frames G..G+F-1, last = frames 0..F-1 (same device arrays, shifted by G frames)
        tr.search_by_projection_last_dev(F, kps.data_ptr() + G * cap * 28, desc.data_ptr() + G * cap * 32, nout.data_ptr() + G * 4, cap, cs.data_ptr() + G * 12501 * 2, ci.data_ptr() + G * cap * 2,
                                         taken.data_ptr() + G * cap, Tcw.data_ptr(), kps.data_ptr(), nout.data_ptr(), cap, has.data_ptr(), Xw.data_ptr(), desc.data_ptr(), obs.data_ptr(),
                                         650, 650, cth, 15.0, 1, match.data_ptr(), nmatch.data_ptr())
        tr.gather_pose_inputs_dev(F, match.data_ptr(), kps.data_ptr() + G * cap * 28, nout.data_ptr() + G * 4, cap, rays.data_ptr() + G * cap * 12, cth, Xw.data_ptr(), cap, invs2.data_ptr(),
                                  cXw.data_ptr(), ckp.data_ptr(), cw.data_ptr(), ccount.data_ptr())
        os_.wait_stream(ts)
        op.pose_optimization_dev(F, cap, ccount.data_ptr(), Tcw.data_ptr(), cXw.data_ptr(), ckp.data_ptr(), cw.data_ptr(), 650, 650, outl.data_ptr(), inl.data_ptr())
    one_pass(); op.sync(); tr.sync()
    Tcw.copy_(T0); torch.cuda.synchronize()
    l0 = fe.launches + tr.launches + op.launches
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    steps = 2
    e0.record(fs)
    for _ in range(steps):
        one_pass()
    e1.record(os_)
    op.sync(); tr.sync(); fe.sync(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out = {"config": "configs[4]: warp + extract + frame index + SearchByProjection(last frame, th 15) + PoseOptimization, %d synthetic 1280x1024 frame tasks per GPU, device-resident" % F,
           "frames_per_s_per_gpu": round(F / (ms / 1e3), 1), "ms_per_pass": round(ms, 3), "gpu_launches_per_pass": int((fe.launches + tr.launches + op.launches - l0) / steps),
           "mean_keypoints": round(float(nout.float().mean().item()), 1), "mean_projection_matches": round(float(nmatch.float().mean().item()), 1),
           "mean_pose_inliers": round(float(inl.float().mean().item()), 1),
           "roofline": {"bound": "front end (see the headline roofline); matching + pose-only BA add launches, not bytes", "frames_per_s_frontend_only": None}}
    fe.close(); tr.close(); op.close()
    return out


def _mapping_leg(args, local):
    """LocalMapping feature operations through the host calls (the only form they have): batched ComputeDistinctiveDescriptors, Fuse search,
    SearchForTriangulation over 20 key-frame pairs; oracle on the same inputs as the CPU baseline (one thread: the mapping thread)."""
    from cubemapslam_b200.mapper import Mapper
    import oracle as orc
    m = Mapper(device=local)
    rng = np.random.default_rng(5)
    out = {}
    # ---- distinctive descriptors: 20 000 MapPoints x 3..15 observations
    P = 20000
    nobs = rng.integers(3, 16, P); off = np.concatenate([[0], np.cumsum(nobs)]).astype(np.int32)
    base = rng.integers(0, 256, (P, 32), dtype=np.uint8)
    desc = np.repeat(base, nobs, axis=0) ^ np.packbits(rng.random((int(off[-1]), 256)) < 0.08, axis=1, bitorder="little")
    m.ComputeDistinctiveDescriptors(desc[:off[100]], off[:101])
    t0 = time.perf_counter(); g = m.ComputeDistinctiveDescriptors(desc, off); dt = time.perf_counter() - t0
    sub = 2000
    t0 = time.perf_counter(); r = orc.distinctive_descriptors(desc[:off[sub]], off[:sub + 1]); dtc = time.perf_counter() - t0
    out["distinctive_descriptors"] = {"config": "%d MapPoints, %d observation descriptors, one host call" % (P, int(off[-1])), "points_per_s": round(P / dt, 1), "seconds": round(dt, 5),
                                      "e2e": {"points_per_s": round(P / dt, 1), "h2d_bytes": int(off[-1]) * 32 + 4 * (P + 1), "d2h_bytes": 4 * P},
                                      "roofline": {"bound": "latency / PCIe of one small call (9 N^2 Hamming per point; %.1f MB in)" % (int(off[-1]) * 32 / 1e6)},
                                      "cpu_baseline": {"points_per_s": round(sub / dtc, 1), "cores": 1, "kind": "port", "sample": "%d points, one thread" % sub},
                                      "equal_to_oracle": bool(np.array_equal(g[:sub], r))}
    # ---- Fuse search + SearchForTriangulation on synthetic key-frame pairs
    ss = [synth.mapping_pair(40 + i, n=2000, faceW=650) for i in range(4)]
    s = ss[0]; n = len(s["Xw"])
    valid = np.ones(n, np.uint8); level = np.clip(s["kLast"]["octave"], 0, 7).astype(np.int32)
    m.FuseSearch(s["kCur"], s["dCur"], s["TcwCur"], valid, s["Xw"], level, s["dLast"], 3.0, s["scale"], s["inv_level_sigma2"], 650, 650)
    reps = 20
    t0 = time.perf_counter()
    for _ in range(reps):
        gi, gd = m.FuseSearch(s["kCur"], s["dCur"], s["TcwCur"], valid, s["Xw"], level, s["dLast"], 3.0, s["scale"], s["inv_level_sigma2"], 650, 650)
    dt = (time.perf_counter() - t0) / reps
    grid = orc.FrameGrid(s["kCur"], 650, 650)
    t0 = time.perf_counter(); bi, bd = grid.fuse_search(s["dCur"], s["TcwCur"], s["scale"], s["inv_level_sigma2"], valid, s["Xw"], level, s["dLast"], 3.0); dtc = time.perf_counter() - t0
    out["fuse_search"] = {"config": "Fuse(pKF, %d MapPoints, th 3) against a key frame of %d key points, one host call" % (n, len(s["kCur"])), "calls_per_s": round(1 / dt, 1),
                          "map_points_per_s": round(n / dt, 1), "e2e": {"calls_per_s": round(1 / dt, 1), "note": "host arrays in / out, frame index built inside the call"},
                          "roofline": {"bound": "latency (3 launches + 9 small copies per call)"},
                          "cpu_baseline": {"calls_per_s": round(1 / dtc, 1), "cores": 1, "kind": "port", "sample": "the same call, one thread (grid prebuilt)"},
                          "equal_to_oracle": bool(np.array_equal(gi, bi) and np.array_equal(gd, bd))}
    P2 = 20
    s1 = max(len(q["kCur"]) for q in ss); s2 = max(len(q["kLast"]) for q in ss)
    KP = ss[0]["kCur"].dtype
    k1 = np.zeros((P2, s1), KP); d1 = np.zeros((P2, s1, 32), np.uint8); r1 = np.zeros((P2, s1, 3), np.float32); h1 = np.zeros((P2, s1), np.uint8); nd1 = np.zeros((P2, s1), np.int32)
    k2 = np.zeros((P2, s2), KP); d2 = np.zeros((P2, s2, 32), np.uint8); r2 = np.zeros((P2, s2, 3), np.float32); h2 = np.zeros((P2, s2), np.uint8); nd2 = np.zeros((P2, s2), np.int32)
    n1 = np.zeros(P2, np.int32); n2 = np.zeros(P2, np.int32); Ow = np.zeros((P2, 3), np.float32); T2 = np.zeros((P2, 16), np.float32); E = np.zeros((P2, 9), np.float32)
    rays = [(orc.key_point_rays(q["kCur"], 650, 650)[0], orc.key_point_rays(q["kLast"], 650, 650)[0]) for q in ss]
    for p in range(P2):
        q = ss[p % len(ss)]; ra, rb = rays[p % len(ss)]
        a, b = len(q["kCur"]), len(q["kLast"]); n1[p] = a; n2[p] = b
        k1[p, :a] = q["kCur"]; d1[p, :a] = q["dCur"]; r1[p, :a] = ra; h1[p, :a] = q["hasMPCur"]; nd1[p, :a] = q["nodeCur"]
        k2[p, :b] = q["kLast"]; d2[p, :b] = q["dLast"]; r2[p, :b] = rb; h2[p, :b] = q["hasMPObs"]; nd2[p, :b] = q["nodeObs"]
        T1 = q["TcwCur"].astype(np.float64)
        Ow[p] = (-T1[:3, :3].T @ T1[:3, 3]).astype(np.float32); T2[p] = q["TcwLast"].reshape(16); E[p] = q["E12"].reshape(9)
    m.SearchForTriangulation(k1[:2], d1[:2], r1[:2], h1[:2], nd1[:2], n1[:2], k2[:2], d2[:2], r2[:2], h2[:2], nd2[:2], n2[:2], Ow[:2], T2[:2], E[:2], s["scale"], s["level_sigma2"], 650, 650)
    t0 = time.perf_counter()
    nm, mm = m.SearchForTriangulation(k1, d1, r1, h1, nd1, n1, k2, d2, r2, h2, nd2, n2, Ow, T2, E, s["scale"], s["level_sigma2"], 650, 650)
    dt = time.perf_counter() - t0
    t0 = time.perf_counter()
    w = [orc.search_for_triangulation(ss[p]["kCur"], ss[p]["dCur"], rays[p][0], ss[p]["hasMPCur"], ss[p]["nodeCur"], ss[p]["kLast"], ss[p]["dLast"], rays[p][1], ss[p]["hasMPObs"],
                                      ss[p]["nodeObs"], Ow[p], T2[p], E[p], s["scale"], s["level_sigma2"], 650, 650, False) for p in range(len(ss))]
    dtc = time.perf_counter() - t0
    out["search_for_triangulation"] = {"config": "%d key-frame pairs x ~%d features per key frame, one host call (LocalMapping::CreateNewMapPoints matches ~20 neighbours)" % (P2, s1),
                                       "pairs_per_s": round(P2 / dt, 1), "e2e": {"pairs_per_s": round(P2 / dt, 1), "note": "host arrays in / out"},
                                       "mean_matches": round(float(nm.mean()), 1), "roofline": {"bound": "latency (one CTA per pair: in-CTA sort + node scans)"},
                                       "cpu_baseline": {"pairs_per_s": round(len(ss) / dtc, 1), "cores": 1, "kind": "port", "sample": "%d pairs, one thread" % len(ss)},
                                       "equal_to_oracle": bool(all(nm[p] == w[p][0] and np.array_equal(mm[p, :n1[p]], w[p][1]) for p in range(len(ss))))}
    m.close()
    return out


def _latency_leg(args, local):
    """What ONE frame costs through the reference-facing calls (the drop-in ORBextractor::operator() / System::CvtFisheyeToCubeMap... make exactly these):
    host image in, host key points + descriptors out, batch 1. The throughput legs batch frames; a SLAM front end that feeds one frame at a time sees this."""
    from cubemapslam_b200.frontend import FrontEnd
    cfg = config.front_1024(); mask = config.load_mask("gray_cubemap_front_mask_650")
    fe = FrontEnd(cfg, mask, max_batch=1, device=local)
    frames = [synth.fisheye_frame(cfg, i) for i in range(4)]
    canv = [fe.warp(f) for f in frames]
    for f in frames:
        fe.run(f)
    def med(fn, n=40):
        t = []
        for i in range(n):
            t0 = time.perf_counter(); fn(i); t.append(time.perf_counter() - t0)
        return float(np.median(t)) * 1e3
    ms_run = med(lambda i: fe.run(frames[i % 4]))
    ms_ext = med(lambda i: fe.extract(canv[i % 4]))
    ms_warp = med(lambda i: fe.warp(frames[i % 4]))
    fe.close()
    return {"config": "one 1280x1024 frame per call, 650-px faces, 3000 features, host buffers in / out (median of 40 calls)",
            "warp_plus_extract_ms": round(ms_run, 3), "extract_only_ms": round(ms_ext, 3), "warp_only_ms": round(ms_warp, 3),
            "frames_per_s_single_stream": round(1e3 / ms_run, 1),
            "note": "cslam_frontend_run / cslam_orb_extract / cslam_warp with batch 1: launch latency of ~37 kernels + 1.3 MB H2D + 0.2 MB D2H per frame"}


def run(args, local):
    import torch
    dev = torch.device("cuda", local)
    out = {}
    for name, fn in (("match", lambda: _match_leg(torch, dev, args, local)), ("local_ba", lambda: _ba_leg(args, local)), ("local_ba_dense", lambda: _ba_leg(args, local, dense=True)),
                     ("pose_optimization", lambda: _pose_leg(args, local)), ("tracking", lambda: _tracking_leg(torch, dev, args, local)),
                     ("local_mapping", lambda: _mapping_leg(args, local)), ("single_frame_latency", lambda: _latency_leg(args, local))):
        try:
            out[name] = fn()
        except Exception as e:  # the headline metric must still be reported
            out[name] = {"error": repr(e)[:400]}
        torch.cuda.empty_cache()
    return out


def _sharded_ba(local, rank, world, dist, torch, dense=False):
    """LocalBA with landmarks sharded over all ranks; rank 0 also solves the problem alone and reports parity of the sharded result against it."""
    from cubemapslam_b200.optimizer import Optimizer
    p = synth.ba_problem(nKF=50, nMP=4000, kmin=50, kmax=50, radius=9.0) if dense else synth.ba_problem()
    a = (p["Tcw"], p["kf_fixed"], p["pts"], p["eMP"], p["eKF"], p["kpxy"], p["inv_sigma2"], 650, 650)
    o = Optimizer(device=local)
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.from_numpy(Optimizer.nccl_unique_id()))
    dist.broadcast(idt, 0)
    o.init_nccl(idt.cpu().numpy(), rank, world)
    o.LocalBundleAdjustment(*a, its1=1, its2=0)
    best = None
    for _ in range(2):
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        g = o.LocalBundleAdjustment(*a)
        torch.cuda.synchronize(); dist.barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        best = float(tt.item()) if best is None else min(best, float(tt.item()))
    o.close()
    res = {"config": ("dense 50 KF x 4000 pts x 50 obs" if dense else "configs[3]") + " landmark-sharded over %d GPUs, one exchange of [S|g|bpr] per LM trial" % world, "edges": int(len(p["eMP"])),
           "lm_iters": int(g["iters"]), "seconds": round(best, 4), "lm_iters_per_s": round(g["iters"] / best, 2),
           "exchange": "one-shot NVLink kernel (peer pointers)" if not os.environ.get("CSLAM_BA_NCCL_ONLY") else "ncclAllReduce"}
    if rank == 0:
        ref = Optimizer(device=local)
        ref.LocalBundleAdjustment(*a, its1=1, its2=0)
        t0 = time.perf_counter(); r = ref.LocalBundleAdjustment(*a); d1 = time.perf_counter() - t0
        ref.close()
        rel = lambda x, y: float(np.max(np.abs(x - y)) / max(np.max(np.abs(y)), 1e-30))
        res.update({"single_gpu_lm_iters_per_s": round(r["iters"] / d1, 2), "speedup_vs_single_gpu": round(d1 / best, 3), "rel_pose_vs_single_gpu": rel(g["pose64"], r["pose64"]),
                    "rel_points_vs_single_gpu": rel(g["pts64"], r["pts64"]), "outliers_equal": bool(np.array_equal(g["outlier"], r["outlier"])),
                    "log_equal": bool(np.array_equal(g["log"][:, 2:], r["log"][:, 2:]))})
    return res


def run_multi(args, local, rank, world, dist, torch):
    """N > 1: every rank takes part; rank 0 gets the merged dictionary."""
    dev = torch.device("cuda", local)
    out = {}

    def allsum(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.SUM); return float(t.item())
    try:
        out["local_ba_sharded"] = _sharded_ba(local, rank, world, dist, torch)
    except Exception as e:
        out["local_ba_sharded"] = {"error": repr(e)[:400]}
    try:
        out["local_ba_dense_sharded"] = _sharded_ba(local, rank, world, dist, torch, dense=True)
    except Exception as e:
        out["local_ba_dense_sharded"] = {"error": repr(e)[:400]}
    try:
        m = _match_leg(torch, dev, args, local, rank, world)
        out["match_sharded"] = {"config": "configs[2], %d pairs per GPU, pair-sharded over %d GPUs (no collective)" % (args.match_pairs, world),
                                "bruteforce_pairs_per_s_all_gpus": round(allsum(m["bruteforce"]["pairs_per_s"]), 1),
                                "search_by_bow_pairs_per_s_all_gpus": round(allsum(m["search_by_bow"]["pairs_per_s"]), 1), "rank0": m}
    except Exception as e:
        out["match_sharded"] = {"error": repr(e)[:400]}
    try:
        t = _tracking_leg(torch, dev, args, local, rank, world)
        out["tracking_sharded"] = {"config": "configs[4], %d frame tasks per GPU, frame-sharded over %d GPUs (no collective)" % (args.track_frames, world),
                                   "frames_per_s_all_gpus": round(allsum(t["frames_per_s_per_gpu"]), 1), "rank0": t}
    except Exception as e:
        out["tracking_sharded"] = {"error": repr(e)[:400]}
    return out
