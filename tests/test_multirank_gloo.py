"""N>1 host logic on CPU (gloo, world_size 2): (1) the frame shard of bench.py covers every frame exactly once;
(2) landmark-sharded reduced camera systems, all-reduced, equal the unsharded system - the identity the multi-GPU
LocalBA relies on (rank r owns landmarks l % N == r; poses replicated; sum of [S | g] over ranks)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def shard_system(p, rank, world, orc):
    """Partial [S | g] of one rank at lambda = 0 built from the oracle's edge evaluation (numpy)."""
    free = [k for k in range(len(p["kf_fixed"])) if not p["kf_fixed"][k]]
    pidx = {k: i for i, k in enumerate(free)}
    n = 6 * len(free)
    S = np.zeros((n, n)); g = np.zeros(n)
    Hll = {}; bl = {}; B = {}
    for e in range(len(p["eMP"])):
        k, l = int(p["eKF"][e]), int(p["eMP"][e])
        if l % world != rank:
            continue
        err, Jp, Jx, face = orc.edge_eval(p["Tcw"][k], p["pts"][l].astype(np.float64), p["kpxy"][e, 0], p["kpxy"][e, 1], p["faceW"], p["faceH"])
        w = float(p["inv_sigma2"][e])
        Hll[l] = Hll.get(l, np.zeros((3, 3))) + w * Jx.T @ Jx
        bl[l] = bl.get(l, np.zeros(3)) - w * Jx.T @ err
        if k in pidx:
            i = pidx[k]
            S[6 * i:6 * i + 6, 6 * i:6 * i + 6] += w * Jp.T @ Jp
            g[6 * i:6 * i + 6] -= w * Jp.T @ err
            B.setdefault(l, []).append((i, w * Jp.T @ Jx))
    for l, blocks in B.items():
        Dinv = np.linalg.inv(Hll[l] + 1e-3 * np.eye(3))
        for i1, B1 in blocks:
            g[6 * i1:6 * i1 + 6] -= B1 @ Dinv @ bl[l]
            for i2, B2 in blocks:
                S[6 * i1:6 * i1 + 6, 6 * i2:6 * i2 + 6] -= B1 @ Dinv @ B2.T
    return S, g


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as orc
    from cubemapslam_b200 import synth
    # (1) frame shard
    frames = 37
    mine = torch.zeros(frames, dtype=torch.int32)
    mine[rank::world] = 1
    dist.all_reduce(mine)
    ok1 = bool((mine == 1).all())
    # (2) landmark-sharded Schur system
    p = synth.ba_problem(nKF=5, nMP=60, kmin=2, kmax=5, faceW=450, seed=8, radius=1.5)
    S, g = shard_system(p, rank, world, orc)
    t = torch.from_numpy(np.concatenate([S.ravel(), g]))
    dist.all_reduce(t)
    Sf, gf = shard_system(p, 0, 1, orc)
    full = np.concatenate([Sf.ravel(), gf])
    ok2 = bool(np.allclose(t.numpy(), full, rtol=1e-12, atol=1e-9))
    if rank == 0:
        q.put((ok1, ok2, float(np.abs(t.numpy() - full).max())))
    dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = q.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert res[0] and res[1], res
