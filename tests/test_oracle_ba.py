"""Oracle BA (fp64 g2o restatement) against numpy: analytic Jacobians vs finite differences, one damped Gauss-Newton
step vs a dense numpy solve, and convergence to the planted ground truth (reference src/Optimizer.cpp:48-451,
src/g2o_cubemap_vertices_edges.cpp:61-233)."""
import numpy as np

from cubemapslam_b200 import synth


def quat_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


R_LOCAL = {0: np.eye(3), 1: np.array([[0, 0, 1], [0, 1, 0], [-1, 0, 0.]]), 2: np.array([[0, 0, -1], [0, 1, 0], [1, 0, 0.]]),
           3: np.array([[1, 0, 0], [0, 0, 1], [0, -1, 0.]]), 4: np.array([[1, 0, 0], [0, 0, -1], [0, 1, 0.]])}


def project(T, X, face, f):
    Xc = T[:3, :3] @ X + T[:3, 3]
    L = R_LOCAL[face] @ Xc
    return np.array([L[0] * f / L[2] + f, L[1] * f / L[2] + f])


def test_se3_exp_matches_rodrigues(oracle):
    rng = np.random.default_rng(0)
    for scale in (1e-7, 1e-3, 0.3, 2.0):
        u = rng.normal(0, scale, 6)
        out = oracle.se3_exp(u)
        R = quat_R(out[3:])
        assert np.allclose(R, synth._rodrigues(u[:3]), atol=1e-12 if scale > 1e-6 else 1e-10)
        assert abs(np.linalg.norm(out[3:]) - 1) < 1e-15 and out[6] >= 0


def test_edge_jacobians_finite_difference(oracle):
    p = synth.ba_problem(nKF=4, nMP=40, kmin=2, kmax=4, faceW=450, seed=3, outlier_frac=0, radius=1.5)
    faces = set()
    for e in range(0, len(p["eMP"]), 3):
        T32 = p["Tcw"][p["eKF"][e]]; X = p["pts"][p["eMP"][e]].astype(np.float64)
        kx, ky = p["kpxy"][e]
        err, Jp, Jx, face = oracle.edge_eval(T32, X, kx, ky, 450, 450)
        faces.add(face)
        # pose given as float32 matrix -> the oracle re-orthonormalises through the quaternion
        T = T32.astype(np.float64)
        m = np.array([kx - (kx // 450) * 450, ky - (ky // 450) * 450], np.float64)
        e0 = m - project(T, X, face, 225.0)
        assert np.allclose(err, e0, atol=2e-3)              # float32 projection + quaternion renormalisation
        h = 1e-6
        for j in range(3):
            d = np.zeros(3); d[j] = h
            num = ((m - project(T, X + d, face, 225.0)) - (m - project(T, X - d, face, 225.0))) / (2 * h)
            assert np.allclose(Jx[:, j], num, rtol=1e-4, atol=1e-4)
        for j in range(6):
            d = np.zeros(6); d[j] = h
            Tp = np.eye(4); Tp[:3, :3] = synth._rodrigues(d[:3]); Tp[:3, 3] = d[3:]
            Tm = np.eye(4); Tm[:3, :3] = synth._rodrigues(-d[:3]); Tm[:3, 3] = -d[3:]
            num = ((m - project(Tp @ T, X, face, 225.0)) - (m - project(Tm @ T, X, face, 225.0))) / (2 * h)
            assert np.allclose(Jp[:, j], num, rtol=1e-4, atol=1e-4)
    assert len(faces) >= 3


def test_local_ba_converges_to_truth(oracle):
    p = synth.ba_problem(nKF=8, nMP=400, kmin=3, kmax=6, faceW=450, seed=21, radius=1.5)
    r = oracle.local_ba(p["Tcw"], p["kf_fixed"], p["pts"], p["eMP"], p["eKF"], p["kpxy"], p["inv_sigma2"], 450, 450)
    assert r["iters"] >= 4 and r["log"][0, 0] > r["log"][-1, 0]
    assert np.all(np.diff(r["log"][:5, 0]) <= 1e-9)          # accepted LM steps never increase chi2 within one optimize()
    t_err0 = np.linalg.norm(p["Tcw"][:, :3, 3] - p["Tcw_true"][:, :3, 3], axis=1).mean()
    t_err1 = np.linalg.norm(r["Tcw"][:, :3, 3] - p["Tcw_true"][:, :3, 3], axis=1).mean()
    assert t_err1 < 0.5 * t_err0
    frac = r["outlier"].mean()
    assert 0.02 < frac < 0.12                                 # ~5 % planted gross outliers
    assert np.allclose(r["Tcw"][0], p["Tcw"][0], atol=1e-6)   # fixed KF: only the quaternion round trip of the write-back


def test_first_lm_step_matches_dense_numpy(oracle):
    """One LM trial (its1=1, its2=0) must equal the dense damped normal-equation solve built from the oracle's own edge
    evaluations (checks accumulation, Huber weights, Schur elimination, back-substitution and the update)."""
    p = synth.ba_problem(nKF=5, nMP=60, kmin=2, kmax=5, faceW=450, seed=8, radius=1.5)
    r = oracle.local_ba(p["Tcw"], p["kf_fixed"], p["pts"], p["eMP"], p["eKF"], p["kpxy"], p["inv_sigma2"], 450, 450, its1=1, its2=0)
    assert r["iters"] == 1 and r["log"][0, 2] == 1 and r["log"][0, 3] == 1
    nKF, nMP = 5, 60
    free = [k for k in range(nKF) if not p["kf_fixed"][k]]
    pidx = {k: i for i, k in enumerate(free)}
    used = sorted(set(p["eMP"].tolist())); lidx = {l: i for i, l in enumerate(used)}
    n = 6 * len(free) + 3 * len(used)
    H = np.zeros((n, n)); b = np.zeros(n)
    delta = float(np.float32(np.sqrt(5.991)))
    for e in range(len(p["eMP"])):
        k, l = int(p["eKF"][e]), int(p["eMP"][e])
        err, Jp, Jx, face = oracle.edge_eval(p["Tcw"][k], p["pts"][l].astype(np.float64), p["kpxy"][e, 0], p["kpxy"][e, 1], 450, 450)
        w = float(p["inv_sigma2"][e]); chi = w * err @ err
        rho1 = 1.0 if chi <= delta * delta else delta / np.sqrt(chi)
        J = np.zeros((2, n))
        if k in pidx:
            J[:, 6 * pidx[k]:6 * pidx[k] + 6] = Jp
        o = 6 * len(free) + 3 * lidx[l]
        J[:, o:o + 3] = Jx
        H += rho1 * w * J.T @ J; b -= rho1 * w * J.T @ err
    lam = 1e-5 * np.abs(np.diag(H)).max()
    x = np.linalg.solve(H + lam * np.eye(n), b)
    for k in free:
        d = x[6 * pidx[k]:6 * pidx[k] + 6]
        T0 = p["Tcw"][k].astype(np.float64)
        q = oracle.se3_exp(d)
        Tn = np.eye(4); Tn[:3, :3] = quat_R(q[3:]) @ T0[:3, :3]; Tn[:3, 3] = quat_R(q[3:]) @ T0[:3, 3] + q[:3]
        got = np.eye(4); got[:3, :3] = quat_R(r["pose64"][k, 3:]); got[:3, 3] = r["pose64"][k, :3]
        assert np.allclose(got, Tn, atol=5e-7)                # T0 is a float32 matrix, the oracle re-normalises its rotation
    for l in used:
        assert np.allclose(r["pts64"][l], p["pts"][l].astype(np.float64) + x[6 * len(free) + 3 * lidx[l]:][:3], atol=1e-9)


def test_pose_optimization(oracle):
    q = synth.pose_problem(n=300, faceW=450, seed=5, outlier_frac=0.15)
    o = oracle.pose_opt(q["Tcw"], q["Xw"], q["kpxy"], q["inv_sigma2"], 450, 450)
    assert 0.7 * 300 < o["inliers"] < 300 and o["inliers"] == 300 - int(o["outlier"].sum())
    e0 = np.linalg.norm(q["Tcw"][:3, 3] - q["Tcw_true"][:3, 3]); e1 = np.linalg.norm(o["Tcw"][:3, 3] - q["Tcw_true"][:3, 3])
    assert e1 < 0.3 * e0
    o2 = oracle.pose_opt(q["Tcw"], q["Xw"][:2], q["kpxy"][:2], q["inv_sigma2"][:2], 450, 450)
    assert o2["inliers"] == 0 and np.array_equal(o2["Tcw"], q["Tcw"])     # < 3 correspondences: untouched (src/Optimizer.cpp:133-134)


def test_golden_ba(oracle):
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ba_small.npz"))
    p = synth.ba_problem(nKF=8, nMP=300, kmin=2, kmax=6, faceW=450, seed=11, radius=1.5)
    r = oracle.local_ba(p["Tcw"], p["kf_fixed"], p["pts"], p["eMP"], p["eKF"], p["kpxy"], p["inv_sigma2"], 450, 450)
    assert np.allclose(r["pose64"], g["pose64"], rtol=0, atol=1e-12) and np.allclose(r["pts64"], g["pts64"], rtol=0, atol=1e-12)
    assert np.array_equal(r["outlier"], g["outlier"]) and np.allclose(r["log"], g["log"], rtol=1e-12)
    q = synth.pose_problem(n=200, faceW=450, seed=5)
    o = oracle.pose_opt(q["Tcw"], q["Xw"], q["kpxy"], q["inv_sigma2"], 450, 450)
    assert o["inliers"] == int(g["po_inliers"]) and np.array_equal(o["outlier"], g["po_outlier"])
    assert np.allclose(o["pose64"], g["po_pose64"], rtol=0, atol=1e-12)
