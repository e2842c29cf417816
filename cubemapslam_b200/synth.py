"""Seeded synthetic inputs for the parity tests and bench.py (recipes: SURVEY.md §8d). numpy only."""
import math
import numpy as np


def _horner(coeffs, x):
    r = 0.0
    for c in reversed(coeffs):
        r = r * x + c
    return r


def _resize_bilinear_u8(src, dw, dh):
    """Plain float bilinear upsample (half-pixel centres); only used to make a smooth synthetic background."""
    sh, sw = src.shape
    ys = np.clip((np.arange(dh) + 0.5) * sh / dh - 0.5, 0, sh - 1); xs = np.clip((np.arange(dw) + 0.5) * sw / dw - 0.5, 0, sw - 1)
    y0 = np.floor(ys).astype(int); x0 = np.floor(xs).astype(int)
    y1 = np.minimum(y0 + 1, sh - 1); x1 = np.minimum(x0 + 1, sw - 1)
    fy = (ys - y0)[:, None]; fx = (xs - x0)[None, :]
    s = src.astype(np.float64)
    out = (s[y0][:, x0] * (1 - fy) * (1 - fx) + s[y0][:, x1] * (1 - fy) * fx + s[y1][:, x0] * fy * (1 - fx) + s[y1][:, x1] * fy * fx)
    return np.rint(out).astype(np.int32)


def fisheye_frame(cfg, frame_idx):
    """Synthetic fisheye frame: smooth base + 400 rectangles + +-3 noise, zero outside the image circle."""
    Iw, Ih = int(cfg["Camera.Iw"]), int(cfg["Camera.Ih"])
    rng = np.random.default_rng(1000 + frame_idx)
    base = _resize_bilinear_u8(rng.integers(0, 256, (Ih // 8, Iw // 8), dtype=np.uint8), Iw, Ih)
    for _ in range(400):
        x = int(rng.integers(0, Iw - 64)); y = int(rng.integers(0, Ih - 64))
        w, h = (int(v) for v in rng.integers(8, 64, 2))
        base[y:y + h, x:x + w] = int(rng.integers(0, 256))
    base = base + rng.integers(-3, 4, (Ih, Iw))
    img = np.clip(base, 0, 255).astype(np.uint8)
    invp = [cfg.get("Camera.pol%d" % i, 0.0) for i in range(int(cfg["Camera.nrinvpol"]))]
    fov = cfg["Camera.fov"] / 2.0 * math.pi / 180.0
    theta = math.atan(-math.cos(fov) / math.sin(fov))
    rho_max = _horner(invp, theta)
    yy, xx = np.mgrid[0:Ih, 0:Iw]
    img[(xx - cfg["Camera.u0"]) ** 2 + (yy - cfg["Camera.v0"]) ** 2 > rho_max ** 2] = 0
    return img


def descriptor_pair(pair_idx, n=2000):
    """Config-3 pair: B = permuted A with 8 % bit flips on 70 % of rows, fresh random on 30 %."""
    rng = np.random.default_rng(2000 + pair_idx)
    A = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    perm = rng.permutation(n)
    B = A[perm].copy()
    fresh = rng.random(n) < 0.3
    flips = np.packbits(rng.random((n, 256)) < 0.08, axis=1, bitorder="little")
    B ^= flips
    B[fresh] = rng.integers(0, 256, (int(fresh.sum()), 32), dtype=np.uint8)
    angA = rng.uniform(0, 360, n).astype(np.float32)
    angB = np.mod(angA[perm] + rng.uniform(-6, 6, n).astype(np.float32), np.float32(360)).astype(np.float32)
    angB[angB >= 360] = 0
    return A, angA, B, angB, perm


# ----------------------------------------------------------------------------- bundle adjustment problem
_FACE_TILE = {0: (1, 1), 1: (0, 1), 2: (2, 1), 3: (1, 0), 4: (1, 2)}  # face -> (col,row) of the canvas tile


def _rays_to_cubemap(x, y, z, W):
    """Ordered face tests of CamModelGeneral::TransformRaysToCubemap (reference src/CamModelGeneral.cpp:95-154)."""
    f = W / 2.0
    def proj(lx, ly, lz):
        return lx * f / lz + f, ly * f / lz + f
    if z > 0 and abs(x / z) <= 1 and abs(y / z) <= 1:
        face, (u, v) = 0, proj(x, y, z)
    elif x > 0 and abs(y / x) <= 1 and abs(z / x) <= 1:
        face, (u, v) = 2, proj(-z, y, x)
    elif x < 0 and abs(y / x) <= 1 and abs(z / x) <= 1:
        face, (u, v) = 1, proj(z, y, -x)
    elif y > 0 and abs(x / y) <= 1 and abs(z / y) <= 1:
        face, (u, v) = 4, proj(x, -z, y)
    elif y < 0 and abs(x / y) <= 1 and abs(z / y) <= 1:
        face, (u, v) = 3, proj(x, z, -y)
    else:
        return None
    if not (0 <= u < W and 0 <= v < W):
        return None
    c, r = _FACE_TILE[face]
    return u + c * W, v + r * W


def _rodrigues(w):
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + math.sin(th) / th * K + (1 - math.cos(th)) / th ** 2 * K @ K


def ba_problem(nKF=50, nMP=20000, kmin=2, kmax=16, faceW=650, seed=4242, fov_deg=190.0, outlier_frac=0.05, radius=5.0):
    """Config-4 style local BA window: KFs on a circle looking inward, points in a cube, windowed co-visibility.

    Returns float32 Tcw (nKF,4,4), kf_fixed, pts (nMP,3), edges (eMP,eKF), kpxy (canvas px), inv_sigma2 and the
    ground truth. All I/O arrays are float32 like the reference's cv::Mat boundary. radius=5 is the SURVEY §8d recipe
    (everything lands on the front face); a radius inside the point cube (e.g. 1.5) puts observations on all 5 faces."""
    rng = np.random.default_rng(seed)
    cos_fov = math.cos(fov_deg / 2 * math.pi / 180)
    Twc = []
    for k in range(nKF):
        a = 2 * math.pi * k / nKF
        c = np.array([radius * math.cos(a), 0.0, radius * math.sin(a)]) + rng.uniform(-0.2, 0.2, 3)
        zc = -c / np.linalg.norm(c)
        up = np.array([0.0, 1.0, 0.0])
        xc = np.cross(up, zc); xc /= np.linalg.norm(xc)
        yc = np.cross(zc, xc)
        R = np.stack([xc, yc, zc], 1)  # columns = camera axes in world
        Twc.append((R, c))
    Tcw_true = np.zeros((nKF, 4, 4))
    for k, (R, c) in enumerate(Twc):
        Tcw_true[k, :3, :3] = R.T; Tcw_true[k, :3, 3] = -R.T @ c; Tcw_true[k, 3, 3] = 1
    pts_true = rng.uniform(-3, 3, (nMP, 3))
    eMP, eKF, kp, octs = [], [], [], []
    for l in range(nMP):
        k = int(rng.integers(kmin, kmax + 1)); start = int(rng.integers(0, nKF))
        for kk in sorted((start + i) % nKF for i in range(min(k, nKF))):
            Xc = Tcw_true[kk, :3, :3] @ pts_true[l] + Tcw_true[kk, :3, 3]
            if Xc[2] / np.linalg.norm(Xc) < cos_fov:
                continue
            uv = _rays_to_cubemap(Xc[0], Xc[1], Xc[2], faceW)
            if uv is None:
                continue
            noise = rng.normal(0, 1, 2)
            if rng.random() < outlier_frac:
                noise = noise + rng.uniform(-30, 30, 2)
            u, v = uv[0] + noise[0], uv[1] + noise[1]
            c0, r0 = int(uv[0] // faceW), int(uv[1] // faceW)
            u = min(max(u, c0 * faceW + 0.01), (c0 + 1) * faceW - 0.01)   # keep the observation on its face tile
            v = min(max(v, r0 * faceW + 0.01), (r0 + 1) * faceW - 0.01)
            eMP.append(l); eKF.append(kk); kp.append((u, v)); octs.append(int(rng.integers(0, 8)))
    scale = np.float32(1.0); inv_sigma2_tab = []
    for i in range(8):
        inv_sigma2_tab.append(np.float32(1.0) / (scale * scale)); scale = scale * np.float32(1.2)
    inv_sigma2 = np.array([inv_sigma2_tab[o] for o in octs], np.float32)
    Tcw0 = np.zeros((nKF, 4, 4), np.float32)
    for k in range(nKF):
        d = rng.normal(0, 0.01, 6) if k > 0 else np.zeros(6)
        Rn = _rodrigues(d[:3]) @ Tcw_true[k, :3, :3]
        Tcw0[k, :3, :3] = Rn; Tcw0[k, :3, 3] = Tcw_true[k, :3, 3] + d[3:]; Tcw0[k, 3, 3] = 1
    pts0 = (pts_true + rng.normal(0, 0.02, pts_true.shape)).astype(np.float32)
    kf_fixed = np.zeros(nKF, np.uint8); kf_fixed[0] = 1
    return dict(Tcw=Tcw0, kf_fixed=kf_fixed, pts=pts0, eMP=np.array(eMP, np.int32), eKF=np.array(eKF, np.int32),
                kpxy=np.array(kp, np.float32).reshape(-1, 2), inv_sigma2=inv_sigma2, faceW=faceW, faceH=faceW,
                Tcw_true=Tcw_true, pts_true=pts_true)


def pose_problem(n=600, faceW=650, seed=77, outlier_frac=0.1, radius=1.5):
    """One PoseOptimization instance: known 3-D points, noisy cubemap observations, perturbed prior pose."""
    p = ba_problem(nKF=2, nMP=n * 3, kmin=2, kmax=2, faceW=faceW, seed=seed, outlier_frac=outlier_frac, radius=radius)
    sel = p["eKF"] == 1
    idx = np.nonzero(sel)[0][:n]
    Xw = p["pts_true"][p["eMP"][idx]].astype(np.float32)
    return dict(Tcw=p["Tcw"][1].copy(), Xw=Xw, kpxy=p["kpxy"][idx].copy(), inv_sigma2=p["inv_sigma2"][idx].copy(), faceW=faceW, faceH=faceW,
                Tcw_true=p["Tcw_true"][1])


# ----------------------------------------------------------------------------- tracking (projection matching) scenario
def _pixel_to_ray(px, py, W):
    """Unit bearing vector of a canvas pixel (plain float64 pinhole per face; generator only, the exact reference arithmetic lives in the library)."""
    f = W / 2.0
    c, r = int(px // W), int(py // W)
    face = {(1, 1): 0, (0, 1): 1, (2, 1): 2, (1, 0): 3, (1, 2): 4}.get((c, r))
    if face is None:
        return None
    lx, ly = (px - c * W - f) / f, (py - r * W - f) / f
    v = {0: (lx, ly, 1.0), 1: (-1.0, ly, lx), 2: (1.0, ly, -lx), 4: (lx, 1.0, -ly), 3: (lx, -1.0, ly)}[face]
    v = np.array(v)
    return v / np.linalg.norm(v)


def tracking_pair(seed=0, n=2000, faceW=650, motion=0.02, rot=0.01, mp_frac=0.8, extra_frac=0.15):
    """A (LastFrame, CurrentFrame) pair for ORBMatcher::SearchByProjection: LastFrame key points spread over all five faces with MapPoints at
    random depth, CurrentFrame = the same points seen after a small motion (+ pixel noise, descriptor bit flips, unrelated extra features).
    All arrays float32 / uint8 / int32 like the reference's containers. kps use the cv::KeyPoint field order (x, y, size, angle, response, octave, class_id)."""
    rng = np.random.default_rng(7000 + seed)
    KP = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
    W = faceW
    tiles = [(1, 1), (0, 1), (2, 1), (1, 0), (1, 2)]
    kL = np.zeros(n, KP); Xw = np.zeros((n, 3), np.float32)
    TcwL = np.eye(4, dtype=np.float32)
    Rc = _rodrigues(rng.normal(0, rot, 3)); tc = rng.normal(0, motion, 3)
    TcwC = np.eye(4, dtype=np.float32); TcwC[:3, :3] = Rc; TcwC[:3, 3] = tc
    rays = []
    for i in range(n):
        c, r = tiles[rng.integers(0, 5)]
        # a third of the points hug the face borders so that search windows cross face seams
        u, v = rng.uniform(1, W - 1, 2)
        if i % 3 == 0:
            u = rng.choice([rng.uniform(0.5, 25), rng.uniform(W - 25, W - 0.5)])
        if i % 6 == 0:
            v = rng.choice([rng.uniform(0.5, 25), rng.uniform(W - 25, W - 0.5)])
        kL["x"][i] = c * W + u; kL["y"][i] = r * W + v
        ray = _pixel_to_ray(float(kL["x"][i]), float(kL["y"][i]), W)
        rays.append(ray)
        Xw[i] = (ray * rng.uniform(2.0, 12.0)).astype(np.float32)
    kL["octave"] = rng.integers(0, 8, n); kL["angle"] = rng.uniform(0, 360, n).astype(np.float32); kL["size"] = 31; kL["class_id"] = -1
    dL = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    hasMP = (rng.random(n) < mp_frac).astype(np.uint8)
    # current frame: project, perturb
    kC = []; dC = []; src = []
    for i in range(n):
        Xc = Rc @ Xw[i].astype(np.float64) + tc
        uv = _rays_to_cubemap(Xc[0], Xc[1], Xc[2], W)
        if uv is None or rng.random() < 0.1:
            continue
        x, y = uv[0] + rng.normal(0, 1.5), uv[1] + rng.normal(0, 1.5)
        if _pixel_to_ray(x, y, W) is None:
            continue
        d = dL[i] ^ np.packbits(rng.random(256) < 0.06, bitorder="little")
        oct_ = int(np.clip(kL["octave"][i] + rng.integers(-1, 2), 0, 7))
        kC.append((x, y, 31, float((kL["angle"][i] + rng.uniform(-5, 5)) % 360), 0, oct_, -1)); dC.append(d); src.append(i)
    for _ in range(int(extra_frac * n)):
        c, r = tiles[rng.integers(0, 5)]
        kC.append((c * W + rng.uniform(0.5, W - 0.5), r * W + rng.uniform(0.5, W - 0.5), 31, float(rng.uniform(0, 360)), 0, int(rng.integers(0, 8)), -1))
        dC.append(rng.integers(0, 256, 32, dtype=np.uint8)); src.append(-1)
    perm = rng.permutation(len(kC))
    kC = np.array([kC[j] for j in perm], KP); dC = np.stack([dC[j] for j in perm]); src = np.array([src[j] for j in perm], np.int32)
    order = np.lexsort((kC["x"], kC["y"], kC["octave"]))          # extractor order: level-major
    kC, dC, src = kC[order], dC[order], src[order]
    scale = np.ones(8, np.float32)
    for l in range(1, 8):
        scale[l] = scale[l - 1] * np.float32(1.2)
    mpObs = (rng.random(n) < 0.9).astype(np.int32)
    curTaken = (rng.random(len(kC)) < 0.03).astype(np.uint8)
    return dict(kLast=kL, dLast=dL, TcwLast=TcwL, hasMP=hasMP, Xw=Xw, mpObs=mpObs, kCur=kC, dCur=dC, TcwCur=TcwC, src=src, scale=scale, curTaken=curTaken, faceW=W)


def mapping_pair(seed=0, n=1500, faceW=650):
    """Two key frames for the LocalMapping feature operations (Fuse, SearchForTriangulation), built on tracking_pair: KF "obs" = its LastFrame
    (identity pose, MapPoint m seen by key point m), KF "cur" = its CurrentFrame. Adds vocabulary-node ids shared by true correspondences, MapPoint
    presence flags and the essential matrix E12 (x_cur^T E12 x_obs = 0 for bearing vectors), float32 like the reference's cv::Mat."""
    s = tracking_pair(100 + seed, n=n, faceW=faceW, motion=0.25, rot=0.03)
    rng = np.random.default_rng(9000 + seed)
    nO, nC = len(s["kLast"]), len(s["kCur"])
    nodeO = rng.integers(0, 120, nO).astype(np.int32)
    nodeC = np.where(s["src"] >= 0, nodeO[np.maximum(s["src"], 0)], rng.integers(0, 120, nC)).astype(np.int32)
    flip = rng.random(nC) < 0.1
    nodeC[flip] = rng.integers(0, 120, int(flip.sum()))
    T1 = s["TcwCur"].astype(np.float64); T2 = s["TcwLast"].astype(np.float64)     # KF1 = cur, KF2 = obs
    R1, t1, R2, t2 = T1[:3, :3], T1[:3, 3], T2[:3, :3], T2[:3, 3]
    R12 = R1 @ R2.T; t12 = -R12 @ t2 + t1
    tx = np.array([[0, -t12[2], t12[1]], [t12[2], 0, -t12[0]], [-t12[1], t12[0], 0]])
    E12 = (tx @ R12).astype(np.float32)
    level_sigma2 = (s["scale"] * s["scale"]).astype(np.float32)
    inv_level_sigma2 = (np.float32(1.0) / level_sigma2).astype(np.float32)
    s.update(nodeObs=nodeO, nodeCur=nodeC, hasMPCur=(rng.random(nC) < 0.4).astype(np.uint8), hasMPObs=(rng.random(nO) < 0.4).astype(np.uint8), E12=E12,
             level_sigma2=level_sigma2, inv_level_sigma2=inv_level_sigma2)
    return s


# ----------------------------------------------------------------------------- vocabulary tree (DBoW2 text format)
def vocabulary(k=10, L=4, seed=0, stop_frac=0.01):
    """A synthetic vocabulary tree in ORBvoc.txt's node order (breadth-first per parent, like DBoW2's saveToTextFile): arrays for nodes 1..n
    (parent, is_leaf, 32-byte descriptor, weight). Children descriptors are the parent's with ~12 % of the bits flipped (a hierarchical
    k-medians-like structure); leaf weights are idf-like positive numbers, `stop_frac` of the words have weight 0 (stopped words)."""
    rng = np.random.default_rng(9000 + seed)
    parent = []; leaf = []; desc = []; weight = []
    frontier = [(0, rng.integers(0, 256, 32, dtype=np.uint8))]
    for level in range(1, L + 1):
        nxt = []
        for pid, pd in frontier:
            for _ in range(k):
                d = pd ^ np.packbits(rng.random(256) < 0.12, bitorder="little")
                parent.append(pid); leaf.append(1 if level == L else 0); desc.append(d)
                weight.append(0.0 if (level == L and rng.random() < stop_frac) else (float(rng.uniform(0.5, 9.0)) if level == L else 0.0))
                nxt.append((len(parent), d))
        frontier = nxt
    return dict(k=k, L=L, parent=np.array(parent, np.int32), is_leaf=np.array(leaf, np.uint8), desc=np.stack(desc), weight=np.array(weight, np.float64))


def write_vocabulary_text(voc, path):
    """ORBvoc.txt format (reference ThirdParty/DBoW2/DBoW2/TemplatedVocabulary.h:1337-1415): 'k L scoring weighting' then one line per node.
    No trailing newline: the reference's `while(!f.eof())` loop would turn it into one more (garbage) child of the root."""
    lines = ["%d %d 0 0" % (voc["k"], voc["L"])]
    for i in range(len(voc["parent"])):
        lines.append("%d %d %s %r" % (voc["parent"][i], voc["is_leaf"][i], " ".join(str(int(b)) for b in voc["desc"][i]), float(voc["weight"][i])))
    with open(path, "w") as f:
        f.write("\n".join(lines))
