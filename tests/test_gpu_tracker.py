"""GPU parity of the per-frame indexing and the projection matchers against the oracle, bit-exact, through the C ABI."""
import numpy as np
import pytest

from cubemapslam_b200 import config, synth

pytestmark = pytest.mark.gpu


def cos_fov_th(fov_deg):
    """CamModelGeneral::SetCosFovTh (reference include/CamModelGeneral.h:224-229): float argument, double cos, float result."""
    a = np.float32(fov_deg) / np.float32(2) * (np.float32(3.1415926535897932384626) / np.float32(180))
    return np.float32(np.cos(np.float64(a)))


@pytest.fixture(scope="module")
def trk():
    from cubemapslam_b200.tracker import Tracker
    t = Tracker(max_frames=4, max_features=4096)
    yield t
    t.close()


def test_frame_index_rays_and_grid(oracle, trk):
    from cubemapslam_b200.frontend import FrontEnd
    cfg = config.lafida_450(); mask = config.load_mask("gray_lafida_cubemap_mask_450")
    fe = FrontEnd(cfg, mask, max_batch=2)
    res = fe.run(np.stack([synth.fisheye_frame(cfg, i) for i in (0, 4)]))
    fe.close()
    stride = max(len(r[0]) for r in res) + 5
    kps = np.zeros((2, stride), res[0][0].dtype); n = np.zeros(2, np.int32)
    for f in range(2):
        n[f] = len(res[f][0]); kps[f, :n[f]] = res[f][0]
    rays, cs, ci = trk.frame_index(kps, 450, 450, n=n)
    for f in range(2):
        r_or, _ = oracle.key_point_rays(res[f][0], 450, 450)
        assert np.array_equal(rays[f, :n[f]], r_or)
        start, idx = oracle.FrameGrid(res[f][0], 450, 450).csr()
        assert np.array_equal(cs[f].astype(np.int32), start) and np.array_equal(ci[f, :n[f]].astype(np.int32), idx)


@pytest.mark.parametrize("seed,th,ori", [(0, 15.0, True), (1, 7.0, True), (2, 15.0, False), (3, 30.0, True)])
def test_search_by_projection_last_frame(oracle, trk, seed, th, ori):
    s = synth.tracking_pair(seed, n=1500, faceW=650)
    cth = cos_fov_th(config.front_1024()["Camera.fov"])
    g = oracle.FrameGrid(s["kCur"], 650, 650)
    n1, m1 = g.search_by_projection_last(s["dCur"], s["TcwCur"], s["scale"], s["kLast"], s["hasMP"], s["Xw"], s["dLast"], s["mpObs"], s["curTaken"], cth, th, ori)
    n2, m2 = trk.SearchByProjection_last(s["kCur"], s["dCur"], s["TcwCur"], s["kLast"], s["hasMP"], s["Xw"], s["dLast"], s["mpObs"], s["curTaken"], 650, 650, cth, th, ori)
    assert n1 == n2 and n1 > 300 and np.array_equal(m1, np.where(m2 == -2, -1, m2))


def test_search_by_projection_batched_ragged(oracle, trk):
    pairs = [synth.tracking_pair(20 + i, n=n, faceW=650) for i, n in enumerate((900, 1500, 40))]
    cth = cos_fov_th(190.0)
    cs = max(len(p["kCur"]) for p in pairs) + 3; ls = max(len(p["kLast"]) for p in pairs) + 2
    P = len(pairs)
    kC = np.zeros((P, cs), pairs[0]["kCur"].dtype); dC = np.zeros((P, cs, 32), np.uint8); tk = np.zeros((P, cs), np.uint8); T = np.zeros((P, 4, 4), np.float32)
    kL = np.zeros((P, ls), kC.dtype); has = np.zeros((P, ls), np.uint8); Xw = np.zeros((P, ls, 3), np.float32); dM = np.zeros((P, ls, 32), np.uint8); ob = np.zeros((P, ls), np.int32)
    nC = np.zeros(P, np.int32); nL = np.zeros(P, np.int32)
    for i, p in enumerate(pairs):
        a, b = len(p["kCur"]), len(p["kLast"]); nC[i] = a; nL[i] = b
        kC[i, :a] = p["kCur"]; dC[i, :a] = p["dCur"]; tk[i, :a] = p["curTaken"]; T[i] = p["TcwCur"]
        kL[i, :b] = p["kLast"]; has[i, :b] = p["hasMP"]; Xw[i, :b] = p["Xw"]; dM[i, :b] = p["dLast"]; ob[i, :b] = p["mpObs"]
    nm, match = trk.SearchByProjection_last(kC, dC, T, kL, has, Xw, dM, ob, tk, 650, 650, cth, 15.0, True, nCur=nC, nLast=nL)
    for i, p in enumerate(pairs):
        g = oracle.FrameGrid(p["kCur"], 650, 650)
        n1, m1 = g.search_by_projection_last(p["dCur"], p["TcwCur"], p["scale"], p["kLast"], p["hasMP"], p["Xw"], p["dLast"], p["mpObs"], p["curTaken"], cth, 15.0, True)
        assert nm[i] == n1 and np.array_equal(np.where(match[i, :nC[i]] == -2, -1, match[i, :nC[i]]), m1), i


@pytest.mark.parametrize("seed,th,nn", [(0, 1.0, 0.8), (1, 3.0, 0.8), (2, 5.0, 0.6)])
def test_search_by_projection_local_map(oracle, trk, seed, th, nn):
    s = synth.tracking_pair(10 + seed, n=1500, faceW=650)
    rng = np.random.default_rng(seed)
    has = s["src"] >= 0
    nMP = 1200
    pick = rng.choice(np.nonzero(has)[0], nMP, replace=False)
    proj = np.stack([s["kCur"]["x"][pick], s["kCur"]["y"][pick]], 1).astype(np.float32) + rng.normal(0, 2.0, (nMP, 2)).astype(np.float32)
    lvl = np.clip(s["kCur"]["octave"][pick] + rng.integers(0, 2, nMP), 0, 7).astype(np.int32)
    vcos = rng.choice(np.array([0.9999, 0.99, 0.7], np.float32), nMP)
    dMP = s["dLast"][s["src"][pick]]
    inView = (rng.random(nMP) < 0.9).astype(np.uint8); obs = (rng.random(nMP) < 0.9).astype(np.int32)
    taken = (rng.random(len(s["kCur"])) < 0.05).astype(np.uint8)
    g = oracle.FrameGrid(s["kCur"], 650, 650)
    n1, m1 = g.search_by_projection_local(s["dCur"], s["scale"], inView, proj, lvl, vcos, dMP, obs, taken, th, nn)
    n2, m2 = trk.SearchByProjection_local(s["kCur"], s["dCur"], inView, proj, lvl, vcos, dMP, obs, taken, 650, 650, th, nn)
    assert n1 == n2 and n1 > 200 and np.array_equal(m1, m2)
