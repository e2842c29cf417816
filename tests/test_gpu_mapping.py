"""GPU parity of the LocalMapping feature operations (cubemapslam_b200/csrc/mapping.cu) against the oracle (oracle/mapping.h, itself pinned to the
compiled reference in tests/test_oracle_mapping.py): integer / index results, bit-exact."""
import numpy as np
import pytest

from cubemapslam_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mapper():
    from cubemapslam_b200.mapper import Mapper
    m = Mapper()
    yield m
    m.close()


def test_distinctive_descriptors_batch(oracle, mapper):
    rng = np.random.default_rng(0)
    descs = []; off = [0]
    for p in range(3000):
        N = int(rng.choice([0, 1, 2, 3, 4, 5, 8, 13, 21, 40, 130, 300], p=[.02, .08, .15, .15, .15, .15, .1, .1, .05, .03, .01, .01]))
        base = rng.integers(0, 256, 32, dtype=np.uint8)
        d = np.stack([base ^ np.packbits(rng.random(256) < rng.uniform(0.0, 0.3), bitorder="little") for _ in range(N)]) if N else np.zeros((0, 32), np.uint8)
        if N > 3 and p % 5 == 0:
            d[2] = d[0]
        descs.append(d); off.append(off[-1] + N)
    desc = np.concatenate(descs); off = np.array(off, np.int32)
    want = oracle.distinctive_descriptors(desc, off)
    got = mapper.ComputeDistinctiveDescriptors(desc, off)
    assert np.array_equal(got, want)
    assert (want == -1).sum() > 10 and (want > 0).sum() > 1000
    # more observations than the shared-memory staging holds: same answer from the global-memory path
    big = rng.integers(0, 256, (1500, 32), dtype=np.uint8)
    assert mapper.ComputeDistinctiveDescriptors(big, [0, 1500])[0] == oracle.distinctive_descriptors(big, np.array([0, 1500], np.int32))[0]


@pytest.mark.parametrize("seed,th", [(0, 3.0), (1, 3.0), (2, 6.0), (3, 12.0)])
def test_fuse_search(oracle, mapper, seed, th):
    s = synth.mapping_pair(seed, n=1500, faceW=650)
    rng = np.random.default_rng(seed)
    n = len(s["Xw"])
    valid = (rng.random(n) < 0.9).astype(np.uint8)
    level = np.clip(s["kLast"]["octave"] + rng.integers(-1, 2, n), 0, 7).astype(np.int32)
    g = oracle.FrameGrid(s["kCur"], 650, 650)
    bi, bd = g.fuse_search(s["dCur"], s["TcwCur"], s["scale"], s["inv_level_sigma2"], valid, s["Xw"], level, s["dLast"], th)
    gi, gd = mapper.FuseSearch(s["kCur"], s["dCur"], s["TcwCur"], valid, s["Xw"], level, s["dLast"], th, s["scale"], s["inv_level_sigma2"], 650, 650)
    assert np.array_equal(gi, bi) and np.array_equal(gd, bd)
    assert (bd <= 50).sum() > 150 and (bi[valid == 0] == -1).all()


@pytest.mark.parametrize("ori", [False, True])
def test_search_for_triangulation_batch(oracle, mapper, ori):
    P = 6
    ss = [synth.mapping_pair(20 + p, n=1200 + 100 * p, faceW=650) for p in range(P)]
    s1 = max(len(s["kCur"]) for s in ss); s2 = max(len(s["kLast"]) for s in ss)
    KP = ss[0]["kCur"].dtype
    k1 = np.zeros((P, s1), KP); d1 = np.zeros((P, s1, 32), np.uint8); r1 = np.zeros((P, s1, 3), np.float32); h1 = np.zeros((P, s1), np.uint8); nd1 = np.zeros((P, s1), np.int32)
    k2 = np.zeros((P, s2), KP); d2 = np.zeros((P, s2, 32), np.uint8); r2 = np.zeros((P, s2, 3), np.float32); h2 = np.zeros((P, s2), np.uint8); nd2 = np.zeros((P, s2), np.int32)
    n1 = np.zeros(P, np.int32); n2 = np.zeros(P, np.int32); Ow = np.zeros((P, 3), np.float32); T2 = np.zeros((P, 16), np.float32); E = np.zeros((P, 9), np.float32)
    want = []
    for p, s in enumerate(ss):
        a, b = len(s["kCur"]), len(s["kLast"]); n1[p] = a; n2[p] = b
        ra, _ = oracle.key_point_rays(s["kCur"], 650, 650); rb, _ = oracle.key_point_rays(s["kLast"], 650, 650)
        k1[p, :a] = s["kCur"]; d1[p, :a] = s["dCur"]; r1[p, :a] = ra; h1[p, :a] = s["hasMPCur"]; nd1[p, :a] = s["nodeCur"]
        k2[p, :b] = s["kLast"]; d2[p, :b] = s["dLast"]; r2[p, :b] = rb; h2[p, :b] = s["hasMPObs"]; nd2[p, :b] = s["nodeObs"]
        T1 = s["TcwCur"].astype(np.float64)
        Ow[p] = (-T1[:3, :3].T @ T1[:3, 3]).astype(np.float32); T2[p] = s["TcwLast"].reshape(16); E[p] = s["E12"].reshape(9)
        want.append(oracle.search_for_triangulation(s["kCur"], s["dCur"], ra, s["hasMPCur"], s["nodeCur"], s["kLast"], s["dLast"], rb, s["hasMPObs"], s["nodeObs"], Ow[p], T2[p], E[p],
                                                    s["scale"], s["level_sigma2"], 650, 650, ori))
    nm, m = mapper.SearchForTriangulation(k1, d1, r1, h1, nd1, n1, k2, d2, r2, h2, nd2, n2, Ow, T2, E, ss[0]["scale"], ss[0]["level_sigma2"], 650, 650, ori)
    for p in range(P):
        assert nm[p] == want[p][0] and nm[p] > 80
        assert np.array_equal(m[p, :n1[p]], want[p][1])
