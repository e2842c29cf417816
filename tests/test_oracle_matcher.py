"""Oracle Hamming matcher: known-answer tests + numpy / pure-Python transcriptions (reference src/ORBMatcher.cpp:409-539,905-967)."""
import math
import numpy as np

from cubemapslam_b200 import synth

POP = np.array([bin(i).count("1") for i in range(256)], np.int32)


def hamming_matrix(A, B):
    return POP[A[:, None, :] ^ B[None, :, :]].sum(-1)


def test_distance_kats(oracle):
    z = np.zeros(32, np.uint8); o = np.full(32, 255, np.uint8)
    assert oracle.descriptor_distance(z, o) == 256 and oracle.descriptor_distance(o, o) == 0
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, 32, dtype=np.uint8)
    for bit in (0, 7, 100, 255):
        b = a.copy(); b[bit // 8] ^= 1 << (bit % 8)
        assert oracle.descriptor_distance(a, b) == 1
    b = rng.integers(0, 256, 32, dtype=np.uint8)
    assert oracle.descriptor_distance(a, b) == int(POP[a ^ b].sum())


def three_maxima(h):
    m = [0, 0, 0]; ind = [-1, -1, -1]
    for i, s in enumerate(h):
        if s > m[0]:
            m = [s, m[0], m[1]]; ind = [i, ind[0], ind[1]]
        elif s > m[1]:
            m = [m[0], s, m[1]]; ind = [ind[0], i, ind[1]]
        elif s > m[2]:
            m[2] = s; ind[2] = i
    if m[1] < np.float32(0.1) * np.float32(m[0]):
        ind[1] = ind[2] = -1
    elif m[2] < np.float32(0.1) * np.float32(m[0]):
        ind[2] = -1
    return ind


def rot_bin(a, b):
    rot = np.float32(a) - np.float32(b)
    if rot < 0:
        rot = np.float32(rot + np.float32(360))
    v = float(np.float32(rot * (np.float32(1.0) / np.float32(12))))
    bn = int(math.floor(abs(v) + 0.5)) * (1 if v >= 0 else -1)     # C round(): half away from zero
    return 0 if bn == 30 else bn


def py_bruteforce(A, angA, B, angB, nnratio, th, ori):
    D = hamming_matrix(A, B)
    match = np.full(len(A), -1, np.int32); hist = [[] for _ in range(30)]
    best = D.argmin(1); d1 = D[np.arange(len(A)), best]
    D2 = D.copy(); D2[np.arange(len(A)), best] = 9999
    d2 = np.minimum(D2.min(1), 256)
    for i in range(len(A)):
        if d1[i] <= th and np.float32(d1[i]) < np.float32(nnratio) * np.float32(d2[i]):
            match[i] = best[i]
            hist[rot_bin(angA[i], angB[best[i]])].append(i)
    if ori:
        keep = three_maxima([len(h) for h in hist])
        for b, h in enumerate(hist):
            if b not in keep:
                match[h] = -1
    return match, d1, d2


def test_bruteforce_vs_numpy(oracle):
    for pair in range(3):
        A, angA, B, angB, perm = synth.descriptor_pair(pair, n=300)
        n, m, d, s = oracle.match_bruteforce(A, angA, B, angB, 0.6, 50, True)
        rm, rd, rs = py_bruteforce(A, angA, B, angB, 0.6, 50, True)
        assert np.array_equal(m, rm) and np.array_equal(d, rd) and np.array_equal(s, rs) and n == int((rm >= 0).sum())
        assert n > 100
        inv = np.empty_like(perm); inv[perm] = np.arange(len(perm))
        ok = m >= 0
        assert (m[ok] == inv[ok]).mean() > 0.99       # matches recover the planted permutation


def py_bow(dKF, aKF, valid, nKF_node, dF, aF, nF_node, nnratio, ori):
    match = np.full(len(dF), -1, np.int32); hist = [[] for _ in range(30)]; n = 0
    for node in sorted(set(nKF_node.tolist()) & set(nF_node.tolist())):
        iK = np.nonzero(nKF_node == node)[0]; iF = np.nonzero(nF_node == node)[0]
        for k in iK:
            if not valid[k]:
                continue
            b1 = b2 = 256; bi = -1
            for f in iF:
                if match[f] >= 0:
                    continue
                dist = int(POP[dKF[k] ^ dF[f]].sum())
                if dist < b1:
                    b2 = b1; b1 = dist; bi = f
                elif dist < b2:
                    b2 = dist
            if b1 <= 50 and np.float32(b1) < np.float32(nnratio) * np.float32(b2):
                match[bi] = k; n += 1
                hist[rot_bin(aKF[k], aF[bi])].append(bi)
    if ori:
        keep = three_maxima([len(h) for h in hist])
        for b, h in enumerate(hist):
            if b not in keep:
                for f in h:
                    match[f] = -1; n -= 1
    return n, match


def test_bow_vs_python(oracle):
    rng = np.random.default_rng(4)
    A, angA, B, angB, perm = synth.descriptor_pair(7, n=400)
    nodeA = rng.integers(0, 30, 400).astype(np.int32); nodeB = nodeA[perm].copy()
    nodeB[rng.random(400) < 0.1] = 77                      # node missing on the KF side -> lower_bound skip path
    nodeA[rng.random(400) < 0.05] = 99                     # node missing on the F side
    valid = (rng.random(400) < 0.8).astype(np.uint8)
    for ratio, ori in ((0.7, True), (0.9, False)):
        n, m = oracle.search_by_bow(A, angA, valid, nodeA, B, angB, nodeB, ratio, ori)
        rn, rm = py_bow(A, angA, valid, nodeA, B, angB, nodeB, ratio, ori)
        assert n == rn and np.array_equal(m, rm) and n > 50


def test_empty_inputs(oracle):
    A = np.zeros((0, 32), np.uint8); ang = np.zeros(0, np.float32)
    B, angB = synth.descriptor_pair(1, n=16)[2:4]
    n, m, d, s = oracle.match_bruteforce(A, ang, B, angB)
    assert n == 0 and len(m) == 0
    n, m, d, s = oracle.match_bruteforce(B, angB, A, ang)
    assert n == 0 and np.all(m == -1) and np.all(d == 256)


def test_golden_match(oracle):
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "match_pair0_n500.npz"))
    A, angA, B, angB, perm = synth.descriptor_pair(0, n=500)
    n, m, d, s = oracle.match_bruteforce(A, angA, B, angB, 0.6, 50, True)
    assert n == int(g["bf_n"]) and np.array_equal(m, g["bf_match"]) and np.array_equal(d, g["bf_dist"]) and np.array_equal(s, g["bf_second"])
    nb, mf = oracle.search_by_bow(A, angA, g["valid"], g["nodeA"], B, angB, g["nodeB"], 0.7, True)
    assert nb == int(g["bow_n"]) and np.array_equal(mf, g["bow_match"])


def test_bow_keyframe_pair_properties(oracle):
    rng = np.random.default_rng(3)
    A, angA, B, angB, perm = synth.descriptor_pair(5, n=400)
    node1 = rng.integers(0, 30, 400).astype(np.int32); node2 = node1[perm].copy()
    v1 = (rng.random(400) < 0.8).astype(np.uint8); v2 = (rng.random(400) < 0.8).astype(np.uint8)
    n, m = oracle.search_by_bow_kf(A, angA, v1, node1, B, angB, v2, node2, 0.75, True)
    ok = m >= 0
    assert n == int(ok.sum()) and n > 50
    assert np.all(v1[ok] == 1) and np.all(v2[m[ok]] == 1)                     # both sides need a good MapPoint
    assert len(set(m[ok].tolist())) == n                                        # features of KF2 are consumed once
    assert np.all(node1[ok] == node2[m[ok]])                                    # same vocabulary node
    assert all(int(POP[A[i] ^ B[m[i]]].sum()) < 50 for i in np.nonzero(ok)[0])  # strict TH_LOW
