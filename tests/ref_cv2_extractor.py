"""Independent Python transcription of the reference extractor control flow (src/ORBExtractor.cpp:739-926) that calls
the REAL OpenCV primitives (cv2 4.13: resize, FAST, GaussianBlur, fastAtan2). Test infrastructure: it referees the
C++ oracle on small images. Quadtree tie-break and sin/cos follow the oracle's documented definitions."""
import math
import numpy as np
import cv2

EDGE = 19
HALF = 15


def f32(x):
    return np.float32(x)


class Node:
    __slots__ = ("UL", "UR", "BL", "BR", "keys", "nomore", "seq")

    def __init__(self):
        self.keys = []; self.nomore = False; self.seq = 0


def divide(n):
    halfX = int(math.ceil(f32(n.UR[0] - n.UL[0]) / f32(2))); halfY = int(math.ceil(f32(n.BR[1] - n.UL[1]) / f32(2)))
    c = [Node() for _ in range(4)]
    c[0].UL = n.UL; c[0].UR = (n.UL[0] + halfX, n.UL[1]); c[0].BL = (n.UL[0], n.UL[1] + halfY); c[0].BR = (n.UL[0] + halfX, n.UL[1] + halfY)
    c[1].UL = c[0].UR; c[1].UR = n.UR; c[1].BL = c[0].BR; c[1].BR = (n.UR[0], n.UL[1] + halfY)
    c[2].UL = c[0].BL; c[2].UR = c[0].BR; c[2].BL = n.BL; c[2].BR = (c[0].BR[0], n.BL[1])
    c[3].UL = c[2].UR; c[3].UR = c[1].BR; c[3].BL = c[2].BR; c[3].BR = n.BR
    for k in n.keys:
        if k[0] < c[0].UR[0]:
            (c[0] if k[1] < c[0].BR[1] else c[2]).keys.append(k)
        elif k[1] < c[0].BR[1]:
            c[1].keys.append(k)
        else:
            c[3].keys.append(k)
    for x in c:
        if len(x.keys) == 1:
            x.nomore = True
    return c


def distribute(keys, minX, maxX, minY, maxY, N):
    """keys: list of (x, y, response); returns list in the reference's output order."""
    seq = [0]
    root = Node(); root.UL = (0, 0); root.UR = (maxX - minX, 0); root.BL = (0, maxY - minY); root.BR = (maxX - minX, maxY - minY)
    root.keys = list(keys)
    assert round((maxX - minX) / (maxY - minY)) == 1
    nodes = [root]
    if len(root.keys) == 1:
        root.nomore = True
    elif not root.keys:
        nodes = []
    finish = False

    def push(children, rec, counter):
        for ch in children:
            if ch.keys:
                ch.seq = seq[0] = seq[0] + 1
                nodes.insert(0, ch)
                if len(ch.keys) > 1:
                    counter[0] += 1
                    rec.append(ch)

    rec = []
    while not finish:
        prev = len(nodes)
        work = [n for n in nodes]
        rec = []; cnt = [0]; ndiv = 0
        for n in work:
            if n.nomore:
                continue
            ndiv += 1
            push(divide(n), rec, cnt)
            nodes.remove(n)
        if len(nodes) >= N or (len(nodes) == prev and len(nodes) >= N // 100):
            finish = True
        elif ndiv == 0:
            finish = True      # defined behaviour (the reference would spin forever), see oracle/orb_extractor.h
        elif len(nodes) + cnt[0] * 3 > N:
            while not finish:
                prev = len(nodes)
                prevrec = sorted(rec, key=lambda n: (len(n.keys), n.seq))
                rec = []
                for n in reversed(prevrec):
                    push(divide(n), rec, [0])
                    nodes.remove(n)
                    if len(nodes) >= N:
                        break
                if len(nodes) >= N or len(nodes) == prev:
                    finish = True
    out = []
    for n in nodes:
        best = n.keys[0]
        for k in n.keys[1:]:
            if k[2] > best[2]:
                best = k
        out.append(best)
    return out


def umax_table():
    um = [0] * 16
    vmax = int(math.floor(HALF * math.sqrt(2.0) / 2 + 1)); vmin = int(math.ceil(HALF * math.sqrt(2.0) / 2))
    for v in range(vmax + 1):
        um[v] = int(round(math.sqrt(HALF * HALF - v * v)))
    v0 = 0
    for v in range(HALF, vmin - 1, -1):
        while um[v0] == um[v0 + 1]:
            v0 += 1
        um[v] = v0; v0 += 1
    return um


def ic_angle(img, x, y, um):
    m01 = 0; m10 = 0
    for u in range(-HALF, HALF + 1):
        m10 += u * int(img[y, x + u])
    for v in range(1, HALF + 1):
        d = um[v]; vs = 0
        for u in range(-d, d + 1):
            p = int(img[y + v, x + u]); m = int(img[y - v, x + u])
            vs += p - m; m10 += u * (p + m)
        m01 += v * vs
    return cv2.fastAtan2(float(m01), float(m10))


def extract_stages(image, nfeatures, scaleFactor, nlevels, iniTh, minTh):
    """Returns (pyramid, candidates per level [(x,y,resp)...], distributed per level [(x,y,resp,angle)], blurred)."""
    sf = [f32(1)]
    for i in range(1, nlevels):
        sf.append(f32(sf[-1] * f32(scaleFactor)))
    inv = [f32(1) / s for s in sf]
    factor = f32(1) / f32(scaleFactor)
    nd = f32(nfeatures) * (f32(1) - factor) / (f32(1) - f32(math.pow(float(factor), float(nlevels))))
    per = []; tot = 0
    for l in range(nlevels - 1):
        per.append(int(np.rint(nd))); tot += per[-1]; nd = f32(nd * factor)
    per.append(max(nfeatures - tot, 0))
    pyr = []
    for l in range(nlevels):
        sw = int(np.rint(f32(image.shape[1]) * inv[l])); sh = int(np.rint(f32(image.shape[0]) * inv[l]))
        pyr.append(image.copy() if l == 0 else cv2.resize(pyr[l - 1], (sw, sh), interpolation=cv2.INTER_LINEAR))
    um = umax_table()
    cands = []; dist = []; blurred = []
    for l in range(nlevels):
        img = pyr[l]
        minB = EDGE - 3; maxBX = img.shape[1] - EDGE + 3; maxBY = img.shape[0] - EDGE + 3
        width = f32(maxBX - minB); height = f32(maxBY - minB)
        nCols = int(width / f32(30)); nRows = int(height / f32(30))
        wCell = int(math.ceil(width / f32(nCols))); hCell = int(math.ceil(height / f32(nRows)))
        keys = []
        for i in range(nRows):
            iniY = minB + i * hCell; maxY = iniY + hCell + 6
            if iniY >= maxBY - 3:
                continue
            maxY = min(maxY, maxBY)
            for j in range(nCols):
                iniX = minB + j * wCell; maxX = iniX + wCell + 6
                if iniX >= maxBX - 6:
                    continue
                maxX = min(maxX, maxBX)
                roi = img[iniY:maxY, iniX:maxX]
                k = cv2.FastFeatureDetector_create(iniTh, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16).detect(roi)
                if not k:
                    k = cv2.FastFeatureDetector_create(minTh, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16).detect(roi)
                for kp in k:
                    keys.append((kp.pt[0] + j * wCell, kp.pt[1] + i * hCell, kp.response))
        cands.append(keys)
        d = distribute(keys, minB, maxBX, minB, maxBY, per[l])
        d = [(x + minB, y + minB, r) for (x, y, r) in d]
        dist.append([(x, y, r, ic_angle(img, int(x), int(y), um)) for (x, y, r) in d])
        blurred.append(cv2.GaussianBlur(img.copy(), (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101))
    return pyr, cands, dist, blurred, per
