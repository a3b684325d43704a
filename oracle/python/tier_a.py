"""TEST INFRASTRUCTURE ONLY — Tier A oracle: the reference path restated on REAL OpenCV primitives (cv2 4.13).

OpenCV is the un-vendored dependency that holds most of the hot-path arithmetic; the reference pins
3.3.1, the only runnable build here is the cv2 4.13.0 wheel, so parity is defined against 4.13.
Tier A validates the dependency-free C++ restatement (Tier B, oracle/cpp) stage by stage and generates the golden
vectors under tests/golden/ (oracle/python/make_golden.py).
"""
import math
import numpy as np
import cv2

cv2.setNumThreads(1)
EDGE = 19


def pyramid(img, n_levels=8, scale_factor=1.2):
    """ORBextractor::ComputePyramid, src/ORBextractor.cc:1152-1177 (level images without the border)."""
    sf = np.float32(1.0)
    inv = [np.float32(1.0)]
    for _ in range(1, n_levels):
        sf = np.float32(sf * np.float32(scale_factor))
        inv.append(np.float32(np.float32(1.0) / sf))
    out = [img.copy()]
    h, w = img.shape
    for l in range(1, n_levels):
        lw = int(np.rint(np.float32(w) * inv[l]))
        lh = int(np.rint(np.float32(h) * inv[l]))
        out.append(cv2.resize(out[-1], (lw, lh), interpolation=cv2.INTER_LINEAR))
    return out


def level_candidates(lvl, ini_th=20, min_th=7):
    """Cell loop of ComputeKeyPointsOctTree, src/ORBextractor.cc:769-854, with cv2 FAST run per cell window."""
    f_ini = cv2.FastFeatureDetector_create(ini_th, True)
    f_min = cv2.FastFeatureDetector_create(min_th, True)
    h, w = lvl.shape
    minBX = minBY = EDGE - 3
    maxBX, maxBY = w - EDGE + 3, h - EDGE + 3
    width, height = float(maxBX - minBX), float(maxBY - minBY)
    nCols, nRows = int(width / 30), int(height / 30)
    wCell, hCell = int(math.ceil(width / nCols)), int(math.ceil(height / nRows))
    out = []
    for i in range(nRows):
        iniY = minBY + i * hCell
        maxY = iniY + hCell + 6
        if iniY >= maxBY - 3:
            continue
        maxY = min(maxY, maxBY)
        for j in range(nCols):
            iniX = minBX + j * wCell
            maxX = iniX + wCell + 6
            if iniX >= maxBX - 6:
                continue
            maxX = min(maxX, maxBX)
            sub = np.ascontiguousarray(lvl[iniY:maxY, iniX:maxX])
            k = f_ini.detect(sub)
            if len(k) == 0:
                k = f_min.detect(sub)
            for p in k:
                out.append((p.pt[0] + j * wCell + minBX, p.pt[1] + i * hCell + minBY, p.response))
    return np.array(out, np.float32).reshape(-1, 3)


class _Node:
    __slots__ = ("ulx", "uly", "brx", "bry", "keys", "no_more", "seq")


def distribute_octree(pts, minX, maxX, minY, maxY, N):
    """ORBextractor::DistributeOctTree, src/ORBextractor.cc:537-761, as a python list emulation of std::list.
    pts: (n,3) float32 of x,y (relative to minX,minY),response.  Size ties in the careful phase: newest node first."""
    f32 = np.float32
    nIni = int(math.floor(float(f32(maxX - minX) / f32(maxY - minY)) + 0.5))
    hX = f32(maxX - minX) / f32(nIni)
    seq = [0]

    def mk(ulx, uly, brx, bry):
        n = _Node()
        n.ulx, n.uly, n.brx, n.bry = ulx, uly, brx, bry
        n.keys, n.no_more = [], False
        n.seq = seq[0]
        seq[0] += 1
        return n

    nodes = [mk(int(hX * f32(i)), 0, int(hX * f32(i + 1)), maxY - minY) for i in range(nIni)]
    for k in range(len(pts)):
        nodes[int(f32(pts[k, 0]) / hX)].keys.append(k)
    nodes = [n for n in nodes if n.keys]
    for n in nodes:
        n.no_more = len(n.keys) == 1

    def divide(n):
        halfX = int(math.ceil(float(f32(n.brx - n.ulx) / f32(2))))
        halfY = int(math.ceil(float(f32(n.bry - n.uly) / f32(2))))
        mx, my = n.ulx + halfX, n.uly + halfY
        c = [mk(n.ulx, n.uly, mx, my), mk(mx, n.uly, n.brx, my), mk(n.ulx, my, mx, n.bry), mk(mx, my, n.brx, n.bry)]
        for k in n.keys:
            x, y = pts[k, 0], pts[k, 1]
            if x < mx:
                c[0 if y < my else 2].keys.append(k)
            else:
                c[1 if y < my else 3].keys.append(k)
        c = [q for q in c if q.keys]
        for q in c:
            q.no_more = len(q.keys) == 1
        return c

    finish = False
    while not finish:
        prev = len(nodes)
        created = []
        kept = []
        for n in nodes:
            if n.no_more:
                kept.append(n)
            else:
                created.extend(divide(n))
        nodes = created[::-1] + kept
        expand = [q for q in created if len(q.keys) > 1]
        if len(nodes) >= N or len(nodes) == prev:
            finish = True
        elif len(nodes) + 3 * len(expand) > N:
            while not finish:
                prev = len(nodes)
                order = sorted(expand, key=lambda q: (len(q.keys), q.seq))
                expand = []
                for q in reversed(order):
                    ch = divide(q)
                    nodes.remove(q)
                    nodes = ch[::-1] + nodes
                    expand.extend(c for c in ch if len(c.keys) > 1)
                    if len(nodes) >= N:
                        break
                if len(nodes) >= N or len(nodes) == prev:
                    finish = True
    out = []
    for n in nodes:
        best = n.keys[0]
        for k in n.keys[1:]:
            if pts[k, 2] > pts[best, 2]:
                best = k
        out.append(best)
    return np.array(out, np.int64)


def umax_table():
    HP = 15
    umax = [0] * 16
    vmax = int(math.floor(HP * math.sqrt(2.0) / 2 + 1))
    vmin = int(math.ceil(HP * math.sqrt(2.0) / 2))
    for v in range(vmax + 1):
        umax[v] = int(np.rint(math.sqrt(HP * HP - v * v)))
    v0 = 0
    for v in range(HP, vmin - 1, -1):
        while umax[v0] == umax[v0 + 1]:
            v0 += 1
        umax[v] = v0
        v0 += 1
    return umax


def ic_angle(lvl, x, y, umax):
    """IC_Angle, src/ORBextractor.cc:75-102, with cv2.fastAtan2."""
    m01 = m10 = 0
    row = lvl[y].astype(np.int64)
    for u in range(-15, 16):
        m10 += u * int(row[x + u])
    for v in range(1, 16):
        d = umax[v]
        p = lvl[y + v, x - d:x + d + 1].astype(np.int64)
        m = lvl[y - v, x - d:x + d + 1].astype(np.int64)
        us = np.arange(-d, d + 1)
        m01 += v * int((p - m).sum())
        m10 += int((us * (p + m)).sum())
    return cv2.fastAtan2(float(m01), float(m10))


def orb_descriptor(blur, x, y, angle_deg, pattern):
    """computeOrbDescriptor, src/ORBextractor.cc:106-145 (float32 arithmetic, round-half-even)."""
    f32 = np.float32
    ang = f32(angle_deg) * f32(math.pi / 180.0)
    a, b = f32(math.cos(float(ang))), f32(math.sin(float(ang)))   # double libm rounded to float
    pat = pattern.reshape(256, 4).astype(np.float32)
    r0 = np.rint(pat[:, 0] * b + pat[:, 1] * a).astype(np.int64)
    c0 = np.rint(pat[:, 0] * a - pat[:, 1] * b).astype(np.int64)
    r1 = np.rint(pat[:, 2] * b + pat[:, 3] * a).astype(np.int64)
    c1 = np.rint(pat[:, 2] * a - pat[:, 3] * b).astype(np.int64)
    bits = (blur[y + r0, x + c0] < blur[y + r1, x + c1]).astype(np.uint8)
    return np.packbits(bits.reshape(32, 8), axis=1, bitorder="little").reshape(32)


def load_pattern(path):
    txt = open(path).read()
    vals = [int(t) for line in txt.splitlines() if not line.startswith("//") for t in line.replace(",", " ").split()]
    assert len(vals) == 1024
    return np.array(vals, np.int32)


def lsd_segments(img, scale=1.2):
    """cv::createLineSegmentDetector(0, 1.2, 0.6, 2.0, 22.5, 1.0, 0.6, 1024)->detect, as called from
    LSDDetector_custom.cpp:246-262."""
    lsd = cv2.createLineSegmentDetector(0, scale, 0.6, 2.0, 22.5, 1.0, 0.6, 1024)
    lines = lsd.detect(img)[0]
    if lines is None:
        return np.zeros((0, 4), np.float32)
    return lines.reshape(-1, 4).astype(np.float32)
