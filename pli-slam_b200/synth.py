"""Synthetic EuRoC-shape stereo pairs: there are no datasets offline.

Pure numpy, deterministic in (W, H, seed).  Left image: grey background, 150 axis-aligned rectangles (even index
outlined with thickness 3, odd filled), 40 straight segments of thickness 1-3, a 3x3 Gaussian (sigma 0.8) and
additive uniform noise in [-4, 4].  Right image: the same objects, each shifted left by its own disparity
d in {2..60}, with an independent noise stream.
"""
import numpy as np

N_RECT, N_SEG = 150, 40


def _draw_rect(img, x0, y0, x1, y1, g, filled):
    H, W = img.shape
    xa, xb = sorted((x0, x1))
    ya, yb = sorted((y0, y1))
    if filled:
        img[max(ya, 0):max(yb + 1, 0), max(xa, 0):max(xb + 1, 0)] = g
        return
    t = 1   # thickness 3 = centre +-1
    for (ra, rb, ca, cb) in ((ya - t, ya + t, xa - t, xb + t), (yb - t, yb + t, xa - t, xb + t),
                             (ya - t, yb + t, xa - t, xa + t), (ya - t, yb + t, xb - t, xb + t)):
        ra, ca = max(ra, 0), max(ca, 0)
        rb, cb = max(rb + 1, 0), max(cb + 1, 0)
        if ra < rb and ca < cb:
            img[ra:rb, ca:cb] = g


def _draw_seg(img, x0, y0, x1, y1, g, thick):
    H, W = img.shape
    r = thick / 2.0
    xa, xb = int(np.floor(min(x0, x1) - r - 1)), int(np.ceil(max(x0, x1) + r + 1))
    ya, yb = int(np.floor(min(y0, y1) - r - 1)), int(np.ceil(max(y0, y1) + r + 1))
    xa, ya = max(xa, 0), max(ya, 0)
    xb, yb = min(xb, W - 1), min(yb, H - 1)
    if xa > xb or ya > yb:
        return
    ys, xs = np.mgrid[ya:yb + 1, xa:xb + 1].astype(np.float64)
    dx, dy = x1 - x0, y1 - y0
    L2 = dx * dx + dy * dy
    if L2 == 0:
        return
    t = np.clip(((xs - x0) * dx + (ys - y0) * dy) / L2, 0.0, 1.0)
    d2 = (xs - (x0 + t * dx)) ** 2 + (ys - (y0 + t * dy)) ** 2
    sub = img[ya:yb + 1, xa:xb + 1]
    sub[d2 <= r * r + 1e-9] = g


def _blur3(img):
    k = np.exp(-np.array([-1.0, 0.0, 1.0]) ** 2 / (2 * 0.8 * 0.8))
    k /= k.sum()
    f = img.astype(np.float64)
    p = np.pad(f, 1, mode="reflect")
    f = k[0] * p[1:-1, :-2] + k[1] * p[1:-1, 1:-1] + k[2] * p[1:-1, 2:]
    p = np.pad(f, 1, mode="reflect")
    f = k[0] * p[:-2, 1:-1] + k[1] * p[1:-1, 1:-1] + k[2] * p[2:, 1:-1]
    return f


def synth_pair(W=752, H=480, seed=1):
    rng = np.random.default_rng(seed)
    rx = rng.integers(0, W, size=(N_RECT, 2))
    ry = rng.integers(0, H, size=(N_RECT, 2))
    rg = rng.integers(30, 255, size=N_RECT)
    rd = rng.integers(2, 61, size=N_RECT)
    sx = rng.integers(0, W, size=(N_SEG, 2))
    sy = rng.integers(0, H, size=(N_SEG, 2))
    sg = rng.integers(30, 255, size=N_SEG)
    st = rng.integers(1, 4, size=N_SEG)
    sd = rng.integers(2, 61, size=N_SEG)
    out = []
    for side in (0, 1):
        img = np.full((H, W), 90, np.uint8)
        for i in range(N_RECT):
            d = int(rd[i]) * side
            _draw_rect(img, int(rx[i, 0]) - d, int(ry[i, 0]), int(rx[i, 1]) - d, int(ry[i, 1]), int(rg[i]), i % 2 == 1)
        for i in range(N_SEG):
            d = int(sd[i]) * side
            _draw_seg(img, float(sx[i, 0]) - d, float(sy[i, 0]), float(sx[i, 1]) - d, float(sy[i, 1]), int(sg[i]),
                      int(st[i]))
        f = _blur3(img)
        nrng = np.random.default_rng(seed + 1 if side else seed + 7919)
        f = f + nrng.integers(-4, 5, size=f.shape)
        out.append(np.clip(np.rint(f), 0, 255).astype(np.uint8))
    return out[0], out[1]


def synth_batch(W, H, seeds):
    L = np.empty((len(seeds), H, W), np.uint8)
    R = np.empty((len(seeds), H, W), np.uint8)
    for i, s in enumerate(seeds):
        L[i], R[i] = synth_pair(W, H, int(s))
    return L, R


# EuRoC MAV stereo calibration (Examples/Stereo/Config/EuRoC.yaml:42-84 of the reference: LEFT/RIGHT .K .D .R .P)
EUROC_CALIB = {
    0: dict(K=[458.654, 0.0, 367.215, 0.0, 457.296, 248.375, 0.0, 0.0, 1.0],
            D=[-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0],
            R=[0.999966347530033, -0.001422739138722922, 0.008079580483432283, 0.001365741834644127,
               0.9999741760894847, 0.007055629199258132, -0.008089410156878961, -0.007044357138835809,
               0.9999424675829176],
            P=[435.2046959714599, 0, 367.4517211914062, 0, 0, 435.2046959714599, 252.2008514404297, 0, 0, 0, 1, 0]),
    1: dict(K=[457.587, 0.0, 379.999, 0.0, 456.134, 255.238, 0.0, 0.0, 1],
            D=[-0.28368365, 0.07451284, -0.00010473, -3.555907e-05, 0.0],
            R=[0.9999633526194376, -0.003625811871560086, 0.007755443660172947, 0.003680398547259526,
               0.9999684752771629, -0.007035845251224894, -0.007729688520722713, 0.007064130529506649,
               0.999945173484644],
            P=[435.2046959714599, 0, 367.4517211914062, -47.90639384423901, 0, 435.2046959714599,
               252.2008514404297, 0, 0, 0, 1, 0]),
}


def rectify_maps(W=752, H=480, side=0):
    """Undistort-rectify maps (float32 [H, W] x, y) of one EuRoC camera: the standard pinhole + radial/tangential model
    that cv::initUndistortRectifyMap(K, D, R, P[:3,:3], size, CV_32F) evaluates (stereo_euroc.cc:117-118), in numpy
    double precision.  Used to exercise the rectification kernel with realistic maps without needing cv2."""
    c = EUROC_CALIB[side]
    K = np.array(c["K"], np.float64).reshape(3, 3)
    k1, k2, p1, p2, k3 = c["D"]
    R = np.array(c["R"], np.float64).reshape(3, 3)
    Pn = np.array(c["P"], np.float64).reshape(3, 4)[:3, :3]
    iR = np.linalg.inv(Pn @ R)
    u, v = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    X = iR[0, 0] * u + iR[0, 1] * v + iR[0, 2]
    Y = iR[1, 0] * u + iR[1, 1] * v + iR[1, 2]
    Wd = iR[2, 0] * u + iR[2, 1] * v + iR[2, 2]
    x, y = X / Wd, Y / Wd
    r2 = x * x + y * y
    kr = 1 + ((k3 * r2 + k2) * r2 + k1) * r2
    xd = x * kr + p1 * 2 * x * y + p2 * (r2 + 2 * x * x)
    yd = y * kr + p1 * (r2 + 2 * y * y) + p2 * 2 * x * y
    return (K[0, 0] * xd + K[0, 2]).astype(np.float32), (K[1, 1] * yd + K[1, 2]).astype(np.float32)


def synth_curvy(W=752, H=480, seed=1):
    """A second kind of test content (parity coverage beyond axis-aligned rectangles): smooth shading from a few random
    sinusoids, 60 discs / rings / ellipses with soft and hard edges, 25 thick polylines, a mild vignette and noise in
    [-3, 3].  Curved and slanted edges give region growing long chains of slowly turning angles, T-junctions and many
    small rejected regions.  Pure numpy, deterministic in (W, H, seed); returns one uint8 image."""
    rng = np.random.default_rng(seed)
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float64)
    img = np.full((H, W), 110.0)
    for _ in range(6):
        fx, fy, ph, am = rng.uniform(0.002, 0.02), rng.uniform(0.002, 0.02), rng.uniform(0, 6.28), rng.uniform(5, 25)
        img += am * np.sin(xs * fx * 6.28 + ys * fy * 6.28 + ph)
    for i in range(60):
        cx, cy = rng.uniform(0, W), rng.uniform(0, H)
        a, b = rng.uniform(8, 90), rng.uniform(8, 90)
        th = rng.uniform(0, 3.14159)
        g = rng.uniform(20, 240)
        u = (xs - cx) * np.cos(th) + (ys - cy) * np.sin(th)
        v = -(xs - cx) * np.sin(th) + (ys - cy) * np.cos(th)
        d = np.sqrt((u / a) ** 2 + (v / b) ** 2)
        if i % 3 == 0:        # ring
            m = np.clip(1.5 - np.abs(d - 1.0) * min(a, b) / 2.0, 0, 1)
        elif i % 3 == 1:      # hard disc
            m = (d <= 1.0).astype(np.float64)
        else:                 # soft disc
            m = np.clip((1.0 - d) * min(a, b) / 6.0, 0, 1)
        img = img * (1 - m) + g * m
    for _ in range(25):
        n = int(rng.integers(2, 6))
        pts = np.stack([rng.uniform(0, W, n), rng.uniform(0, H, n)], 1)
        g, r = rng.uniform(10, 250), rng.uniform(0.8, 3.0)
        for (x0, y0), (x1, y1) in zip(pts[:-1], pts[1:]):
            dx, dy = x1 - x0, y1 - y0
            L2 = dx * dx + dy * dy
            if L2 < 1:
                continue
            t = np.clip(((xs - x0) * dx + (ys - y0) * dy) / L2, 0, 1)
            dist = np.sqrt((xs - (x0 + t * dx)) ** 2 + (ys - (y0 + t * dy)) ** 2)
            m = np.clip(r + 0.5 - dist, 0, 1)
            img = img * (1 - m) + g * m
    img *= 1.0 - 0.25 * (((xs - W / 2) / W) ** 2 + ((ys - H / 2) / H) ** 2)
    img += rng.integers(-3, 4, (H, W))
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def synth_vocabulary(k=10, L=4, seed=0, ragged=0.05, stop=0.02):
    """A DBoW2-shaped vocabulary tree for tests (the real ORBvoc cannot travel): k children per inner node, at most L
    levels below the root, node descriptors = the parent's with a level-dependent share of random bit flips (so the
    descent is meaningful), a share `ragged` of the inner nodes cut short into leaves, a share `stop` of the words with
    weight 0 (stopped words).  Returns the flat arrays plf_bow_set_vocabulary takes."""
    rng = np.random.default_rng(seed)
    if ragged == 0:          # complete tree, built level by level (fast enough for the 10^6-word shape of ORBvoc)
        levels = [rng.integers(0, 256, (1, 32), dtype=np.uint8)]
        for l in range(L):
            par = np.repeat(levels[-1], k, axis=0)
            flips = np.packbits(rng.random((par.shape[0], 256)) < (0.5 / (l + 1)), axis=1)
            levels.append(par ^ flips)
        desc = np.concatenate(levels)
        n = desc.shape[0]
        start = np.cumsum([0] + [k ** l for l in range(L + 1)])          # first node id of every level
        first = np.zeros(n, np.int32); count = np.zeros(n, np.int32)
        inner = start[L]                                                   # nodes of levels 0 .. L-1
        count[:inner] = k
        first[:inner] = np.arange(inner, dtype=np.int64) * k               # children lists are contiguous in BFS order
        child = np.arange(1, n, dtype=np.int32)
        word = np.full(n, -1, np.int32); weight = np.zeros(n, np.float64)
        nw = n - inner
        word[inner:] = np.arange(nw, dtype=np.int32)
        weight[inner:] = np.where(rng.random(nw) < stop, 0.0, rng.uniform(0.5, 9.0, nw))
        return dict(levels=L, child_first=first, child_count=count, child=child, desc=desc, word_id=word, weight=weight)
    desc = [rng.integers(0, 256, 32, dtype=np.uint8)]
    first, count, child, level = [0], [0], [], [0]
    queue = [0]
    while queue:
        i = queue.pop(0)
        if level[i] == L or (level[i] > 0 and rng.random() < ragged):
            continue
        first[i], count[i] = len(child), k
        for _ in range(k):
            j = len(desc)
            flips = rng.random(256) < (0.5 / (level[i] + 1))
            d = np.unpackbits(desc[i]) ^ flips.astype(np.uint8)
            desc.append(np.packbits(d))
            first.append(0); count.append(0); level.append(level[i] + 1)
            child.append(j)
            queue.append(j)
    n = len(desc)
    word = np.full(n, -1, np.int32)
    weight = np.zeros(n, np.float64)
    leaves = [i for i in range(n) if count[i] == 0]
    for w, i in enumerate(leaves):
        word[i] = w
        weight[i] = 0.0 if rng.random() < stop else float(rng.uniform(0.5, 9.0))
    return dict(levels=L, child_first=np.array(first, np.int32), child_count=np.array(count, np.int32),
                child=np.array(child, np.int32), desc=np.stack(desc), word_id=word, weight=weight)
