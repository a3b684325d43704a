"""Host-side logic on CPU: parameter defaults against the reference's EuRoC.yaml, scale tables / quotas, Hamming
distance, matcher edge cases (empty, one-row train set, ties), reference UB rules, stream sharding incl. a
world_size-2 gloo run."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_default_params_match_euroc_yaml(plf, oracle):
    p = oracle.default_params()
    # Examples/Stereo/Config/EuRoC.yaml:91-104,130-175 and src/Config.cpp:100-108 of the reference
    assert (p.width, p.height, p.n_features, p.n_levels, p.ini_th_fast, p.min_th_fast) == (752, 480, 1200, 8, 20, 7)
    assert abs(p.scale_factor - 1.2) < 1e-7
    assert (p.lsd_nfeatures, p.lsd_refine, p.lsd_n_bins, p.matching_s_ws, p.best_lr_matches) == (500, 0, 1024, 10, 1)
    assert (p.lsd_scale, p.lsd_sigma_scale, p.lsd_quant, p.lsd_ang_th) == (1.2, 0.6, 2.0, 22.5)
    assert (p.min_ratio_12_l, p.line_sim_th, p.min_disp, p.line_horiz_th) == (0.9, 0.75, 1.0, 0.1)
    assert (p.stereo_overlap_th, p.ls_min_disp_ratio, p.min_line_length) == (0.75, 0.7, 0.025)
    assert abs(p.fx - 435.2047) < 1e-3 and abs(p.bf - 47.90639) < 1e-4


def test_scale_tables_and_quotas(plf, oracle):
    f = plf.Frontend(oracle)
    s, inv, s2, inv2, n = f.scale_tables()
    assert list(n) == [261, 217, 181, 151, 126, 105, 87, 72]
    assert s[0] == 1.0 and np.allclose(s[1:] / s[:-1], 1.2, rtol=1e-6)
    assert np.array_equal(inv, np.float32(1.0) / s) and np.array_equal(s2, s * s)
    f2 = plf.Frontend(oracle, n_features=2000)
    assert list(f2.scale_tables()[4]) == [434, 362, 302, 251, 209, 175, 145, 122]


def test_unsupported_parameters_fail(plf, oracle):
    with pytest.raises(plf.PlfError):
        plf.Frontend(oracle, lsd_refine=2)       # ADVANCED (NFA rectangle improvement) is not built
    with pytest.raises(plf.PlfError):
        plf.Frontend(oracle, width=16, height=16)


def test_matchers_edge_cases(plf, oracle):
    f = plf.Frontend(oracle)
    rng = np.random.default_rng(3)
    d1 = rng.integers(0, 256, (40, 32), dtype=np.uint8)
    # identical sets: every row matches itself (distance 0 < d1*nnr)
    n, m = f.match_nnr(d1, d1, 0.9)
    assert n == 40 and np.array_equal(m, np.arange(40))
    # train set with a single row: no second neighbour -> no matches (declared rule, LineMatcher.cpp:152)
    n, m = f.match_nnr(d1, d1[:1], 0.9)
    assert n == 0 and (m == -1).all()
    # empty query
    n, m = f.match_nnr(np.zeros((0, 32), np.uint8), d1, 0.9)
    assert n == 0 and len(m) == 0
    # duplicated best neighbour: tie -> ratio test fails for nnr < 1
    d2 = np.concatenate([d1[:1], d1[:1], d1[5:]])
    n, m = f.match_nnr(d1[:1], d2, 0.9)
    assert n == 0
    # mutual-best drops one-sided matches
    n1, m1 = f.match(d1, d1[::-1].copy(), 0.9, True)
    assert n1 == 40 and np.array_equal(m1, np.arange(40)[::-1])


def test_empty_frame_rules(plf, oracle):
    """A textureless pair: no keypoints, no lines, and the matchers must return cleanly (Frame.cc:146-149)."""
    f = plf.Frontend(oracle)
    flat = np.full((1, 480, 752), 127, np.uint8)
    r = f.frontend_batch(flat, flat)
    assert r.n_kp_left[0] == 0 and r.n_kl_left[0] == 0


def test_lapping_area_row_order(plf, oracle, pair1):
    """ORBextractor::operator() writes rows inside vLappingArea back to front (src/ORBextractor.cc:1135-1144)."""
    L, _ = pair1
    f = plf.Frontend(oracle)
    mono0, k0, d0 = f.orb_extract(0, L, (0, 0))
    mono1, k1, d1 = f.orb_extract(0, L, (0, 1000))      # monocular ctor: everything is "lapping"
    assert mono0 == len(k0) and mono1 == 0
    assert np.array_equal(k1[::-1], k0) and np.array_equal(d1[::-1], d0)
    mono2, k2, d2 = f.orb_extract(0, L, (300, 500))
    inside = (k0["x"] >= 300) & (k0["x"] <= 500)
    assert mono2 == int((~inside).sum())
    assert np.array_equal(k2[:mono2], k0[~inside]) and np.array_equal(k2[mono2:][::-1], k0[inside])


def test_shard_streams():
    sys.path.insert(0, ROOT)
    import bench
    for world in (1, 2, 4, 8):
        got = sorted(s for r in range(world) for s in bench.shard_streams(512, world, r))
        assert got == list(range(512))
        assert all(len(bench.shard_streams(512, world, r)) == 512 // world for r in range(world))
    assert bench.stream_seed(0, 0) == 10_000 and bench.stream_seed(511, 63) == 5_120_063


GLOO_WORKER = r"""
import os, sys, json
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
import bench, plf
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
streams = bench.shard_streams(4, world, rank)
orc = plf.load_oracle()
f = plf.Frontend(orc, width=752, height=480, max_batch=len(streams), has_lines=0, n_features=300)
L, R = plf.synth_batch(752, 480, [bench.stream_seed(s, 0) for s in streams])
r = f.frontend_batch(L, R)
cnt = torch.tensor([float(r.n_kp_left[:len(streams)].sum()), float(len(streams))])
dist.all_reduce(cnt)                       # bookkeeping only: the data path has no collective
t = torch.tensor([1.0 + rank]); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"kps": cnt[0].item(), "pairs": cnt[1].item(), "tmax": t.item(), "streams": streams}))
dist.destroy_process_group()
"""


def test_two_rank_gloo_sharding(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(GLOO_WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["pairs"] == 4 and res["tmax"] == 2.0 and res["streams"] == [0, 2] and res["kps"] > 4 * 250


def test_feature_grid_and_area_lookup_oracle(plf, oracle):
    """Frame::AssignFeaturesToGrid / GetFeaturesInArea (src/Frame.cc:451-482, 774-855): the oracle's CSR grid and the
    header's lookup against a direct Python restatement of the reference loops."""
    import math
    W, H = 752, 480
    L, R = plf.synth_pair(W, H, 3)
    o = plf.Frontend(oracle, max_batch=1)
    res = o.frontend_batch(L[None], R[None])
    n = int(res.n_kp_left[0])
    kps = res.kp_left[0, :n]
    st, ix = o.feature_grid(0, 1)
    st, ix = st[0], ix[0]
    invw, invh = np.float32(64) / np.float32(W), np.float32(48) / np.float32(H)

    def rnd(v):      # C round(): half away from zero
        return int(math.floor(abs(v) + 0.5)) * (1 if v >= 0 else -1)
    grid = {}
    for i in range(n):
        px, py = rnd(float(np.float32(kps["x"][i]) * invw)), rnd(float(np.float32(kps["y"][i]) * invh))
        if 0 <= px < 64 and 0 <= py < 48:
            grid.setdefault(px * 48 + py, []).append(i)
    assert st[-1] == sum(len(v) for v in grid.values()) and st[-1] >= n - 5
    for c in range(64 * 48):
        assert list(ix[st[c]:st[c + 1]]) == grid.get(c, []), c
    rng = np.random.default_rng(1)
    for _ in range(60):
        x, y = np.float32(rng.uniform(-30, W + 30)), np.float32(rng.uniform(-30, H + 30))
        r = np.float32(rng.choice([3.0, 10.0, 17.5, 40.0]))
        lo, hi = [(-1, -1), (0, 2), (1, -1), (3, 7)][int(rng.integers(0, 4))]
        want = []
        x0 = max(0, math.floor(float((x - r) * invw))); x1 = min(63, math.ceil(float((x + r) * invw)))
        y0 = max(0, math.floor(float((y - r) * invh))); y1 = min(47, math.ceil(float((y + r) * invh)))
        if x0 < 64 and x1 >= 0 and y0 < 48 and y1 >= 0:
            for cx in range(x0, x1 + 1):
                for cy in range(y0, y1 + 1):
                    for i in grid.get(cx * 48 + cy, []):
                        if lo > 0 or hi >= 0:
                            if kps["octave"][i] < lo or (hi >= 0 and kps["octave"][i] > hi):
                                continue
                        if abs(np.float32(kps["x"][i]) - x) < r and abs(np.float32(kps["y"][i]) - y) < r:
                            want.append(i)
        got = o.features_in_area(kps, st, ix, x, y, r, lo, hi)
        assert list(got) == want


def test_backproject_oracle_against_cv2_gemm_and_numpy(plf, oracle):
    """Frame::UnprojectStereo / Frame::backProjection (src/Frame.cc:1332-1358): the oracle against cv2.gemm (the cv::Mat
    expression mRwc*x3Dc+mOw) for the points and against the double expression written out in numpy for the lines."""
    cv2 = pytest.importorskip("cv2")
    W, H = 752, 480
    L, R = plf.synth_pair(W, H, 3)
    o = plf.Frontend(oracle, max_batch=1)
    res = o.frontend_batch(L[None], R[None])
    n, nl = int(res.n_kp_left[0]), int(res.n_kl_left[0])
    rng = np.random.default_rng(2)
    A = rng.normal(size=(3, 3)); Q, _ = np.linalg.qr(A)
    Rwc = Q.astype(np.float32); Ow = rng.normal(size=3).astype(np.float32) * 2
    fx, fy, cx, cy = np.float32(o.params.fx), np.float32(435.2047), np.float32(367.4517), np.float32(252.2009)
    x3d, l3d = o.backproject(Rwc[None], Ow[None], fy, cx, cy)
    invfx, invfy = np.float32(1) / fx, np.float32(1) / fy
    kps, depth = res.kp_left[0], res.depth[0]
    hits = 0
    for i in range(n):
        z = np.float32(depth[i])
        if not z > 0:
            assert not x3d[0, i].any()
            continue
        x = np.float32(np.float32(np.float32(kps["x"][i]) - cx) * z) * invfx
        y = np.float32(np.float32(np.float32(kps["y"][i]) - cy) * z) * invfy
        want = cv2.gemm(Rwc, np.array([[x], [y], [z]], np.float32), 1.0, Ow.reshape(3, 1), 1.0).ravel()
        assert np.array_equal(x3d[0, i], want), i
        hits += 1
    assert hits > 100 and not x3d[0, n:].any()
    mb = np.float32(o.params.bf) / np.float32(o.params.fx)
    kls, disp = res.kl_left[0], res.disp_se[0]
    Rd, Od = Rwc.astype(np.float64), Ow.astype(np.float64)
    hits = 0
    for i in range(nl):
        d0, d1 = disp[i]
        if not (d0 > 0 and d1 > 0):
            assert not l3d[0, i].any()
            continue
        for e, (u, v, d) in enumerate(((kls["startPointX"][i], kls["startPointY"][i], d0), (kls["endPointX"][i], kls["endPointY"][i], d1))):
            bd = np.float64(mb) / np.float64(d)
            P = np.array([bd * (np.float64(u) - np.float64(cx)), bd * (np.float64(v) - np.float64(cy)), bd * np.float64(fx)])
            want = ((Rd[:, 0] * P[0] + Rd[:, 1] * P[1]) + Rd[:, 2] * P[2]) + Od
            assert np.array_equal(l3d[0, i, 3 * e:3 * e + 3], want), (i, e)
        hits += 1
    assert hits > 20


def _bow_python(voc, descs, levelsup):
    """DBoW2 transform(features, BowVector, FeatureVector, levelsup) restated directly (TemplatedVocabulary.h:1139-1270,
    BowVector.cpp:34-84): TF_IDF weighting, L1 norm."""
    cf, cc, ch = voc["child_first"], voc["child_count"], voc["child"]
    nd = np.unpackbits(voc["desc"], axis=1)
    words, weights, nodes = [], [], []
    bow, fv = {}, {}
    for i, f in enumerate(np.unpackbits(descs, axis=1)):
        node, level, nid = 0, 0, 0
        while True:
            level += 1
            kids = ch[cf[node]:cf[node] + cc[node]]
            d = (nd[kids] ^ f).sum(axis=1)
            node = int(kids[int(np.argmin(d))])          # argmin returns the first minimum = strict '<' scan
            if level == voc["levels"] - levelsup:
                nid = node
            if cc[node] == 0:
                break
        w = float(voc["weight"][node])
        words.append(int(voc["word_id"][node])); weights.append(w); nodes.append(nid)
        if w > 0:
            bow[words[-1]] = bow.get(words[-1], 0.0) + w
            fv.setdefault(nid, []).append(i)
    norm = 0.0
    for k in sorted(bow):
        norm += abs(bow[k])
    if norm > 0:
        bow = {k: v / norm for k, v in bow.items()}
    return words, weights, nodes, bow, fv


@pytest.mark.parametrize("which,levelsup", [(0, 2), (1, 4), (0, 0)])
def test_bow_oracle_against_python(plf, oracle, which, levelsup):
    """Frame::ComputeBoW's per-feature descent and the BowVector / FeatureVector construction on a synthetic DBoW2-shaped
    vocabulary (ragged tree, some stopped words): oracle and header inline against the direct restatement."""
    voc = plf.synth_vocabulary(k=10, L=4, seed=3 + which) if levelsup else plf.synth_vocabulary(k=6, L=3, seed=9, ragged=0.0, stop=0.1)
    L, R = plf.synth_pair(752, 480, 5)
    o = plf.Frontend(oracle, max_batch=1)
    res = o.frontend_batch(L[None], R[None])
    n = int(res.n_kl_left[0]) if which else int(res.n_kp_left[0])
    descs = (res.ldesc_left if which else res.desc_left)[0, :n]
    o.bow_set_vocabulary(which, voc)
    w, v, nd = o.bow_transform(which, 1, 0, levelsup)
    words, weights, nodes, bow, fv = _bow_python(voc, descs, levelsup)
    assert list(w[0, :n]) == words and list(v[0, :n]) == weights and list(nd[0, :n]) == nodes
    assert np.all(w[0, n:] == -1) and not v[0, n:].any()
    bw, bv, fvec = o.bow_build(w[0, :n], v[0, :n], nd[0, :n])
    assert list(bw) == sorted(bow) and [bow[k] for k in sorted(bow)] == list(bv)
    assert fvec == fv and abs(bv.sum() - 1.0) < 1e-12 and len(bw) > 50
    if voc["levels"] - levelsup <= 0:
        assert set(fvec) == {0}


def _proj_queries(plf, res, b, rng, n_extra=150):
    """Map points as SearchByProjection sees them: every left keypoint of frame b re-observed with a small projection
    error, a possibly different predicted level and a few flipped descriptor bits, plus some unrelated points; shuffled."""
    n = int(res.n_kp_left[b])
    kps, desc, ur = res.kp_left[b, :n], res.desc_left[b, :n], res.u_right[b, :n]
    m = n + n_extra
    q = np.zeros(m, plf.PROJ_QUERY_DT)
    src = np.concatenate([np.arange(n), rng.integers(0, n, n_extra)])
    q["proj_x"] = kps["x"][src] + rng.normal(0, 1.5, m).astype(np.float32)
    q["proj_y"] = kps["y"][src] + rng.normal(0, 1.5, m).astype(np.float32)
    q["proj_xr"] = np.where(ur[src] > 0, ur[src] + rng.normal(0, 1.0, m), -1).astype(np.float32)
    q["view_cos"] = rng.choice(np.array([0.9995, 0.99, 0.7], np.float32), m)
    q["level"] = np.clip(kps["octave"][src] + rng.integers(-1, 2, m), 0, 7)
    q["skip"] = (rng.random(m) < 0.05).astype(np.int32)
    d = desc[src].copy()
    flips = rng.random((m, 256)) < 0.04
    d ^= np.packbits(flips, axis=1)
    d[n:] = rng.integers(0, 256, (n_extra, 32), dtype=np.uint8)
    q["desc"] = d
    q["proj_x"][n:] = rng.uniform(-20, 780, n_extra); q["proj_y"][n:] = rng.uniform(-20, 500, n_extra)
    return q[rng.permutation(m)]


def _search_by_projection_python(q, kps, desc, ur, scale, occupied, th, nn_ratio, th_high, W=752, H=480):
    """ORBmatcher::SearchByProjection(F, vpMapPoints, th) restated loop for loop (src/ORBmatcher.cc:44-130, Nleft == -1)."""
    import math
    f32 = np.float32
    invw, invh = f32(64) / f32(W), f32(48) / f32(H)

    def rnd(v):
        return int(math.floor(abs(v) + 0.5)) * (1 if v >= 0 else -1)
    grid = {}
    for i in range(len(kps)):
        px, py = rnd(float(f32(kps["x"][i]) * invw)), rnd(float(f32(kps["y"][i]) * invh))
        if 0 <= px < 64 and 0 <= py < 48:
            grid.setdefault(px * 48 + py, []).append(i)
    bits = np.unpackbits(desc, axis=1)
    match = np.full(len(q), -1, np.int32)
    nm = 0
    for i, qq in enumerate(q):
        if qq["skip"]:
            continue
        r = f32(2.5) if qq["view_cos"] > f32(0.998) else f32(4.0)
        if th != 1.0:
            r = f32(r * f32(th))
        rad = f32(r * scale[qq["level"]])
        lo, hi = qq["level"] - 1, qq["level"]
        x, y = f32(qq["proj_x"]), f32(qq["proj_y"])
        x0 = max(0, math.floor(float(f32(f32(x - rad) * invw)))); x1 = min(63, math.ceil(float(f32(f32(x + rad) * invw))))
        y0 = max(0, math.floor(float(f32(f32(y - rad) * invh)))); y1 = min(47, math.ceil(float(f32(f32(y + rad) * invh))))
        if not (x0 < 64 and x1 >= 0 and y0 < 48 and y1 >= 0):
            continue
        qb = np.unpackbits(qq["desc"])
        best, bl, best2, bl2, bidx = 256, -1, 256, -1, -1
        for cx in range(x0, x1 + 1):
            for cy in range(y0, y1 + 1):
                for idx in grid.get(cx * 48 + cy, []):
                    o = int(kps["octave"][idx])
                    if o < lo or o > hi:
                        continue
                    if not (abs(f32(kps["x"][idx]) - x) < rad and abs(f32(kps["y"][idx]) - y) < rad):
                        continue
                    if occupied[idx]:
                        continue
                    if ur[idx] > 0 and abs(f32(qq["proj_xr"]) - f32(ur[idx])) > rad:
                        continue
                    dist = int((bits[idx] ^ qb).sum())
                    if dist < best:
                        best2, best, bl2, bl, bidx = best, dist, bl, o, idx
                    elif dist < best2:
                        bl2, best2 = o, dist
        if best <= th_high:
            if bl == bl2 and f32(best) > f32(nn_ratio) * f32(best2):
                continue
            match[i] = bidx
            occupied[bidx] = 1
            nm += 1
    return match, nm


@pytest.mark.parametrize("th", [1.0, 3.0])
def test_search_by_projection_oracle_against_python(plf, oracle, th):
    """The oracle's SearchByProjection against the direct Python restatement, including pre-occupied features and the
    order dependence (a feature taken by an earlier map point is skipped by the later ones)."""
    L, R = plf.synth_pair(752, 480, 8)
    o = plf.Frontend(oracle, max_batch=1)
    res = o.frontend_batch(L[None], R[None])
    n = int(res.n_kp_left[0])
    rng = np.random.default_rng(11)
    q = _proj_queries(plf, res, 0, rng, 60)[:500]
    scale = o.scale_tables()[0] if hasattr(o, "scale_tables") else np.float32(1.2) ** np.arange(8, dtype=np.float32)
    occ0 = (rng.random(n) < 0.1).astype(np.uint8)
    occ_a, occ_b = occ0.copy(), occ0.copy()
    got, nm = o.search_by_projection(q, occ_a, th=th)
    sc = np.ones(8, np.float32)
    for i in range(1, 8):
        sc[i] = np.float32(sc[i - 1] * np.float32(1.2))
    want, nm2 = _search_by_projection_python(q, res.kp_left[0, :n], res.desc_left[0, :n], res.u_right[0, :n], sc, occ_b, th, 0.8, 100)
    assert nm == nm2 and np.array_equal(got, want) and np.array_equal(occ_a, occ_b)
    assert nm > 100 and len(np.unique(got[got >= 0])) == nm          # a feature is given to one map point only


def _frame_queries(plf, res, b, rng, mode="around"):
    """LastFrame map points projected into the current frame b: every keypoint re-observed with a projection error, the
    GetFeaturesInArea level arguments of one of the three motion cases, a rotated angle, some points without
    observations (temporal points), some skipped, a block of unrelated points at the end; shuffled."""
    n = int(res.n_kp_left[b])
    kps, desc, ur = res.kp_left[b, :n], res.desc_left[b, :n], res.u_right[b, :n]
    m = n + 120
    q = np.zeros(m, plf.FRAME_QUERY_DT)
    src = np.concatenate([np.arange(n), rng.integers(0, n, 120)])
    q["u"] = kps["x"][src] + rng.normal(0, 2.0, m).astype(np.float32)
    q["v"] = kps["y"][src] + rng.normal(0, 2.0, m).astype(np.float32)
    q["ur"] = np.where(ur[src] > 0, ur[src] + rng.normal(0, 1.5, m), q["u"] - 20).astype(np.float32)
    oct_ = kps["octave"][src].astype(np.int32)
    sc = np.ones(8, np.float32)
    for i in range(1, 8):
        sc[i] = np.float32(sc[i - 1] * np.float32(1.2))
    q["radius"] = np.float32(7.0) * sc[oct_]
    if mode == "forward":
        q["min_level"], q["max_level"] = oct_, -1
    elif mode == "backward":
        q["min_level"], q["max_level"] = 0, oct_
    else:
        q["min_level"], q["max_level"] = oct_ - 1, oct_ + 1
    q["skip"] = (rng.random(m) < 0.05).astype(np.int32)
    q["has_observations"] = (rng.random(m) < 0.7).astype(np.int32)
    rot = rng.choice(np.array([3.0, 5.0, 8.0, 100.0, 200.0], np.float32), m, p=[0.45, 0.3, 0.15, 0.05, 0.05])
    q["angle"] = np.mod(kps["angle"][src] + rot + rng.normal(0, 1.0, m), 360).astype(np.float32)
    d = desc[src].copy()
    d ^= np.packbits(rng.random((m, 256)) < 0.05, axis=1)
    d[n:] = rng.integers(0, 256, (120, 32), dtype=np.uint8)
    q["desc"] = d
    return q[rng.permutation(m)]


def _search_frame_python(q, kps, desc, ur, occupied, th_high, check, W=752, H=480, stereo=True, holder_by_obs=True):
    """src/ORBmatcher.cc:2229-2317 restated loop for loop, incl. ComputeThreeMaxima (:2449-2490)."""
    import math
    f32 = np.float32
    invw, invh = f32(64) / f32(W), f32(48) / f32(H)

    def rnd(v):
        return int(math.floor(abs(v) + 0.5)) * (1 if v >= 0 else -1)
    grid = {}
    for i in range(len(kps)):
        px, py = rnd(float(f32(kps["x"][i]) * invw)), rnd(float(f32(kps["y"][i]) * invh))
        if 0 <= px < 64 and 0 <= py < 48:
            grid.setdefault(px * 48 + py, []).append(i)
    bits = np.unpackbits(desc, axis=1)
    N = len(kps)
    fq, m12 = np.full(N, -1, np.int32), np.full(N, -1, np.int32)
    hist = [[] for _ in range(30)]
    factor = f32(1.0) / f32(30)
    nm = 0
    for i, qq in enumerate(q):
        if qq["skip"]:
            continue
        u, v, rad = f32(qq["u"]), f32(qq["v"]), f32(qq["radius"])
        lo, hi = int(qq["min_level"]), int(qq["max_level"])
        x0 = max(0, math.floor(float(f32(f32(u - rad) * invw)))); x1 = min(63, math.ceil(float(f32(f32(u + rad) * invw))))
        y0 = max(0, math.floor(float(f32(f32(v - rad) * invh)))); y1 = min(47, math.ceil(float(f32(f32(v + rad) * invh))))
        if not (x0 < 64 and x1 >= 0 and y0 < 48 and y1 >= 0):
            continue
        qb = np.unpackbits(qq["desc"])
        best, bidx = 256, -1
        check_lv = lo > 0 or hi >= 0
        for cx in range(x0, x1 + 1):
            for cy in range(y0, y1 + 1):
                for idx in grid.get(cx * 48 + cy, []):
                    o = int(kps["octave"][idx])
                    if check_lv and (o < lo or (hi >= 0 and o > hi)):
                        continue
                    if not (abs(f32(kps["x"][idx]) - u) < rad and abs(f32(kps["y"][idx]) - v) < rad):
                        continue
                    if occupied[idx]:
                        continue
                    if stereo and ur[idx] > 0 and abs(f32(qq["ur"]) - f32(ur[idx])) > rad:
                        continue
                    dist = int((bits[idx] ^ qb).sum())
                    if dist < best:
                        best, bidx = dist, idx
        if bidx >= 0 and best <= th_high:
            fq[bidx] = i
            occupied[bidx] = 1 if (qq["has_observations"] or not holder_by_obs) else 0
            nm += 1
            if m12[bidx] < 0:
                m12[bidx] = i
            if check:
                rot = f32(f32(qq["angle"]) - f32(kps["angle"][bidx]))
                if rot < 0:
                    rot = f32(rot + f32(360))
                b = rnd(float(f32(rot * factor)))
                if b == 30:
                    b = 0
                hist[b].append(bidx)
    if check:
        i1 = i2 = i3 = -1
        m1 = m2 = m3 = 0
        for i in range(30):
            s = len(hist[i])
            if s > m1:
                m3, m2, m1, i3, i2, i1 = m2, m1, s, i2, i1, i
            elif s > m2:
                m3, m2, i3, i2 = m2, s, i2, i
            elif s > m3:
                m3, i3 = s, i
        if m2 < f32(0.1) * f32(m1):
            i2 = i3 = -1
        elif m3 < f32(0.1) * f32(m1):
            i3 = -1
        for i in range(30):
            if i not in (i1, i2, i3):
                for f in hist[i]:
                    fq[f] = -1; occupied[f] = 0; nm -= 1; m12[f] = -1
    return fq, m12, nm


@pytest.mark.parametrize("mode,check", [("around", True), ("forward", True), ("backward", False)])
def test_search_by_projection_frame_oracle_against_python(plf, oracle, mode, check):
    """The frame-to-frame overload (window search, temporal points that do not block a feature, rotation histogram with
    its three maxima) — oracle against the direct Python restatement."""
    L, R = plf.synth_pair(752, 480, 8)
    o = plf.Frontend(oracle, max_batch=1)
    res = o.frontend_batch(L[None], R[None])
    n = int(res.n_kp_left[0])
    rng = np.random.default_rng(13)
    q = _frame_queries(plf, res, 0, rng, mode)[:600]
    occ0 = (rng.random(n) < 0.08).astype(np.uint8)
    oa, ob = occ0.copy(), occ0.copy()
    fq, m12, nm = o.search_by_projection_frame(q, oa, 100, check)
    wfq, wm12, wnm = _search_frame_python(q, res.kp_left[0, :n], res.desc_left[0, :n], res.u_right[0, :n], ob, 100, check)
    assert nm == wnm and np.array_equal(fq, wfq) and np.array_equal(m12, wm12) and np.array_equal(oa, ob)
    assert nm > 150
    if check:
        assert (fq >= 0).sum() < 600 - 30          # the histogram removed the matches with the odd rotations


@pytest.mark.parametrize("check", [True, False])
def test_search_by_projection_reloc_oracle_against_python(plf, oracle, check):
    """The relocalisation overload (src/ORBmatcher.cc:2325-2447): no stereo check, any holder blocks a feature."""
    L, R = plf.synth_pair(752, 480, 9)
    o = plf.Frontend(oracle, max_batch=1)
    res = o.frontend_batch(L[None], R[None])
    n = int(res.n_kp_left[0])
    rng = np.random.default_rng(21)
    q = _frame_queries(plf, res, 0, rng, "around")[:700]
    q["has_observations"] = 0                      # must not matter for this overload
    occ0 = (rng.random(n) < 0.1).astype(np.uint8)
    oa, ob = occ0.copy(), occ0.copy()
    fq, nm = o.search_by_projection_reloc(q, oa, 64, check)
    wfq, _, wnm = _search_frame_python(q, res.kp_left[0, :n], res.desc_left[0, :n], res.u_right[0, :n], ob, 64, check,
                                       stereo=False, holder_by_obs=False)
    assert nm == wnm and np.array_equal(fq, wfq) and np.array_equal(oa, ob)
    assert nm > 150 and not np.any((fq >= 0) & (occ0 != 0))


@pytest.mark.parametrize("ratio", [1.0, 0.64])
def test_search_by_projection_loop_oracle_against_python(plf, oracle, ratio):
    """The loop-closing overloads (src/ORBmatcher.cc:473-704): levels [pred - 1, pred], float threshold TH_LOW * ratio."""
    L, R = plf.synth_pair(752, 480, 10)
    o = plf.Frontend(oracle, max_batch=1)
    res = o.frontend_batch(L[None], R[None])
    n = int(res.n_kp_left[0])
    rng = np.random.default_rng(22)
    q = _frame_queries(plf, res, 0, rng, "backward")[:700]
    q["min_level"] = q["max_level"] - 1
    occ0 = (rng.random(n) < 0.1).astype(np.uint8)
    oa, ob = occ0.copy(), occ0.copy()
    fq, nm = o.search_by_projection_loop(q, oa, 50, ratio)
    wfq, _, wnm = _search_frame_python(q, res.kp_left[0, :n], res.desc_left[0, :n], res.u_right[0, :n], ob,
                                       np.float32(50) * np.float32(ratio), False, stereo=False, holder_by_obs=False)
    assert nm == wnm and np.array_equal(fq, wfq) and np.array_equal(oa, ob)
    assert nm > 100


def _bow_case(plf, res, rng, b=0, n_nodes=60):
    """A keyframe made of the frame's own features (noisy descriptors, rotated angles) plus strangers, with DBoW2-like
    node ids on both sides (some features in no node: stopped words)."""
    n = int(res.n_kp_left[b])
    kps, desc = res.kp_left[b, :n], res.desc_left[b, :n]
    m = 900
    src = rng.integers(0, n, m)
    kf_desc = desc[src].copy()
    kf_desc ^= np.packbits(rng.random((m, 256)) < 0.04, axis=1)
    kf_desc[-100:] = rng.integers(0, 256, (100, 32), dtype=np.uint8)
    f_node = rng.integers(0, n_nodes, n).astype(np.int32) * 7 + 100
    kf_node = f_node[src].copy()
    kf_node[rng.random(m) < 0.1] = rng.integers(0, n_nodes, int((rng.random(m) < 0.1).sum()) or 1)[0] * 7 + 100
    f_node[rng.random(n) < 0.05] = -1
    kf_node[rng.random(m) < 0.05] = -1
    rot = rng.choice(np.array([4.0, 9.0, 150.0], np.float32), m, p=[0.6, 0.3, 0.1])
    kf_angle = np.mod(kps["angle"][src] + rot + rng.normal(0, 1.0, m), 360).astype(np.float32)
    kf_valid = (rng.random(m) < 0.85).astype(np.uint8)
    return kf_desc, kf_angle, kf_node.astype(np.int32), kf_valid, f_node


def _search_bow_python(kf_desc, kf_angle, kf_node, kf_valid, f_node, kps, desc, th_low, nn_ratio, check):
    """src/ORBmatcher.cc:269-471 (F.Nleft == -1) restated: merge join of the two FeatureVectors, best / second best,
    TH_LOW and mfNNratio, rotation histogram."""
    import math
    f32 = np.float32

    def rnd(v):
        return int(math.floor(abs(v) + 0.5)) * (1 if v >= 0 else -1)
    fv_kf, fv_f = {}, {}
    for i, nd in enumerate(kf_node):
        if nd >= 0:
            fv_kf.setdefault(int(nd), []).append(i)
    for i, nd in enumerate(f_node[:len(kps)]):
        if nd >= 0:
            fv_f.setdefault(int(nd), []).append(i)
    bits_f = np.unpackbits(desc, axis=1)
    bits_kf = np.unpackbits(kf_desc, axis=1)
    match = np.full(len(f_node), -1, np.int32)
    hist = [[] for _ in range(30)]
    factor = f32(1.0) / f32(30)
    nm = 0
    for nd in sorted(set(fv_kf) & set(fv_f)):
        for ikf in fv_kf[nd]:
            if not kf_valid[ikf]:
                continue
            b1, b2, bidx = 256, 256, -1
            for f in fv_f[nd]:
                if match[f] >= 0:
                    continue
                d = int((bits_kf[ikf] ^ bits_f[f]).sum())
                if d < b1:
                    b2, b1, bidx = b1, d, f
                elif d < b2:
                    b2 = d
            if b1 <= th_low and f32(b1) < f32(nn_ratio) * f32(b2):
                match[bidx] = ikf
                nm += 1
                if check:
                    rot = f32(f32(kf_angle[ikf]) - f32(kps["angle"][bidx]))
                    if rot < 0:
                        rot = f32(rot + f32(360))
                    b = rnd(float(f32(rot * factor)))
                    if b == 30:
                        b = 0
                    hist[b].append(bidx)
    if check:
        i1 = i2 = i3 = -1
        m1 = m2 = m3 = 0
        for i in range(30):
            s = len(hist[i])
            if s > m1:
                m3, m2, m1, i3, i2, i1 = m2, m1, s, i2, i1, i
            elif s > m2:
                m3, m2, i3, i2 = m2, s, i2, i
            elif s > m3:
                m3, i3 = s, i
        if m2 < f32(0.1) * f32(m1):
            i2 = i3 = -1
        elif m3 < f32(0.1) * f32(m1):
            i3 = -1
        for i in range(30):
            if i not in (i1, i2, i3):
                for f in hist[i]:
                    match[f] = -1; nm -= 1
    return match, nm


@pytest.mark.parametrize("check,ratio", [(True, 0.7), (False, 0.9)])
def test_search_by_bow_oracle_against_python(plf, oracle, check, ratio):
    L, R = plf.synth_pair(752, 480, 11)
    o = plf.Frontend(oracle, max_batch=1)
    res = o.frontend_batch(L[None], R[None])
    n = int(res.n_kp_left[0])
    kf_desc, kf_angle, kf_node, kf_valid, f_node = _bow_case(plf, res, np.random.default_rng(31))
    m, nm = o.search_by_bow(kf_desc, kf_angle, kf_node, kf_valid, f_node, 50, ratio, check)
    wm, wnm = _search_bow_python(kf_desc, kf_angle, kf_node, kf_valid, f_node, res.kp_left[0, :n], res.desc_left[0, :n], 50, ratio, check)
    assert nm == wnm and np.array_equal(m, wm)
    assert nm > 200
    e, enm = o.search_by_bow(kf_desc[:0], kf_angle[:0], kf_node[:0], kf_valid[:0], f_node)
    assert enm == 0 and np.all(e == -1)


def _track_lines_case(plf, res, rng, mode):
    """Current frame = the slot's left lines; first set = the same lines moved / rotated a little (some a lot), plus
    strangers; some without map line, some current lines without stereo or already held."""
    nl = int(res.n_kl_left[0])
    kl2 = res.kl_left[0, :nl].copy()
    desc2 = res.ldesc_left[0, :nl].copy()
    disp2 = res.disp_se[0, :nl].copy()
    m = nl + 40
    src = np.concatenate([rng.permutation(nl), rng.integers(0, nl, 40)])
    lines1 = np.zeros(m, plf.TRACK_LINE_DT)
    big = rng.random(m) < 0.15
    shift = np.where(big[:, None], rng.normal(0, 60.0, (m, 4)), rng.normal(0, 6.0, (m, 4))).astype(np.float32)
    lines1["sx"] = kl2["startPointX"][src] + shift[:, 0]; lines1["sy"] = kl2["startPointY"][src] + shift[:, 1]
    lines1["ex"] = kl2["endPointX"][src] + shift[:, 2]; lines1["ey"] = kl2["endPointY"][src] + shift[:, 3]
    dang = np.where(rng.random(m) < 0.15, rng.uniform(-3.2, 3.2, m), rng.normal(0, 0.1, m))
    lines1["angle"] = (kl2["angle"][src] + dang).astype(np.float32)
    lines1["eligible"] = (rng.random(m) < 0.85).astype(np.int32)      # mode 0: has a map line; mode 1: the map line has observations
    desc1 = desc2[src].copy()
    desc1 ^= np.packbits(rng.random((m, 256)) < 0.03, axis=1)
    desc1[-40:] = rng.integers(0, 256, (40, 32), dtype=np.uint8)
    held2 = (rng.random(nl) < 0.1).astype(np.uint8) if mode == 1 else None
    return desc1, lines1, desc2, kl2, disp2, held2


def _match_lines_tracked_python(oracle_front, mode, desc1, lines1, desc2, kl2, disp2, held2, nnr, bounds):
    """src/Tracking.cc:3055-3099 / :3879-3919 restated on top of match() / matchNNR() (themselves tested above): mode 1
    matches one way only (src/LineMatcher.cpp:161-170 returns after matchNNR), so the loop is order-dependent."""
    import math
    f32 = np.float32
    if mode == 0:
        _, m12 = oracle_front.match(desc1, desc2, nnr, True)
    else:
        _, m12 = oracle_front.match_nnr(desc1, desc2, nnr)
    m12 = m12.copy()
    asg = np.full(len(desc1), -1, np.int32)
    holder = {}                                     # i2 -> ("old", has_obs) | ("new", i1)
    if mode == 1 and held2 is not None:
        for i2 in np.nonzero(held2)[0]:
            holder[int(i2)] = ("old", True)
    dw = float(f32(bounds[1]) - f32(bounds[0])) * 0.1
    dh = float(f32(bounds[3]) - f32(bounds[2])) * 0.1
    for i1 in range(len(desc1)):
        if mode == 0 and not lines1["eligible"][i1]:
            continue
        i2 = int(m12[i1])
        if i2 < 0:
            continue
        if disp2[i2, 0] < 0 or disp2[i2, 1] < 0:
            continue
        if mode == 1 and i2 in holder:
            kind, v = holder[i2]
            if (kind == "old" and v) or (kind == "new" and lines1["eligible"][v]):
                continue
        if mode == 0:
            th = float(f32(kl2["angle"][i2]) - f32(lines1["angle"][i1]))
            if th < -math.pi:
                th += 2 * math.pi
            elif th > math.pi:
                th -= 2 * math.pi
            if abs(th) > math.pi / 8.0:
                m12[i1] = -1
                continue
        if (abs(float(f32(kl2["startPointX"][i2]) - f32(lines1["sx"][i1]))) > dw or abs(float(f32(kl2["endPointX"][i2]) - f32(lines1["ex"][i1]))) > dw or
                abs(float(f32(kl2["startPointY"][i2]) - f32(lines1["sy"][i1]))) > dh or abs(float(f32(kl2["endPointY"][i2]) - f32(lines1["ey"][i1]))) > dh):
            m12[i1] = -1
            continue
        if i2 in holder and holder[i2][0] == "new":
            asg[holder[i2][1]] = -1
        holder[i2] = ("new", i1)
        asg[i1] = i2
    return m12, asg, int((asg >= 0).sum())


@pytest.mark.parametrize("mode", [0, 1])
def test_match_lines_tracked_oracle_against_python(plf, oracle, mode):
    L, R = plf.synth_pair(752, 480, 12)
    o = plf.Frontend(oracle, max_batch=1)
    res = o.frontend_batch(L[None], R[None])
    case = _track_lines_case(plf, res, np.random.default_rng(41 + mode), mode)
    bounds = (0.0, 752.0, 0.0, 480.0)
    m12, asg, na = o.match_lines_tracked(mode, *case, 0.9, bounds)
    wm, wa, wna = _match_lines_tracked_python(o, mode, *case, 0.9, bounds)
    assert na == wna and np.array_equal(m12, wm) and np.array_equal(asg, wa)
    assert na > 40 and (m12 >= 0).sum() > na            # some matches pass without an assignment (no stereo / not eligible / held)
    nm0, plain = o.match(case[0], case[2], 0.9, True) if mode == 0 else o.match_nnr(case[0], case[2], 0.9)
    assert (plain >= 0).sum() > (m12 >= 0).sum()        # and the gates removed some


def test_header_inlines_edge_cases(plf, oracle):
    """The host inlines of include/plf_b200.h on their edges: a bag of words with nothing but stopped words or no
    features at all, an area lookup whose window misses the image or covers all of it."""
    o = plf.Frontend(oracle, max_batch=1)
    bw, bv, fv = o.bow_build(np.array([3, 1, 3], np.int32), np.zeros(3), np.array([7, 7, 9], np.int32))
    assert len(bw) == 0 and len(bv) == 0 and fv == {}
    bw, bv, fv = o.bow_build(np.zeros(0, np.int32), np.zeros(0), np.zeros(0, np.int32))
    assert len(bw) == 0 and fv == {}
    bw, bv, fv = o.bow_build(np.array([5, 2, 5, 9], np.int32), np.array([1.0, 3.0, 0.5, 0.0]), np.array([40, 41, 40, 42], np.int32))
    assert list(bw) == [2, 5] and list(bv) == [3.0 / 4.5, 1.5 / 4.5] and fv == {40: [0, 2], 41: [1]}
    L, R = plf.synth_pair(752, 480, 4)
    res = o.frontend_batch(L[None], R[None])
    n = int(res.n_kp_left[0])
    st, ix = o.feature_grid(0, 1)
    kps = res.kp_left[0, :n]
    assert len(o.features_in_area(kps, st[0], ix[0], -500.0, 100.0, 20.0)) == 0
    assert len(o.features_in_area(kps, st[0], ix[0], 100.0, 5000.0, 20.0)) == 0
    everything = o.features_in_area(kps, st[0], ix[0], 376.0, 240.0, 2000.0)
    assert sorted(everything) == sorted(ix[0, :st[0, -1]].tolist()) and len(everything) >= n - 5
    lvl = o.features_in_area(kps, st[0], ix[0], 376.0, 240.0, 2000.0, 2, 3)
    assert len(lvl) > 0 and set(kps["octave"][lvl]) <= {2, 3}
