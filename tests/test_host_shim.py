"""The C++ host mirror of the reference classes (pli-slam_b200/host/plf_frontend.hpp): compiles on CPU against the
product library; on the GPU box it runs one frame the way Frame::Frame(stereo) does and must agree with the oracle."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host_shim_test.cpp")


def _build(tmp_path):
    import __graft_entry__ as g
    if not os.path.exists(g.LIB):
        g.build()
    exe = str(tmp_path / "host_shim_test")
    subprocess.check_call(["g++", "-std=c++17", "-O1", SRC, "-o", exe, g.LIB, "-Wl,-rpath," + os.path.dirname(g.LIB)])
    return exe


def test_shim_compiles_and_fails_loudly_without_gpu(tmp_path, plf, pair1):
    import torch
    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    L, R = pair1
    L.tofile(tmp_path / "l.raw"); R.tofile(tmp_path / "r.raw")
    out = subprocess.run([exe, str(tmp_path / "l.raw"), str(tmp_path / "r.raw"), "752", "480"], capture_output=True, text=True)
    assert out.returncode == 1 and "no CPU path" in out.stderr      # no fallback: the shim throws


@pytest.mark.gpu
def test_shim_matches_oracle(tmp_path, plf, oracle, pair1):
    exe = _build(tmp_path)
    L, R = pair1
    L.tofile(tmp_path / "l.raw"); R.tofile(tmp_path / "r.raw")
    mx, my = plf.rectify_maps(752, 480, 0)
    mx.tofile(tmp_path / "mx.f32"); my.tofile(tmp_path / "my.f32")
    out = subprocess.run([exe, str(tmp_path / "l.raw"), str(tmp_path / "r.raw"), "752", "480", str(tmp_path / "mx.f32"),
                          str(tmp_path / "my.f32")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    got = json.loads(out.stdout.strip().splitlines()[-1])
    o = plf.Frontend(oracle)
    ml, k, d = o.orb_extract(0, L)
    mr, kr, dr = o.orb_extract(1, R)
    kl, ld = o.line_extract(0, L)
    klr, ldr = o.line_extract(1, R)
    u, _ = o.stereo_match_points(len(k))
    disp, _, _ = o.stereo_match_lines(len(kl))
    nnr, _ = o.match_nnr(ld, ldr, 0.9)
    h = 1469598103934665603
    for b in d.tobytes():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    assert got["N"] == len(k) and got["Nr"] == len(kr) and got["mono"] == [ml, mr]
    assert got["Nl"] == len(kl) and got["Nlr"] == len(klr)
    assert got["stereo_pts"] == int((u >= 0).sum()) and abs(got["sum_u"] - float(u[u >= 0].astype(np.float64).sum())) < 1e-2
    assert got["stereo_lines"] == int((disp[:, 0] >= 0).sum()) and got["nnr"] == nnr
    assert got["desc_fnv"] == h
    assert got["hamming01"] == int(np.unpackbits(d[0] ^ d[1]).sum()) and got["empty"] == -1

    def fnv(values):
        x = 1469598103934665603
        for v in values:
            x = ((x ^ int(v)) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        return x
    # FrameTail: grid + area lookup + back-projection against the oracle
    st, ix = o.feature_grid(0, 1)
    area = o.features_in_area(k, st[0], ix[0], 376.0, 240.0, 60.0, 0, 3)
    assert got["area_n"] == len(area) > 5 and got["area_fnv"] == fnv(area)
    Rwc = np.array([0.96, -0.28, 0, 0.28, 0.96, 0, 0, 0, 1], np.float32).reshape(1, 3, 3)
    x3d, l3d = o.backproject(Rwc, np.array([[0.5, -1.25, 2.0]], np.float32), 435.2047, 367.4517, 252.2008)
    assert abs(got["sum_x3d"] - float(x3d[0, :len(k)].astype(np.float64).sum())) < 1e-3 * max(1.0, abs(got["sum_x3d"]))
    assert abs(got["sum_l3d"] - float(l3d[0, :len(kl)].sum())) < 1e-6 * max(1.0, abs(got["sum_l3d"]))
    # gated line matching and SearchByBoW through the shim, against the oracle's entry points
    lines1 = np.zeros(len(kl), plf.TRACK_LINE_DT)
    lines1["sx"] = kl["startPointX"] + np.float32(4); lines1["sy"] = kl["startPointY"]
    lines1["ex"] = kl["endPointX"] + np.float32(4); lines1["ey"] = kl["endPointY"]
    lines1["angle"] = kl["angle"]; lines1["eligible"] = 1
    _, _, na = o.match_lines_tracked(0, ld, lines1, ld, kl, disp, None, 0.9, (0.0, 752.0, 0.0, 480.0))
    assert got["tracked"] == na > 20
    nodes = (np.arange(len(k)) % 8).astype(np.int32)
    m, nb = o.search_by_bow(d, k["angle"], nodes, np.ones(len(k), np.uint8), nodes, 50, 0.7, True)
    assert got["bow"] == [nb, int((m == np.arange(len(k))).sum())] and nb > 500
    # lapping area {0, 1000}: every row back to front, monoIndex 0; padded rows give the dense result
    assert got["mono_lap"] == o.orb_extract(0, L, lapping=(0, 1000))[0] == 0 and got["lap_reversed"] == 1
    assert got["pad_same"] == 1
    # Rectifier
    o.rectify_set_maps(0, mx, my)
    assert got["rect_fnv"] == fnv(o.rectify(0, L).tobytes())
