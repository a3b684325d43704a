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
    out = subprocess.run([exe, str(tmp_path / "l.raw"), str(tmp_path / "r.raw"), "752", "480"], capture_output=True, text=True)
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
