"""Stage times (CUDA events between the stage marks of one context) and wall time of one batched call at small batch sizes:
where the latency of a single stereo pair goes.  python tools/latency_stages.py [pairs ...]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, plf
sizes = [int(a) for a in sys.argv[1:]] or [1, 2, 8, 64]
D = max(sizes)
L, R = plf.synth_batch(752, 480, [3000 + i for i in range(min(D, 64))])
for B in sizes:
    idx = np.arange(B) % L.shape[0]
    f = plf.Frontend(plf.load_product(), max_batch=B, lsd_nfeatures=300)
    out = f.new_result(B)
    Lb, Rb = np.ascontiguousarray(L[idx]), np.ascontiguousarray(R[idx])
    for _ in range(3):
        f.frontend_batch(Lb, Rb, out)
    t0 = time.perf_counter()
    for _ in range(5):
        f.frontend_batch(Lb, Rb, out)
    wall = (time.perf_counter() - t0) / 5 * 1e3
    f.set_stage_timing(True)
    for _ in range(2):
        f.frontend_batch(Lb, Rb, out)
    ms = f.stage_ms()
    print("B=%d wall %.2f ms; stages (sum %.2f): %s" % (B, wall, sum(ms.values()), ", ".join("%s %.2f" % (k, v) for k, v in ms.items())), flush=True)
    del f

# the reference's own five signatures for ONE pair, called one after the other on one context:
# ORBextractor::operator() x 2, Lineextractor::operator() x 2, ComputeStereoMatches, ComputeStereoMatches_Lines
f = plf.Frontend(plf.load_product(), max_batch=1, lsd_nfeatures=300)
def five():
    m, k, d = f.orb_extract(0, L[0]); m2, k2, d2 = f.orb_extract(1, R[0])
    kl, ld = f.line_extract(0, L[0]); klr, ldr = f.line_extract(1, R[0])
    u, dep = f.stereo_match_points(len(k)); disp, le, m12 = f.stereo_match_lines(len(kl))
    return len(k), len(kl)
for _ in range(3):
    five()
t0 = time.perf_counter()
for _ in range(5):
    nk, nl = five()
print("five reference signatures, one pair: %.2f ms per pair (%d keypoints, %d lines)" % ((time.perf_counter() - t0) / 5 * 1e3, nk, nl), flush=True)
