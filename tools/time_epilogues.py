"""Wall time of the §8(f) entry points over one 512-pair batch (API calls incl. their D2H copies and the sync)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, plf
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
W, H = 752, 480
L, R = plf.synth_batch(W, H, [1000 + i for i in range(16)])
idx = np.arange(B) % 16
f = plf.Frontend(plf.load_product(), max_batch=B, lsd_nfeatures=300)
for side in (0, 1):
    f.rectify_set_maps(side, *plf.rectify_maps(W, H, side))
out = f.new_result(B)
f.frontend_batch(L[idx], R[idx], out)
f.bow_set_vocabulary(0, plf.synth_vocabulary(10, 6, seed=1, ragged=0.0, stop=0.0))     # ORBvoc shape: 10^6 words
f.bow_set_vocabulary(1, plf.synth_vocabulary(10, 5, seed=2, ragged=0.0, stop=0.0))
Rwc = np.tile(np.eye(3, dtype=np.float32), (B, 1, 1)); Ow = np.zeros((B, 3), np.float32)


def timed(name, fn, reps=5):
    fn()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    dt = (time.perf_counter() - t) / reps
    print("%-34s %8.2f ms per %d-pair batch  (%.1f us per pair)" % (name, dt * 1e3, B, dt * 1e6 / B))


timed("feature_grid", lambda: f.feature_grid(0, B))
timed("backproject (points + lines)", lambda: f.backproject(Rwc, Ow, 435.2, 367.4, 252.2))
timed("bow_transform ORB (k=10, L=6)", lambda: f.bow_transform(0, B, 0, 4))
timed("bow_transform lines (k=10, L=5)", lambda: f.bow_transform(1, B, 0, 4))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_host_logic import _proj_queries
q = _proj_queries(plf, out, 0, np.random.default_rng(1))
occ = np.zeros(int(out.n_kp_left[0]), np.uint8)
t = time.perf_counter()
for _ in range(20):
    occ[:] = 0
    m, nm = f.search_by_projection(q, occ, th=1.0, slot=0)
print("%-34s %8.2f ms per frame (%d map points, %d matches)" % ("search_by_projection th=1", (time.perf_counter() - t) / 20 * 1e3, len(q), nm))
f.set_stage_timing(True)
f.batch_upload_raw(L[idx], R[idx]); f.batch_run(B); f.batch_download(B, out)
print("rectify kernel stage: %.2f ms per %d-pair batch" % (f.stage_ms().get("rectify", float("nan")), B))
