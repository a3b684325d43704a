"""One small pass through every entry point of the C ABI (developer tool: run under compute-sanitizer on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, plf
W, H, B = 752, 480, 2
L, R = plf.synth_batch(W, H, [1, 2])
f = plf.Frontend(plf.load_product(), max_batch=B, lsd_refine=int(sys.argv[1]) if len(sys.argv) > 1 else 0)
for side in (0, 1):
    f.rectify_set_maps(side, *plf.rectify_maps(W, H, side))
img = f.rectify(0, L[0])
out = f.new_result(B)
f.batch_upload_raw(L, R); f.batch_run(B); f.batch_download(B, out)
res = f.frontend_batch(L, R)
st, ix = f.feature_grid(0, B)
x3d, l3d = f.backproject(np.tile(np.eye(3, dtype=np.float32), (B, 1, 1)), np.zeros((B, 3), np.float32), 435.2, 367.4, 252.2)
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_host_logic import _proj_queries, _frame_queries, _bow_case, _track_lines_case
rng = np.random.default_rng(0)
n0 = int(res.n_kp_left[0])
mq, nmq = f.search_by_projection(_proj_queries(plf, res, 0, rng), np.zeros(n0, np.uint8), th=3.0)
fq, m12, nfq = f.search_by_projection_frame(_frame_queries(plf, res, 0, rng), np.zeros(n0, np.uint8))
fr, nfr = f.search_by_projection_reloc(_frame_queries(plf, res, 0, rng), np.zeros(n0, np.uint8), 64)
ql = _frame_queries(plf, res, 0, rng, "backward"); ql["min_level"] = ql["max_level"] - 1
fl, nfl = f.search_by_projection_loop(ql, np.zeros(n0, np.uint8), 50, 0.8)
mb, nmb = f.search_by_bow(*_bow_case(plf, res, rng), 50, 0.7, True)
for mode in (0, 1):
    tm, ta, tn = f.match_lines_tracked(mode, *_track_lines_case(plf, res, rng, mode), 0.9, (0.0, float(W), 0.0, float(H)))
for rep in range(3):                      # the second and third call replay the CUDA graph of plf_batch_run
    res = f.frontend_batch(L, R)
f.bow_set_vocabulary(0, plf.synth_vocabulary(10, 4, seed=1)); f.bow_set_vocabulary(1, plf.synth_vocabulary(6, 3, seed=2, ragged=0.0))
bw = f.bow_transform(0, B); bl = f.bow_transform(1, B)
m, k, d = f.orb_extract(0, L[0]); m2, k2, d2 = f.orb_extract(1, R[0])
kl, ld = f.line_extract(0, L[0]); klr, ldr = f.line_extract(1, R[0])
u, dep = f.stereo_match_points(len(k)); disp, le, m12 = f.stereo_match_lines(len(kl))
n1, _ = f.match_nnr(ld, ldr, 0.9); n2, _ = f.match(ld, ldr, 0.9, 1)
print("ok", nmq, nfq, nfr, nfl, nmb, tn, int((bw[0] >= 0).sum()), int(res.n_kp_left[0]), int(res.n_kl_left[0]), int(st[0, -1]), int((x3d[0] != 0).any(axis=1).sum()), n1, n2, img.mean().round(1))
