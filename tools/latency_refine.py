"""Wall time of one batched call with lsd_refine = 1 at small batch sizes (streaming grower; PLF_LSD_GROWER=seq for the sequential one)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, plf
sizes = [int(a) for a in sys.argv[1:]] or [1, 8, 64]
L, R = plf.synth_batch(752, 480, [3000 + i for i in range(min(max(sizes), 64))])
for B in sizes:
    idx = np.arange(B) % L.shape[0]
    f = plf.Frontend(plf.load_product(), max_batch=B, lsd_nfeatures=300, lsd_refine=1)
    out = f.new_result(B)
    Lb, Rb = np.ascontiguousarray(L[idx]), np.ascontiguousarray(R[idx])
    for _ in range(3):
        f.frontend_batch(Lb, Rb, out)
    t0 = time.perf_counter()
    for _ in range(5):
        f.frontend_batch(Lb, Rb, out)
    print("refine 1, B=%d: %.2f ms per call (%s)" % (B, (time.perf_counter() - t0) / 5 * 1e3, os.environ.get("PLF_LSD_GROWER", "auto")), flush=True)
    del f
