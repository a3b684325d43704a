#!/usr/bin/env python
"""Headline benchmark: stereo frames/s of the point+line frontend on synthetic EuRoC-shape 752x480 pairs.

Workload = BASELINE.json configs[1]: stereo stream 752x480, 1200 ORB features / 8 levels, LSD+LBD lines (top 300 by
response), stereo point + line matching, batch 64 pairs per call.  One process per GPU; streams (= independent stereo
sequences) are sharded across ranks with no data-path collective (replicas only, weak scaling); torch.distributed is
used for the barrier and the max-over-ranks of the device time only.

Each camera stream contributes a batch of 64 consecutive frames; a rank serves `streams` independent streams per
context, so one call of the C ABI processes streams x 64 pairs (the region-growing kernel is latency-bound with one
warp per image, hence many images per launch), and `contexts` such calls are kept in flight on their own CUDA streams
so that the copies and the wide kernels of one overlap the narrow kernels of the other.
A "step" = one pass of the hot path over contexts x streams x 64 pairs on this rank.

  value : pairs/s with the inputs already resident in HBM (plf_batch_run only), all ranks
  e2e   : pairs/s through the C ABI with pinned HOST buffers: H2D of the images + kernels + D2H of every result array

  python bench.py --gpus 1 --steps 4 --warmup 3
  python -m torch.distributed.run --nproc-per-node 8 ... bench.py --gpus 8 ...
  python bench.py --impl reference       # the CPU port (oracle) on the host cores, bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 752, 480
WORKLOAD = dict(n_features=1200, n_levels=8, lsd_nfeatures=300)
# BASELINE.json configs that are bench lines (the others are parity-test cases).  bytes_per_pair = SURVEY.md 8d's algorithmic
# bytes of one stereo pair through the configured path; dominant = the kernel the roofline object describes, with its own
# algorithmic bytes per image (8d) - region growing reads angle + modgrad (2 x f32) and the u8 used map once: 9 S.
CONFIGS = {
    "c2": dict(name="euroc_752x480_stereo_pointline_batch64", W=752, H=480, params=dict(n_features=1200, n_levels=8, lsd_nfeatures=300),
               bytes_per_pair=50_661_326, dominant=("lsd_grow_kernel", "lsd_grow", 9 * 902 * 576), seed0=10_000),
    "c3": dict(name="orb_only_752x480_2000feat", W=752, H=480, params=dict(n_features=2000, n_levels=8, has_lines=0),
               bytes_per_pair=18_551_470, dominant=("fast_score_kernel+fast_cells_kernel", "orb_fast", 2 * 1_117_367), seed0=1000),
    "c4": dict(name="line_path_1280x720_lsd_lbd_matchnnr", W=1280, H=720, params=dict(n_features=1200, n_levels=8, lsd_nfeatures=500, has_points=0),
               bytes_per_pair=82_240_000, dominant=("lsd_grow_kernel", "lsd_grow", 9 * 1536 * 864), seed0=2000),
}
CONFIGS["c5"] = dict(CONFIGS["c2"], name="euroc_752x480_512_streams_sharded")
# c2 with TRUE 64-pair calls (one camera stream per call, 16 calls in flight): launches of 128 images go through the streaming
# multi-warp region grower (1.9x the instructions of the one-warp-per-image kernel, a sixth of its latency)
CONFIGS["c2_batch64"] = dict(CONFIGS["c2"], name="euroc_752x480_stereo_pointline_true_batch64",
                             dominant=("lsd_grow_sw_kernel", "lsd_grow", 9 * 902 * 576))
BYTES_PER_PAIR = CONFIGS["c2"]["bytes_per_pair"]


def shard_streams(n_streams, world_size, rank):
    """Stream s -> rank s mod world_size: the product's own sharder (pli-slam_b200/pool.py)."""
    import plf
    return plf.shard_streams(n_streams, world_size, rank)


def stream_seed(stream, frame):
    """Seed convention of SURVEY §8d for C5: stream s, frame f -> 10000*(s+1)+f (pli-slam_b200/pool.py)."""
    import plf
    return plf.stream_seed(stream, frame)


def _gen_pair(args):
    import plf
    seed, w, h = args
    return plf.synth_pair(w, h, seed)


def make_inputs(seeds, W=752, H=480):
    """Distinct synthetic pairs, generated on the host cores in parallel (pure numpy, deterministic per seed)."""
    from concurrent.futures import ProcessPoolExecutor
    L = np.empty((len(seeds), H, W), np.uint8)
    R = np.empty((len(seeds), H, W), np.uint8)
    seeds = [(s, W, H) for s in seeds]
    workers = max(1, min(len(seeds), (os.cpu_count() or 2)))
    try:
        with ProcessPoolExecutor(max_workers=workers) as ex:
            for i, (l, r) in enumerate(ex.map(_gen_pair, seeds, chunksize=2)):
                L[i], R[i] = l, r
    except Exception:
        for i, s_ in enumerate(seeds):
            L[i], R[i] = _gen_pair(s_)
    return L, R


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libplf_ref_o3.so")


def _ref_lib():
    """oracle/_ref: the reference's OWN frontend sources (ORBextractor.cc, LineExtractor.cc, the line_descriptor library,
    the stereo matchers of Frame.cc, LineMatcher.cpp ...) compiled where they lie by oracle/build_ref.py, -O3
    -march=x86-64-v3, over the cv2-pinned restatements of the OpenCV primitives.  Returns None when it did not travel."""
    import ctypes as C
    if not os.path.exists(REF_LIB):
        return None
    lib = C.CDLL(REF_LIB)
    lib.ref_frame_create.restype = C.c_void_p
    return lib


def reference_cpu(cfg, L, R, workers, pairs_per_worker):
    """The reference's stereo Frame constructor on the host: `workers` pairs in flight, each with the reference's own
    four extraction threads (src/Frame.cc:128-135) followed by the two stereo matchers.  Returns (pairs/s, seconds)."""
    import ctypes as C
    lib = _ref_lib()
    p = dict(n_features=1200, scale_factor=1.2, n_levels=8, ini=20, mn=7, lsd_nfeatures=500, has_lines=1, has_points=1)
    p.update(cfg["params"])
    w, h = cfg["W"], cfg["H"]
    use_threads = 1 | (0 if p["has_points"] else 2) | (0 if p["has_lines"] else 4)

    def make():
        return C.c_void_p(lib.ref_frame_create(p["n_features"], C.c_float(1.2), p["n_levels"], 20, 7, p["lsd_nfeatures"], C.c_double(0.025), 0,
                                               C.c_double(1.2), C.c_double(0.6), C.c_double(2.0), C.c_double(22.5), C.c_double(1.0),
                                               C.c_double(0.6), 1024, C.c_float(47.90639), C.c_float(435.2047)))
    handles = [make() for _ in range(workers)]
    done = [0.0] * workers

    def work(k):
        cnt = (C.c_int * 6)()
        for i in range(pairs_per_worker):
            j = (k * pairs_per_worker + i) % len(L)
            lib.ref_frame_run(handles[k], L[j].ctypes.data_as(C.c_void_p), R[j].ctypes.data_as(C.c_void_p), w, h, w, use_threads, cnt)
        done[k] = time.perf_counter()
    ts = [threading.Thread(target=work, args=(k,)) for k in range(workers)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.perf_counter() - t0
    for hnd in handles:
        lib.ref_frame_destroy(hnd)
    return workers * pairs_per_worker / dt, dt


def port_cpu(cfg, L, R, threads):
    """Fallback when oracle/_ref did not travel: the restated oracle (kind "port"), `threads` workers over independent pairs."""
    import plf
    orc = plf.load_oracle()
    n = len(L)
    f = plf.Frontend(orc, width=cfg["W"], height=cfg["H"], max_batch=n, **cfg["params"])
    orc.dll.plf_cpu_set_threads(f.ctx, threads)
    out = f.new_result(n)
    f.batch_upload(L, R)
    t0 = time.perf_counter()
    f.batch_run(n)
    dt = time.perf_counter() - t0
    f.batch_download(n, out)
    return n / dt, dt


def cpu_cross_checks(cfg, img):
    """Single-thread times of the dominant CPU stage beside real OpenCV (cv2 wheel of this image), ms per image."""
    out = {}
    try:
        import ctypes as C
        import plf
        o = plf.load_oracle().dll
        h, w = img.shape
        seg = np.zeros((20000, 4), np.float32)
        n = C.c_int(0)
        P = lambda a: a.ctypes.data_as(C.c_void_p)
        t = time.perf_counter()
        for _ in range(3):
            o.plf_cpu_prim_lsd(P(img), w, h, C.c_double(1.2), 1, P(seg), 20000, C.byref(n))
        out["port_lsd_ms"] = round((time.perf_counter() - t) / 3 * 1e3, 1)
        import cv2
        cv2.setNumThreads(1)
        lsd = cv2.createLineSegmentDetector(0, 1.2, 0.6, 2.0, 22.5, 1.0, 0.6, 1024)
        lsd.detect(img)
        t = time.perf_counter()
        for _ in range(3):
            lsd.detect(img)
        out["cv2_lsd_ms"] = round((time.perf_counter() - t) / 3 * 1e3, 1)
    except Exception as e:          # cv2 missing on the box: the cross-check is optional
        out["cross_check_error"] = str(e)[:80]
    return out


def cpu_baseline(cfg, budget_s=12.0):
    """Reference CPU leg on a bounded sample: throughput-saturated (cores/2 pairs in flight x the reference's 4 threads) and
    the reference's own shape (one pair at a time, 4 threads).  Returns the cpu_baseline object of the JSON line."""
    cores = os.cpu_count() or 1
    workers = max(1, cores // 2)
    seeds = [cfg["seed0"] + i for i in range(max(workers, 4))]
    L, R = make_inputs(seeds, cfg["W"], cfg["H"])
    if _ref_lib() is None:
        v, dt = port_cpu(cfg, L, R, cores)
        return {"value": v, "unit": "stereo pairs/s", "cores": cores, "kind": "port",
                "sample": "%d pairs, %d worker threads over independent pairs, %.1f s (oracle/_ref absent)" % (len(L), cores, dt)}
    _, d1 = reference_cpu(cfg, L, R, 1, 2)                              # warm-up + latency of the reference's own shape
    lat_ms = d1 / 2 * 1e3
    per_worker = max(2, int(budget_s / max(d1 / 2, 1e-3) / 2))
    per_worker = min(per_worker, 16)
    v, dt = reference_cpu(cfg, L, R, workers, per_worker)
    out = {"value": v, "unit": "stereo pairs/s", "cores": cores, "kind": "reference",
           "sample": "%d pairs: %d pairs in flight x the reference's 4 extraction threads, %.1f s; reference sources compiled by "
                     "oracle/build_ref.py (-O3 -march=x86-64-v3) over cv2-pinned OpenCV primitives" % (workers * per_worker, workers, dt),
           "latency_ms_one_pair_4_threads": round(lat_ms, 1), "pairs_in_flight": workers}
    out.update(cpu_cross_checks(cfg, L[0]))
    return out


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (oracle/_ref) with every host thread it can use,
    on the configured workload, each step a bounded sample.  Rank 0 alone runs it."""
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    cores = os.cpu_count() or 1
    workers = max(1, cores // 2)
    have_ref = _ref_lib() is not None
    seeds = [cfg["seed0"] + i for i in range(max(workers * 2, 8))]
    L, R = make_inputs(seeds, cfg["W"], cfg["H"])
    per_worker = 2
    run = (lambda: reference_cpu(cfg, L, R, workers, per_worker)) if have_ref else (lambda: port_cpu(cfg, L, R, cores))
    for _ in range(min(args.warmup, 1)):
        run()
    pairs = 0
    t_all = 0.0
    for _ in range(args.steps):
        v, dt = run()
        pairs += v * dt
        t_all += dt
    value = pairs / t_all
    base = cpu_baseline(cfg, budget_s=4.0) if have_ref else {"kind": "port", "cores": cores, "sample": "oracle port, all cores"}
    base["value"] = value
    line = {"impl": "reference", "metric": "stereo_frames_per_sec", "value": value, "unit": "stereo pairs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_all / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": cfg["name"], "width": cfg["W"], "height": cfg["H"], **cfg["params"],
                       "pairs_per_step": int(round(pairs / args.steps)), "timing": "host wall clock, inputs in host memory"},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": "stereo pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="frames per camera stream per call (BASELINE config C2)")
    ap.add_argument("--streams", type=int, default=8, help="independent camera streams served by one context")
    ap.add_argument("--rectify", action="store_true",
                    help="e2e leg starts from RAW frames: upload + cv::remap rectification on the device (SURVEY 8f rank 2)")
    ap.add_argument("--contexts", type=int, default=8, help="calls kept in flight per GPU (one CUDA stream each; 8 x 16.5 GB of buffers for 512-pair calls)")
    ap.add_argument("--distinct", type=int, default=512, help="distinct synthetic pairs generated per rank (every launch sees pairs_per_call distinct pairs)")
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS), help="BASELINE.json config: c2 headline, c3 ORB only, c4 1280x720 line path, c5 512 streams")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import plf

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    # native libraries (NCCL's version banner, for one) printf to stdout: keep fd 1 for the one JSON line
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    prod = plf.load_product()
    cfg = CONFIGS[args.config]
    global W, H, WORKLOAD
    W, H, WORKLOAD = cfg["W"], cfg["H"], dict(cfg["params"])
    if args.config == "c2_batch64" and args.streams == 8 and args.contexts == 8:
        args.streams, args.contexts = 1, 16
    if args.config == "c4" and args.streams == 8 and args.contexts == 8:
        args.streams, args.contexts = 3, 8                # 1280x720: 2.6x the pixels and device memory per pair (192-pair calls)
    B, C = args.batch * args.streams, args.contexts       # pairs per call, calls in flight
    # this rank's camera streams (global ids), args.streams of them per context: stream s -> rank s mod world
    n_streams_total = 512 if args.config == "c5" else C * args.streams * world
    my_streams = shard_streams(max(n_streams_total, C * args.streams * world), world, rank)
    distinct = min(args.distinct, B)
    # distinct pairs of this rank: frames f of its first streams, seed convention of SURVEY 8d (c2/c5) or the config's own
    if args.config in ("c2", "c5", "c2_batch64"):
        seeds = [stream_seed(my_streams[(i // args.batch) % len(my_streams)], i % args.batch) for i in range(distinct)]
    else:
        seeds = [cfg["seed0"] + rank * distinct + i for i in range(distinct)]
    Ld, Rd = make_inputs(seeds, W, H)
    ctxs, hostL, hostR, results = [], [], [], []
    pool = plf.DevicePool(prod, local_rank, C, args.streams, args.batch, width=W, height=H, **WORKLOAD)
    for ci, f in enumerate(pool):
        # every call covers `distinct` different pairs (all of them when distinct == pairs_per_call); the contexts rotate them
        idx = (np.arange(B) + ci * (distinct // max(C, 1) + 1)) % distinct
        l = torch.from_numpy(np.ascontiguousarray(Ld[idx])).pin_memory()
        r = torch.from_numpy(np.ascontiguousarray(Rd[idx])).pin_memory()
        if args.rectify:
            for side in (0, 1):
                f.rectify_set_maps(side, *plf.rectify_maps(W, H, side))
        ctxs.append(f); hostL.append(l); hostR.append(r); results.append(f.new_result(B, pinned=True))
    ext = [torch.cuda.ExternalStream(f.stream(), device=local_rank) for f in ctxs]
    main_stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        for f in ctxs:
            f.batch_run(B)

    pending = [False] * C

    def step_e2e():
        # per context: collect the previous call's results (D2H + sync), then queue the next upload + kernels, so the
        # other contexts keep the GPU busy while the host waits
        for i, (f, l, r, out) in enumerate(zip(ctxs, hostL, hostR, results)):
            if pending[i]:
                f.batch_download(B, out)      # D2H of every result array + stream sync
            if args.rectify:
                f.batch_upload_raw_ptr(l.data_ptr(), r.data_ptr(), B, W)
            else:
                f.batch_upload_ptr(l.data_ptr(), r.data_ptr(), B, W)
            f.batch_run(B)
            pending[i] = True

    def drain_e2e():
        for i, (f, out) in enumerate(zip(ctxs, results)):
            if pending[i]:
                f.batch_download(B, out)
                pending[i] = False

    def timed(step_fn, steps):
        """K steps bracketed by barrier+synchronize; device time by CUDA events fanned out to / joined from every
        context stream; returns seconds (max over ranks)."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main_stream)
        for s in ext:
            s.wait_event(e0)
        for _ in range(steps):
            step_fn()
        for s in ext:
            d = torch.cuda.Event()
            d.record(s)
            main_stream.wait_event(d)
        e1.record(main_stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / 1e3

    # inputs resident for the `value` leg
    for f, l, r in zip(ctxs, hostL, hostR):
        f.batch_upload_ptr(l.data_ptr(), r.data_ptr(), B, W)
    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local_rank)
    sampler.start()
    # stage marks (CUDA events on the launching stream) of context 0 stay on during the timed region: the dominant
    # kernel's duration is taken live, from the last timed step, with the other contexts running beside it
    ctxs[0].set_stage_timing(True)
    t_res = timed(step_resident, args.steps)
    sampler.stop_flag = True
    ctxs[0].batch_download(B, results[0])              # outside the timed region: joins the stream, evaluates the marks
    live_ms = dict(ctxs[0].stage_ms())
    ctxs[0].set_stage_timing(False)
    launches = sum(f.launch_count() for f in ctxs) * args.steps
    for _ in range(2):
        step_e2e()
    drain_e2e()

    def e2e_steps():
        step_e2e()

    barrier()
    t0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(main_stream)
    for s_ in ext:
        s_.wait_event(ev0)
    for _ in range(args.steps):
        step_e2e()
    drain_e2e()                               # every result of the K steps is on the host when the clock stops
    for s_ in ext:
        d_ = torch.cuda.Event()
        d_.record(s_)
        main_stream.wait_event(d_)
    ev1.record(main_stream)
    barrier()
    t_e2e = ev0.elapsed_time(ev1) / 1e3
    t_e2e_wall = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([t_e2e], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_e2e = float(tt.item())
    sampler.join(timeout=2)

    # c2_batch64 only: the same resident steps with every context told to favour throughput (one warp per image even for
    # small launches) - what a caller gets who keeps many small calls in flight and does not care about the latency of one
    value_tp = None
    if args.config == "c2_batch64":
        for f in ctxs:
            f.set_grower_policy(1)
        for _ in range(3):
            step_resident()
        value_tp = B * C * args.steps * world / timed(step_resident, args.steps)
        for f in ctxs:
            f.set_grower_policy(0)
        for _ in range(2):
            step_resident()
        barrier()

    # per-stage device time and the dominant kernel's own duration (CUDA events on the launching stream, one context
    # alone so that stages do not overlap each other)
    f0 = ctxs[0]
    f0.set_stage_timing(True)
    stage_acc = {}
    reps = 3
    for _ in range(reps):
        if args.rectify:
            f0.batch_upload_raw_ptr(hostL[0].data_ptr(), hostR[0].data_ptr(), B, W)
        f0.batch_run(B)
        f0.batch_download(B, results[0])
        for k, v in f0.stage_ms().items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v / reps
    # one warp per image: the grow launch lasts as long as its slowest image - per-image run times
    grow_img = None
    try:
        ns = f0.grow_ns(2 * B).astype(np.float64)
        if ns.max() > 0:
            grow_img = {"max": round(float(ns.max()) / 1e6, 2), "mean": round(float(ns.mean()) / 1e6, 2),
                        "min": round(float(ns.min()) / 1e6, 2), "p99": round(float(np.percentile(ns, 99)) / 1e6, 2), "images": int(2 * B)}
    except Exception:
        pass
    f0.set_stage_timing(False)
    barrier()

    # latency of ONE pair through the batched call on an otherwise idle GPU (host buffers in, results out)
    lat_ms = lat5_ms = None
    try:
        f1 = plf.Frontend(prod, device=local_rank, width=W, height=H, max_batch=1, **WORKLOAD)
        o1 = f1.new_result(1, pinned=True)
        for _ in range(3):
            f1.frontend_batch(Ld[:1], Rd[:1], o1)
        t1 = time.perf_counter()
        for k in range(5):
            f1.frontend_batch(Ld[k % distinct:k % distinct + 1], Rd[k % distinct:k % distinct + 1], o1)
        lat_ms = (time.perf_counter() - t1) / 5 * 1e3
        # ... and through the reference's own five signatures, one call after the other on one context (c2-like configs)
        if WORKLOAD.get("has_points", 1) and WORKLOAD.get("has_lines", 1):
            def five(k):
                kp = f1.orb_extract(0, Ld[k])[1]; f1.orb_extract(1, Rd[k])
                kl = f1.line_extract(0, Ld[k])[0]; f1.line_extract(1, Rd[k])
                f1.stereo_match_points(len(kp)); f1.stereo_match_lines(len(kl))
            for k in range(2):
                five(k % distinct)
            t1 = time.perf_counter()
            for k in range(4):
                five(k % distinct)
            lat5_ms = (time.perf_counter() - t1) / 4 * 1e3
        f1.close()
    except Exception:
        pass
    barrier()

    pairs_per_step = B * C * world
    value = pairs_per_step * args.steps / t_res
    e2e = pairs_per_step * args.steps / t_e2e
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    dom_kernel, dom_stage, dom_bytes = cfg["dominant"]
    if dom_stage == "lsd_grow":      # launches of at most 296 images go through the streaming multi-warp grower
        dom_kernel = "lsd_grow_sw_kernel" if 2 * B <= 296 else "lsd_grow_kernel"
    dom_alone_ms = stage_acc.get(dom_stage, 0.0)
    dom_ms = live_ms.get(dom_stage, 0.0) or dom_alone_ms
    achieved = (dom_bytes * 2 * B / (dom_ms * 1e-3) / 1e9) if dom_ms > 0 else 0.0
    achieved_alone = (dom_bytes * 2 * B / (dom_alone_ms * 1e-3) / 1e9) if dom_alone_ms > 0 else 0.0
    live_total = sum(v for k, v in live_ms.items() if k not in ("h2d", "d2h")) or 1.0
    h2d, d2h = f0.io_bytes()
    traffic = None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel (ncu --set full, profiles/)
        prof = json.load(open(os.path.join(ROOT, "profiles", "r02_dominant_traffic.json")))[args.config]
        if prof.get("images_per_launch") == 2 * B:
            traffic = prof["dram_bytes_per_launch"]
    except Exception:
        pass
    if rank == 0:
        line = {
            "metric": "stereo_frames_per_sec", "value": value, "unit": "stereo pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * t_res / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": cfg["name"], "baseline_config": args.config, "width": W, "height": H, **WORKLOAD,
                       "frames_per_stream": args.batch, "streams_per_context": args.streams, "pairs_per_call": B,
                       "contexts_in_flight": C, "pairs_per_step": pairs_per_step,
                       "e2e_input": "raw frames, rectified on the device" if args.rectify else "rectified frames",
                       "distinct_pairs_per_rank": distinct, "distinct_images_per_launch": 2 * min(distinct, B),
                       "l2": "inputs larger than L2: %d contexts x %.0f MB of resident images" % (C, 2 * B * W * H / 1e6),
                       "parallelism": "replicas%d (streams sharded, no collective)" % world},
            "e2e": {"value": e2e, "unit": "stereo pairs/s", "h2d_bytes_per_step": h2d * B * C * world,
                    "d2h_bytes_per_step": d2h * B * C * world, "host_wall_s": round(t_e2e_wall, 4)},
            "gpu_launches": launches,
            **({"grower_policy": {"auto_pairs_per_s": value, "throughput_pairs_per_s": value_tp}} if value_tp else {}),
            "latency_ms_single_pair": None if lat_ms is None else round(lat_ms, 2),
            "latency_ms_five_signatures": None if lat5_ms is None else round(lat5_ms, 2),
            "ms_per_stage": {k: round(v, 4) for k, v in stage_acc.items()},
            "grow_ms_per_image": grow_img,
            "roofline": {"bound": "hbm", "kernel": dom_kernel, "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak if peak else None, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                         "algorithmic_bytes_per_launch": dom_bytes * 2 * B,
                         "launch_ms_live": round(dom_ms, 3), "launch_ms_alone": round(dom_alone_ms, 3),
                         "achieved_alone": achieved_alone, "share_of_step_live": round(dom_ms / live_total, 4),
                         "whole_path_frac": cfg["bytes_per_pair"] * value / world / (peak * 1e9)},
            "clocks": sampler.summary(),
        }
        if not args.no_cpu_baseline and world == 1:      # reported on rank 0 at N = 1 only
            line["cpu_baseline"] = cpu_baseline(cfg)
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
