#!/usr/bin/env python
"""bench.py — frames/sec of the filtering + segmentation hot path (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W     # the reference's CPU path (oracle port)
    torchrun --nproc-per-node N bench.py --gpus N ...        # one rank per GPU

A step is one pass of the hot path over one chunk of synthetic video: by
default BASELINE.json configs[1] (1080p full frame, N=5 median, 3x3 opening,
int32 labels) in chunks of 1024 frames (6.4 GB of BGR, >> the 126 MB L2, so no
L2 flush is needed between steps).  Inputs are generated on the device by the
seeded CUDA generator before the timed region.  `value` is device-timed (CUDA
events on the launch stream, max over ranks); `e2e` runs the same chunks
through the C ABI from pinned HOST buffers (H2D copy and D2H of the segment
table inside the timed region; two contexts alternate so that the copy of the
next chunk is queued while the table of the previous one is collected).
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

CONFIGS = {
    # BASELINE.json configs[1]
    "1080p_full_n5_open3": dict(H=1080, W=1920, roi=None, N=5, se=3, do_close=False, birds=300, chunk=1024),
    # BASELINE.json configs[2]
    "4k_full_n9_oc5": dict(H=2160, W=3840, roi=None, N=9, se=5, do_close=True, birds=600, chunk=256),
    # BASELINE.json configs[0] geometry (the reference's CPU-runnable case)
    "1080p_roi320x240_n5_open3": dict(H=1080, W=1920, roi=[(800, 400), (1120, 640)], N=5, se=3,
                                      do_close=False, birds=300, chunk=2048, dropin=True, cpu_in_also=True),
    # BASELINE.json configs[4] filtering part (dense swarm, ~500 segments/frame)
    "1080p_dense_n5_open3": dict(H=1080, W=1920, roi=None, N=5, se=3, do_close=False, birds=2500, chunk=512),
    # BASELINE.json configs[4]: filter + label + batched segment classification (SqueezeNet1.0 as in the
    # reference, random-init weights: model.pt is not redistributable), ~500 segments per frame
    "1080p_swarm500_n5_open3": dict(H=1080, W=1920, roi=None, N=5, se=3, do_close=False, birds=500, chunk=1024),
    "1080p_swarm500_classify": dict(H=1080, W=1920, roi=None, N=5, se=3, do_close=False, birds=500, chunk=16,
                                    classify=True),
}
# the reference's own background model (rpca + bilateral) on its own batch size (21 frames)
CONFIGS["1080p_roi320x160_rpca21"] = dict(H=1080, W=1920, roi=[(800, 400), (1120, 560)], N=1, se=3, do_close=False,
                                          birds=300, chunk=21, bg_model="rpca")
CONFIGS["1080p_full_rpca21"] = dict(H=1080, W=1920, roi=None, N=1, se=3, do_close=False, birds=300, chunk=21,
                                    bg_model="rpca")
# BASELINE.json configs[3]: 16 videos, each with its own chimney ROI, processed concurrently
CONFIGS["16x1080p_rois_n5_open3"] = dict(H=1080, W=1920, roi=None, N=5, se=3, do_close=False, birds=300, chunk=512,
                                         videos=16)
DEFAULT_CONFIG = "1080p_full_n5_open3"
SEED = 2
# Sub-lines of the default run ("also": [...]): every other BASELINE.json config, measured in the same process
# right after the headline config (fewer steps; the CPU arm only where it is cheap).
ALSO = [("4k_full_n9_oc5", "i32"), ("1080p_roi320x240_n5_open3", "i32"), ("1080p_full_n5_open3", "u8"),
        ("16x1080p_rois_n5_open3", "i32"), ("1080p_swarm500_classify", "i32")]


class Env:
    """Process-group state shared by every config of one bench.py run."""

    def __init__(self):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None

    def init(self):
        import torch
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist

    def barrier(self):
        import torch
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, x):
        import torch
        if self.dist is None:
            return float(x)
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_ok(self, ok):
        return self.max_over_ranks(0.0 if ok else 1.0) == 0.0

    def close(self):
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()


def parity_frames(T, ts, n):
    """Frames to compare with the oracle after the timed region: first, last, both sides of the first
    temporal sub-chunk boundary of the filtering kernel, then a spread."""
    picks = [0, T - 1]
    if 0 < ts < T:
        picks += [ts - 1, ts]
    k = 1
    while len(set(picks)) < min(n, T):
        picks.append((k * 2654435761) % T)
        k += 1
    return sorted(set(picks))[:max(n, 1)] if len(set(picks)) > n else sorted(set(picks))


def parity_check(ctx, dev, halo, par, n_frames, roi_w):
    """Untimed result guard: the submit the bench just timed, compared with the oracle (masks, labels,
    table rows: bit-exact) on a few frames.  Returns {"frames": n, "ok": bool, ...}."""
    from oracle import reference_path as rp
    from swiftwatcher_b200.pipeline import centroids
    ctx.submit(dev, n_halo=halo)
    rows, counts = ctx.collect()
    T = len(counts)
    offs = np.concatenate([[0], np.cumsum(counts)])
    picks = parity_frames(T, ctx.last_subchunk(), n_frames)
    bad = []
    for t in picks:
        host = dev[t:t + halo + 1].cpu().numpy()
        rec = rp.run_path(host[halo:], par, history=list(host[:halo]))[0]
        r = rows[offs[t]:offs[t + 1]]
        exp = rp.props_table(rec["props"])
        ok = np.array_equal(ctx.masks(t, 1)[0], rec["mask"]) and np.array_equal(ctx.labels(t, 1)[0], rec["labels"]) \
            and len(r) == len(exp)
        if ok and len(r):
            got = np.zeros((len(r), 8))
            got[:, 0], got[:, 1], got[:, 2:6] = r["label"], r["area"], r["bbox"]
            got[:, 6:8] = centroids(r)
            ok = np.array_equal(got, exp)
        if not ok:
            bad.append(int(t))
    return {"frames": len(picks), "ok": not bad, "frame_list": [int(t) for t in picks], "mismatches": bad,
            "what": "masks, labels and table rows of these frames of the timed submit == oracle/reference_path.run_path"}


def h2d_attainable(env, host, dev, reps):
    """What the host -> device fabric gives this run: every rank copies its pinned chunk with one plain
    cudaMemcpyAsync per step, all ranks at once (GB/s summed over ranks, max-over-ranks time)."""
    import torch
    dev.copy_(host, non_blocking=True)
    env.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    dt = env.max_over_ranks(time.perf_counter() - t0)
    return env.world * reps * host.numel() / dt / 1e9


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: NVML polled every ~2 ms from a
    thread (the timed region is tens of milliseconds; nvidia-smi -lms cannot sample that fast).
    Falls back to one nvidia-smi query when NVML is unavailable."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []
        self.reasons = set()
        self.stop_flag = False
        self.thread = None
        self.max_mhz = None

    def _poll(self):
        import pynvml as nv
        h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = get_reasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None

    def stop(self):
        if self.thread is None:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=clocks.sm,clocks.max.sm",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout
                a, b = [float(x) for x in out.strip().split(",")]
                return {"sm_mhz": a, "sm_max_mhz": b, "samples": 1, "reasons": [], "source": "nvidia-smi (after)"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["unavailable"]}
        self.stop_flag = True
        self.thread.join(timeout=1.0)
        sm = self.samples
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz,
                "samples": len(sm), "reasons": sorted(self.reasons), "source": "nvml, 2 ms poll"}


def algorithmic_bytes(cfg, h, w, label_bytes=4):
    """Per OUTPUT frame.  Whole path: read BGR once,
    write the uint8 mask once, write int32 labels once = 8*h*w (A' = 5*h*w with the
    reference's uint8 labels)."""
    hw = h * w
    return {
        "path": (4 + label_bytes) * hw,
        "fg_bits": 3 * hw + hw // 8,            # BGR read + 1 bit/px write
        "morph_mask": hw // 8 + hw // 8 + hw,   # bits read, bits write, uint8 mask write
        "ccl_merge": hw // 8, "ccl_rank": hw // 8, "ccl_label": hw // 8,
        "write_labels": hw // 8 + label_bytes * hw,       # bits read + label write
    }


def cpu_baseline_sample(cfg, n_frames, threads=None):
    """The reference's CPU path (oracle port: the same cv2/scipy calls the
    reference makes + np.median/absdiff) on `n_frames` frames of the workload."""
    import cv2
    from oracle import reference_path as rp
    from oracle import synth
    if threads:
        cv2.setNumThreads(threads)
    H, W = cfg["H"], cfg["W"]
    roi = cfg["roi"] or [(0, 0), (W, H)]
    halo = cfg["N"] - 1
    frames = synth.synth_video(SEED, 0, 1000, halo + n_frames, H, W, cfg["birds"])
    par = rp.PathParams(roi, cfg["N"], 15, cfg["se"], True, cfg["do_close"], "u8")
    ref_clf = None
    if cfg.get("classify"):
        # the reference's per-segment classifier loop (segment_classification.py:27-45) on the host
        import torch
        from oracle import reference_classifier as rc
        from swiftwatcher_b200.segment_classification import setup_model
        torch.manual_seed(SEED)
        ref_clf = rc.RefSegmentClassifier(setup_model(2, "cpu").state_dict(), "cpu")
    t0 = time.perf_counter()
    if cfg.get("bg_model") == "rpca":
        out = rp.run_path_rpca(frames[halo:], par, want_images=True)   # the reference as written: IALM + bilateral
    else:
        out = rp.run_path(frames[halo:], par, history=list(frames[:halo]), want_images=True)
    if ref_clf is not None:
        class _Seg:
            pass
        for o in out:
            segs_o = []
            for im in o["crops"]:
                if im.size:
                    sg = _Seg()
                    sg.segment_image, sg.label = im, 0
                    segs_o.append(sg)
            ref_clf(segs_o)
    dt = time.perf_counter() - t0
    segs = sum(len(o["props"]) for o in out)
    return n_frames / dt, dt, segs, cv2.getNumThreads()


def _cpu_chunk_worker(cfg, n_frames, t0, n_steps, barrier, queue):
    """One host process of the parallel CPU arm: its own temporal chunk (with the N-1 frames of
    halo, exactly the multi-GPU partition of DESIGN.md §6), one cv2 thread."""
    import cv2
    from oracle import reference_path as rp
    from oracle import synth
    cv2.setNumThreads(1)
    H, W = cfg["H"], cfg["W"]
    roi = cfg["roi"] or [(0, 0), (W, H)]
    halo = cfg["N"] - 1
    frames = synth.synth_video(SEED, 0, t0 - halo, halo + n_frames, H, W, cfg["birds"])
    par = rp.PathParams(roi, cfg["N"], 15, cfg["se"], True, cfg["do_close"], "u8")
    rp.run_path(frames[halo:halo + 1], par, history=list(frames[:halo]), want_images=True)   # warm the libraries
    segs = 0
    for _ in range(n_steps):
        barrier.wait()
        out = rp.run_path(frames[halo:], par, history=list(frames[:halo]), want_images=True)
        segs += sum(len(o["props"]) for o in out)
        barrier.wait()
    queue.put(segs)


def cpu_baseline_parallel(cfg, frames_per_proc, procs=None, steps=1):
    """The CPU port on ALL host cores: one process per core, each filtering its own temporal
    chunk of the video (frames are independent given their N-1 predecessors).  A step is timed
    from the moment every process holds its frames until the last one is done.
    Returns (frames/s, seconds over all steps, segments, processes)."""
    import multiprocessing as mp
    procs = procs or os.cpu_count() or 1
    mpc = mp.get_context("spawn")          # the parent may hold a CUDA context: no fork
    barrier = mpc.Barrier(procs + 1)
    queue = mpc.Queue()
    ps = [mpc.Process(target=_cpu_chunk_worker,
                      args=(cfg, frames_per_proc, 1000 + i * frames_per_proc, steps, barrier, queue))
          for i in range(procs)]
    for p_ in ps:
        p_.start()
    dt = 0.0
    for _ in range(steps):
        barrier.wait(timeout=1800)
        t0 = time.perf_counter()
        barrier.wait(timeout=1800)
        dt += time.perf_counter() - t0
    segs = sum(queue.get() for _ in ps)
    for p_ in ps:
        p_.join()
    return steps * procs * frames_per_proc / dt, dt, segs, procs


def _parallel_ok(cfg):
    return not cfg.get("classify") and cfg.get("bg_model") != "rpca"


def run_reference(args, cfg, name):
    """--impl reference: the reference's own CPU implementation of the path
    (oracle port; /root/reference is pure Python over cv2/scipy/skimage and does
    not exist on the GPU box), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = max(2, min(8, int(24 // max(args.steps, 1)) or 2)) if cfg["H"] >= 1080 and cfg["roi"] is None else 64
    if cfg.get("classify"):
        sample = 1
    if cfg.get("bg_model") == "rpca":
        sample = cfg["chunk"]
    parallel = _parallel_ok(cfg)
    how = "cv2 thread pool, numpy/scipy single-threaded"
    if parallel:
        # one process per host core, each on its own temporal chunk (+ N-1 halo frames)
        per_proc = 4 if cfg["H"] >= 1080 and cfg["roi"] is None else 64
        if cfg["H"] > 1080:
            per_proc = 2
        how = "one process per host core, each on its own temporal chunk of %d frames + halo, one cv2 thread each" % per_proc
    t_total, n_total, segs, threads = 0.0, 0, 0, 0
    if parallel:
        # the processes warm their libraries on one frame before the first timed step
        fps, t_total, segs, threads = cpu_baseline_parallel(cfg, per_proc, steps=args.steps)
        sample = per_proc * threads
        n_total = sample * args.steps
    else:
        for _ in range(min(args.warmup, 1)):
            cpu_baseline_sample(cfg, 2)
        for _ in range(args.steps):
            fps, dt, s, threads = cpu_baseline_sample(cfg, sample)
            n_total += sample
            t_total += dt
            segs += s
    value = n_total / t_total
    line = {
        "impl": "reference", "metric": "frames/sec (filter + label hot path)", "value": value,
        "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": name, "frames_per_step": sample, "median_n": cfg["N"], "morph": cfg["se"]},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": "%d steps x %d frames of %s on the host CPU (oracle port of the reference's "
                                   "cv2/scipy path; %s)" % (args.steps, sample, name, how)},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host_cpus": os.cpu_count(),
    }
    print(json.dumps(line))


def video_roi(v, H, W):
    """ROI of video v: a 2:1 rectangle (generate_crop_region, image_filtering.py:48-51) between
    200x100 and 640x320, placed by an integer hash so that every run sees the same rectangles."""
    hsh = (v * 2654435761 + 0x9E3779B9) & 0xFFFFFFFF
    w = 200 + 8 * (hsh % 56)                       # 200 .. 640, multiples of 8
    h = w // 2
    x0 = (hsh >> 8) % (W - w)
    y0 = (hsh >> 20) % (H - h)
    return [(int(x0), int(y0)), (int(x0 + w), int(y0 + h))]


def run_multi_video(args, cfg, name, env, label_mode="i32", secondary=False):
    """configs[3]: whole videos are dealt round-robin to the GPUs (no temporal split, no
    collective); on each GPU every video has its own context and CUDA stream, so the small ROI
    kernels of different videos overlap.  A step = one chunk of every video of this rank."""
    import torch
    import swiftwatcher_b200 as swb
    from swiftwatcher_b200.pipeline import synth_frames

    world, rank, local_rank = env.world, env.rank, env.local_rank
    steps = max(3, args.steps // 2) if secondary else args.steps
    e2e_steps = 2 if secondary else args.e2e_steps
    H, W, N, T = cfg["H"], cfg["W"], cfg["N"], cfg["chunk"]
    halo = N - 1
    mine = [v for v in range(cfg["videos"]) if v % world == rank]
    main_stream = torch.cuda.Stream()
    vids = []
    for v in mine:
        roi = video_roi(v, H, W)
        x = torch.empty((halo + T, H, W, 3), dtype=torch.uint8, device="cuda")
        synth_frames(SEED, v, 1000 - halo, halo + T, H, W, cfg["birds"], device=local_rank, out=x)
        ctx = swb.FilterContext((H, W, 3), roi, median_n=N, threshold=15, morph_size=cfg["se"], do_open=True,
                                do_close=cfg["do_close"], label_mode=label_mode, max_frames=T,
                                max_segments=T * 1024, device=local_rank, gpu_share=len(mine))
        st = torch.cuda.Stream()
        ctx.set_stream(st.cuda_stream)
        vids.append(dict(v=v, roi=roi, frames=x, ctx=ctx, stream=st, done=torch.cuda.Event()))
    torch.cuda.synchronize()

    def step(fork):
        for d in vids:
            d["stream"].wait_event(fork)
            d["ctx"].submit(d["frames"], n_halo=halo)

    for _ in range(args.warmup):
        fork = torch.cuda.Event()
        fork.record(main_stream)
        step(fork)
    torch.cuda.synchronize()
    segs = sum(float(d["ctx"].collect()[1].mean()) for d in vids) / max(len(vids), 1)

    sampler = ClockSampler(local_rank)
    launches0 = sum(d["ctx"].launch_count() for d in vids)
    env.barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(main_stream)
    for _ in range(steps):
        step(ev0)
    for d in vids:
        d["done"].record(d["stream"])
        main_stream.wait_event(d["done"])
    ev1.record(main_stream)
    env.barrier()
    clocks = sampler.stop()
    ms = env.max_over_ranks(ev0.elapsed_time(ev1))
    launches = sum(d["ctx"].launch_count() for d in vids) - launches0
    frames_total = cfg["videos"] * T * steps
    fps = frames_total / (ms * 1e-3)
    peak, peak_src = measured_peak()
    px = [(video_roi(v, H, W)[1][0] - video_roi(v, H, W)[0][0]) * (video_roi(v, H, W)[1][1] - video_roi(v, H, W)[0][1])
          for v in range(cfg["videos"])]
    alg_per_step = 8 * sum(px) * T                      # all videos, one chunk each
    path_gbs = alg_per_step * steps / (ms * 1e-3) / 1e9 / world
    # one video alone on one stream, for comparison (what concurrency buys)
    solo = None
    if vids:
        d = vids[0]
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(d["stream"])
        for _ in range(steps):
            d["ctx"].submit(d["frames"], n_halo=halo)
        b.record(d["stream"])
        torch.cuda.synchronize()
        solo = T * steps / (a.elapsed_time(b) * 1e-3)

    # result guard: this rank's first video against the oracle
    check = {"frames": 0, "ok": True}
    if vids and not args.no_parity:
        from oracle import reference_path as rp
        d = vids[0]
        par = rp.PathParams(d["roi"], N, 15, cfg["se"], True, cfg["do_close"], label_mode)
        check = parity_check(d["ctx"], d["frames"], halo, par, 3, W)
    check["ok"] = env.all_ok(check["ok"])

    # end to end: pinned host frames of every video, ROI staged by swb_submit, tables read back
    e2e = None
    if not args.no_e2e:
        Te = min(T, 64)
        hosts = []
        for d in vids:
            hbuf = torch.empty((halo + Te, H, W, 3), dtype=torch.uint8, pin_memory=True)
            hbuf.copy_(d["frames"][:halo + Te])
            hosts.append(hbuf)
        torch.cuda.synchronize()
        h2d = d2h = 0
        for rep_i in range(e2e_steps + 1):
            if rep_i == 1:
                env.barrier()
                t0 = time.perf_counter()
                h2d = d2h = 0
            for d, hbuf in zip(vids, hosts):
                d["ctx"].submit(hbuf, n_halo=halo)
            for d in vids:
                rows_h, counts_h = d["ctx"].collect()
                d2h += rows_h.nbytes + counts_h.nbytes + 4 * (Te + 1)
                (x0, y0), (x1, y1) = d["roi"]
                x0a = x0 & ~31
                wa = ((x1 + 31) & ~31) - x0a
                h2d += (halo + Te) * (y1 - y0) * min(wa, W - x0a) * 3
        torch.cuda.synchronize()
        dt = env.max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": cfg["videos"] * Te * e2e_steps / dt, "unit": "frames/s",
               "h2d_bytes_per_step": h2d // max(e2e_steps, 1), "d2h_bytes_per_step": d2h // max(e2e_steps, 1),
               "steps": e2e_steps, "frames_per_video_per_step": Te,
               "note": "pinned host frames of every video -> swb_submit (ROI staged over PCIe) -> swb_collect"}
        del hosts

    cpu = None
    if not args.no_cpu and world == 1 and not secondary:
        c1 = dict(cfg)
        c1["roi"] = video_roi(0, H, W)
        n_cpu = args.cpu_frames or 200
        cpu_fps, cpu_dt, _, threads = cpu_baseline_sample(c1, n_cpu)
        cpu = {"value": cpu_fps, "unit": "frames/s", "cores": threads, "kind": "port",
               "sample": "%d frames of video 0 (ROI %r, %.1f s) through oracle/reference_path.py; the reference "
                         "processes its videos one after another" % (n_cpu, c1["roi"], cpu_dt)}
    line = {
        "metric": "frames/sec (filter + label hot path)", "value": fps, "unit": "frames/s", "n_gpus": world,
        "steps": steps, "warmup": args.warmup, "ms_per_step": ms / steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": name, "frame": [H, W, 3], "videos": cfg["videos"],
                   "videos_per_gpu": len(mine), "rois": [video_roi(v, H, W) for v in range(cfg["videos"])],
                   "median_n": N, "threshold": 15, "morph": cfg["se"], "labels": label_mode,
                   "frames_per_video_per_step": T, "segments_per_frame": round(segs, 1),
                   "l2": "full 1080p frames of all videos resident (%.1f GB on this GPU) >> 126 MB L2; no flush"
                         % (len(mine) * (halo + T) * H * W * 3 / 1e9),
                   "partition": "whole videos round-robin over GPUs, one context + stream per video, no collective"},
        "roofline": {"bound": "hbm", "kernel": "whole path (launch/latency bound at ROI size)",
                     "achieved": round(path_gbs, 1), "peak": peak, "unit": "GB/s", "frac": round(path_gbs / peak, 4),
                     "path_frac": round(path_gbs / peak, 4),
                     "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_per_step,
                     "single_video_fps_alone": solo},
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "parity_check": check,
    }
    for d in vids:
        d["ctx"].close()
    del vids
    torch.cuda.empty_cache()
    return line


def dropin_latency(cfg, env, batches=40):
    """The reference's own unit of work through the reference's own API: 21-frame batches pushed into
    ``FrameQueue`` -> preprocess_queue -> segment_queue (masks, labels, Segment objects with their colour
    crops attached to every Frame: data_structures.py:171-217), frames popped oldest first."""
    import swiftwatcher_b200.data_structures as ds
    from swiftwatcher_b200.pipeline import synth_frames
    H, W = cfg["H"], cfg["W"]
    roi = cfg["roi"] or [(0, 0), (W, H)]
    queue = ds.FrameQueue(queue_size=21, median_n=cfg["N"], morph_size=cfg["se"], do_close=cfg["do_close"],
                          device=env.local_rank)
    video = synth_frames(SEED, 0, 1000, 63, H, W, cfg["birds"], device=env.local_rank)
    stamps = ["00:00:00.000"] * 21
    segs = 0

    def one(b):
        nonlocal segs
        batch = queue.pinned_batch((H, W, 3), 21)
        np.copyto(batch, video[21 * (b % 3):21 * (b % 3) + 21])       # stands in for the decoder writing the frames
        t0 = time.perf_counter()
        queue.push_list_of_frames(list(batch), list(range(21 * b, 21 * b + 21)), stamps)
        queue.preprocess_queue(roi, (300, 150))
        queue.segment_queue((24, 24), roi)
        dt = time.perf_counter() - t0
        while not queue.is_empty():
            segs += queue.pop_frame().get_num_segments()
        return dt
    for b in range(5):
        one(b)
    segs = 0
    dts = [one(b) for b in range(batches)]
    dt = float(np.median(dts))
    return {"value": 21 / dt, "unit": "frames/s", "us_per_21_frame_batch": round(dt * 1e6, 1),
            "batches": batches, "segments_per_frame": round(segs / (21.0 * batches), 2),
            "h2d_bytes_per_step": 21 * (roi[1][1] - roi[0][1]) * ((((roi[1][0] + 31) & ~31) - (roi[0][0] & ~31)) * 3),
            "d2h_bytes_per_step": 21 * (roi[1][1] - roi[0][1]) * (roi[1][0] - roi[0][0]) * 2,
            "note": "FrameQueue.push_list_of_frames + preprocess_queue + segment_queue per 21-frame batch (median of "
                    "%d batches): masks, uint8 labels, Segment objects and crops attached to the Frames" % batches}


def partition_check(env, cfg, frames_per_rank=48):
    """N > 1 only, untimed: ONE video split by temporal chunk over the N ranks (chunking.run_rank: every
    rank its range + the N-1 halo frames), tables gathered over NCCL (chunking.gather_tables), and the
    result compared with rank 0 processing the whole range alone."""
    import hashlib
    import torch
    import swiftwatcher_b200 as swb
    from swiftwatcher_b200 import chunking
    from swiftwatcher_b200.pipeline import synth_frames
    H, W, N = cfg["H"], cfg["W"], cfg["N"]
    total = frames_per_rank * env.world + 5            # uneven split: the first ranks get one frame more
    chunk = 32

    def read(a, b):
        x = torch.empty((b - a, H, W, 3), dtype=torch.uint8, device="cuda")
        synth_frames(SEED, 7, a, b - a, H, W, cfg["birds"], device=env.local_rank, out=x)
        return x
    with swb.FilterContext((H, W, 3), cfg["roi"], median_n=N, threshold=15, morph_size=cfg["se"], do_open=True,
                           do_close=cfg["do_close"], label_mode="i32", max_frames=chunk, max_segments=chunk * 4096,
                           device=env.local_rank) as ctx:
        rows, counts, t0, t1 = chunking.run_rank(ctx, read, total, env.rank, env.world, chunk_frames=chunk)
        all_rows, all_counts = chunking.gather_tables(rows, counts)
        ok, sha = True, None
        if env.rank == 0:
            ctx.reset()
            solo_rows, solo_counts, _, _ = chunking.run_rank(ctx, read, total, 0, 1, chunk_frames=chunk)
            ok = np.array_equal(solo_rows, all_rows) and np.array_equal(solo_counts, all_counts)
            sha = hashlib.sha256(all_rows.tobytes() + all_counts.tobytes()).hexdigest()[:16]
    ok = env.all_ok(ok)
    return {"ok": ok, "frames": total, "ranks": env.world, "rows": int(len(all_rows)), "table_sha16": sha,
            "what": "one %dx%d video of %d frames split by chunking.run_rank over %d ranks (+%d halo frames each), "
                    "chunking.gather_tables over NCCL == rank 0 alone" % (W, H, total, env.world, N - 1)}


def run_single(args, cfg, name, env, label_mode="i32", secondary=False):
    """One config on this rank's GPU: device-timed steps, per-kernel roofline, result guard, end-to-end
    from pinned host memory, the CPU arm.  Returns the JSON line (identical on every rank)."""
    import torch
    import swiftwatcher_b200 as swb
    from swiftwatcher_b200.pipeline import synth_frames

    world, rank, local_rank = env.world, env.rank, env.local_rank
    steps = max(3, args.steps // 2) if secondary else args.steps
    e2e_steps = 2 if secondary else args.e2e_steps
    H, W, N, T = cfg["H"], cfg["W"], cfg["N"], cfg["chunk"]
    roi = cfg["roi"]
    halo = N - 1
    rh, rw = (H, W) if roi is None else (roi[1][1] - roi[0][1], roi[1][0] - roi[0][0])
    frame_bytes = H * W * 3
    rpca_mode = cfg.get("bg_model") == "rpca"

    # ---- inputs resident in HBM: n_buf distinct chunks of this rank's part of the video
    n_buf = max(1, min(4, int(24e9 // ((halo + T) * frame_bytes))))
    bufs = []
    for b in range(n_buf):
        t0 = 1000 + (rank * 64 + b) * T          # each rank works on its own temporal chunk
        x = torch.empty((halo + T, H, W, 3), dtype=torch.uint8, device="cuda")
        synth_frames(SEED, 0, t0 - halo, halo + T, H, W, cfg["birds"], device=local_rank, out=x)
        bufs.append(x)
    torch.cuda.synchronize()

    ctx = swb.FilterContext((H, W, 3), roi, median_n=N, threshold=15, morph_size=cfg["se"], do_open=True,
                            do_close=cfg["do_close"], label_mode=label_mode, max_frames=T,
                            max_segments=T * 4096, device=local_rank, bg_model=cfg.get("bg_model", "median"))
    stream = torch.cuda.Stream()          # a real (non-default) stream: the library launches on it and the
    ctx.set_stream(stream.cuda_stream)    # CUDA events below are recorded on it

    clf = None
    if cfg.get("classify"):
        from swiftwatcher_b200.segment_classification import SegmentClassifier, setup_model
        torch.manual_seed(SEED)
        clf = SegmentClassifier(setup_model(2, "cpu").state_dict(), device="cuda:%d" % local_rank, batch_size=2048)
    stats = {"segments": 0, "kept": 0}

    def run_step(context, frames):
        """One pass of the hot path over one chunk (+ the classifier where the config has it)."""
        context.submit(frames, n_halo=halo)
        if clf is not None:
            rows_s, _ = context.collect()
            with torch.cuda.stream(stream):
                keep = clf.classify_submit(context, len(rows_s), empty="drop")
                stats["kept"] += int(keep.sum().item())      # device -> host read of the result
            stats["segments"] += len(rows_s)

    for i in range(args.warmup):
        run_step(ctx, bufs[i % n_buf])
    ctx.sync()
    rows, counts = ctx.collect()
    segs_per_frame = float(counts.mean())

    # ---- timed region: K steps, device-timed on the launch stream
    sampler = ClockSampler(local_rank)
    launches0 = ctx.launch_count()
    env.barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if clf is None and not rpca_mode:
        # ~10 ms of spinning on the launch stream in front of the first event: the host enqueues the K steps
        # behind it, so the device-timed region holds back-to-back launches whatever the host thread is doing
        try:
            with torch.cuda.stream(stream):
                torch.cuda._sleep(20_000_000)
        except Exception:                                  # private torch helper: the gate is optional
            pass
    ev0.record(stream)
    for i in range(steps):
        run_step(ctx, bufs[i % n_buf])
    ev1.record(stream)
    env.barrier()
    clocks = sampler.stop()
    ms = env.max_over_ranks(ev0.elapsed_time(ev1))
    launches = ctx.launch_count() - launches0
    fps = world * steps * T / (ms * 1e-3)

    # ---- per-kernel durations (CUDA events between launches, same stream)
    ctx.enable_timing(True)
    per_kernel = {}
    reps = min(steps, 8)
    for i in range(reps):
        ctx.submit(bufs[i % n_buf], n_halo=halo)
        ctx.sync()
        for k, v in ctx.timing().items():
            per_kernel[k] = per_kernel.get(k, 0.0) + v / reps
    ctx.enable_timing(False)
    peak, peak_src = measured_peak()
    alg = algorithmic_bytes(cfg, rh, rw, 1 if label_mode == "u8" else 4)
    dom = max(per_kernel, key=per_kernel.get)
    kernels = {k: {"ms": round(v, 4), "alg_gbs": round(alg[k] * T / (v * 1e-3) / 1e9, 1) if v > 0 else None}
               for k, v in per_kernel.items()}
    dom_ach = alg[dom] * T / (per_kernel[dom] * 1e-3) / 1e9
    path_ach = alg["path"] * fps / world / 1e9
    # `frac` answers BASELINE.json's target (whole path: 8 h w bytes per frame over the step time);
    # the dominant kernel's own figure sits beside it
    roofline = {"bound": "hbm", "kernel": "whole path", "achieved": round(path_ach, 1), "peak": peak, "unit": "GB/s",
                "frac": round(path_ach / peak, 4), "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg["path"] * T,
                "path_achieved": round(path_ach, 1), "path_frac": round(path_ach / peak, 4),
                "dominant_kernel": {"kernel": dom, "achieved": round(dom_ach, 1), "frac": round(dom_ach / peak, 4),
                                    "algorithmic_bytes_per_launch": alg[dom] * T, "traffic": None},
                "kernels": kernels}
    if rpca_mode:
        roofline = rpca_roofline(ctx, roofline, per_kernel, rh, rw, T, peak)
    if clf is not None:
        # the classifier (cuDNN convolutions, library code) dominates this config: time it alone
        crops = torch.randint(0, 256, (4096, 24, 24, 3), dtype=torch.uint8, device="cuda")
        with torch.cuda.stream(stream):
            clf.predict(crops[:2048])
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record(stream)
            clf.predict(crops)
            c1.record(stream)
        torch.cuda.synchronize()
        roofline["classifier"] = {"crops_per_s": round(4096 / (c0.elapsed_time(c1) * 1e-3), 1),
                                  "segments_per_frame": round(segs_per_frame, 1),
                                  "tf32": bool(torch.backends.cudnn.allow_tf32),
                                  "note": "torchvision SqueezeNet1.0 weights and layers as in the reference (float32, "
                                          "cuDNN, library code), evaluated on the window of positions the device-gathered "
                                          "24x24 crop can influence + cached blank-canvas activations in persistent patch buffers (WindowedSqueezeNet: "
                                          "same scores as the padded 224x224 forward pass to 1e-4); the hot-path kernels "
                                          "in `kernels` are the filtering + labelling part of the step"}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file) and label_mode == "i32" and not args.chunk:   # captured at the config's own chunk size
        try:
            tr = json.load(open(traffic_file)).get(name, {})
            roofline["dominant_kernel"]["traffic"] = tr.get(dom)
            if all(k in tr for k in per_kernel):
                roofline["traffic"] = sum(tr[k] for k in per_kernel)
        except Exception:
            pass

    # ---- result guard (untimed): the timed submit against the oracle on a few frames
    check = {"frames": 0, "ok": True, "what": "skipped"}
    if not args.no_parity and not rpca_mode:
        from oracle import reference_path as rp
        par = rp.PathParams(roi or [(0, 0), (W, H)], N, 15, cfg["se"], True, cfg["do_close"], label_mode)
        n_chk = (3 if H > 1080 else 6) if not secondary else (2 if H > 1080 else 3)
        check = parity_check(ctx, bufs[0], halo, par, n_chk if rank == 0 else min(n_chk, 2), rw)
    check["ok"] = env.all_ok(check["ok"])

    # ---- end to end through the C ABI with HOST buffers
    e2e = None
    if not args.no_e2e:
        ctx_h = swb.FilterContext((H, W, 3), roi, median_n=N, threshold=15, morph_size=cfg["se"], do_open=True,
                                  do_close=cfg["do_close"], label_mode=label_mode, max_frames=T,
                                  max_segments=T * 4096, device=local_rank, bg_model=cfg.get("bg_model", "median"))
        host = torch.empty((halo + T, H, W, 3), dtype=torch.uint8, pin_memory=True)
        host.copy_(bufs[0])
        torch.cuda.synchronize()
        two = clf is None and not rpca_mode         # two contexts alternate: chunk i+1 is on its way while chunk i is collected
        ctxs = [ctx_h]
        if two:
            ctxs.append(swb.FilterContext((H, W, 3), roi, median_n=N, threshold=15, morph_size=cfg["se"], do_open=True,
                                          do_close=cfg["do_close"], label_mode=label_mode, max_frames=T,
                                          max_segments=T * 4096, device=local_rank))
        else:
            ctx_h.set_stream(stream.cuda_stream)
        for c in ctxs:                                # untimed: staging buffers, pinned table buffers
            run_step(c, host)
            rows_h, counts_h = c.collect()
        d2h = 0
        env.barrier()
        t0 = time.perf_counter()
        if two:
            for i in range(e2e_steps):
                ctxs[i % 2].submit(host, n_halo=halo)
                if i > 0:
                    rows_h, counts_h = ctxs[(i - 1) % 2].collect()
            rows_h, counts_h = ctxs[(e2e_steps - 1) % 2].collect()
            d2h = rows_h.nbytes + counts_h.nbytes + 4 * (T + 1)
        else:
            for i in range(e2e_steps):
                run_step(ctx_h, host)
                rows_h, counts_h = ctx_h.collect()
                d2h = rows_h.nbytes + counts_h.nbytes + 4 * (T + 1) + (8 if clf is not None else 0)
        torch.cuda.synchronize()
        dt = env.max_over_ranks(time.perf_counter() - t0)
        if roi is None:
            h2d = (halo + T) * frame_bytes
        else:
            x0a = roi[0][0] & ~31
            wa = ((roi[1][0] + 31) & ~31) - x0a
            h2d = (halo + T) * rh * min(wa, W - x0a) * 3
        e2e = {"value": world * e2e_steps * T / dt, "unit": "frames/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
               "note": "pinned host frames -> swb_submit (H2D) -> swb_collect (D2H segment table)"
                       + (" -> device crops -> classifier -> kept count (D2H)" if clf is not None else
                          "; PCIe-bound; two contexts alternate, so that the copy of chunk i+1 is queued while the "
                          "table of chunk i is collected" if two else "; PCIe-bound")}
        for c in ctxs:
            c.close()
        if roi is None and clf is None:
            # the roof of this number: the same bytes with one plain cudaMemcpyAsync per step on every rank at once
            gbs = h2d_attainable(env, host, bufs[0], max(e2e_steps, 2))
            e2e["h2d_gbs"] = round(e2e["value"] * h2d / T / 1e9, 1)
            e2e["attainable_h2d_gbs"] = round(gbs, 1)
            e2e["attainable"] = round(gbs * 1e9 / (h2d / T), 1)
            e2e["frac"] = round(e2e["value"] / e2e["attainable"], 4)
            e2e["attainable_note"] = ("%d rank(s) copying their pinned chunk with one plain cudaMemcpyAsync per step at the "
                                      "same time (sum over ranks / max-over-ranks time): the host-memory / PCIe fabric's "
                                      "limit for this run" % world)
        del host
    e2e_dropin = None
    if cfg.get("dropin") and not args.no_e2e and rank == 0:
        e2e_dropin = dropin_latency(cfg, env)

    # ---- the reference's CPU path on this box's host cores (rank 0, N=1 only)
    cpu = None
    if not args.no_cpu and world == 1 and (not secondary or cfg.get("cpu_in_also")):
        n_cpu = args.cpu_frames or (64 if roi is None and H >= 1080 else 300)
        if H > 1080:
            n_cpu = args.cpu_frames or 12
        if cfg.get("classify"):
            n_cpu = args.cpu_frames or 2
        if rpca_mode:
            n_cpu = T                              # one batch: the decomposition cannot be sampled
            if roi is None and not args.cpu_frames:
                n_cpu = 0                          # ~86 s per 21-frame 1080p batch: only with --cpu-frames
        if n_cpu > 0 and _parallel_ok(cfg):
            # all host cores: one process per core on its own temporal chunk; the single-process figure
            # (the reference is a single-threaded script) is kept beside it
            one_fps, one_dt, _, one_threads = cpu_baseline_sample(cfg, max(2, n_cpu // 8))
            per_proc = max(2, n_cpu // 8) if roi is None else 64
            try:
                cpu_fps, cpu_dt, _, procs = cpu_baseline_parallel(cfg, per_proc)
            except Exception as e:                       # a host that cannot spawn: keep the single-process figure
                sys.stderr.write("parallel CPU arm failed (%r); reporting the single-process sample\n" % (e,))
                cpu_fps, cpu_dt, procs, per_proc = one_fps, one_dt, 1, max(2, n_cpu // 8)
            cpu = {"value": cpu_fps, "unit": "frames/s", "cores": procs, "kind": "port",
                   "sample": "%d processes x %d frames of %s (%.1f s) through oracle/reference_path.py: the reference's "
                             "cv2/scipy calls + np.median/absdiff, each process on its own temporal chunk (+ %d halo frames), "
                             "one cv2 thread each; host has %d CPUs" % (procs, per_proc, name, cpu_dt, halo,
                                                                        os.cpu_count()),
                   "single_process": {"value": one_fps, "cv2_threads": one_threads,
                                      "sample": "%d frames (%.1f s)" % (max(2, n_cpu // 8), one_dt)}}
        elif n_cpu > 0:
            cpu_fps, cpu_dt, _, threads = cpu_baseline_sample(cfg, n_cpu)
            cpu = {"value": cpu_fps, "unit": "frames/s", "cores": threads, "kind": "port",
                   "sample": "%d frames of %s (%.1f s) through oracle/reference_path.py: the reference's cv2/scipy calls "
                             "+ %s; cv2 pool of %d threads, numpy/scipy stages single-threaded; host has %d CPUs"
                             % (n_cpu, name, cpu_dt,
                                "its IALM rpca (numpy/LAPACK svd)" if rpca_mode else "np.median/absdiff",
                                threads, os.cpu_count())}
        else:
            cpu = {"value": None, "unit": "frames/s", "cores": 0, "kind": "port",
                   "sample": "skipped by default (SURVEY.md measured 86 s per 21-frame 1080p batch = 0.24 frames/s for "
                             "the reference as written); pass --cpu-frames 21 to time it here"}

    line = {
        "metric": "frames/sec (filter + label hot path%s)" % (" + segment classification" if clf is not None else ""),
        "value": fps, "unit": "frames/s",
        "n_gpus": world, "steps": steps, "warmup": args.warmup, "ms_per_step": ms / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": name, "frame": [H, W, 3], "roi": roi, "median_n": N,
                   "threshold": 15, "morph": cfg["se"], "open": True, "close": cfg["do_close"],
                   "labels": label_mode, "frames_per_step": T, "halo_frames": halo,
                   "resident_chunks": n_buf, "segments_per_frame": round(segs_per_frame, 1),
                   "l2": "inputs (%.1f GB/step) >> 126 MB L2; no flush" % ((halo + T) * frame_bytes / 1e9),
                   "partition": "temporal chunks, %d-frame halo, no collective" % halo,
                   **({"classifier": "squeezenet1_0 (2 classes), random-init weights, float32, eval mode, TF32 %s"
                                     % ("on" if torch.backends.cudnn.allow_tf32 else "off"),
                       "kept_fraction": round(stats["kept"] / max(stats["segments"], 1), 4)}
                      if clf is not None else {})},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": clocks, "parity_check": check,
    }
    if e2e_dropin is not None:
        line["e2e_dropin"] = e2e_dropin
    ctx.close()
    del bufs, ctx
    torch.cuda.empty_cache()
    return line


def rpca_roofline(ctx, roofline, per_kernel, h, w, T, peak):
    """bg_model = rpca: the step is the IALM iteration loop, not the median kernel.  Algorithmic bytes of one
    iteration's fused pass: read X 1 + read A 8 + write A' 8 + read Y 8 + write Y 8 + write the uint8 image of
    E 1 = 34 bytes per matrix element, n x P elements.  The `fg_bits` timer of the per-kernel table covers
    crop + gray, the whole loop (every iteration: Gram reduce, 21 x 21 eigenproblem, fused pass, stopping test)
    and bilateral + threshold, so time / iterations is an UPPER bound of the fused pass's duration and the
    fraction below a lower bound."""
    stats = ctx.rpca_stats()
    iters = max(stats["iterations"], 1)
    P = h * w
    loop_ms = per_kernel.get("fg_bits", 0.0)
    it_bytes = 34 * T * P
    per_it = loop_ms / iters
    ach = it_bytes / (per_it * 1e-3) / 1e9 if per_it > 0 else None
    roofline = dict(roofline)
    roofline["kernels"] = dict(roofline["kernels"])
    roofline["kernels"]["fg_bits"] = {"ms": round(loop_ms, 4),
                                      "note": "crop+gray, the whole IALM loop and bilateral+threshold"}
    roofline["dominant_kernel"] = {"kernel": "k_rpca_apply_pair (fused apply + next Gram pass of one IALM iteration)",
                                   "achieved": round(ach, 1) if ach else None,
                                   "frac": round(ach / peak, 4) if ach else None,
                                   "algorithmic_bytes_per_launch": it_bytes, "traffic": None,
                                   "ms_per_iteration_upper_bound": round(per_it, 4)}
    roofline["rpca"] = {"iterations": stats["iterations"], "jacobi_sweeps": stats["jacobi_sweeps"],
                        "device_loop": stats["device_loop"], "loop_ms": round(loop_ms, 4)}
    # the whole path in this mode: every iteration streams the batch once (34 B per element) + 8 B per pixel
    total_bytes = iters * it_bytes + 8 * T * P
    step_ms = sum(per_kernel.values())
    roofline["kernel"] = "IALM iterations (34 B per element and iteration) + filtering / labelling (8 B per pixel)"
    roofline["algorithmic_bytes_per_launch"] = total_bytes
    roofline["achieved"] = round(total_bytes / (step_ms * 1e-3) / 1e9, 1)
    roofline["frac"] = round(roofline["achieved"] / peak, 4)
    roofline["path_achieved"], roofline["path_frac"] = roofline["achieved"], roofline["frac"]
    return roofline


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default=DEFAULT_CONFIG, choices=sorted(CONFIGS))
    ap.add_argument("--chunk", type=int, default=0, help="frames per step (0 = config default)")
    ap.add_argument("--label-mode", default="i32", choices=["i32", "u8"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="only the headline config (no `also` sub-lines)")
    ap.add_argument("--no-parity", action="store_true", help="skip the untimed oracle comparison")
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--cpu-frames", type=int, default=0)
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.chunk:
        cfg["chunk"] = args.chunk
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        if cfg.get("videos"):
            cfg["roi"] = video_roi(0, cfg["H"], cfg["W"])
        run_reference(args, cfg, args.config)
        return

    import swiftwatcher_b200 as swb
    if swb.device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device: libswb200 has no CPU path")
    env = Env()
    env.init()
    t_start = time.perf_counter()

    def run(name, c, mode, secondary):
        fn = run_multi_video if c.get("videos") else run_single
        return fn(args, c, name, env, label_mode=mode, secondary=secondary)

    line = run(args.config, cfg, args.label_mode, False)
    default_run = (args.config == DEFAULT_CONFIG and args.label_mode == "i32" and not args.chunk)
    if default_run and not args.no_also:
        line["also"] = []
        for name, mode in ALSO:
            line["also"].append(run(name, dict(CONFIGS[name]), mode, True))
    if env.world > 1 and not args.no_parity:
        line["partition_check"] = partition_check(env, cfg if not cfg.get("videos") else CONFIGS[DEFAULT_CONFIG])
    line["bench_wall_s"] = round(time.perf_counter() - t_start, 1)
    ok = line["parity_check"]["ok"] and all(s["parity_check"]["ok"] for s in line.get("also", [])) and \
        line.get("partition_check", {"ok": True})["ok"]
    if env.rank == 0:
        print(json.dumps(line))
    env.close()
    if not ok:
        sys.stderr.write("bench.py: PARITY CHECK FAILED (see parity_check / partition_check in the line)\n")
        sys.exit(3)


if __name__ == "__main__":
    main()
