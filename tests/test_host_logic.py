"""CPU suite: host-side logic (chunk planning, table handling, multi-rank
gather over gloo) with the oracle standing in for the GPU context."""
import os
import sys

import numpy as np
import pytest

from oracle import reference_path as rp
from oracle import synth
from swiftwatcher_b200 import chunking
from swiftwatcher_b200._lib import SEGMENT_DTYPE
from swiftwatcher_b200.image_filtering import crop_frame, expand_bbox, extract_segment_images
from swiftwatcher_b200.pipeline import centroids, props_from_rows

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleContext:
    """Stands in for FilterContext on CPU (tests only): same submit/collect
    contract, computed by the oracle."""

    def __init__(self, crop_region, median_n, max_frames):
        self.params = rp.PathParams(crop_region, median_n, 15, 3, True, False, "i32")
        self.median_n = median_n
        self.max_frames = max_frames
        self._out = None

    def submit(self, frames, n_halo=0):
        self._out = rp.run_path(frames[n_halo:], self.params, history=list(frames[:n_halo]))

    def collect(self):
        rows = []
        counts = []
        for t, rec in enumerate(self._out):
            lab = rec["labels"]
            counts.append(len(rec["props"]))
            for p in rec["props"]:
                ys, xs = np.nonzero(lab == p.label)
                r = np.zeros((), SEGMENT_DTYPE)
                r["frame"], r["label"], r["area"], r["bbox"] = t, p.label, p.area, p.bbox
                r["sum_row"], r["sum_col"] = ys.sum(), xs.sum()
                rows.append(r)
        rows = np.array(rows, dtype=SEGMENT_DTYPE) if rows else np.empty(0, SEGMENT_DTYPE)
        return rows, np.array(counts, np.int32)


def test_rank_range_partitions_exactly():
    for total in (0, 1, 7, 100, 54000):
        for world in (1, 2, 3, 8):
            spans = [chunking.rank_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_plan_chunks_halo():
    plan = list(chunking.plan_chunks(0, 50, 21, 5))
    assert plan == [(0, 0, 0, 21), (17, 4, 21, 21), (38, 4, 42, 8)]
    plan = list(chunking.plan_chunks(2, 10, 100, 9))
    assert plan == [(0, 2, 2, 8)]


def test_centroid_from_integer_sums_is_numpy_mean():
    rng = np.random.default_rng(0)
    lab = np.zeros((300, 500), np.int32)
    lab[rng.random(lab.shape) < 0.4] = 1
    ys, xs = np.nonzero(lab)
    row = np.zeros(1, SEGMENT_DTYPE)
    row["area"], row["sum_row"], row["sum_col"] = len(ys), ys.sum(), xs.sum()
    want = np.stack([ys, xs], 1).mean(axis=0)
    assert np.array_equal(centroids(row)[0], want)
    p = props_from_rows(row)[0]
    assert p.centroid == tuple(want)


def test_expand_bbox_and_crops_match_oracle():
    rng = np.random.default_rng(1)
    frame = rng.integers(0, 256, (120, 200, 3), dtype=np.uint8)
    region = [(10, 20), (150, 100)]
    boxes = [(0, 0, 3, 5), (70, 130, 80, 140), (5, 5, 45, 40), (78, 138, 80, 140), (2, 3, 7, 10)]
    segs = [rp.RegionProperties(i + 1, 1, b, (0.0, 0.0)) for i, b in enumerate(boxes)]
    for b in boxes:
        assert expand_bbox(b, (24, 24), region) == rp.expand_bbox(b, (24, 24), region)
    mine = extract_segment_images(segs, frame, (24, 24), region)
    want = rp.extract_segment_images(segs, frame, (24, 24), region)
    assert all(np.array_equal(a, b) for a, b in zip(mine, want))
    assert np.shares_memory(crop_frame(frame, region), frame)


def _video():
    return synth.synth_video(11, 0, 0, 23, 96, 160, 25)


def test_chunked_run_equals_whole():
    frames = _video()
    region = [(8, 4), (150, 90)]
    whole_ctx = OracleContext(region, 5, 64)
    whole_ctx.submit(frames)
    rows_w, counts_w = whole_ctx.collect()
    for world in (1, 2, 3):
        parts = []
        for r in range(world):
            ctx = OracleContext(region, 5, 7)
            rows, counts, t0, t1 = chunking.run_rank(ctx, lambda a, b: frames[a:b], len(frames), r, world)
            parts.append((rows, counts))
        rows = np.concatenate([p[0] for p in parts])
        counts = np.concatenate([p[1] for p in parts])
        assert np.array_equal(counts, counts_w)
        assert np.array_equal(rows, rows_w)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = _video()
    region = [(8, 4), (150, 90)]
    ctx = OracleContext(region, 5, 6)
    rows, counts, _, _ = chunking.run_rank(ctx, lambda a, b: frames[a:b], len(frames), rank, world)
    rows, counts = chunking.gather_tables(rows, counts)
    q.put((rank, rows.tobytes(), counts.tobytes()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather_restores_global_order():
    import torch.multiprocessing as mp
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctxm.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    frames = _video()
    ctx = OracleContext([(8, 4), (150, 90)], 5, 64)
    ctx.submit(frames)
    rows_w, counts_w = ctx.collect()
    for _, rb, cb in got:
        assert np.array_equal(np.frombuffer(rb, SEGMENT_DTYPE), rows_w)
        assert np.array_equal(np.frombuffer(cb, np.int32), counts_w)


def test_bench_parallel_cpu_arm_two_processes():
    """bench.py's CPU reference arm: one process per core, each on its own temporal chunk
    (+ N-1 halo frames); here two processes on a tiny ROI workload."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    cfg = dict(H=96, W=160, roi=[(16, 8), (144, 88)], N=5, se=3, do_close=False, birds=20, chunk=8)
    fps, dt, segs, procs = bench.cpu_baseline_parallel(cfg, 6, procs=2, steps=2)
    assert procs == 2 and dt > 0 and fps == pytest.approx(2 * 2 * 6 / dt)
    # the segments the two processes found are the oracle's for the same frames (two steps each)
    par = rp.PathParams(cfg["roi"], 5, 15, 3, True, False, "u8")
    want = 0
    for i in range(2):
        fr = synth.synth_video(bench.SEED, 0, 1000 + 6 * i - 4, 4 + 6, 96, 160, 20)
        want += sum(len(o["props"]) for o in rp.run_path(fr[4:], par, history=list(fr[:4]), want_images=True))
    assert segs == 2 * want
