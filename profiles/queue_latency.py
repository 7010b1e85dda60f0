"""Latency of the reference-sized unit of work: one FrameQueue batch (21 frames) of a 320x240
chimney ROI cut from 1080p host frames.

  * C ABI only: swb_submit + swb_collect (table), swb_submit + swb_collect_all (table + masks + labels
    into page-locked memory, one synchronisation), and the round-1 sequence (collect, get_masks, get_labels:
    three synchronisations, pageable destinations) for comparison;
  * the drop-in itself: FrameQueue.push_list_of_frames + preprocess_queue + segment_queue (masks, labels,
    Segment objects with their crops on every Frame), frames decoded in place into a pinned batch.

    python profiles/queue_latency.py
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import swiftwatcher_b200 as swb                      # noqa: E402
import swiftwatcher_b200.data_structures as ds       # noqa: E402
from swiftwatcher_b200._lib import pinned_empty      # noqa: E402
from swiftwatcher_b200.pipeline import synth_frames  # noqa: E402

H, W, T = 1080, 1920, 21
roi = [(800, 400), (1120, 640)]
frames = pinned_empty((T, H, W, 3))
frames[:] = synth_frames(2, 0, 100, T, H, W, 300)
N = 100


def timed(fn):
    for _ in range(10):
        fn()
    ts = []
    for _ in range(N):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts))


for model in ("median", "rpca"):
    with swb.FilterContext((H, W, 3), roi, label_mode="u8", max_frames=T, bg_model=model) as ctx:
        m = pinned_empty((T, ctx.roi_h, ctx.roi_w), np.uint8)
        l = pinned_empty((T, ctx.roi_h, ctx.roi_w), np.uint8)
        rows = [None]

        def table():
            ctx.submit(frames)
            rows[0] = ctx.collect()[0]

        def old_three_calls():
            ctx.submit(frames)
            ctx.collect()
            ctx.masks()
            ctx.labels()

        def one_call():
            ctx.submit(frames)
            ctx.collect_all(m, l)
        for name, fn in (("table", table), ("table+masks+labels, r01 (3 syncs)", old_three_calls),
                         ("table+masks+labels, collect_all", one_call)):
            dt = timed(fn)
            print("%-6s %-36s %8.1f us per 21-frame batch  = %9.0f frames/s  (%d segments)"
                  % (model, name, dt * 1e6, T / dt, len(rows[0])))

queue = ds.FrameQueue(queue_size=T)
stamps = ["00:00:00.000"] * T
state = {"b": 0, "segs": 0}


def dropin():
    batch = queue.pinned_batch((H, W, 3), T)
    np.copyto(batch, frames)                          # the decoder's work, part of neither path's filtering time
    t0 = time.perf_counter()
    queue.push_list_of_frames(list(batch), list(range(T)), stamps)
    queue.preprocess_queue(roi, (300, 150))
    queue.segment_queue((24, 24), roi)
    dt = time.perf_counter() - t0
    state["segs"] = 0
    while not queue.is_empty():
        state["segs"] += queue.pop_frame().get_num_segments()
    return dt


for _ in range(10):
    dropin()
dt = float(np.median([dropin() for _ in range(N)]))
print("median FrameQueue push + preprocess_queue + segment_queue   %8.1f us per 21-frame batch  = %9.0f frames/s  "
      "(%d Segment objects with crops)" % (dt * 1e6, T / dt, state["segs"]))
queue.close()
