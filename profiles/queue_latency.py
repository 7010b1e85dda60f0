"""Latency of the reference-sized unit of work: one FrameQueue batch (21 frames) of a 320x240
chimney ROI cut from 1080p host frames -> swb_submit + swb_collect (+ masks / labels read-back),
median and RPCA background models.

    python profiles/queue_latency.py
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import swiftwatcher_b200 as swb                      # noqa: E402
from swiftwatcher_b200._lib import pinned_empty      # noqa: E402
from swiftwatcher_b200.pipeline import synth_frames  # noqa: E402

H, W, T = 1080, 1920, 21
roi = [(800, 400), (1120, 640)]
frames = pinned_empty((T, H, W, 3))
frames[:] = synth_frames(2, 0, 100, T, H, W, 300)
for model in ("median", "rpca"):
    with swb.FilterContext((H, W, 3), roi, label_mode="u8", max_frames=T, bg_model=model) as ctx:
        for _ in range(5):
            ctx.submit(frames)
            ctx.collect()
        for what in ("table", "table+masks+labels"):
            t0 = time.perf_counter()
            n = 50
            for _ in range(n):
                ctx.submit(frames)
                rows, counts = ctx.collect()
                if what != "table":
                    ctx.masks()
                    ctx.labels()
            dt = (time.perf_counter() - t0) / n
            print("%-6s %-20s %8.1f us per 21-frame batch  = %9.0f frames/s  (%d segments)"
                  % (model, what, dt * 1e6, T / dt, len(rows)))
