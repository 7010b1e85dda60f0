"""Small end-to-end runs for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck python profiles/sanitize_small.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import swiftwatcher_b200 as swb                      # noqa: E402
from swiftwatcher_b200.pipeline import synth_frames  # noqa: E402

frames = synth_frames(3, 0, 0, 26, 90, 170, 25)
for (n, se, close, mode, roi) in [(5, 3, False, "i32", [(7, 3), (160, 85)]), (9, 5, True, "u8", None),
                                  (3, 3, True, "i32", [(32, 0), (128, 64)]), (7, 3, False, "u8", [(1, 1), (33, 40)])]:
    with swb.FilterContext(frames.shape[1:], roi, median_n=n, morph_size=se, do_close=close, label_mode=mode,
                           max_frames=26) as ctx:
        ctx.submit(frames[:20], n_halo=0)
        rows, counts = ctx.collect()
        ctx.submit(frames[20:])                      # carried history
        rows2, counts2 = ctx.collect()
        ctx.masks(); ctx.labels()
        print("median n=%d se=%d %s: %d + %d segments" % (n, se, mode, len(rows), len(rows2)))
noise = np.random.default_rng(0).integers(0, 256, size=(6, 64, 96, 3), dtype=np.uint8)
with swb.FilterContext(noise.shape[1:], None, label_mode="i32", max_frames=6, max_segments=6 * 64 * 96) as ctx:
    ctx.submit(noise, n_halo=0)                      # dense tiles: the listed-tile path of the labeller
    print("noise: %d segments" % len(ctx.collect()[0]))
with swb.FilterContext(frames.shape[1:], [(10, 6), (150, 80)], label_mode="u8", max_frames=21, bg_model="rpca") as ctx:
    ctx.submit(frames[:21], n_halo=0)
    print("rpca: %d segments" % len(ctx.collect()[0]))
