"""Time swb_stage_rpca (IALM on the GPU) per batch of 21 gray frames at several frame sizes: shows the fixed cost per
iteration (launches, two stream syncs, the 21 x 21 eigenproblem on the host) against the streaming cost.

    python profiles/rpca_timing.py
"""
import sys, time
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as np, ctypes as C
from swiftwatcher_b200 import _lib
from swiftwatcher_b200._lib import ptr, check
from oracle import synth
lib = _lib.load()
for (h, w) in [(8, 16), (160, 320), (540, 960), (1080, 1920)]:
    fr = synth.synth_video(2, 0, 0, 21, h, w, max(3, h * w // 7000))
    g = np.ascontiguousarray(fr[..., 1])
    out = np.empty_like(g); it = C.c_int32(0)
    check(lib.swb_stage_rpca(0, ptr(g), 21, h, w, ptr(out), C.byref(it)))
    t = time.perf_counter()
    for _ in range(3):
        check(lib.swb_stage_rpca(0, ptr(g), 21, h, w, ptr(out), C.byref(it)))
    dt = (time.perf_counter() - t) / 3
    print(h, w, "iters", it.value, "ms per batch %.2f" % (dt * 1e3), "us per iteration %.1f" % (dt * 1e6 / max(it.value, 1)))
