"""Short end-to-end probe of the TMA-staged morphology kernel (run under `timeout 60`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import swiftwatcher_b200 as swb
from swiftwatcher_b200.pipeline import synth_frames
from oracle import reference_path as rp
for (H, W, roi, se, close) in [(150, 2048, None, 3, False), (150, 2100, [(37, 3), (1300, 148)], 5, True), (150, 2100, [(845, 0), (2100, 150)], 3, True)]:
    T = 7
    dev = torch.empty((T, H, W, 3), dtype=torch.uint8, device="cuda")
    synth_frames(3, 0, 0, T, H, W, 300, out=dev)
    frames = dev.cpu().numpy()
    region = roi or [(0, 0), (W, H)]
    want = rp.run_path(frames, rp.PathParams(region, 5, 15, se, True, close, "i32"))
    with swb.FilterContext((H, W, 3), roi, morph_size=se, do_close=close, label_mode="i32", max_frames=T) as ctx:
        ctx.submit(dev, n_halo=0)
        rows, counts = ctx.collect()
        m = ctx.masks()
        print(H, W, roi, se, close, "masks equal:", all(np.array_equal(m[t], want[t]["mask"]) for t in range(T)), flush=True)
