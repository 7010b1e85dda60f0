"""Pin the oracle against the reference itself and write tests/golden/*.npz.

TEST INFRASTRUCTURE ONLY.  Runs only where ``/root/reference`` exists (the
build container); the fixtures it writes are committed so that the CPU and GPU
test suites never need the reference at run time.

    python -m oracle.make_golden

What it does
1. imports the reference's own ``swiftwatcher/image_filtering.py`` UNMODIFIED
   from /root/reference (scikit-image is absent: ``oracle/_shim`` provides a
   ``skimage.measure`` whose ``regionprops`` is the oracle restatement);
2. asserts that every restated function in ``oracle/reference_path.py`` equals
   the reference function on seeded inputs (bit-exact);
3. writes golden input/output vectors produced BY THE REFERENCE FUNCTIONS:
   ``stage_kats.npz`` (per-stage known answers) and ``path_*.npz`` (the
   north_star composition on seeded synthetic video: reference functions for
   crop/gray/threshold/opening/labelling/regionprops/crops, oracle functions
   for the stages the reference lacks: median, absdiff, closing).
"""

import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("SWB_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")


def import_reference():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(HERE, "_shim"))
    sys.path.insert(0, REF)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")          # `is 3` SyntaxWarning
        import swiftwatcher.image_filtering as ref_img
    return ref_img


def sparse_blobs(rng, h, w, n, maxs=6):
    img = np.zeros((h, w), np.uint8)
    for _ in range(n):
        y, x = rng.integers(0, h), rng.integers(0, w)
        a, b = rng.integers(1, maxs + 1, 2)
        img[y:y + a, x:x + b] = rng.integers(16, 256)
    return img


def stage_kats(ref, rp):
    rng = np.random.default_rng(20240607)
    out = {}
    # a2 gray
    bgr = rng.integers(0, 256, (96, 128, 3), dtype=np.uint8)
    g_ref = ref.convert_grayscale(bgr)
    assert np.array_equal(g_ref, rp.convert_grayscale(bgr))
    assert np.array_equal(g_ref, rp.gray_fixed_point(bgr))
    out["gray_in"], out["gray_out"] = bgr, g_ref
    # a1 crop
    region = [(17, 9), (101, 77)]
    assert np.array_equal(ref.crop_frame(bgr, region), rp.crop_frame(bgr, region))
    out["crop_region"] = np.array(region)
    out["crop_out"] = np.ascontiguousarray(ref.crop_frame(bgr, region))
    # a5 threshold
    g = rng.integers(0, 64, (64, 80), dtype=np.uint8)
    t_ref = ref.thresh_to_zero(g, 15)
    assert np.array_equal(t_ref, rp.thresh_to_zero(g, 15))
    assert t_ref[g == 15].max(initial=0) == 0 and np.all(t_ref[g == 16] == 16)
    out["thresh_in"], out["thresh_out"] = g, t_ref
    # a6 opening 3x3 (the reference call site) on thresholded noise + blobs
    m = sparse_blobs(rng, 72, 100, 60)
    m[rng.random(m.shape) < 0.02] = 200
    o_ref = ref.grayscale_opening(m, (3, 3))
    assert np.array_equal(o_ref, rp.grayscale_opening(m, (3, 3)))
    out["open_in"], out["open3_out"] = m, o_ref
    o5 = ref.grayscale_opening(m, (5, 5))
    assert np.array_equal(o5, rp.grayscale_opening(m, (5, 5)))
    out["open5_out"] = o5
    out["close3_out"] = rp.grayscale_closing(m, (3, 3))     # no reference fn
    out["close5_out"] = rp.grayscale_closing(m, (5, 5))     # no reference fn
    # a7 labelling: diagonal pair, random blobs, >255 components (uint8 wrap)
    diag = np.zeros((4, 4), np.uint8); diag[1, 1] = 255; diag[2, 2] = 255
    l_ref = ref.cc_labeling(diag, 4)
    assert l_ref.max() == 1, "reference labels 8-connected despite connectivity=4"
    out["cc_diag_in"], out["cc_diag_out"] = diag, l_ref
    blobs = sparse_blobs(rng, 90, 121, 120)
    l_ref = ref.cc_labeling(blobs, 4)
    assert np.array_equal(l_ref, rp.cc_labeling(blobs, 4))
    assert np.array_equal(rp.cc_labeling_i32(blobs), rp.label_order_spec(blobs))
    out["cc_blobs_in"], out["cc_blobs_out"] = blobs, l_ref
    many = np.zeros((120, 160), np.uint8)
    many[rng.random(many.shape) < 0.06] = 255
    l_ref = ref.cc_labeling(many, 4)
    l32 = rp.cc_labeling_i32(many)
    assert l32.max() > 255 and np.array_equal(l_ref, l32.astype(np.uint8))
    assert np.array_equal(l32, rp.label_order_spec(many))
    out["cc_many_in"], out["cc_many_out"], out["cc_many_i32"] = many, l_ref, l32
    # a8 regionprops through the reference entry point (shim -> restatement)
    props = ref.get_segment_properties(l_ref)
    out["props_many"] = rp.props_table(props)
    props_b = ref.get_segment_properties(out["cc_blobs_out"])
    out["props_blobs"] = rp.props_table(props_b)
    # a9 crops: shapes incl. the top-left wrap and bottom-right truncation
    frame = rng.integers(0, 256, (90 + 20, 121 + 30, 3), dtype=np.uint8)
    region = [(10, 5), (10 + 121, 5 + 90)]
    crops_ref = ref.extract_segment_images(props_b, frame, (24, 24), region)
    crops_or = rp.extract_segment_images(props_b, frame, (24, 24), region)
    assert all(np.array_equal(a, b) for a, b in zip(crops_ref, crops_or))
    out["crops_frame"] = frame
    out["crops_region"] = np.array(region)
    out["crops_shapes"] = np.array([c.shape for c in crops_ref])
    out["crops_sums"] = np.array([int(c.sum()) for c in crops_ref])
    np.savez_compressed(os.path.join(GOLD, "stage_kats.npz"), **out)
    print("stage_kats.npz:", len(out), "arrays")


PATH_CASES = {
    # name: synth(seed, video, H, W, birds, T), roi [(x0,y0),(x1,y1)], N, thresh, se, open, close
    "path_roi_n5_open3": dict(seed=1, video=0, H=180, W=320, birds=60, T=14,
                              roi=[(37, 21), (37 + 200, 21 + 120)], N=5, thresh=15,
                              se=3, do_open=1, do_close=0),
    "path_full_n9_oc5": dict(seed=3, video=2, H=135, W=240, birds=50, T=14,
                             roi=[(0, 0), (240, 135)], N=9, thresh=15,
                             se=5, do_open=1, do_close=1),
    "path_dense_n5_open3": dict(seed=5, video=1, H=270, W=480, birds=700, T=8,
                                roi=[(0, 0), (480, 270)], N=5, thresh=15,
                                se=3, do_open=1, do_close=0),
}


def path_case(ref, rp, synth, name, c):
    frames = synth.synth_video(c["seed"], c["video"], 0, c["T"], c["H"], c["W"], c["birds"])
    roi = c["roi"]
    grays = [ref.convert_grayscale(ref.crop_frame(f, roi)) for f in frames]
    masks, lab8, lab32, tabs8, tabs32, counts8, counts32, crop_shapes = [], [], [], [], [], [], [], []
    for t in range(c["T"]):
        win = [grays[i] for i in rp.window_indices(t, c["N"])]
        fg = rp.absdiff(win[-1], rp.temporal_median(win))
        x = ref.thresh_to_zero(fg, c["thresh"])
        if c["do_open"]:
            x = ref.grayscale_opening(x, (c["se"], c["se"]))
        if c["do_close"]:
            x = rp.grayscale_closing(x, (c["se"], c["se"]))
        l8 = ref.cc_labeling(x, 4)
        l32 = rp.cc_labeling_i32(x)
        p8 = ref.get_segment_properties(l8)
        p32 = rp.regionprops(l32)
        crops = ref.extract_segment_images(p8, frames[t], (24, 24), roi)
        masks.append(np.packbits(x > 0, axis=1, bitorder="little"))
        lab8.append(l8); lab32.append(l32.astype(np.int32))
        tabs8.append(rp.props_table(p8)); tabs32.append(rp.props_table(p32))
        counts8.append(len(p8)); counts32.append(len(p32))
        crop_shapes.append(np.array([cc.shape for cc in crops]).reshape(-1, 3))
    # whole-path oracle must reproduce the reference-function composition
    par = rp.PathParams(roi, c["N"], c["thresh"], c["se"], bool(c["do_open"]),
                        bool(c["do_close"]), "u8")
    got = rp.run_path(frames, par)
    for t in range(c["T"]):
        assert np.array_equal(got[t]["labels"], lab8[t]), (name, t)
        assert np.array_equal(rp.props_table(got[t]["props"]), tabs8[t]), (name, t)
    out = dict(
        cfg=np.array([c["seed"], c["video"], c["H"], c["W"], c["birds"], c["T"],
                      roi[0][0], roi[0][1], roi[1][0], roi[1][1], c["N"],
                      c["thresh"], c["se"], c["do_open"], c["do_close"]]),
        frame_sums=frames.reshape(c["T"], -1).sum(axis=1),
        masks_packed=np.stack(masks), labels_u8=np.stack(lab8),
        labels_i32=np.stack(lab32),
        counts_u8=np.array(counts8), counts_i32=np.array(counts32),
        props_u8=np.concatenate(tabs8) if sum(counts8) else np.zeros((0, 8)),
        props_i32=np.concatenate(tabs32) if sum(counts32) else np.zeros((0, 8)),
        crop_shapes=np.concatenate(crop_shapes) if sum(counts8) else np.zeros((0, 3), int),
    )
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(name, "segments/frame (i32):", counts32)


RPCA_CASES = {
    # the reference's real pipeline (rpca + bilateral) on one 21-frame batch, and a short batch
    "rpca_roi_batch21": dict(seed=7, video=0, H=96, W=160, birds=40, T=21, roi=[(16, 8), (144, 88)]),
    "rpca_full_batch8": dict(seed=8, video=1, H=54, W=100, birds=25, T=8, roi=[(0, 0), (100, 54)]),
}


def rpca_case(ref, rp, synth, name, c):
    """data_structures.py:171-217 with the reference's own functions, one batch, oldest first on disk."""
    frames = synth.synth_video(c["seed"], c["video"], 0, c["T"], c["H"], c["W"], c["birds"])
    roi = c["roi"]
    grays = [ref.convert_grayscale(ref.crop_frame(f, roi)) for f in frames]
    sparse_newest_first = ref.rpca(grays[::-1])                     # the queue holds the newest frame at index 0
    mine, iters = rp.rpca(grays[::-1], want_iters=True)
    assert all(np.array_equal(a, b) for a, b in zip(sparse_newest_first, mine)), name
    sparse = sparse_newest_first[::-1]
    bil, masks, lab8, tabs, counts = [], [], [], [], []
    for sp in sparse:
        b = ref.bilateral_blur(sp, 7, 15, 1)
        assert np.array_equal(b, rp.bilateral_blur(sp, 7, 15, 1))
        x = ref.grayscale_opening(ref.thresh_to_zero(b, 15), (3, 3))
        l8 = ref.cc_labeling(x, 4)
        p8 = ref.get_segment_properties(l8)
        bil.append(b); masks.append(np.packbits(x > 0, axis=1, bitorder="little")); lab8.append(l8)
        tabs.append(rp.props_table(p8)); counts.append(len(p8))
    got = rp.run_path_rpca(frames, rp.PathParams(roi, 1, 15, 3, True, False, "u8"))
    for t in range(c["T"]):
        assert np.array_equal(got[t]["labels"], lab8[t]) and np.array_equal(got[t]["rpca"], sparse[t]), (name, t)
    # how far the scalar definition of the bilateral filter is from the (SIMD) library on these images
    bs = sum(int((rp.bilateral_scalar(sp) != b).sum()) for sp, b in zip(sparse, bil))
    out = dict(cfg=np.array([c["seed"], c["video"], c["H"], c["W"], c["birds"], c["T"], roi[0][0], roi[0][1],
                             roi[1][0], roi[1][1], iters]),
               gray=np.stack(grays), rpca=np.stack(sparse), bilateral=np.stack(bil), masks_packed=np.stack(masks),
               labels_u8=np.stack(lab8), counts_u8=np.array(counts),
               props_u8=np.concatenate(tabs) if sum(counts) else np.zeros((0, 8)))
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(name, "IALM iterations:", iters, "segments/frame:", counts, "scalar-vs-cv2 bilateral px:", bs)


def main(only=None):
    os.makedirs(GOLD, exist_ok=True)
    ref = import_reference()
    from oracle import reference_path as rp
    from oracle import synth
    if only in (None, "rpca"):
        for name, c in RPCA_CASES.items():
            rpca_case(ref, rp, synth, name, c)
    if only == "rpca":
        return
    stage_kats(ref, rp)
    for name, c in PATH_CASES.items():
        path_case(ref, rp, synth, name, c)
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None)
