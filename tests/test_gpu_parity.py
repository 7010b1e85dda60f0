"""GPU suite (B200): the CUDA path through the C ABI against the oracle on the
same seeded inputs, against the committed golden vectors, and through
size-independent properties at BASELINE.json's full sizes."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import reference_path as rp
from oracle import synth

import swiftwatcher_b200 as swb
from swiftwatcher_b200 import chunking
from swiftwatcher_b200 import data_structures as ds
from swiftwatcher_b200 import image_filtering as img
from swiftwatcher_b200.pipeline import centroids, synth_frames


@pytest.fixture(scope="module", autouse=True)
def _need_gpu(lib):
    assert swb.device_count() > 0, "GPU tests need a CUDA device (no CPU fallback exists)"


def table_from_rows(rows):
    cen = centroids(rows) if len(rows) else np.zeros((0, 2))
    tab = np.zeros((len(rows), 8))
    tab[:, 0] = rows["label"]
    tab[:, 1] = rows["area"]
    tab[:, 2:6] = rows["bbox"]
    tab[:, 6:8] = cen
    return tab


def check_against_oracle(frames, region, n=5, thresh=15, se=3, do_open=True, do_close=False,
                         mode="i32", history=None, ctx=None, n_halo=None, submit_frames=None,
                         max_segments=0):
    """Run the CUDA path and the oracle on the same frames; compare masks and
    labels bit-exactly, tables exactly (centroids within 1e-5 relative is the
    stated tolerance; integer sums make them equal)."""
    par = rp.PathParams(region, n, thresh, se, do_open, do_close, mode)
    want = rp.run_path(frames, par, history=history)
    own = ctx is None
    if own:
        ctx = swb.FilterContext(frames.shape[1:], region, median_n=n, threshold=thresh,
                                morph_size=se, do_open=do_open, do_close=do_close,
                                label_mode=mode, max_frames=max(len(frames), 1), max_segments=max_segments)
    try:
        if submit_frames is None:
            if history is not None:
                submit_frames = np.ascontiguousarray(np.concatenate([np.stack(history), frames]))
                n_halo = len(history)
            else:
                submit_frames, n_halo = frames, (0 if n_halo is None else n_halo)
        ctx.submit(submit_frames, n_halo=n_halo)
        rows, counts = ctx.collect()
        masks, labels = ctx.masks(), ctx.labels()
    finally:
        if own:
            ctx.close()
    assert len(want) == len(counts)
    o = 0
    for t, rec in enumerate(want):
        assert np.array_equal(masks[t], rec["mask"]), "mask differs at frame %d" % t
        assert labels.dtype == rec["labels"].dtype
        assert np.array_equal(labels[t], rec["labels"]), "labels differ at frame %d" % t
        assert counts[t] == len(rec["props"]), "segment count differs at frame %d" % t
        got = table_from_rows(rows[o:o + counts[t]])
        exp = rp.props_table(rec["props"])
        assert np.array_equal(got[:, :6], exp[:, :6]), "label/area/bbox differ at frame %d" % t
        np.testing.assert_allclose(got[:, 6:], exp[:, 6:], rtol=1e-5, atol=0)   # north_star tolerance
        assert np.array_equal(got[:, 6:], exp[:, 6:])                           # and in fact bit-exact
        assert np.all(rows["frame"][o:o + counts[t]] == t)
        o += counts[t]
    return want


# ---------------------------------------------------------------------------------
# synthetic generator twin
# ---------------------------------------------------------------------------------
def test_synth_cuda_equals_numpy():
    for (seed, video, t0, n, h, w, birds) in [(1, 0, 0, 3, 72, 128, 20), (7, 3, 1000, 2, 50, 77, 9),
                                              (2, 1, 54000, 1, 33, 40, 0)]:
        got = synth_frames(seed, video, t0, n, h, w, birds)
        want = synth.synth_video(seed, video, t0, n, h, w, birds)
        assert np.array_equal(got, want)


# ---------------------------------------------------------------------------------
# single-stage drop-ins against the golden vectors produced by the reference
# ---------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def kats(golden_dir):
    return np.load(os.path.join(golden_dir, "stage_kats.npz"))


def test_stage_gray(kats):
    assert np.array_equal(img.convert_grayscale(kats["gray_in"]), kats["gray_out"])
    g = kats["gray_out"]
    assert img.convert_grayscale(g) is g
    rng = np.random.default_rng(0)
    big = rng.integers(0, 256, (257, 513, 3), dtype=np.uint8)
    assert np.array_equal(img.convert_grayscale(big), rp.convert_grayscale(big))


def test_stage_threshold(kats):
    assert np.array_equal(img.thresh_to_zero(kats["thresh_in"], 15), kats["thresh_out"])
    probe = np.array([[14, 15, 16, 255]], np.uint8)
    assert img.thresh_to_zero(probe, 15).tolist() == [[0, 0, 16, 255]]


def test_stage_morphology(kats):
    m = kats["open_in"]
    assert np.array_equal(img.grayscale_opening(m, (3, 3)), kats["open3_out"])
    assert np.array_equal(img.grayscale_opening(m, (5, 5)), kats["open5_out"])
    assert np.array_equal(img.grayscale_closing(m, (3, 3)), kats["close3_out"])
    assert np.array_equal(img.grayscale_closing(m, (5, 5)), kats["close5_out"])
    assert np.array_equal(img.grayscale_opening(m, (3, 5)), rp.grayscale_opening(m, (3, 5)))


def test_stage_median_absdiff():
    rng = np.random.default_rng(4)
    stack = rng.integers(0, 256, (9, 61, 83), dtype=np.uint8)
    for n in (1, 3, 5, 7, 9):
        assert np.array_equal(img.temporal_median(list(stack[:n])), rp.temporal_median(list(stack[:n])))
    assert np.array_equal(img.absdiff(stack[0], stack[1]), rp.absdiff(stack[0], stack[1]))


def test_stage_cc_labeling(kats):
    assert np.array_equal(img.cc_labeling(kats["cc_diag_in"], 4), kats["cc_diag_out"])
    assert np.array_equal(img.cc_labeling(kats["cc_blobs_in"], 4), kats["cc_blobs_out"])
    assert np.array_equal(img.cc_labeling(kats["cc_many_in"], 4), kats["cc_many_out"])
    assert np.array_equal(img.cc_labeling_i32(kats["cc_many_in"]), kats["cc_many_i32"])


def test_stage_cc_labeling_random_shapes():
    rng = np.random.default_rng(5)
    for _ in range(12):
        h, w = int(rng.integers(1, 90)), int(rng.integers(1, 200))
        a = (rng.random((h, w)) < rng.uniform(0.02, 0.7)).astype(np.uint8) * 200
        assert np.array_equal(img.cc_labeling_i32(a), rp.cc_labeling_i32(a)), (h, w)
    for a in (np.zeros((5, 7), np.uint8), np.full((64, 96), 9, np.uint8), np.eye(33, dtype=np.uint8)):
        assert np.array_equal(img.cc_labeling_i32(a), rp.cc_labeling_i32(a))


def test_stage_cc_labeling_serpentine():
    """Long snake: deep union-find chains across many words and rows."""
    a = np.zeros((101, 300), np.uint8)
    a[::4, :] = 1
    a[2::8, -1] = 1
    a[6::8, 0] = 1
    a[1::8, -1] = 1; a[3::8, -1] = 1
    a[5::8, 0] = 1; a[7::8, 0] = 1
    assert np.array_equal(img.cc_labeling_i32(a), rp.cc_labeling_i32(a))


def test_stage_regionprops(kats):
    for key, tab in (("cc_blobs_out", "props_blobs"), ("cc_many_out", "props_many")):
        props = img.get_segment_properties(kats[key])
        got = np.array([(p.label, p.area, *p.bbox, *p.centroid) for p in props])
        assert np.array_equal(got, kats[tab])
    props = img.get_segment_properties(kats["cc_many_i32"])
    want = rp.regionprops(kats["cc_many_i32"])
    assert [(p.label, p.area, p.bbox, p.centroid) for p in props] == \
           [(p.label, p.area, p.bbox, tuple(p.centroid)) for p in want]
    assert img.get_segment_properties(np.zeros((4, 4), np.uint8)) == []


# ---------------------------------------------------------------------------------
# the fused path against the golden vectors (reference functions' outputs)
# ---------------------------------------------------------------------------------
PATH_CASES = ["path_roi_n5_open3", "path_full_n9_oc5", "path_dense_n5_open3"]


@pytest.mark.parametrize("name", PATH_CASES)
@pytest.mark.parametrize("mode", ["u8", "i32"])
def test_fused_path_against_golden(golden_dir, name, mode):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    c = [int(v) for v in z["cfg"]]
    seed, video, H, W, birds, T = c[:6]
    roi = [(c[6], c[7]), (c[8], c[9])]
    frames = synth_frames(seed, video, 0, T, H, W, birds)      # CUDA generator
    assert np.array_equal(frames.reshape(T, -1).sum(axis=1), z["frame_sums"])
    with swb.FilterContext(frames.shape[1:], roi, median_n=c[10], threshold=c[11], morph_size=c[12],
                           do_open=bool(c[13]), do_close=bool(c[14]), label_mode=mode,
                           max_frames=T) as ctx:
        ctx.submit(frames, n_halo=0)
        rows, counts = ctx.collect()
        masks, labels, bits = ctx.masks(), ctx.labels(), ctx.mask_bits()
    w = roi[1][0] - roi[0][0]
    want_bits = np.unpackbits(z["masks_packed"], axis=2, bitorder="little")[:, :, :w].astype(bool)
    assert np.array_equal(masks > 0, want_bits)
    assert set(np.unique(masks)) <= {0, 255}
    got_bits = np.unpackbits(bits.view(np.uint8), axis=2, bitorder="little")[:, :, :w].astype(bool)
    assert np.array_equal(got_bits, want_bits)
    assert np.array_equal(labels, z["labels_" + mode])
    assert np.array_equal(counts, z["counts_" + mode])
    assert np.array_equal(table_from_rows(rows), z["props_" + mode])


# ---------------------------------------------------------------------------------
# the fused path against the oracle: edge cases
# ---------------------------------------------------------------------------------
def test_unaligned_roi_and_odd_frame_width():
    # W % 16 != 0 -> guarded loads; ROI x0 odd -> bit realignment; ROI touching borders
    frames = synth.synth_video(21, 0, 0, 9, 75, 203, 40)
    for region in ([(33, 7), (190, 70)], [(0, 0), (203, 75)], [(1, 1), (34, 74)], [(170, 3), (203, 75)]):
        check_against_oracle(frames, region, n=5, mode="i32")


def test_aligned_width_unaligned_roi():
    frames = synth.synth_video(22, 0, 0, 8, 64, 256, 40)
    for region in ([(37, 5), (229, 60)], [(32, 0), (256, 64)], [(95, 9), (97, 11)]):
        check_against_oracle(frames, region, n=5, mode="u8")


@pytest.mark.parametrize("n", [1, 3, 5, 7, 9])
def test_median_windows(n):
    frames = synth.synth_video(23, 1, 0, 2 * n + 3, 48, 96, 20)
    check_against_oracle(frames, [(0, 0), (96, 48)], n=n, mode="i32")


@pytest.mark.parametrize("se,do_open,do_close", [(0, False, False), (3, True, True), (3, False, True),
                                                  (5, True, False), (5, True, True)])
def test_morphology_variants(se, do_open, do_close):
    frames = synth.synth_video(24, 0, 0, 8, 70, 130, 45)
    check_against_oracle(frames, [(3, 2), (128, 69)], se=se or 3, do_open=do_open, do_close=do_close)


def test_thresholds():
    frames = synth.synth_video(25, 0, 0, 7, 40, 64, 12)
    for th in (0, 3, 15, 99, 255):
        check_against_oracle(frames, [(0, 0), (64, 40)], thresh=th)


def test_random_noise_frames_many_tiny_components():
    rng = np.random.default_rng(11)
    frames = rng.integers(0, 256, (7, 66, 131, 3), dtype=np.uint8)
    check_against_oracle(frames, [(0, 0), (131, 66)], thresh=60, se=0, do_open=False, mode="i32")
    check_against_oracle(frames, [(0, 0), (131, 66)], thresh=60, se=0, do_open=False, mode="u8")
    check_against_oracle(frames, [(2, 3), (130, 61)], thresh=40, se=3, do_open=True, do_close=True)


def test_wide_frame_uses_global_union_find():
    """Frames wider than 4096 pixels take the non-tiled labelling path."""
    rng = np.random.default_rng(13)
    frames = np.zeros((6, 24, 4300), np.uint8)
    frames[5] = (rng.random((24, 4300)) < 0.3) * 200
    frames[5, 10, :] = 200                       # one component spanning the whole width
    check_against_oracle(frames, [(0, 0), (4300, 24)], thresh=50, se=0, do_open=False, mode="i32")


def test_tall_narrow_frames_cross_tile_boundaries():
    """Narrow ROIs use tall shared-memory tiles (up to 256 block rows): components
    crossing several tile boundaries, U shapes joined only through a lower tile."""
    h, w = 1300, 40
    frames = np.zeros((6, h, w), np.uint8)
    img0 = np.zeros((h, w), np.uint8)
    img0[5:1290, 3] = 200                        # long vertical bars
    img0[100:1200, 20] = 200
    img0[1200, 3:21] = 200                       # joined near the bottom: a tall U
    img0[300:305, 30:35] = 200
    img0[511:514, 8:12] = 200                    # straddles the first tile boundary (row 512)
    rng = np.random.default_rng(14)
    img0[rng.random((h, w)) < 0.03] = 200
    frames[5] = img0
    check_against_oracle(frames, [(0, 0), (w, h)], thresh=50, se=0, do_open=False, mode="i32")
    for ww in (70, 200, 300, 600):
        fr = np.zeros((6, 700, ww), np.uint8)
        fr[5] = (rng.random((700, ww)) < 0.25) * 255
        check_against_oracle(fr, [(0, 0), (ww, 700)], thresh=50, se=0, do_open=False, mode="i32",
                             max_segments=100000)


@pytest.mark.parametrize("se,do_close", [(3, False), (5, True), (3, True)])
def test_wide_rois_span_several_morphology_slabs(se, do_close):
    """ROIs wider than one 30-word slab of the morphology kernel, several strips of rows tall: unaligned ROI origin
    (bit realignment across the slab's edge word), ROI touching the frame's borders, a frame width that is not a
    multiple of 16 pixels behind the ROI."""
    frames = synth.synth_video(35, 0, 0, 7, 150, 2100, 400)
    for region in ([(37, 3), (1300, 148)], [(0, 0), (2048, 150)], [(845, 0), (2100, 150)]):
        check_against_oracle(frames, region, n=5, se=se, do_close=do_close, mode="i32", max_segments=7 * 4096)


def test_gray_input_frames():
    rng = np.random.default_rng(12)
    frames = rng.integers(0, 256, (6, 50, 90), dtype=np.uint8)
    check_against_oracle(frames, [(5, 5), (85, 45)], thresh=50)


def test_blank_and_full_foreground():
    frames = np.zeros((7, 40, 70, 3), np.uint8)
    check_against_oracle(frames, [(0, 0), (70, 40)])
    frames[5:] = 255          # every pixel becomes foreground: one frame-filling component
    check_against_oracle(frames, [(0, 0), (70, 40)], mode="i32")


def test_explicit_halo_equals_whole():
    frames = synth.synth_video(26, 0, 0, 16, 60, 100, 30)
    region = [(4, 4), (99, 58)]
    whole = check_against_oracle(frames, region, n=5)
    part = check_against_oracle(frames[9:], region, n=5, history=list(frames[5:9]))
    for a, b in zip(whole[9:], part):
        assert np.array_equal(a["labels"], b["labels"])
    # a short halo replicates the earliest supplied frame (oracle window policy)
    check_against_oracle(frames[9:], region, n=5, history=list(frames[7:9]))


def test_carried_history_across_submits():
    """Stateful streaming: submits with SWB_HALO_CARRY continue the rolling
    median exactly as one long submit would (FrameQueue's 21-frame batches)."""
    frames = synth.synth_video(27, 0, 0, 19, 54, 120, 30)
    region = [(7, 3), (117, 50)]
    par = rp.PathParams(region, 5, 15, 3, True, False, "i32")
    want = rp.run_path(frames, par)
    with swb.FilterContext(frames.shape[1:], region, label_mode="i32", max_frames=8) as ctx:
        t = 0
        for n in (1, 2, 8, 3, 5):
            ctx.submit(np.ascontiguousarray(frames[t:t + n]))     # n_halo = CARRY
            labels = ctx.labels()
            for i in range(n):
                assert np.array_equal(labels[i], want[t + i]["labels"]), (t, i)
            t += n
        ctx.reset()
        ctx.submit(np.ascontiguousarray(frames[10:14]))
        fresh = rp.run_path(frames[10:14], par)
        assert np.array_equal(ctx.labels()[3], fresh[3]["labels"])


@pytest.mark.parametrize("n,sizes", [(9, (1, 2, 3, 11, 4, 9, 7)), (5, (3, 1, 1, 6, 2, 7)), (3, (1, 1, 5)), (7, (2, 9, 3))])
def test_carried_history_all_windows(n, sizes):
    """SWB_HALO_CARRY with submits shorter and longer than the window, for every median
    kernel variant (grouped N = 5 / 9 loops have partial last groups here)."""
    total = sum(sizes)
    frames = synth.synth_video(31, 0, 0, total, 40, 96, 25)
    region = [(0, 0), (96, 40)]
    par = rp.PathParams(region, n, 15, 3, True, False, "i32")
    want = rp.run_path(frames, par)
    with swb.FilterContext(frames.shape[1:], region, median_n=n, label_mode="i32", max_frames=max(sizes)) as ctx:
        t = 0
        for k in sizes:
            ctx.submit(np.ascontiguousarray(frames[t:t + k]))     # n_halo = CARRY
            masks = ctx.masks()
            for i in range(k):
                assert np.array_equal(masks[i], want[t + i]["mask"]), (t, i)
            t += k


@pytest.mark.parametrize("n,T", [(5, 67), (9, 71), (9, 150), (5, 97)])
def test_many_temporal_subchunks(n, T):
    """Small frames with many frames per submit: K1 splits the submit into temporal sub-chunks
    (multiples of 6 frames) whose last one is a partial group; every frame must still match."""
    frames = synth.synth_video(32, 0, 0, T, 36, 64, 12)
    region = [(0, 0), (64, 36)]
    par = rp.PathParams(region, n, 15, 3, True, False, "i32")
    want = rp.run_path(frames, par)
    with swb.FilterContext(frames.shape[1:], region, median_n=n, label_mode="i32", max_frames=T) as ctx:
        ctx.submit(frames, n_halo=0)
        masks = ctx.masks()
        for i in range(T):
            assert np.array_equal(masks[i], want[i]["mask"]), i
        # and the history left behind continues the stream
        more = synth.synth_video(32, 0, T, 5, 36, 64, 12)
        want2 = rp.run_path(np.concatenate([frames, more]), par)
        ctx.submit(more)
        masks = ctx.masks()
        for i in range(5):
            assert np.array_equal(masks[i], want2[T + i]["mask"]), i


@pytest.mark.parametrize("n,T,se,close,mode", [(9, 71, 5, True, "i32"), (5, 50, 3, False, "u8"), (5, 26, 3, False, "i32"),
                                               (3, 31, 3, True, "i32")])
def test_sub_batches_equal_one_batch(n, T, se, close, mode):
    """A host submit cut into sub-batches that are filtered while later frames are still being
    copied (forced here with the "sub_batch_min_px" option; real frames reach it at >= 64 Mpx per
    sub-batch) gives exactly the oracle's masks, labels and table, the same as the one-batch
    device-resident submit, and leaves the right history behind."""
    import torch
    frames = synth.synth_video(33, 0, 0, T + 6, 44, 100, 25)
    region = [(3, 2), (99, 43)]
    par = rp.PathParams(region, n, 15, se, True, close, mode)
    want = rp.run_path(frames, par)
    with swb.FilterContext(frames.shape[1:], region, median_n=n, morph_size=se, do_close=close, label_mode=mode,
                           max_frames=T) as ctx:
        ctx.set_option("sub_batch_min_px", 1)
        check_against_oracle(frames[:T], region, n=n, se=se, do_close=close, mode=mode, ctx=ctx, n_halo=0)
        ctx.submit(np.ascontiguousarray(frames[T:]))            # CARRY: history written by the last sub-batch
        labels = ctx.labels()
        for i in range(6):
            assert np.array_equal(labels[i], want[T + i]["labels"]), i
        dev = torch.from_numpy(frames).cuda()
        ctx.reset()
        ctx.submit(dev[:T], n_halo=0)
        rows_d, counts_d = ctx.collect()
        labels_d = ctx.labels()
    with swb.FilterContext(frames.shape[1:], region, median_n=n, morph_size=se, do_close=close, label_mode=mode,
                           max_frames=T) as ctx:
        ctx.set_option("host_pipeline", 0)                      # one batch on one stream
        ctx.submit(frames[:T], n_halo=0)
        rows_1, counts_1 = ctx.collect()
        assert np.array_equal(rows_1, rows_d) and np.array_equal(counts_1, counts_d)
        assert np.array_equal(ctx.labels(), labels_d)


def test_fuzz_small_shapes_and_parameters():
    """Seeded sweep over odd frame / ROI shapes (down to one pixel), every window length,
    both structuring elements, open / close combinations, thresholds, host and device frames."""
    import torch
    rng = np.random.default_rng(20241018)
    for case in range(28):
        H, W = int(rng.integers(1, 90)), int(rng.integers(1, 150))
        T = int(rng.integers(1, 11))
        n = int(rng.choice([1, 3, 5, 7, 9]))
        se = int(rng.choice([3, 5]))
        do_open, do_close = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
        if not (do_open or do_close):
            do_open = True
        x0, y0 = int(rng.integers(0, W)), int(rng.integers(0, H))
        x1, y1 = int(rng.integers(x0 + 1, W + 1)), int(rng.integers(y0 + 1, H + 1))
        region = [(x0, y0), (x1, y1)]
        frames = synth.synth_video(100 + case, case % 3, int(rng.integers(0, 50)), T, H, W, int(rng.integers(0, 25)))
        if case % 4 == 0:
            frames = rng.integers(0, 256, size=frames.shape, dtype=np.uint8)          # pure noise: dense labelling
        mode = "u8" if case % 2 else "i32"
        thresh = int(rng.choice([0, 7, 15, 40]))
        submit = torch.from_numpy(frames).cuda() if case % 3 == 0 else None
        try:
            check_against_oracle(frames, region, n=n, thresh=thresh, se=se, do_open=do_open, do_close=do_close,
                                 mode=mode, submit_frames=submit, n_halo=0, max_segments=T * (H * W + 16))
        except AssertionError as e:
            raise AssertionError("case %d: H=%d W=%d T=%d n=%d se=%d open=%s close=%s roi=%r mode=%s thresh=%d: %s"
                                 % (case, H, W, T, n, se, do_open, do_close, region, mode, thresh, e))


@pytest.mark.parametrize("n", [5, 9, 3])
def test_device_roi_tiles_through_the_tensor_map(n):
    """Device-resident frames + a cropped ROI: K1's producer fetches whole-row tiles with TMA tensor
    loads (rows of the ROI are not adjacent in memory).  ROIs whose last tile is partial, one-row
    tiles (wide ROI), ROIs at the frame borders; carried history across submits."""
    import torch
    frames = synth.synth_video(50 + n, 0, 0, 22, 150, 704, 60)
    dev = torch.from_numpy(frames).cuda()
    for region in ([(64, 10), (384, 131)], [(37, 0), (650, 150)], [(0, 3), (40, 147)], [(600, 50), (704, 149)],
                   [(96, 20), (160, 21)]):
        par = rp.PathParams(region, n, 15, 3, True, False, "i32")
        want = rp.run_path(frames, par)
        with swb.FilterContext(frames.shape[1:], region, median_n=n, label_mode="i32", max_frames=16) as ctx:
            t = 0
            for k in (16, 5, 1):
                ctx.submit(dev[t:t + k])                        # carried history
                rows, counts = ctx.collect()
                labels = ctx.labels()
                for i in range(k):
                    assert np.array_equal(labels[i], want[t + i]["labels"]), (region, t + i)
                    assert counts[i] == len(want[t + i]["props"])
                t += k


@pytest.mark.parametrize("n", [5, 9])
def test_device_roi_tiles_over_several_temporal_subchunks(n):
    """Cropped ROI in device-resident frames, long enough that K1 cuts the submit into several temporal
    sub-chunks (the one-wave sizing of pick_ts_roi: every sub-chunk re-reads its n - 1 warm-up frames through
    the tensor map), also with the sub-chunk length forced; every frame against the oracle."""
    import torch
    T = 90
    frames = synth.synth_video(70 + n, 1, 0, T, 150, 704, 60)
    dev = torch.from_numpy(frames).cuda()
    region = [(64, 10), (384, 131)]
    want = rp.run_path(frames, rp.PathParams(region, n, 15, 3, True, False, "i32"))
    for forced in (0, 36):
        with swb.FilterContext(frames.shape[1:], region, median_n=n, label_mode="i32", max_frames=T) as ctx:
            if forced:
                ctx.set_option("temporal_subchunk", forced)
            ctx.submit(dev, n_halo=0)
            rows, counts = ctx.collect()
            ts = ctx.last_subchunk()
            assert 0 < ts < T and (not forced or ts == forced), ts      # several sub-chunks
            labels = ctx.labels()
            for t in range(T):
                assert np.array_equal(labels[t], want[t]["labels"]), (n, forced, t)
                assert counts[t] == len(want[t]["props"])


def test_concurrent_contexts_on_their_own_streams():
    """BASELINE configs[3]: several videos per GPU, one context + CUDA stream each, submits
    interleaved without synchronising in between — every context must still produce exactly
    what it produces alone."""
    import torch
    videos = []
    for v, (H, W, region, n) in enumerate([(96, 160, [(10, 8), (150, 90)], 5), (120, 200, [(33, 20), (180, 100)], 9),
                                            (80, 128, [(0, 0), (128, 80)], 5), (64, 256, [(40, 3), (250, 60)], 3)]):
        frames = synth.synth_video(40 + v, v, 0, 24, H, W, 30)
        par = rp.PathParams(region, n, 15, 3, True, False, "i32")
        videos.append(dict(frames=torch.from_numpy(frames).cuda(), want=rp.run_path(frames, par), region=region, n=n,
                           shape=frames.shape[1:], stream=torch.cuda.Stream()))
    ctxs = [swb.FilterContext(d["shape"], d["region"], median_n=d["n"], label_mode="i32", max_frames=8) for d in videos]
    try:
        for c, d in zip(ctxs, videos):
            c.set_stream(d["stream"].cuda_stream)
        got = [[] for _ in videos]
        for t0 in (0, 8, 16):                                   # three rounds of interleaved asynchronous submits
            for c, d in zip(ctxs, videos):
                c.submit(d["frames"][t0:t0 + 8])                # history carried; no sync between contexts
            for k, c in enumerate(ctxs):
                rows, counts = c.collect()
                got[k].append((c.labels(), counts))
        for k, d in enumerate(videos):
            labels = np.concatenate([g[0] for g in got[k]])
            counts = np.concatenate([g[1] for g in got[k]])
            for t, rec in enumerate(d["want"]):
                assert np.array_equal(labels[t], rec["labels"]), (k, t)
                assert counts[t] == len(rec["props"]), (k, t)
    finally:
        for c in ctxs:
            c.close()


def test_gpu_share_only_sizes_grids():
    """swb_config.gpu_share (contexts working side by side on one GPU) changes the temporal
    sub-chunking of the filtering kernel, never the results."""
    frames = synth.synth_video(61, 0, 0, 8 + 300, 96, 224, 40)
    for region, n in (([(0, 0), (224, 96)], 5), ([(37, 9), (200, 90)], 9)):
        for share in (2, 16, 64):
            ctx = swb.FilterContext(frames.shape[1:], region, median_n=n, label_mode="i32", max_frames=300,
                                    gpu_share=share)
            check_against_oracle(frames[n - 1:n - 1 + 300], region, n=n, history=list(frames[:n - 1]), ctx=ctx)
            ctx.close()


def test_device_resident_input_zero_copy():
    import torch
    frames = synth.synth_video(28, 0, 0, 10, 72, 160, 30)
    dev = torch.from_numpy(frames).cuda()
    for region in ([(0, 0), (160, 72)], [(19, 5), (150, 70)]):
        par = rp.PathParams(region, 5, 15, 3, True, False, "i32")
        want = rp.run_path(frames[4:], par, history=list(frames[:4]))
        with swb.FilterContext(frames.shape[1:], region, label_mode="i32", max_frames=6) as ctx:
            ctx.submit(dev, n_halo=4)
            labels = ctx.labels()
            rows, counts = ctx.collect()
        for t, rec in enumerate(want):
            assert np.array_equal(labels[t], rec["labels"])
            assert counts[t] == len(rec["props"])


def test_capacity_errors_are_loud():
    frames = synth.synth_video(29, 0, 0, 6, 40, 64, 30)
    with swb.FilterContext(frames.shape[1:], None, label_mode="i32", max_frames=4, max_segments=3) as ctx:
        with pytest.raises(swb.SwbError) as e:
            ctx.submit(frames)
        assert e.value.code == -3
        ctx.submit(np.ascontiguousarray(frames[:4]), n_halo=0)
        with pytest.raises(swb.SwbError) as e:
            ctx.collect()
        assert e.value.code == -3


def _blob_video(H, W, blobs, seed, T=7):
    """Static random texture; from frame 5 on, dark textured rectangles (y, x, h, w) appear."""
    rng = np.random.default_rng(seed)
    base = rng.integers(100, 256, (H, W, 3), dtype=np.uint8)
    frames = np.repeat(base[None], T, axis=0)
    for (y, x, h, w) in blobs:
        frames[5:, y:y + h, x:x + w] = rng.integers(0, 60, (h, w, 3), dtype=np.uint8)
    return frames


def _reference_tile(image):
    """What the classifier's first two transforms make of a segment image (segment_classification.py:19-20)."""
    from torchvision import transforms
    return np.asarray(transforms.Resize((24, 24))(transforms.ToPILImage()(np.ascontiguousarray(image))))


def _check_device_crops(frames, region, rows, tiles, rects):
    H, W = frames.shape[1:3]
    kinds = {"exact": 0, "resized": 0, "empty": 0}
    for r, tile, rect in zip(rows, tiles, rects):
        seg = rp.RegionProperties(int(r["label"]), int(r["area"]), tuple(int(v) for v in r["bbox"]), (0, 0))
        want = rp.extract_segment_images([seg], frames[r["frame"]], (24, 24), region)[0]     # the reference's slice
        b = rp.expand_bbox(seg.bbox, (24, 24), region)
        ys, xs = slice(b[0], b[2]).indices(H), slice(b[1], b[3]).indices(W)
        y1, x1 = max(ys[1], ys[0]), max(xs[1], xs[0])
        assert tuple(rect) == (ys[0], xs[0], y1, x1), (tuple(rect), b)
        assert want.shape[:2] == (y1 - ys[0], x1 - xs[0])
        if want.size == 0:
            assert not tile.any()
            kinds["empty"] += 1
        else:
            assert np.array_equal(tile.squeeze(), _reference_tile(want)), (r["bbox"], want.shape)
            kinds["exact" if want.shape[:2] == (24, 24) else "resized"] += 1
    return kinds


def test_device_crops_are_the_reference_segment_images_resized():
    """a9, on the device and exact: every row of the table, whatever its size and position — interior birds
    (24x24 copy), birds larger than 24 px (kept whole by image_filtering.py:349-358, resized by
    segment_classification.py:20), bboxes truncated by the right / bottom frame edge, bboxes within 12 px of
    the top / left edge (empty image: numpy's negative start wraps around)."""
    import torch
    H, W = 200, 320
    blobs = [(60, 50, 9, 14), (100, 100, 40, 30), (20, 180, 30, 100), (150, 20, 45, 25), (3, 3, 8, 8),
             (2, 120, 10, 10), (90, 1, 10, 10), (190, 300, 10, 20), (100, 311, 30, 9), (185, 150, 15, 26),
             (130, 200, 25, 24), (60, 250, 24, 24), (30, 30, 23, 30)]
    frames = _blob_video(H, W, blobs, 40)
    dev = torch.from_numpy(frames).cuda()
    for region in ([(0, 0), (W, H)], [(16, 1), (320, 199)]):
        for mode in ("i32", "u8"):
            with swb.FilterContext(frames.shape[1:], region, label_mode=mode, max_frames=7) as ctx:
                ctx.submit(dev, n_halo=0)
                rows, counts = ctx.collect()
                tiles, rects = ctx.gather_crops(len(rows), 24)
                t_dev = torch.empty((len(rows), 24, 24, 3), dtype=torch.uint8, device="cuda")
                _, r_dev = ctx.gather_crops(len(rows), 24, out=t_dev)
            assert np.array_equal(t_dev.cpu().numpy(), tiles) and np.array_equal(r_dev.cpu().numpy(), rects)
            kinds = _check_device_crops(frames, region, rows, tiles, rects)
            assert kinds["exact"] >= 2 and kinds["resized"] >= 8 and (kinds["empty"] >= 2 or region[0] != (0, 0))


def test_device_crops_tall_thin_and_tiny_frames():
    """Pillow resamples an image more than 100 times taller than wide vertically first; numpy's wrap-around
    of a negative slice start yields a NON-empty image when the frame is narrower than the crop."""
    import torch
    for (H, W) in ((300, 2), (260, 1), (30, 20), (5, 300)):
        frames = np.full((7, H, W), 200, np.uint8)
        frames[5:] = np.random.default_rng(H).integers(0, 100, (2, H, W), dtype=np.uint8)
        region = [(0, 0), (W, H)]
        dev = torch.from_numpy(frames).cuda()
        with swb.FilterContext(frames.shape[1:], region, label_mode="i32", max_frames=7) as ctx:
            ctx.submit(dev, n_halo=0)
            rows, _ = ctx.collect()
            tiles, rects = ctx.gather_crops(len(rows), 24)
        assert len(rows) >= 2
        kinds = _check_device_crops(frames, region, rows, tiles, rects)
        assert kinds["resized"] + kinds["empty"] == len(rows)


# ---------------------------------------------------------------------------------
# the drop-in FrameQueue flow (data_structures.py:171-217, __main__.py:71-96)
# ---------------------------------------------------------------------------------
def test_framequeue_drop_in_flow():
    frames = synth.synth_video(31, 0, 0, 30, 90, 160, 30)
    region = [(16, 8), (150, 80)]
    par = rp.PathParams(region, 5, 15, 3, True, False, "u8")
    want = rp.run_path(frames, par, want_images=True)
    queue = ds.FrameQueue(queue_size=21)
    t = 0
    seen = 0
    while t < len(frames):
        batch = list(frames[t:t + 21])
        queue.push_list_of_frames(batch, list(range(t, t + len(batch))), ["00:00:00.000"] * len(batch))
        queue.preprocess_queue(region, (300, 150))
        queue.segment_queue((24, 24), region)
        while not queue.is_empty():
            f = queue.pop_frame()
            rec = want[f.frame_number]
            assert np.array_equal(f.get_processed_frame("cc_labeling"), rec["labels"])
            assert np.array_equal(f.get_processed_frame("mask"), rec["mask"])
            assert np.shares_memory(f.get_processed_frame("crop"), f.frame)
            assert f.get_num_segments() == len(rec["props"])
            for s, p, c in zip(f.segments, rec["props"], rec["crops"]):
                assert (s.label, s.area, s.bbox) == (p.label, p.area, p.bbox)
                assert s.centroid == tuple(p.centroid)
                assert np.array_equal(s.segment_image, c)
                assert s.parent_frame_number == f.frame_number
            seen += 1
        t += 21
    assert seen == len(frames) and queue.frames_processed == len(frames)


# ---------------------------------------------------------------------------------
# full-size properties (BASELINE.json configs 2 and 3): no oracle at this size
# ---------------------------------------------------------------------------------
def _full_size_properties(H, W, n, se, do_close, T, birds):
    import torch
    halo = n - 1
    dev = torch.empty((halo + T, H, W, 3), dtype=torch.uint8, device="cuda")
    synth_frames(3, 0, 100, halo + T, H, W, birds, out=dev)
    with swb.FilterContext((H, W, 3), None, median_n=n, morph_size=se, do_close=do_close,
                           label_mode="i32", max_frames=T) as ctx:
        ctx.submit(dev, n_halo=halo)
        rows, counts = ctx.collect()
        masks, labels = ctx.masks(), ctx.labels()
        # temporal chunks with halo == one submit (the multi-GPU partition, run on one GPU)
        rows2, counts2, _, _ = chunking.run_rank(
            ctx, lambda a, b: dev[a:b], halo + T, 0, 1, chunk_frames=max(T // 3, 1))
    assert counts.sum() == len(rows) and counts.min() > 0
    assert np.array_equal((labels > 0), (masks == 255))
    assert rows["area"].sum() == int((masks == 255).sum())            # checksum of checksums
    o = 0
    for t in range(T):
        r = rows[o:o + counts[t]]
        assert np.array_equal(r["label"], np.arange(1, counts[t] + 1))
        assert labels[t].max() == counts[t]
        # spot-check a few segments exactly against the dense label image
        for k in (0, counts[t] // 2, counts[t] - 1):
            ys, xs = np.nonzero(labels[t] == r["label"][k])
            assert r["area"][k] == len(ys)
            assert tuple(r["bbox"][k]) == (ys.min(), xs.min(), ys.max() + 1, xs.max() + 1)
            assert r["sum_row"][k] == ys.sum() and r["sum_col"][k] == xs.sum()
        o += counts[t]
    # run_rank processed frames [0, halo+T) of `dev` as a video of its own: its frames
    # halo.. coincide with the submit above (full window available in both)
    keep = rows2["frame"] >= halo
    r2 = rows2[keep].copy()
    r2["frame"] -= halo
    assert np.array_equal(counts2[halo:], counts)
    assert np.array_equal(r2, rows)
    # frame 0 label image against the oracle's labeller (per-frame, affordable)
    assert np.array_equal(labels[0], rp.cc_labeling_i32(masks[0]))


def test_full_size_1080p_n5_open3():
    _full_size_properties(1080, 1920, 5, 3, False, 6, 300)


def test_full_size_4k_n9_openclose5():
    _full_size_properties(2160, 3840, 9, 5, True, 3, 600)
