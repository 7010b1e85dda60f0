"""GPU suite, part 3 (B200): the finished drop-in around the fused path — one-call read-back, the
reference's grey intermediates on demand, overhanging crop regions, the decode-ahead ring with the
batch submitted ahead of time, ``--export``.  Reference: swiftwatcher/data_structures.py:65-217,
swiftwatcher/__main__.py:71-96."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import reference_path as rp
from oracle import synth

import swiftwatcher_b200 as swb
from swiftwatcher_b200 import data_structures as ds
from swiftwatcher_b200._lib import pinned_empty
from swiftwatcher_b200.io_video import ArrayReader, IngestRing


def test_collect_all_equals_the_separate_calls():
    frames = synth.synth_video(71, 0, 0, 21, 96, 160, 40)
    region = [(9, 7), (150, 90)]
    for mode in ("u8", "i32"):
        with swb.FilterContext(frames.shape[1:], region, label_mode=mode, max_frames=21) as ctx:
            ctx.submit(frames, n_halo=0)
            rows, counts = ctx.collect()
            masks, labels = ctx.masks(), ctx.labels()
            m = pinned_empty(masks.shape, np.uint8)
            l = pinned_empty(labels.shape, labels.dtype)
            for _ in range(3):                      # the speculative row copy adapts to the table size
                ctx.submit(frames, n_halo=0)
                rows2, counts2, m2, l2 = ctx.collect_all(m, l)
                assert m2 is m and l2 is l
                assert np.array_equal(rows2, rows) and np.array_equal(counts2, counts)
                assert np.array_equal(m, masks) and np.array_equal(l, labels)
            ctx.submit(frames, n_halo=0)
            ctx.collect_begin(m, l)
            with pytest.raises(swb.SwbError):       # no submit between begin and end
                ctx.submit(frames, n_halo=0)
            rows3, counts3, _, _ = ctx.collect_end()
            assert np.array_equal(rows3, rows) and np.array_equal(counts3, counts)


def test_u8_table_is_merged_on_the_device_like_regionprops_of_the_truncated_image():
    """> 255 components per frame: labels k, k + 256, ... are one region of labels.astype(uint8)
    (image_filtering.py:329,335); also through host sub-batches (chained offsets)."""
    rng = np.random.default_rng(5)
    frames = np.zeros((12, 120, 260), np.uint8)
    frames[5:] = (rng.random((7, 120, 260)) < 0.08) * 200
    region = [(0, 0), (260, 120)]
    par = rp.PathParams(region, 5, 50, 3, False, False, "u8")
    want = rp.run_path(frames, par)
    assert max(len(rp.regionprops(rp.cc_labeling_i32(w["mask"]))) for w in want) > 300
    for min_px in (None, 1):
        with swb.FilterContext(frames.shape[1:], region, threshold=50, do_open=False, morph_size=0, label_mode="u8",
                               max_frames=12, max_segments=12 * 4096) as ctx:
            if min_px:
                ctx.set_option("sub_batch_min_px", min_px)
            ctx.submit(frames, n_halo=0)
            rows, counts, masks, labels = ctx.collect_all()
            o = 0
            for t, rec in enumerate(want):
                assert np.array_equal(labels[t], rec["labels"]), t
                exp = rp.props_table(rec["props"])
                r = rows[o:o + counts[t]]
                assert counts[t] == len(exp) <= 255
                assert np.array_equal(r["label"], exp[:, 0]) and np.array_equal(r["area"], exp[:, 1])
                assert np.array_equal(r["bbox"], exp[:, 2:6])
                assert np.array_equal(r["sum_row"] / r["area"], exp[:, 6]) and np.all(r["frame"] == t)
                o += counts[t]


def test_overhanging_crop_region_is_truncated_like_the_reference():
    """generate_crop_region can run past the right / bottom edge (image_filtering.py:48-51); crop_frame is
    a numpy slice, so the reference silently truncates.  (this used to raise in segment_queue.)"""
    frames = synth.synth_video(72, 0, 0, 21, 90, 160, 30)
    region = [(100, 40), (190, 120)]                       # 30 columns and 30 rows outside the frame
    inside = [(100, 40), (160, 90)]
    want = rp.run_path(frames, rp.PathParams(inside, 5, 15, 3, True, False, "u8"))
    queue = ds.FrameQueue(queue_size=21)
    queue.push_list_of_frames(list(frames), list(range(21)), ["00:00:00.000"] * 21)
    queue.preprocess_queue(region, (300, 150))
    queue.segment_queue((24, 24), region)
    while not queue.is_empty():
        f = queue.pop_frame()
        assert f.get_processed_frame("crop").shape == (50, 60, 3)
        assert np.array_equal(f.get_processed_frame("cc_labeling"), want[f.frame_number]["labels"])
        assert f.get_num_segments() == len(want[f.frame_number]["props"])
    with swb.FilterContext(frames.shape[1:], region, max_frames=21) as ctx:
        assert (ctx.roi_h, ctx.roi_w) == (50, 60) and ctx.crop_region == inside
    with pytest.raises(ValueError):
        swb.FilterContext(frames.shape[1:], [(-4, 0), (50, 50)], max_frames=4)


@pytest.mark.parametrize("n,se,close", [(5, 3, False), (9, 5, True)])
def test_reference_intermediates_on_demand(n, se, close):
    """processed_frames["grayscale" | "thresh_15" | "opened"] (data_structures.py:183-203) are not
    materialised by the fused kernels but are there when somebody asks, across batch borders."""
    frames = synth.synth_video(73, 0, 0, 26, 80, 128, 25)
    region = [(5, 6), (120, 75)]
    queue = ds.FrameQueue(queue_size=13, median_n=n, morph_size=se, do_close=close)
    grays = [rp.convert_grayscale(rp.crop_frame(f, region)) for f in frames]
    for t0 in (0, 13):
        queue.push_list_of_frames(list(frames[t0:t0 + 13]), list(range(t0, t0 + 13)), ["x"] * 13)
        queue.preprocess_queue(region, None)
        queue.segment_queue((24, 24), region)
        assert [k for k in queue[0].processed_frames] == ["crop", "mask", "cc_labeling"]
        assert np.array_equal(queue.get_last_processed_queue()[0], queue[0].processed_frames["cc_labeling"])
        while not queue.is_empty():
            f = queue.pop_frame()
            t = f.frame_number
            if t not in (0, 3, 12, 13, 14, 25):
                continue
            window = [grays[max(k, 0)] for k in range(t - n + 1, t + 1)]
            fg = rp.absdiff(grays[t], rp.temporal_median(window))
            th = rp.thresh_to_zero(fg, 15)
            op = rp.grayscale_opening(th, (se, se))
            if close:
                op = rp.grayscale_closing(op, (se, se))
            st = f.processed_frames
            assert np.array_equal(st["grayscale"], grays[t])
            assert np.array_equal(st["foreground"], fg) and np.array_equal(st["thresh_15"], th)
            assert np.array_equal(st["opened"], op)
            assert np.array_equal(st["mask"], (op > 0).astype(np.uint8) * 255)       # what the fused path produced
            assert list(st) == ["crop", "mask", "cc_labeling"]


def test_ring_feeds_the_queue_in_place_and_batches_are_submitted_ahead(tmp_path):
    """The reference's driver loop (__main__.py:71-96) over an IngestRing: frames decoded into page-locked
    batches by the ring's thread, submitted in place, batch k+1 already on the GPU while batch k is popped;
    --export writes the segments."""
    import time
    total, B = 75, 21
    frames = synth.synth_video(74, 0, 0, total, 90, 160, 30)
    region = [(16, 8), (150, 80)]
    want = rp.run_path(frames, rp.PathParams(region, 5, 15, 3, True, False, "u8"), want_images=True)
    blank = rp.run_path(np.concatenate([frames[-4:], np.zeros((9,) + frames.shape[1:], np.uint8)]),
                        rp.PathParams(region, 5, 15, 3, True, False, "u8"))[4:]
    reader = ArrayReader(frames, fps=30.0, end=total - 1)           # end_frame is inclusive (io_video.py:40)
    ring = IngestRing(reader, batch_frames=B)
    queue = ds.FrameQueue(queue_size=B)
    queue.attach_ring(ring)
    ds.Frame.src_video = "ring"
    seen, ahead_hits, exported, empty_raised = 0, 0, 0, False
    try:
        while queue.frames_processed < total:
            fr, numbers, stamps = ring.get_n_frames(B)
            queue.push_list_of_frames(fr, numbers, stamps)
            was_ahead = queue._inflight is not None
            queue.preprocess_queue(region, (300, 150))
            queue.segment_queue((24, 24), region)
            ahead_hits += was_ahead
            time.sleep(0.02)                                        # "tracking": the ring decodes the next batch
            dummies = 0
            while not queue.is_empty():
                f = queue.pop_frame()
                if f.null:
                    rec = blank[dummies]                            # zero frames still flow through (io_video.py:40-44)
                    dummies += 1
                else:
                    rec = want[f.frame_number]
                    seen += 1
                assert np.array_equal(f.get_processed_frame("cc_labeling"), rec["labels"]), f.frame_number
                assert np.array_equal(f.get_processed_frame("mask"), rec["mask"])
                assert f.get_num_segments() == len(rec["props"])
                if not f.null:
                    for s, p, c in zip(f.segments, rec["props"], rec["crops"]):
                        assert (s.label, s.area, s.bbox, s.centroid) == (p.label, p.area, p.bbox, tuple(p.centroid))
                        assert np.array_equal(s.segment_image, c)
                    sizes = [s.segment_image.size for s in f.segments]
                    if not exported and sizes and min(sizes) > 0:
                        f.export_segments((24, 24), region, tmp_path / "segments")
                        exported = len(list((tmp_path / "segments").glob("*.png")))
                        assert exported == f.get_num_segments() > 0
                    elif sizes and min(sizes) == 0 and not empty_raised:
                        # a bbox within 12 px of the frame's top / left edge slices to an empty image
                        # (image_filtering.py:363-365) and cv2.imwrite refuses it: the reference's behaviour
                        import cv2
                        with pytest.raises(cv2.error):
                            f.export_segments((24, 24), region, tmp_path / "empty")
                        empty_raised = True
    finally:
        queue.close()
        ring.close()
    assert seen == total and ahead_hits >= 2 and exported > 0


def test_queue_without_a_ring_still_takes_plain_arrays():
    frames = synth.synth_video(75, 0, 0, 10, 64, 96, 12)
    region = [(0, 0), (96, 64)]
    want = rp.run_path(frames, rp.PathParams(region, 5, 15, 3, True, False, "u8"))
    queue = ds.FrameQueue(queue_size=5)
    for t0 in (0, 5):
        queue.push_list_of_frames([f.copy() for f in frames[t0:t0 + 5]], list(range(t0, t0 + 5)), ["x"] * 5)
        queue.preprocess_queue(region, None)
        queue.segment_queue((24, 24), region)
        while not queue.is_empty():
            f = queue.pop_frame()
            assert np.array_equal(f.get_processed_frame("cc_labeling"), want[f.frame_number]["labels"])
    queue.close()
