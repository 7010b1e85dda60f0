"""Host-side pieces of the drop-in (no GPU): Frame.export_segments against the reference's own
method, crop-region truncation, the on-demand stage dictionary, the decode-ahead ring."""
import importlib
import os
import sys

import numpy as np
import pytest

from oracle import reference_path as rp
from oracle import reference_reader as rr
from oracle import synth

import swiftwatcher_b200.data_structures as ds
from swiftwatcher_b200.io_video import ArrayReader, IngestRing
from swiftwatcher_b200.pipeline import RegionProperties, clamp_crop_region

REF = "/root/reference"


def _frame_with_segments(cls, seg_cls_props, frame, region, number=7):
    f = cls(frame, number, "00:00:01.000")
    f.processed_frames["crop"] = rp.crop_frame(frame, region)
    props = [seg_cls_props(label, area, bbox, cen) for label, area, bbox, cen in
             [(1, 12, (5, 6, 9, 11), (6.5, 8.0)), (2, 400, (20, 30, 50, 70), (33.0, 51.5)), (3, 3, (0, 0, 2, 3), (0.5, 1.0))]]
    crops = rp.extract_segment_images(props, frame, (24, 24), region)
    f.set_segments(props, crops)
    return f


def test_export_segments_writes_what_the_oracle_describes(tmp_path):
    """data_structures.py:65-113: names, overlay blend and segment crops."""
    import cv2
    frame = synth.synth_video(5, 0, 0, 1, 120, 200, 10)[0]
    region = [(40, 30), (180, 110)]
    ds.Frame.src_video = "clip"
    f = _frame_with_segments(ds.Frame, RegionProperties, frame, region)
    f.export_segments((24, 24), region, tmp_path / "segments")
    names = sorted(p.name for p in (tmp_path / "segments").glob("*.png"))
    assert names == ['"clip"_7_%d_3.png' % k for k in (1, 2, 3)]
    assert sorted(p.name for p in (tmp_path / "segments" / "overlay").glob("*.png")) == names
    crop = f.processed_frames["crop"]
    for seg in f.segments:
        name = '"clip"_7_%d_3.png' % seg.label
        r0, c0, r1, c1 = seg.bbox
        want = crop.copy()
        box = want[r0:r1 + 1, c0:c1 + 1].astype(np.float64)        # cv2.rectangle fills both corners inclusively
        red = np.array([0, 0, 255], np.float64)
        want[r0:r1 + 1, c0:c1 + 1] = np.clip(np.rint(0.6 * red + 0.4 * box), 0, 255).astype(np.uint8)
        got = cv2.imread(str(tmp_path / "segments" / "overlay" / name))
        assert np.abs(got.astype(int) - want.astype(int)).max() <= 1     # addWeighted rounds in float32
        b = rp.expand_bbox(seg.bbox, (24, 24), region)
        assert np.array_equal(cv2.imread(str(tmp_path / "segments" / name)), frame[b[0]:b[2], b[1]:b[3]])


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "swiftwatcher", "data_structures.py")),
                    reason="the reference only exists in the build container")
def test_export_segments_equals_the_reference_method(tmp_path):
    shim = os.path.join(os.path.dirname(rp.__file__), "_shim")
    added = [p for p in (shim, REF) if p not in sys.path]
    for p in added:
        sys.path.insert(0, p)
    try:
        ref_ds = importlib.import_module("swiftwatcher.data_structures")
    finally:
        for p in added:
            sys.path.remove(p)
    frame = synth.synth_video(6, 0, 0, 1, 120, 200, 10)[0]
    region = [(40, 30), (180, 110)]
    ds.Frame.src_video = ref_ds.Frame.src_video = "clip"
    ours = _frame_with_segments(ds.Frame, RegionProperties, frame, region)
    theirs = _frame_with_segments(ref_ds.Frame, rp.RegionProperties, frame, region)
    ours.export_segments((24, 24), region, tmp_path / "a")
    theirs.export_segments((24, 24), region, tmp_path / "b")
    files = sorted(p.relative_to(tmp_path / "b") for p in (tmp_path / "b").rglob("*.png"))
    assert len(files) == 6
    assert files == sorted(p.relative_to(tmp_path / "a") for p in (tmp_path / "a").rglob("*.png"))
    for rel in files:
        assert (tmp_path / "a" / rel).read_bytes() == (tmp_path / "b" / rel).read_bytes(), rel


def test_crop_region_overhang_is_truncated_like_a_numpy_slice():
    frame = np.arange(40 * 60 * 3, dtype=np.uint8).reshape(40, 60, 3)
    for region in ([(10, 5), (70, 30)], [(0, 0), (60, 40)], [(50, 35), (90, 90)], [(3, 4), (20, 41)]):
        got = clamp_crop_region(region, frame.shape)
        view = rp.crop_frame(frame, region)                 # the reference: frame[y0:y1, x0:x1]
        (x0, y0), (x1, y1) = got
        assert (y1 - y0, x1 - x0) == view.shape[:2]
        assert np.array_equal(frame[y0:y1, x0:x1], view)
    for bad in ([(-1, 0), (10, 10)], [(5, 5), (5, 10)], [(70, 0), (80, 10)], [(0, 45), (10, 50)]):
        with pytest.raises(ValueError):
            clamp_crop_region(bad, frame.shape)


def test_stage_dict_keeps_the_reference_ordering():
    d = ds.StageDict()
    d["crop"] = 1
    d["mask"] = 2
    d["cc_labeling"] = 3
    assert next(reversed(d.values())) == 3 and list(d) == ["crop", "mask", "cc_labeling"]
    assert "grayscale" not in d and d.get("grayscale") is None
    with pytest.raises(KeyError):
        d["grayscale"]

    class Batch:
        bg_model = "median"
        calls = []

        def compute(self, name, i, stages):
            self.calls.append((name, i))
            return "%s@%d" % (name, i)
    d.lazy_batch, d.lazy_index = Batch(), 4
    assert "grayscale" in d and "opened" in d and "bilateral" not in d and "nonsense" not in d
    assert d["grayscale"] == "grayscale@4" and d["grayscale"] == "grayscale@4"
    assert Batch.calls == [("grayscale", 4)]                 # computed once
    assert list(d) == ["crop", "mask", "cc_labeling"]         # lazy entries stay out of the stage order


def test_ingest_ring_hands_out_what_the_reader_would():
    frames = synth.synth_video(8, 0, 0, 23, 24, 40, 3)
    bad = {4, 11}

    def read(k):
        return None if k in bad else (frames[k] if 0 <= k < len(frames) else None)
    plain = rr.RefFrameReader(read, 30.0, 0, 22)
    ring = IngestRing(ArrayReader(read, fps=30.0, start=0, end=22), batch_frames=7, depth=4,
                      allocate=lambda shape: np.empty(shape, np.uint8))
    assert ring.total_frames == 22 and ring.fps == 30.0
    batches = []
    try:
        for _ in range(5):                                    # runs past the end: dummy frames, number -1
            got = ring.get_n_frames(7)
            want = plain.get_n_frames(7)
            assert got[1] == want[1] and [str(x) for x in got[2]] == [str(x) for x in want[2]]
            for a, b in zip(got[0], want[0]):
                assert np.array_equal(a, b)
            # the frames of a batch are consecutive rows of one array (submitted in place)
            a0 = got[0][0].__array_interface__["data"][0]
            assert all(f.__array_interface__["data"][0] == a0 + i * f.nbytes for i, f in enumerate(got[0]))
            batches.append([f.copy() for f in got[0]] + [got[0]])
            if len(batches) >= 2:                             # the previous batch is still intact
                prev = batches[-2]
                assert all(np.array_equal(x, y) for x, y in zip(prev[:-1], prev[-1]))
        with pytest.raises(ValueError):
            ring.get_n_frames(5)
    finally:
        ring.close()


def test_ingest_ring_peek_does_not_consume():
    import time
    frames = synth.synth_video(9, 0, 0, 12, 16, 24, 2)
    ring = IngestRing(ArrayReader(frames, fps=25.0), batch_frames=4, allocate=lambda s: np.empty(s, np.uint8),
                      frame_shape=frames.shape[1:])
    try:
        first = ring.get_n_frames()
        ahead = None
        for _ in range(200):
            ahead = ring.peek_next()
            if ahead is not None:
                break
            time.sleep(0.005)
        assert ahead is not None and ahead[1] == [4, 5, 6, 7]
        assert ring.peek_next()[0] is ahead[0]
        nxt = ring.get_n_frames()
        assert nxt[1] == [4, 5, 6, 7] and nxt[0][0].__array_interface__["data"][0] == ahead[0].__array_interface__["data"][0]
        assert np.array_equal(first[0][3], frames[3])
    finally:
        ring.close()


def test_host_tile_gather_equals_numpy_slices():
    """swb_host_gather_tiles (plain host memcpys; what segment_queue cuts the 24x24 segment images with)."""
    from swiftwatcher_b200._lib import gather_tiles
    rng = np.random.default_rng(3)
    frames = rng.integers(0, 256, (3, 50, 70, 3), dtype=np.uint8)
    boxes = [(0, 0, 0), (1, 26, 46), (2, 7, 13), (0, 20, 5)]
    pitch, px = frames.strides[1], frames.strides[2]
    addr = np.array([frames[t].__array_interface__["data"][0] + y * pitch + x * px for t, y, x in boxes], dtype=np.uint64)
    out = np.empty((len(boxes), 24, 24, 3), np.uint8)
    gather_tiles(addr, pitch, 24, 24 * px, out)
    for k, (t, y, x) in enumerate(boxes):
        assert np.array_equal(out[k], frames[t, y:y + 24, x:x + 24])
