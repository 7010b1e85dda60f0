"""Frame ingest: the io_video mirror against the oracle restatement of
FrameReader (and the reference's own base class where /root/reference exists); pinned batches."""
import os

import numpy as np
import pytest

from oracle import reference_reader as rr
from oracle import reference_path as rp
from oracle import synth

REF = "/root/reference"


def source(frames, bad=()):
    def read(k):
        return None if k in bad else (frames[k] if 0 <= k < len(frames) else None)
    return read


def same_triplets(a, b):
    fa, na, ta = a
    fb, nb, tb = b
    assert na == nb and [str(x) for x in ta] == [str(x) for x in tb]
    assert [type(x) for x in ta] == [type(x) for x in tb]
    for x, y in zip(fa, fb):
        assert (x is None and y is None) or np.array_equal(x, y)


def test_reader_semantics_match_the_restatement():
    from swiftwatcher_b200.io_video import ArrayReader
    frames = synth.synth_video(8, 0, 0, 11, 24, 40, 3)
    bad = {0, 4, 5}                                    # read errors: the first one has no last good frame
    ours = ArrayReader(source(frames, bad), fps=29.97, start=2, end=9)
    ref = rr.RefFrameReader(source(frames, bad), 29.97, 2, 9)
    for n in (3, 4, 6):                                # runs past end_frame: dummy frames, number -1
        same_triplets(ours.get_n_frames(n), ref.get_n_frames(n))
    assert (ours.frames_read, ours.read_errors, ours.next_frame_number) == (ref.frames_read, ref.read_errors, ref.next_frame_number)
    assert ours.total_frames == ref.total_frames == 7


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "swiftwatcher", "io_video.py")),
                    reason="the reference only exists in the build container")
def test_reader_semantics_match_the_reference_base_class():
    from swiftwatcher_b200.io_video import ArrayReader
    mod = rr.reference_module(REF)
    frames = synth.synth_video(9, 0, 0, 9, 20, 32, 2)
    bad = {3}

    class Theirs(mod.FrameReader):
        def __init__(self):
            super().__init__()
            self.fps, self.start_frame, self.end_frame = 25.0, 1, 7
            self.next_frame_number, self.total_frames = 1, 6

        def read_frame(self, k, increment=True):
            self.next_frame_number += 1
            return source(frames, bad)(k)

    theirs = Theirs()
    ours = ArrayReader(source(frames, bad), fps=25.0, start=1, end=7)
    ref = rr.RefFrameReader(source(frames, bad), 25.0, 1, 7)
    for n in (2, 5, 3):
        a, b, c = theirs.get_n_frames(n), ours.get_n_frames(n), ref.get_n_frames(n)
        same_triplets(a, b)
        same_triplets(a, c)
    assert (ours.frames_read, ours.read_errors) == (theirs.frames_read, theirs.read_errors)


def test_get_n_frames_into_a_batch_buffer():
    from swiftwatcher_b200.io_video import ArrayReader
    frames = synth.synth_video(10, 0, 0, 6, 16, 24, 2)
    ours = ArrayReader(source(frames, {2}), fps=30.0, start=0, end=5)
    batch = np.zeros((4, 16, 24, 3), np.uint8)
    got, numbers, _ = ours.get_n_frames(4, out=batch)
    assert numbers == [0, 1, 2, 3]
    for i, k in enumerate([0, 1, 1, 3]):              # frame 2 failed: last good frame again
        assert np.shares_memory(got[i], batch) and np.array_equal(batch[i], frames[k])


@pytest.mark.gpu
def test_pinned_batches_feed_the_queue_in_place():
    import swiftwatcher_b200.data_structures as ds
    from swiftwatcher_b200._lib import pinned_empty
    from swiftwatcher_b200.io_video import ArrayReader
    buf = pinned_empty((3, 5, 7), np.uint8)
    buf[:] = 7
    assert buf.sum() == 7 * 105
    frames = synth.synth_video(11, 0, 0, 30, 48, 80, 10)
    region = [(4, 2), (76, 46)]
    par = rp.PathParams(region, 5, 15, 3, True, False, "u8")
    want = rp.run_path(frames, par)
    reader = ArrayReader(frames, fps=30.0, start=0, end=29)
    queue = ds.FrameQueue(queue_size=8)
    t = 0
    prev_batch = None
    while t < 24:
        batch = queue.pinned_batch(frames.shape[1:], 8)
        assert prev_batch is None or not np.shares_memory(batch, prev_batch)   # two buffers alternate
        fr, nums, stamps = reader.get_n_frames(8, out=batch)
        queue.push_list_of_frames(fr, nums, stamps)
        queue.preprocess_queue(region, None)
        queue.segment_queue((24, 24), region)
        cur = queue._pinned[queue._pinned_cur]
        assert np.shares_memory(cur, batch)                                    # submitted in place, no extra batch
        while not queue.is_empty():
            f = queue.pop_frame()
            assert np.array_equal(f.get_processed_frame("cc_labeling"), want[f.frame_number]["labels"])
            assert len(f.segments) == len(want[f.frame_number]["props"])
        prev_batch = batch
        t += 8
