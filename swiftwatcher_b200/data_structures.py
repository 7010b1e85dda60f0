"""Frame / Segment / FrameQueue with the fused CUDA path inside.

Mirror of ``swiftwatcher/data_structures.py`` (types a10 of SURVEY.md §8):
same class and method names, same queue ordering (``appendleft``: index 0 is
the newest frame, data_structures.py:132-135; ``pop_frame`` pops the oldest,
:143-149), same ``processed_frames`` keys where the stage still exists.
``preprocess_queue`` + ``segment_queue`` hand the whole batch to one
``FilterContext`` (swb_submit / swb_collect) instead of eight Python list
comprehensions.

Differences from the reference, all deliberate:
* background model = rolling temporal median + absdiff (BASELINE.json), not
  RPCA + bilateral; the rolling history is carried across batches;
* the stored intermediates are ``"crop"`` (view), ``"mask"`` ({0,255} uint8,
  equal to ``opened > 0``) and ``"cc_labeling"``; the grey-valued
  ``"grayscale"/"thresh_15"/"opened"`` images are not materialised by the
  fused kernels (use ``image_filtering.*`` for them);
* null (dummy, ``frame_number == -1``) frames are still processed, as in the
  reference, but do not pollute the history of a later batch because a video
  ends with them.
"""

from collections import OrderedDict, deque

import numpy as np

from . import image_filtering as img
from ._lib import pinned_empty
from .pipeline import FilterContext, props_from_rows


class Segment:
    """data_structures.py:16-30."""

    def __init__(self, regionprops, frame_number, timestamp, segment_image):
        self.parent_frame_number = frame_number
        self.parent_timestamp = timestamp
        self.segment_image = segment_image
        self.segment_history = []
        self.status = None
        for name in ("label", "area", "bbox", "centroid"):
            setattr(self, name, getattr(regionprops, name, None))


class Frame:
    """data_structures.py:33-63."""

    src_video = None

    def __init__(self, frame=None, frame_number=-1, timestamp="00:00:00.000"):
        self.frame_number = frame_number
        self.timestamp = timestamp
        self.frame = frame
        self.processed_frames = OrderedDict()
        self.segments = []
        self.null = frame_number < 0

    def get_frame(self):
        return self.frame

    def get_processed_frame(self, process_name):
        return self.processed_frames[process_name]

    def get_num_segments(self):
        return len(self.segments)

    def set_segments(self, regionprops_list, segment_images):
        self.segments = [Segment(rp, self.frame_number, self.timestamp, seg)
                         for rp, seg in zip(regionprops_list, segment_images)]


class FrameQueue(deque):
    """data_structures.py:116-217 with the batch handed to the GPU."""

    def __init__(self, queue_size=21, median_n=5, threshold=15, morph_size=3,
                 do_close=False, label_mode="u8", device=0, bg_model="median"):
        deque.__init__(self, maxlen=queue_size)
        self.frames_read = 0
        self.frames_processed = 0
        self._params = dict(median_n=median_n, threshold=threshold, morph_size=morph_size,
                            do_open=True, do_close=do_close, label_mode=label_mode,
                            device=device, bg_model=bg_model)
        self._ctx = None
        self._ctx_key = None
        self._pinned = [None, None]
        self._pinned_cur = 0

    def pinned_batch(self, frame_shape, n=None):
        """A fresh [n, H, W(, 3)] uint8 batch in page-locked host memory (swb_host_alloc), oldest
        frame first: readers decode straight into its rows (``reader.get_n_frames(n, out=batch)``)
        and ``segment_queue`` submits it without another copy.  Two buffers alternate, so the
        frames of the previous batch (e.g. the tracker's cached frame) stay intact."""
        n = self.maxlen if n is None else n
        shape = (self.maxlen,) + tuple(frame_shape)
        self._pinned_cur ^= 1
        k = self._pinned_cur
        if self._pinned[k] is None or self._pinned[k].shape != shape:
            self._pinned[k] = pinned_empty(shape)
        return self._pinned[k][:n]

    def is_empty(self):
        return len(self) == 0

    def push_frame(self, input_frame, frame_number, timestamp):
        super().appendleft(Frame(input_frame, frame_number, timestamp))
        self.frames_read += 1

    def push_list_of_frames(self, frame_list, frame_number_list, timestamp_list):
        for frame, number, stamp in zip(frame_list, frame_number_list, timestamp_list):
            self.push_frame(frame, number, stamp)

    def pop_frame(self):
        popped = super().pop()
        if popped.null is False:
            self.frames_processed += 1
        return popped

    def store_processed_queue(self, processed_frame_list, process_name):
        for pos, frame in enumerate(processed_frame_list):
            self[pos].processed_frames[process_name] = frame

    def store_segmented_queue(self, regionprops_lists, segment_image_list):
        for pos, (props, images) in enumerate(zip(regionprops_lists, segment_image_list)):
            self[pos].set_segments(props, images)

    def get_queue(self):
        return [f.frame for f in self]

    def get_processed_queue(self, process_name):
        return [f.processed_frames[process_name] for f in self]

    def get_last_processed_queue(self):
        return [next(reversed(f.processed_frames.values())) for f in self]

    # -- the hot path -------------------------------------------------------------
    def _context(self, frame_shape, crop_region):
        key = (tuple(frame_shape), tuple(map(tuple, crop_region)))
        if self._ctx is None or self._ctx_key != key:
            if self._ctx is not None:
                self._ctx.close()
            self._ctx = FilterContext(frame_shape, crop_region, max_frames=self.maxlen,
                                      **self._params)
            self._ctx_key = key
        return self._ctx

    def preprocess_queue(self, crop_region, resize_dim):
        """data_structures.py:171-185.  The crop is stored as a view like the
        reference does; grayscale conversion happens inside the fused kernel
        of ``segment_queue`` (``resize_dim`` is dead in the reference too)."""
        self.store_processed_queue([img.crop_frame(f, crop_region) for f in self.get_queue()],
                                   "crop")

    def segment_queue(self, min_seg_size, crop_region):
        """data_structures.py:187-217 for the whole queue in one submit."""
        if self.is_empty():
            return
        frames = self.get_queue()                 # index 0 = newest
        ctx = self._context(frames[0].shape, crop_region)
        # oldest first, in pinned memory; frames that were decoded into the current batch are in place
        def address(a):
            return a.__array_interface__["data"][0]
        cur = self._pinned[self._pinned_cur]
        ordered = frames[::-1]
        in_place = (cur is not None and len(ordered) <= len(cur) and tuple(cur.shape[1:]) == tuple(ordered[0].shape)
                    and all(isinstance(f, np.ndarray) and f.dtype == np.uint8 and f.flags["C_CONTIGUOUS"]
                            and address(f) == address(cur[i]) for i, f in enumerate(ordered)))
        if in_place:
            batch = cur[:len(ordered)]
        else:
            batch = self.pinned_batch(ordered[0].shape, len(ordered))
            for i, f in enumerate(ordered):
                np.copyto(batch[i], f)
        ctx.submit(batch)                          # history carried across batches
        rows, counts = ctx.collect()
        masks = ctx.masks()
        labels = ctx.labels()
        n = len(frames)
        if ctx.bg_model == "rpca":                 # the reference's own intermediate (data_structures.py:191-192)
            sparse = ctx.rpca_images()
            self.store_processed_queue([sparse[n - 1 - pos] for pos in range(n)], "RPCA")
        self.store_processed_queue([masks[n - 1 - pos] for pos in range(n)], "mask")
        self.store_processed_queue([labels[n - 1 - pos] for pos in range(n)], "cc_labeling")
        offs = np.concatenate([[0], np.cumsum(counts)])
        props_lists, image_lists = [], []
        for pos in range(n):
            t = n - 1 - pos
            props = props_from_rows(rows[offs[t]:offs[t + 1]])
            props_lists.append(props)
            image_lists.append(img.extract_segment_images(props, frames[pos], min_seg_size,
                                                          crop_region))
        self.store_segmented_queue(props_lists, image_lists)
