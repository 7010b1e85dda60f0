"""Frame / Segment / FrameQueue with the fused CUDA path inside.

Mirror of ``swiftwatcher/data_structures.py`` (types a10 of SURVEY.md §8):
same class and method names, same queue ordering (``appendleft``: index 0 is
the newest frame, data_structures.py:132-135; ``pop_frame`` pops the oldest,
:143-149), same ``processed_frames`` keys where the stage still exists.
``preprocess_queue`` + ``segment_queue`` hand the whole batch to one
``FilterContext`` (swb_submit / swb_collect) instead of eight Python list
comprehensions.

Differences from the reference, all deliberate:
* the default background model is the rolling temporal median + absdiff of
  BASELINE.json with the history carried across batches; the reference's own
  RPCA + bilateral is ``FrameQueue(bg_model="rpca")``;
* the stored intermediates are ``"crop"`` (view), ``"mask"`` ({0,255} uint8,
  equal to ``opened > 0``) and ``"cc_labeling"``; the grey-valued
  ``"grayscale"/"thresh_15"/"opened"`` images are not materialised by the
  fused kernels: ``processed_frames[...]`` computes them on first access
  (``StageDict``);
* a ``crop_region`` that overhangs the right / bottom frame edge is truncated
  as the reference's numpy slice does;
* null (dummy, ``frame_number == -1``) frames are still processed, as in the
  reference, but do not pollute the history of a later batch because a video
  ends with them.
"""

from collections import OrderedDict, deque
from pathlib import Path

import numpy as np

from . import image_filtering as img
from ._lib import HALO_CARRY, MEM_HOST, gather_tiles, is_pinned, pinned_empty
from .pipeline import FilterContext, centroids, clamp_crop_region, props_from_rows  # noqa: F401


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


class _LazyBatch:
    """What the on-demand stages of one batch need: the crops of the batch (and of the previous batch's
    tail, for the rolling-median window), oldest first."""

    def __init__(self, crops, base, params, bg_model):
        self.crops, self.base, self.params, self.bg_model = crops, base, params, bg_model

    def compute(self, name, i, stages):
        p = self.params
        se = (p["morph_size"], p["morph_size"])
        if name == "grayscale":
            return img.convert_grayscale(self.crops[i])
        if name == "bilateral" and self.bg_model == "rpca":
            return img.bilateral_blur(stages["RPCA"], 7, 15, 1)
        if name == "foreground" and self.bg_model != "rpca":
            n_hist = p["median_n"] - 1
            first = max(i - n_hist, 0)                           # missing history = earliest frame replicated
            window = [img.convert_grayscale(self.crops[max(k, first)]) for k in range(i - n_hist, i + 1)]
            return img.absdiff(window[-1], img.temporal_median(window))
        if name == "thresh_15":
            return img.thresh_to_zero(stages["bilateral" if self.bg_model == "rpca" else "foreground"], p["threshold"])
        if name == "opened":
            x = stages["thresh_15"]
            if p["morph_size"]:
                x = img.grayscale_opening(x, se)
                if p["do_close"]:
                    x = img.grayscale_closing(x, se)
            return x
        raise KeyError(name)


class StageDict(OrderedDict):
    """``Frame.processed_frames``: the stored stage images in insertion order (what
    ``get_last_processed_queue`` relies on, data_structures.py:166-169) plus the reference's grey-valued
    intermediates ON DEMAND — ``"grayscale"``, ``"thresh_15"``, ``"opened"`` (and ``"foreground"``; with
    ``bg_model="rpca"``: ``"bilateral"``) are computed by the single-stage CUDA entry points the first
    time they are read (data_structures.py:183-203 stores them for every frame; the fused kernels never
    materialise them).  Lazy entries do not take part in iteration / ``len`` / "last stage"."""

    LAZY = ("grayscale", "foreground", "bilateral", "thresh_15", "opened")

    def __init__(self):
        super().__init__()
        self.lazy_batch = None
        self.lazy_index = 0
        self.lazy_cache = None

    def __missing__(self, key):
        if self.lazy_batch is None or key not in self.LAZY:
            raise KeyError(key)
        if self.lazy_cache is None:
            self.lazy_cache = {}
        if key not in self.lazy_cache:
            self.lazy_cache[key] = self.lazy_batch.compute(key, self.lazy_index, self)
        return self.lazy_cache[key]

    def __contains__(self, key):
        if OrderedDict.__contains__(self, key):
            return True
        if self.lazy_batch is None or key not in self.LAZY:
            return False
        return (key == "bilateral") == (self.lazy_batch.bg_model == "rpca") or key not in ("bilateral", "foreground")

    def get(self, key, default=None):
        try:
            return self[key]
        except KeyError:
            return default


class Frame:
    """data_structures.py:33-113."""

    src_video = None

    def __init__(self, frame=None, frame_number=-1, timestamp="00:00:00.000"):
        self.frame_number = frame_number
        self.timestamp = timestamp
        self.frame = frame
        self.processed_frames = StageDict()
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

    def export_segments(self, min_seg_size, crop_region, export_dir):
        """data_structures.py:65-113 (``--export``, __main__.py:94-96): per segment one PNG of the cropped
        frame with the segment's bbox shaded red (60 % blend) under ``export_dir/overlay`` and one PNG of
        the segment image — bbox grown to ``min_seg_size``, cut from the FULL frame — under ``export_dir``.
        File names: ``"<src_video>"_<frame number>_<label>_<segments in frame>.png``."""
        import cv2
        export_dir = Path(export_dir)
        (export_dir / "overlay").mkdir(parents=True, exist_ok=True)
        colour = self.processed_frames["crop"]
        crops = img.extract_segment_images(self.segments, self.frame, min_seg_size, crop_region)
        for segment, crop in zip(self.segments, crops):
            name = '"{}"_{}_{}_{}.png'.format(self.src_video, self.frame_number, segment.label, len(self.segments))
            r0, c0, r1, c1 = segment.bbox
            shaded = colour.copy()
            cv2.rectangle(shaded, (c0, r0), (c1, r1), (0, 0, 255), -1)
            cv2.imwrite(str(export_dir / "overlay" / name), cv2.addWeighted(shaded, 0.6, colour, 1 - 0.6, 0))
            cv2.imwrite(str(export_dir / name), crop)


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
        self._out = [None, None]          # page-locked (masks, labels) of the last two batches
        self._out_cur = 0
        self._tail = []                   # the last median_n - 1 crops of the previous batch (lazy stages)
        self._ring = None                 # io_video.IngestRing to look ahead into (attach_ring)
        self._inflight = None             # (address, n) of the batch that was submitted ahead of time

    def pinned_batch(self, frame_shape, n=None):
        """A fresh [n, H, W(, 3)] uint8 batch in page-locked host memory (swb_host_alloc), oldest
        frame first: readers decode straight into its rows (``reader.get_n_frames(n, out=batch)``)
        and ``segment_queue`` submits it without another copy.  Two buffers alternate, so the
        frames of the previous batch (e.g. the tracker's cached frame) stay intact; anything kept
        longer than two batches must be copied (``Segment.segment_image`` already is a copy)."""
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
            if self._inflight is not None:
                raise RuntimeError("FrameQueue: frame shape / crop region changed while a batch submitted ahead "
                                   "of time (attach_ring) is still in flight")
            if self._ctx is not None:
                self._ctx.close()
            self._ctx = FilterContext(frame_shape, crop_region, max_frames=self.maxlen,
                                      **self._params)
            self._ctx_key = key
            self._tail = []
        return self._ctx

    def _outputs(self, ctx, n):
        """Page-locked masks / labels for this batch (two sets alternate, like the input batches)."""
        self._out_cur ^= 1
        k = self._out_cur
        shape = (self.maxlen, ctx.roi_h, ctx.roi_w)
        if self._out[k] is None or self._out[k][0].shape != shape or self._out[k][1].dtype != ctx.label_dtype:
            self._out[k] = (pinned_empty(shape, np.uint8), pinned_empty(shape, ctx.label_dtype))
        return self._out[k][0][:n], self._out[k][1][:n]

    def preprocess_queue(self, crop_region, resize_dim):
        """data_structures.py:171-185.  The crop is stored as a view like the
        reference does; grayscale conversion happens inside the fused kernel
        of ``segment_queue`` (``resize_dim`` is dead in the reference too) and
        ``processed_frames["grayscale"]`` is computed when somebody reads it."""
        self.store_processed_queue([img.crop_frame(f, crop_region) for f in self.get_queue()],
                                   "crop")

    def _attach_lazy_stages(self, ctx, crops_oldest_first):
        """The reference's grey intermediates, on demand (see StageDict)."""
        n_hist = self._params["median_n"] - 1
        seq = self._tail + crops_oldest_first       # crops, oldest first, with the previous batch's tail
        base, n = len(self._tail), len(crops_oldest_first)
        lazy = _LazyBatch(seq, base, self._params, ctx.bg_model)
        for pos in range(n):                          # queue index 0 = newest
            stages = self[pos].processed_frames
            stages.lazy_batch = lazy
            stages.lazy_index = base + (n - 1 - pos)
        self._tail = seq[len(seq) - n_hist:] if n_hist > 0 else []

    def attach_ring(self, ring):
        """Let the queue look ahead into an ``io_video.IngestRing``: right after the results of batch k are
        on the host, batch k+1 — if the ring has already decoded it — is submitted and its read-back queued
        (swb_collect_begin), so the GPU works on k+1 while the caller tracks the frames of k.  The frames
        pushed next MUST then be that batch (``ring.get_n_frames``); anything else raises."""
        self._ring = ring

    def close(self):
        """Drain an early submit that was never consumed and release the context."""
        if self._inflight is not None and self._ctx is not None:
            self._ctx.collect_end()
        self._inflight = None
        if self._ctx is not None:
            self._ctx.close()
            self._ctx = None

    @staticmethod
    def _pinned_block(ordered):
        """(address, bytes per frame) when the frames are consecutive rows of ONE page-locked allocation
        (a ``pinned_batch`` / an ``IngestRing`` batch that was decoded in place), else None."""
        first = ordered[0]
        if not isinstance(first, np.ndarray) or first.dtype != np.uint8 or not first.flags["C_CONTIGUOUS"]:
            return None
        a0, nb = first.__array_interface__["data"][0], first.nbytes
        for i, f in enumerate(ordered):
            if not isinstance(f, np.ndarray) or f.dtype != np.uint8 or f.shape != first.shape or \
                    not f.flags["C_CONTIGUOUS"] or f.__array_interface__["data"][0] != a0 + i * nb:
                return None
        return (a0, nb) if is_pinned(a0, nb * len(ordered)) else None

    def segment_queue(self, min_seg_size, crop_region):
        """data_structures.py:187-217 for the whole queue in one submit and one read-back."""
        if self.is_empty():
            return
        frames = self.get_queue()                 # index 0 = newest
        crop_region = clamp_crop_region(crop_region, frames[0].shape)
        ctx = self._context(frames[0].shape, crop_region)
        ordered = frames[::-1]                    # oldest first
        n = len(frames)
        block = self._pinned_block(ordered)
        if self._inflight is not None:
            # this batch was submitted ahead of time (attach_ring): only wait for its read-back
            if block is None or (block[0], n) != self._inflight:
                raise RuntimeError("FrameQueue: the frames pushed are not the batch the attached IngestRing "
                                   "decoded next (it has already been submitted)")
            self._inflight = None
            rows, counts, masks, labels = ctx.collect_end()
        else:
            if block is not None:                 # decoded in place: DMA straight from where the frames are
                ctx.submit_ptr(block[0], n, HALO_CARRY, MEM_HOST)      # history carried across batches
            else:                                 # anything else: one copy into a page-locked batch
                batch = self.pinned_batch(ordered[0].shape, n)
                for i, f in enumerate(ordered):
                    np.copyto(batch[i], f)
                ctx.submit(batch)
            masks, labels = self._outputs(ctx, n)
            rows, counts, masks, labels = ctx.collect_all(masks, labels)   # one synchronisation, DMA into pinned memory
        if ctx.bg_model == "rpca":                 # the reference's own intermediate (data_structures.py:191-192)
            sparse = ctx.rpca_images()
            self.store_processed_queue([sparse[n - 1 - pos] for pos in range(n)], "RPCA")
        elif self._ring is not None:
            ahead = self._ring.peek_next()
            if ahead is not None and tuple(ahead[0].shape[1:]) == tuple(ordered[0].shape):
                ctx.submit(ahead[0])
                ctx.collect_begin(*self._outputs(ctx, len(ahead[0])))
                self._inflight = (ahead[0].__array_interface__["data"][0], len(ahead[0]))
        self.store_processed_queue([masks[n - 1 - pos] for pos in range(n)], "mask")
        self.store_processed_queue([labels[n - 1 - pos] for pos in range(n)], "cc_labeling")
        self._attach_lazy_stages(ctx, [f.processed_frames["crop"] if "crop" in f.processed_frames
                                       else img.crop_frame(f.frame, crop_region) for f in reversed(self)])
        # Segments straight from the table (no per-row numpy scalar conversions): what Frame.set_segments /
        # Segment.__init__ build (data_structures.py:16-30,61-63), with the crops of extract_segment_images
        # (image_filtering.py:338-369; copied: a segment outlives the pinned batch its frame was decoded into)
        offs = np.concatenate([[0], np.cumsum(counts)]).tolist()
        label, area, bbox = rows["label"].tolist(), rows["area"].tolist(), rows["bbox"].tolist()
        cen = centroids(rows).tolist() if len(rows) else []
        boxes = img.expand_bboxes(rows["bbox"], min_seg_size, crop_region) if len(rows) else np.zeros((0, 4), np.int64)
        box = boxes.tolist()
        # The common case — the grown bbox is exactly min_seg_size and inside the frame — is cut for the whole batch
        # by one library call (row memcpys); anything else (larger birds, frame edges) keeps the reference's slice.
        first = frames[0]
        mh, mw = int(min_seg_size[0]), int(min_seg_size[1])
        tiles, plain_l, tile_of = None, None, None
        uniform = all(isinstance(f, np.ndarray) and f.dtype == np.uint8 and f.shape == first.shape and
                      f.strides == first.strides and f.strides[-1] == 1 for f in frames)
        if uniform and len(rows):
            fh, fw = first.shape[0], first.shape[1]
            pitch, px = first.strides[0], first.strides[1]
            plain = ((boxes[:, 2] - boxes[:, 0] == mh) & (boxes[:, 3] - boxes[:, 1] == mw) & (boxes[:, 0] >= 0) &
                     (boxes[:, 1] >= 0) & (boxes[:, 2] <= fh) & (boxes[:, 3] <= fw))
            sel = np.flatnonzero(plain)
            if len(sel):
                base = np.array([f.__array_interface__["data"][0] for f in ordered], dtype=np.uint64)   # by frame index t
                addr = base[rows["frame"][sel]] + (boxes[sel, 0] * pitch + boxes[sel, 1] * px).astype(np.uint64)
                tiles = np.empty((len(sel), mh, mw) + first.shape[2:], np.uint8)
                gather_tiles(np.ascontiguousarray(addr), pitch, mh, mw * px, tiles)
                plain_l = plain.tolist()
                tile_of = (np.cumsum(plain) - 1).tolist()      # table row -> its tile
        for pos in range(n):
            t = n - 1 - pos
            fr = self[pos]
            full, number, stamp = fr.frame, fr.frame_number, fr.timestamp
            segs = []
            for i in range(offs[t], offs[t + 1]):
                if plain_l is not None and plain_l[i]:
                    image = tiles[tile_of[i]]
                else:
                    b = box[i]
                    image = full[b[0]:b[2], b[1]:b[3]].copy()
                seg = Segment.__new__(Segment)
                seg.__dict__ = {"parent_frame_number": number, "parent_timestamp": stamp, "segment_image": image,
                                "segment_history": [], "status": None, "label": label[i], "area": area[i],
                                "bbox": tuple(bbox[i]), "centroid": tuple(cen[i])}
                segs.append(seg)
            fr.segments = segs
