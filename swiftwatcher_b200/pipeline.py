"""FilterContext — the fused, batched filtering + segmentation path.

Host-side mirror of what ``FrameQueue.preprocess_queue`` + ``segment_queue``
do per batch in the reference (swiftwatcher/data_structures.py:171-217), with
the rolling-median background model BASELINE.json names.  All compute happens
in libswb200 (CUDA); this file only marshals buffers.
"""

import ctypes as C

import numpy as np

from . import _lib
from ._lib import pinned_empty
from ._lib import (BG_MEDIAN, BG_RPCA, HALO_CARRY, LABELS_I32, LABELS_U8, MEM_DEVICE, MEM_HOST, OUT_LABELS,
                   OUT_MASK, SEGMENT_DTYPE, SwbConfig, check, ptr)


class RegionProperties:
    """What swiftwatcher reads from a skimage ``_RegionProperties``
    (image_filtering.py:332-335): label, area, bbox (min_row, min_col, max_row,
    max_col) half-open, centroid (row, col) float64 — as plain attributes so
    that ``Segment.__init__`` (data_structures.py:28-30) can copy them."""

    __slots__ = ("label", "area", "bbox", "centroid")

    def __init__(self, label, area, bbox, centroid):
        self.label = label
        self.area = area
        self.bbox = bbox
        self.centroid = centroid

    def __repr__(self):
        return ("RegionProperties(label=%d, area=%d, bbox=%r, centroid=%r)"
                % (self.label, self.area, self.bbox, self.centroid))


def centroids(rows):
    """(k, 2) float64 centroids: integer coordinate sums / area, one rounding —
    the same value numpy's ``coords.mean(axis=0)`` yields."""
    area = rows["area"].astype(np.float64)
    return np.stack([rows["sum_row"].astype(np.float64) / area,
                     rows["sum_col"].astype(np.float64) / area], axis=1)


def props_from_rows(rows):
    """Segment-table rows (one frame) -> list[RegionProperties], label order."""
    cen = centroids(rows) if len(rows) else np.zeros((0, 2))
    return [RegionProperties(int(r["label"]), int(r["area"]),
                             tuple(int(v) for v in r["bbox"]),
                             (cen[i, 0], cen[i, 1]))
            for i, r in enumerate(rows)]


def clamp_crop_region(crop_region, frame_shape):
    """``crop_frame`` is a numpy slice (image_filtering.py:199-203): a region that runs past the right or
    bottom edge of the frame is silently truncated — ``generate_crop_region`` produces such regions for a
    chimney near the border (image_filtering.py:48-51).  Same here; negative or empty regions raise
    (the reference would wrap around / return an empty image and fail later in cv2)."""
    h, w = int(frame_shape[0]), int(frame_shape[1])
    (x0, y0), (x1, y1) = crop_region
    x0, y0, x1, y1 = int(x0), int(y0), min(int(x1), w), min(int(y1), h)
    if x0 < 0 or y0 < 0 or x1 <= x0 or y1 <= y0:
        raise ValueError("crop_region %r is negative or empty inside a %dx%d frame" % (crop_region, w, h))
    return [(x0, y0), (x1, y1)]


class FilterContext:
    """One context per (GPU, video).  Not thread-safe (include/swb200.h)."""

    def __init__(self, frame_shape, crop_region=None, median_n=5, threshold=15,
                 morph_size=3, do_open=True, do_close=False, label_mode="u8",
                 max_frames=21, max_segments=0, device=0, want_mask=True,
                 want_labels=True, frame_pitch=0, frame_stride=0, bg_model="median", gpu_share=0):
        """frame_shape: (H, W, 3) BGR or (H, W) gray.  crop_region: the
        reference's ``[(x0, y0), (x1, y1)]`` (image_filtering.py:199-203);
        None = whole frame.  bg_model: "median" (rolling temporal median, BASELINE.json) or
        "rpca" (the reference's own rpca + bilateral_blur, image_filtering.py:220-307: every
        submit is one batch of <= 32 frames decomposed on its own).  gpu_share: how many contexts
        work side by side on this GPU (one per video): sizes grids only, results do not change."""
        self._lib = _lib.load()
        h, w = int(frame_shape[0]), int(frame_shape[1])
        ch = int(frame_shape[2]) if len(frame_shape) == 3 else 1
        if crop_region is None:
            crop_region = [(0, 0), (w, h)]
        (x0, y0), (x1, y1) = clamp_crop_region(crop_region, (h, w))
        cfg = SwbConfig()
        cfg.device = device
        cfg.frame_h, cfg.frame_w, cfg.channels = h, w, ch
        cfg.frame_pitch, cfg.frame_stride = frame_pitch, frame_stride
        cfg.roi_x0, cfg.roi_y0, cfg.roi_x1, cfg.roi_y1 = int(x0), int(y0), int(x1), int(y1)
        cfg.median_n, cfg.threshold = median_n, threshold
        cfg.morph_size = morph_size
        cfg.do_open, cfg.do_close = int(bool(do_open)), int(bool(do_close))
        cfg.label_mode = LABELS_U8 if label_mode in ("u8", LABELS_U8) else LABELS_I32
        cfg.out_flags = (OUT_MASK if want_mask else 0) | (OUT_LABELS if want_labels else 0)
        cfg.max_frames = max_frames
        cfg.max_segments = max_segments
        if bg_model not in ("median", "rpca"):
            raise ValueError("bg_model must be 'median' or 'rpca'")
        cfg.bg_model = BG_RPCA if bg_model == "rpca" else BG_MEDIAN
        cfg.gpu_share = int(gpu_share)
        self.bg_model = bg_model
        self.cfg = cfg
        self.frame_shape = (h, w, ch) if ch == 3 else (h, w)
        self.frame_bytes = (frame_stride or (frame_pitch or w * ch) * h)
        self.roi_h, self.roi_w = int(y1) - int(y0), int(x1) - int(x0)
        self.crop_region = [(int(x0), int(y0)), (int(x1), int(y1))]
        self.median_n = median_n
        self.max_frames = max_frames
        self.label_dtype = np.uint8 if cfg.label_mode == LABELS_U8 else np.int32
        self._ctx = C.c_void_p()
        check(self._lib.swb_create(C.byref(cfg), C.byref(self._ctx)))
        self._n_last = 0
        self._keepalive = None
        self._rows_buf = None
        self._fetch = None

    # -- lifetime ---------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._lib.swb_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def reset(self):
        check(self._lib.swb_reset(self._ctx), self._ctx)

    def set_stream(self, cuda_stream):
        """Launch on an external CUDA stream handle (int), e.g.
        ``torch.cuda.current_stream().cuda_stream``."""
        check(self._lib.swb_set_stream(self._ctx, C.c_void_p(cuda_stream or 0)), self._ctx)

    # -- the hot path --------------------------------------------------------
    def submit(self, frames, n_halo=HALO_CARRY, n_frames=None):
        """frames: numpy (n_halo + n, H, W[, 3]) uint8 (host) or a CUDA torch
        tensor of that shape (device, zero-copy).  Asynchronous — except with bg_model="rpca" when the
        iteration loop runs on the host (batches other than 21 frames, or the option "rpca_device_loop"
        set to 0): then the call returns when the decomposition has converged."""
        if isinstance(frames, np.ndarray):
            if frames.dtype != np.uint8 or not frames.flags["C_CONTIGUOUS"]:
                raise ValueError("frames must be C-contiguous uint8")
            kind = MEM_HOST
        elif hasattr(frames, "data_ptr"):
            if not frames.is_contiguous() or frames.element_size() != 1:
                raise ValueError("frames must be a contiguous uint8 tensor")
            kind = MEM_DEVICE if frames.is_cuda else MEM_HOST
        else:
            raise TypeError("frames must be a numpy array or a torch tensor")
        total = int(frames.shape[0])
        halo = 0 if n_halo == HALO_CARRY else int(n_halo)
        n = total - halo if n_frames is None else int(n_frames)
        if tuple(frames.shape[1:]) != tuple(self.frame_shape) and not (self.cfg.frame_pitch or self.cfg.frame_stride):
            raise ValueError("frame shape %r != configured %r" % (tuple(frames.shape[1:]), self.frame_shape))
        self._keepalive = frames
        check(self._lib.swb_submit(self._ctx, ptr(frames), n, n_halo, kind), self._ctx)
        self._n_last = n
        return n

    def submit_ptr(self, address, n_frames, n_halo, mem_kind=MEM_DEVICE):
        check(self._lib.swb_submit(self._ctx, C.c_void_p(address), n_frames, n_halo, mem_kind), self._ctx)
        self._n_last = n_frames

    def sync(self):
        check(self._lib.swb_sync(self._ctx), self._ctx)

    def collect(self, cap=None):
        """-> (rows, counts): the segment table (SEGMENT_DTYPE, ordered by frame
        then label) and segments per frame, for the last submit."""
        n = self._n_last
        cap = int(cap if cap is not None else max(self.cfg.max_segments, 1024 * self.max_frames))
        rows = np.empty(cap, dtype=SEGMENT_DTYPE)
        counts = np.zeros(n, dtype=np.int32)
        n_rows = C.c_int64(0)
        check(self._lib.swb_collect(self._ctx, ptr(rows), cap, C.byref(n_rows), ptr(counts)), self._ctx)
        return rows[:n_rows.value], counts

    def collect_begin(self, masks=None, labels=None, cap=None):
        """Queue every device -> host copy of the last submit (table, masks, labels) behind its kernels and
        return at once (swb_collect_begin); ``collect_end`` waits.  ``masks`` / ``labels``: (n, roi_h, roi_w)
        arrays, page-locked ones (``_lib.pinned_empty``) make the copies asynchronous DMA transfers."""
        n = self._n_last
        cap = int(cap if cap is not None else max(self.cfg.max_segments, 1024 * self.max_frames))
        if self._rows_buf is None or len(self._rows_buf) < cap:
            self._rows_buf = pinned_empty(cap, SEGMENT_DTYPE) if cap * SEGMENT_DTYPE.itemsize <= (8 << 20) \
                else np.empty(cap, dtype=SEGMENT_DTYPE)
        if masks is None:
            masks = np.empty((n, self.roi_h, self.roi_w), dtype=np.uint8)
        if labels is None:
            labels = np.empty((n, self.roi_h, self.roi_w), dtype=self.label_dtype)
        if masks.shape != (n, self.roi_h, self.roi_w) or labels.shape != masks.shape or \
                labels.dtype != self.label_dtype or masks.dtype != np.uint8:
            raise ValueError("masks / labels must be (n, roi_h, roi_w) uint8 / %s" % np.dtype(self.label_dtype))
        check(self._lib.swb_collect_begin(self._ctx, ptr(self._rows_buf), cap, ptr(masks), ptr(labels)), self._ctx)
        self._fetch = (n, masks, labels)

    def collect_end(self):
        """-> (rows, counts, masks, labels) of the ``collect_begin`` in flight."""
        n, masks, labels = self._fetch
        self._fetch = None
        counts = np.zeros(n, dtype=np.int32)
        n_rows = C.c_int64(0)
        check(self._lib.swb_collect_end(self._ctx, C.byref(n_rows), ptr(counts)), self._ctx)
        return self._rows_buf[:n_rows.value].copy(), counts, masks, labels

    def collect_all(self, masks=None, labels=None, cap=None):
        """-> (rows, counts, masks, labels) of the last submit with one stream synchronisation."""
        self.collect_begin(masks, labels, cap)
        return self.collect_end()

    def device_views(self):
        """Zero-copy CUDA torch tensors over the context-owned outputs of the last submit:
        (masks [n, roi_h, pitch] uint8, labels [n, roi_h, pitch] int32/uint8); columns >= roi_w are padding."""
        import torch
        mask, labels = C.c_void_p(), C.c_void_p()
        mp, lp = C.c_int64(0), C.c_int64(0)
        check(self._lib.swb_device_views(self._ctx, C.byref(mask), C.byref(mp), C.byref(labels), C.byref(lp),
                                         None, None), self._ctx)
        dev = "cuda:%d" % self.cfg.device

        def wrap(address, pitch, typestr, dtype):
            if not address:
                return None
            class _Dev:
                pass
            d = _Dev()
            d.__cuda_array_interface__ = {"shape": (self._n_last, self.roi_h, int(pitch)), "typestr": typestr,
                                          "data": (int(address), False), "version": 2, "strides": None}
            return torch.as_tensor(d, device=dev)
        lt = "|u1" if self.label_dtype == np.uint8 else "<i4"
        return wrap(mask.value, mp.value, "|u1", torch.uint8), wrap(labels.value, lp.value, lt, None)

    def set_option(self, name, value):
        """Tuning knobs that never change a result (swb_set_option): "host_pipeline", "sub_batch_min_px"."""
        check(self._lib.swb_set_option(self._ctx, name.encode(), int(value)), self._ctx)

    def last_subchunk(self):
        """Frames per temporal sub-chunk the filtering kernel used for the last submit."""
        n = C.c_int32(0)
        check(self._lib.swb_last_subchunk(self._ctx, C.byref(n)), self._ctx)
        return n.value

    def collect_props(self):
        """-> list (per frame) of list[RegionProperties]."""
        rows, counts = self.collect()
        out, o = [], 0
        for c in counts:
            out.append(props_from_rows(rows[o:o + c]))
            o += c
        return out

    def masks(self, t0=0, n=None):
        n = self._n_last - t0 if n is None else n
        out = np.empty((n, self.roi_h, self.roi_w), dtype=np.uint8)
        check(self._lib.swb_get_masks(self._ctx, t0, n, ptr(out), MEM_HOST), self._ctx)
        return out

    def labels(self, t0=0, n=None):
        n = self._n_last - t0 if n is None else n
        out = np.empty((n, self.roi_h, self.roi_w), dtype=self.label_dtype)
        check(self._lib.swb_get_labels(self._ctx, t0, n, ptr(out), MEM_HOST), self._ctx)
        return out

    def rpca_images(self, t0=0, n=None):
        """bg_model="rpca": the uint8 "RPCA" images (clip(-E, 0, 255)) of the last submit."""
        n = self._n_last - t0 if n is None else n
        out = np.empty((n, self.roi_h, self.roi_w), dtype=np.uint8)
        check(self._lib.swb_get_rpca(self._ctx, t0, n, ptr(out), MEM_HOST), self._ctx)
        return out

    def rpca_stats(self):
        """bg_model="rpca": {"iterations", "jacobi_sweeps", "device_loop"} of the last submit (waits for it)."""
        it, sw, mode = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        check(self._lib.swb_rpca_stats(self._ctx, C.byref(it), C.byref(sw), C.byref(mode)), self._ctx)
        return {"iterations": it.value, "jacobi_sweeps": sw.value, "device_loop": bool(mode.value)}

    def mask_bits(self, t0=0, n=None):
        n = self._n_last - t0 if n is None else n
        wpr = (self.roi_w + 31) // 32
        out = np.empty((n, self.roi_h, wpr), dtype=np.uint32)
        check(self._lib.swb_get_mask_bits(self._ctx, t0, n, ptr(out), MEM_HOST), self._ctx)
        return out

    def gather_crops(self, n_rows, crop=24, out=None, rects=None):
        """-> (tiles, rects) for the last submit's table (device-resident full frames only):
        tiles (n_rows, crop, crop, C) uint8 = ``Resize((crop, crop))(ToPILImage()(segment_image))`` of the
        reference's ``extract_segment_images`` (the crop itself where it already is crop x crop, zeros
        where it is empty); rects (n_rows, 4) int32 = the rectangle (y0, x0, y1, x1) the reference slices
        from the full frame.  ``out`` / ``rects`` may be CUDA torch tensors (then nothing visits the host)."""
        ch = self.cfg.channels
        on_device = hasattr(out, "is_cuda") and out.is_cuda
        if out is None:
            out = np.empty((n_rows, crop, crop, ch), dtype=np.uint8)
        if rects is None:
            if on_device:
                import torch
                rects = torch.empty((n_rows, 4), dtype=torch.int32, device=out.device)
            else:
                rects = np.empty((n_rows, 4), dtype=np.int32)
        if (hasattr(rects, "is_cuda") and rects.is_cuda) != on_device:
            raise ValueError("out and rects must both be host arrays or both CUDA tensors")
        check(self._lib.swb_gather_crops(self._ctx, crop, ptr(out), ptr(rects), MEM_DEVICE if on_device else MEM_HOST),
              self._ctx)
        return out, rects

    # -- instrumentation --------------------------------------------------------
    def enable_timing(self, on=True):
        check(self._lib.swb_enable_timing(self._ctx, int(on)), self._ctx)

    def timing(self):
        names = (C.c_char_p * 8)()
        ms = (C.c_float * 8)()
        n = C.c_int32(0)
        check(self._lib.swb_get_timing(self._ctx, names, ms, 8, C.byref(n)), self._ctx)
        return {names[i].decode(): float(ms[i]) for i in range(n.value)}

    def launch_count(self):
        return int(self._lib.swb_launch_count(self._ctx))


def synth_frames(seed, video, t0, n, h, w, n_birds, device=0, out=None):
    """Synthetic frames from the CUDA generator (twin of oracle/synth.py).
    ``out``: optional CUDA uint8 torch tensor (n, h, w, 3); else numpy."""
    lib = _lib.load()
    if out is None:
        out = np.empty((n, h, w, 3), dtype=np.uint8)
        kind = MEM_HOST
    else:
        kind = MEM_DEVICE if getattr(out, "is_cuda", False) else MEM_HOST
    check(lib.swb_synth_frames(device, ptr(out), kind, seed, video, t0, n, h, w, n_birds))
    return out
