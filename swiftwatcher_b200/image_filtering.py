"""Drop-in for the hot-path functions of ``swiftwatcher/image_filtering.py``.

Same names, argument meaning and return types as the reference; every
function runs its stage as a CUDA kernel through the C ABI (include/swb200.h,
``swb_stage_*``).  These per-function entry points exist so that code written
against the reference keeps working unchanged; the fast path is the fused,
batched ``FilterContext`` (pipeline.py) that ``data_structures.FrameQueue``
uses.  No function here has a CPU fallback.

The reference's own background model is provided too: ``rpca`` and
``bilateral_blur`` (image_filtering.py:220-307; at most 32 frames per batch,
``d <= 7``); structuring elements must be odd-sized.  Reference functions
that are NOT on the per-frame path (generate_regions, generate_roi_mask and
helpers: once per video, image_filtering.py:20-180) are out of scope
 and not provided.
"""

import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import SEGMENT_DTYPE, check, ptr
from .pipeline import RegionProperties, props_from_rows  # noqa: F401

_DEVICE = 0


def set_device(device):
    """CUDA device used by the per-function entry points."""
    global _DEVICE
    _DEVICE = int(device)


def _u8(frame, ndim):
    a = np.ascontiguousarray(frame)
    if a.dtype != np.uint8 or a.ndim != ndim:
        raise ValueError("expected a %d-D uint8 array, got %s %r" % (ndim, a.dtype, a.shape))
    return a


def crop_frame(frame, crop_region):
    """image_filtering.py:199-203 — a numpy view ``frame[y0:y1, x0:x1]``
    (``crop_region = [(x0, y0), (x1, y1)]``); no copy, no bounds checks.  In
    the fused path the crop is just an address offset of the frame loads."""
    return frame[crop_region[0][1]:crop_region[1][1],
                 crop_region[0][0]:crop_region[1][0]]


def convert_grayscale(frame):
    """image_filtering.py:188-196 — BGR -> gray for 3-D input (cv2's 15-bit
    fixed point), identity for 2-D input."""
    if len(frame.shape) == 3:
        a = _u8(frame, 3)
        if a.shape[2] != 3:
            raise ValueError("expected 3 channels")
        out = np.empty(a.shape[:2], dtype=np.uint8)
        check(_lib.load().swb_stage_gray(_DEVICE, ptr(a), a.shape[0], a.shape[1], ptr(out)))
        return out
    return frame


def temporal_median(gray_frames):
    """Per-pixel median over an odd number (<= 9) of gray frames — the rolling
    background model BASELINE.json names (no reference function)."""
    stack = _u8(np.stack(list(gray_frames)), 3)
    out = np.empty(stack.shape[1:], dtype=np.uint8)
    check(_lib.load().swb_stage_median(_DEVICE, ptr(stack), stack.shape[0], stack.shape[1],
                                       stack.shape[2], ptr(out)))
    return out


def absdiff(a, b):
    """``|a - b|`` on uint8 (no reference function)."""
    a, b = _u8(a, 2), _u8(b, 2)
    if a.shape != b.shape:
        raise ValueError("shape mismatch")
    out = np.empty_like(a)
    check(_lib.load().swb_stage_absdiff(_DEVICE, ptr(a), ptr(b), a.shape[0], a.shape[1], ptr(out)))
    return out


def thresh_to_zero(frame, thresh):
    """image_filtering.py:310-316 — ``v if v > thresh else 0``."""
    a = _u8(frame, 2)
    out = np.empty_like(a)
    check(_lib.load().swb_stage_thresh_to_zero(_DEVICE, ptr(a), a.shape[0], a.shape[1],
                                               int(thresh), ptr(out)))
    return out


def _morph(frame, SE, closing):
    a = _u8(frame, 2)
    out = np.empty_like(a)
    check(_lib.load().swb_stage_grey_morph(_DEVICE, ptr(a), a.shape[0], a.shape[1],
                                           int(SE[0]), int(SE[1]), closing, ptr(out)))
    return out


def grayscale_opening(frame, SE):
    """image_filtering.py:319-322 — flat grey opening, ``SE`` = (rows, cols)."""
    return _morph(frame, SE, 0)


def grayscale_closing(frame, SE):
    """Dual of ``grayscale_opening`` (no reference function)."""
    return _morph(frame, SE, 1)


def rpca(frame_list):
    """image_filtering.py:220-253 — robust PCA (IALM) over a batch of gray frames: list of uint8
    "sparse" images, what is darker than the low-rank background.  Columns are taken in list
    order, as in the reference (its queue holds the newest frame first).  float64 on the GPU, SVD
    through the n x n Gram matrix (csrc/rpca.cu); at most 32 frames."""
    stack = np.ascontiguousarray(np.array(frame_list))
    if stack.ndim != 3 or stack.dtype != np.uint8:
        raise ValueError("frame_list must hold 2-D uint8 frames of one shape")
    out = np.empty_like(stack)
    check(_lib.load().swb_stage_rpca(_DEVICE, ptr(stack), stack.shape[0], stack.shape[1], stack.shape[2],
                                     ptr(out), None))
    return [out[i] for i in range(stack.shape[0])]


def bilateral_blur(frame, d, sigmaColor, sigmaSpace):
    """image_filtering.py:304-307 — cv2.bilateralFilter for an 8-bit single-channel frame
    (OpenCV's scalar definition, see csrc/rpca.cu; the reference calls it with (7, 15, 1))."""
    a = _u8(frame, 2)
    out = np.empty_like(a)
    check(_lib.load().swb_stage_bilateral(_DEVICE, ptr(a), a.shape[0], a.shape[1], int(d), float(sigmaColor),
                                          float(sigmaSpace), ptr(out)))
    return out


def cc_labeling(frame, connectivity):
    """image_filtering.py:325-329 — the reference passes ``connectivity``
    into cv2's ``labels`` slot, so labelling is always 8-connected; labels use
    OpenCV's numbering and are truncated to uint8."""
    a = _u8(frame, 2)
    out = np.empty(a.shape, dtype=np.uint8)
    check(_lib.load().swb_stage_cc_label(_DEVICE, ptr(a), a.shape[0], a.shape[1], None, ptr(out), None))
    return out


def cc_labeling_i32(frame):
    """Same labelling without the uint8 truncation (int32 labels 1..n)."""
    a = _u8(frame, 2)
    out = np.empty(a.shape, dtype=np.int32)
    n = C.c_int32(0)
    check(_lib.load().swb_stage_cc_label(_DEVICE, ptr(a), a.shape[0], a.shape[1], ptr(out), None, C.byref(n)))
    return out


def get_segment_properties(frame):
    """image_filtering.py:332-335 — regionprops of a label image: one object
    per label value present, ascending, exposing ``label``, ``area``, ``bbox``
    and ``centroid``."""
    a = np.ascontiguousarray(frame)
    if a.ndim != 2 or a.dtype not in (np.uint8, np.int32):
        raise ValueError("label image must be 2-D uint8 or int32")
    cap = int(a.max()) if a.size else 0
    rows = np.empty(max(cap, 1), dtype=SEGMENT_DTYPE)
    n = C.c_int32(0)
    check(_lib.load().swb_stage_regionprops(_DEVICE, ptr(a), a.dtype.itemsize, a.shape[0], a.shape[1],
                                            ptr(rows), max(cap, 1), C.byref(n)))
    return props_from_rows(rows[:n.value])


def expand_bbox(bbox, min_seg_size, crop_region):
    """bbox arithmetic of image_filtering.py:350-362 (no clamping)."""
    bbox = list(bbox)
    dims = (bbox[2] - bbox[0], bbox[3] - bbox[1])
    if dims[0] < min_seg_size[0]:
        diff = min_seg_size[0] - dims[0]
        bbox[0] -= math.floor(diff / 2)
        bbox[2] += math.ceil(diff / 2)
    if dims[1] < min_seg_size[1]:
        diff = min_seg_size[1] - dims[1]
        bbox[1] -= math.floor(diff / 2)
        bbox[3] += math.ceil(diff / 2)
    oy, ox = crop_region[0][1], crop_region[0][0]
    return [bbox[0] + oy, bbox[1] + ox, bbox[2] + oy, bbox[3] + ox]


def expand_bboxes(bboxes, min_seg_size, crop_region):
    """``expand_bbox`` for a whole table at once: [k, 4] int array -> [k, 4] int64 (full-frame coordinates)."""
    b = np.asarray(bboxes, dtype=np.int64).reshape(-1, 4).copy()
    for lo, hi, want in ((0, 2, int(min_seg_size[0])), (1, 3, int(min_seg_size[1]))):
        diff = np.maximum(want - (b[:, hi] - b[:, lo]), 0)
        b[:, lo] -= diff // 2
        b[:, hi] += (diff + 1) // 2
    b[:, 0::2] += int(crop_region[0][1])
    b[:, 1::2] += int(crop_region[0][0])
    return b


def extract_segment_images(segments, frame, min_seg_size, crop_region):
    """image_filtering.py:338-369 — colour crops from the un-cropped host
    frame: numpy views with the reference's slice semantics.  (Pure indexing
    of a host array; the device-side batched variant for the classifier is
    ``FilterContext.gather_crops``.)"""
    images = []
    for segment in segments:
        b = expand_bbox(segment.bbox, min_seg_size, crop_region)
        images.append(frame[b[0]:b[2], b[1]:b[3]])
    return images
