"""CPU restatement of swiftwatcher's per-frame filtering + segmentation path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Every function cites the reference ``file:line`` it follows (paths relative to
the reference checkout, ``swiftwatcher/...``).  The reference functions on
this path are one-line wrappers around OpenCV / SciPy / scikit-image calls, so
the restatement makes the *same library call with the same arguments*; the
arithmetic therefore lives in the installed third-party libraries:

* opencv-python: reference pins 4.1.0.25 (requirements.txt:8); this image has
  4.13.0.  ``cvtColor(BGR2GRAY)`` on uint8 is the 15-bit fixed point
  ``(3735*B + 19235*G + 9798*R + 16384) >> 15`` in 4.13 (``gray_fixed_point``
  below restates it; tests pin it bit-exactly against cv2).
  ``connectedComponents`` numbers 8-connected components by the rank of the
  component's minimum 2x2-block raster index (``label_order_spec`` restates it;
  tests pin it against cv2).
* scipy: reference pins 1.3.1 (requirements.txt:16); image has 1.18.1.
  ``ndimage.grey_opening`` = min filter then max filter, mode='reflect'.
* scikit-image 0.15.0 (requirements.txt:15) is NOT installed and cannot be
  (no network).  ``regionprops`` below restates skimage 0.15
  ``measure/_regionprops.py`` for the four consumed properties on top of
  ``scipy.ndimage.find_objects`` (skimage's own backend).

Parity pinning: the reference ships no tests, golden vectors or fixtures for
this path.  ``oracle/make_golden.py`` imports the reference's own
``image_filtering`` module from ``/root/reference`` (with a ``skimage.measure``
shim that forwards to ``regionprops`` below) and checks every function here
against it on seeded inputs, then writes ``tests/golden/*.npz``.  The two
stages named by BASELINE.json that the reference does not contain (temporal
median, absdiff; also closing and 5x5 structuring elements) have no reference
function: **their parity is unpinned by the reference** and defined only by
this file.
"""

import math

import cv2
import numpy as np
from scipy import ndimage


# ----------------------------------------------------------------------------
# a1  crop_frame                                image_filtering.py:199-203
# ----------------------------------------------------------------------------
def crop_frame(frame, crop_region):
    """``frame[y0:y1, x0:x1]`` with ``crop_region = [(x0, y0), (x1, y1)]``.
    A numpy view; no bounds checks (image_filtering.py:202-203)."""
    return frame[crop_region[0][1]:crop_region[1][1],
                 crop_region[0][0]:crop_region[1][0]]


# ----------------------------------------------------------------------------
# a2  convert_grayscale                         image_filtering.py:188-196
# ----------------------------------------------------------------------------
def convert_grayscale(frame):
    """cv2 BGR2GRAY for 3-D input, identity for 2-D (image_filtering.py:191-194)."""
    if len(frame.shape) == 3:
        return cv2.cvtColor(np.ascontiguousarray(frame), cv2.COLOR_BGR2GRAY)
    return frame


GRAY_WB, GRAY_WG, GRAY_WR, GRAY_SHIFT = 3735, 19235, 9798, 15


def gray_fixed_point(frame_bgr):
    """Plain-numpy restatement of cv2 4.13 ``cvtColor(BGR2GRAY)`` on uint8:
    ``(3735*B + 19235*G + 9798*R + (1 << 14)) >> 15``.  This is the formula
    the CUDA kernels implement; ``tests/test_oracle.py`` pins it to cv2."""
    f = frame_bgr.astype(np.uint32)
    y = (GRAY_WB * f[..., 0] + GRAY_WG * f[..., 1] + GRAY_WR * f[..., 2]
         + (1 << (GRAY_SHIFT - 1))) >> GRAY_SHIFT
    return y.astype(np.uint8)


# ----------------------------------------------------------------------------
# a3  rolling temporal median                   (no reference function)
# ----------------------------------------------------------------------------
def temporal_median(gray_window):
    """Per-pixel median of an odd number N of uint8 frames (the middle order
    statistic; no rounding question for odd N).  Not in the reference — its
    background model is RPCA (image_filtering.py:220-301).  PARITY UNPINNED."""
    stack = np.stack(list(gray_window))
    assert stack.shape[0] % 2 == 1, "N must be odd"
    return np.median(stack, axis=0).astype(np.uint8)


def window_indices(t, n):
    """History policy frozen by this oracle: the window for output frame ``t``
    is frames ``[t-n+1, t]``; indices before the start of the video are
    clamped to frame 0 (the first frame is replicated)."""
    return [max(i, 0) for i in range(t - n + 1, t + 1)]


# ----------------------------------------------------------------------------
# a4  absdiff                                   (no reference function)
# ----------------------------------------------------------------------------
def absdiff(a, b):
    """Two-sided ``|a - b|`` on uint8 (cv2.absdiff).  PARITY UNPINNED."""
    return cv2.absdiff(a, b)


# ----------------------------------------------------------------------------
# a5  thresh_to_zero                            image_filtering.py:310-316
# ----------------------------------------------------------------------------
def thresh_to_zero(frame, thresh):
    """``v if v > thresh else 0`` (cv2.THRESH_TOZERO, image_filtering.py:311-314)."""
    _, out = cv2.threshold(frame, thresh=thresh, maxval=255,
                           type=cv2.THRESH_TOZERO)
    return out.astype(np.uint8)


# ----------------------------------------------------------------------------
# a6  grayscale_opening / closing               image_filtering.py:319-322
# ----------------------------------------------------------------------------
def grayscale_opening(frame, SE):
    """scipy grey_opening with a flat ``SE``-sized structuring element
    (image_filtering.py:320)."""
    return ndimage.grey_opening(frame, size=SE).astype(np.uint8)


def grayscale_closing(frame, SE):
    """Dual of the opening (scipy grey_closing).  Not in the reference;
    named by BASELINE.json ("morphological open/close").  PARITY UNPINNED."""
    return ndimage.grey_closing(frame, size=SE).astype(np.uint8)


# ----------------------------------------------------------------------------
# a7  cc_labeling                               image_filtering.py:325-329
# ----------------------------------------------------------------------------
def cc_labeling(frame, connectivity):
    """``cv2.connectedComponents(frame, connectivity)`` called positionally
    (image_filtering.py:327): the 2nd positional slot of the Python binding is
    ``labels``, so ``connectivity`` is swallowed and the default (8) applies.
    Labels are then truncated to uint8 (image_filtering.py:329)."""
    _, labeled = cv2.connectedComponents(frame, connectivity)
    return labeled.astype(np.uint8)


def cc_labeling_i32(frame):
    """The same call without the uint8 truncation: int32 labels 1..n,
    8-connected, OpenCV numbering."""
    _, labeled = cv2.connectedComponents(frame)
    return labeled


def label_order_spec(frame):
    """Library-independent statement of OpenCV's numbering for 8-connectivity:
    components are numbered 1..n by ascending *minimum 2x2-block raster index*
    ``(y // 2) * ceil(W / 2) + x // 2`` over the component's pixels.  Uses
    scipy.ndimage.label only to find the components.  Returns int32 labels."""
    fg = frame != 0
    lab, n = ndimage.label(fg, structure=np.ones((3, 3), dtype=bool))
    if n == 0:
        return lab.astype(np.int32)
    h, w = fg.shape
    bw = (w + 1) // 2
    ys, xs = np.nonzero(fg)
    blk = (ys // 2) * bw + xs // 2
    comp = lab[ys, xs]
    minblk = np.full(n + 1, np.iinfo(np.int64).max, dtype=np.int64)
    np.minimum.at(minblk, comp, blk)
    order = np.argsort(minblk[1:], kind="stable")
    new = np.zeros(n + 1, dtype=np.int32)
    new[order + 1] = np.arange(1, n + 1, dtype=np.int32)
    return new[lab]


# ----------------------------------------------------------------------------
# a8  get_segment_properties                    image_filtering.py:332-335
# ----------------------------------------------------------------------------
class RegionProperties:
    """The four properties the rest of swiftwatcher consumes from a skimage
    0.15 ``_RegionProperties`` (label, area, bbox, centroid), as plain
    attributes (data_structures.py:28-30 copies whatever ``dir()`` shows)."""

    def __init__(self, label, area, bbox, centroid):
        self.label = label
        self.area = area
        self.bbox = bbox
        self.centroid = centroid

    def __repr__(self):
        return ("RegionProperties(label=%d, area=%d, bbox=%r, centroid=%r)"
                % (self.label, self.area, self.bbox, self.centroid))


def regionprops(label_image, coordinates=None):
    """Restates skimage 0.15 ``measure.regionprops`` (measure/_regionprops.py)
    for label/area/bbox/centroid: ``objects = ndimage.find_objects(label)``;
    for every non-``None`` slice ``i`` → label ``i+1``, ``image = label[sl] ==
    label``, ``area = image.sum()``, ``bbox = (starts..., stops...)``,
    ``centroid = coords.mean(axis=0)`` with coords in image space.  Regions
    come out in ascending label order.  ``coordinates`` only affects
    moments/orientation in 0.15 and is ignored."""
    out = []
    if label_image.ndim != 2:
        raise TypeError("Only 2-D images supported.")
    objects = ndimage.find_objects(label_image)
    for i, sl in enumerate(objects):
        if sl is None:
            continue
        label = i + 1
        image = label_image[sl] == label
        area = int(image.sum())
        rr, cc = np.nonzero(image)
        coords = np.stack([rr + sl[0].start, cc + sl[1].start], axis=1)
        centroid = tuple(coords.mean(axis=0))
        bbox = (sl[0].start, sl[1].start, sl[0].stop, sl[1].stop)
        out.append(RegionProperties(label, area, bbox, centroid))
    return out


def get_segment_properties(frame):
    """``measure.regionprops(frame, coordinates='xy')`` (image_filtering.py:335)."""
    return regionprops(frame, coordinates='xy')


# ----------------------------------------------------------------------------
# a9  extract_segment_images                    image_filtering.py:338-369
# ----------------------------------------------------------------------------
def expand_bbox(bbox, min_seg_size, crop_region):
    """bbox grown symmetrically to at least ``min_seg_size`` (floor/ceil split,
    image_filtering.py:350-358) and shifted by the crop origin (:361-362).
    Returns full-frame ``[r0, c0, r1, c1]``; no clamping."""
    bbox = list(bbox)
    dims = (bbox[2] - bbox[0], bbox[3] - bbox[1])
    if dims[0] < min_seg_size[0]:
        diff = min_seg_size[0] - dims[0]
        bbox[0] -= math.floor(diff / 2)
        bbox[2] += math.ceil(diff / 2)
    if dims[1] < min_seg_size[1]:
        diff2 = min_seg_size[1] - dims[1]
        bbox[1] -= math.floor(diff2 / 2)
        bbox[3] += math.ceil(diff2 / 2)
    oy, ox = crop_region[0][1], crop_region[0][0]
    return [bbox[0] + oy, bbox[1] + ox, bbox[2] + oy, bbox[3] + ox]


def extract_segment_images(segments, frame, min_seg_size, crop_region):
    """Colour crops from the *un-cropped* frame with numpy slice semantics
    (negative starts wrap, overflow truncates; image_filtering.py:363-365)."""
    images = []
    for segment in segments:
        b = expand_bbox(segment.bbox, min_seg_size, crop_region)
        images.append(frame[b[0]:b[2], b[1]:b[3]])
    return images


# ----------------------------------------------------------------------------
# Composition named by BASELINE.json north_star (stage order of
# data_structures.py:171-217 with RPCA+bilateral replaced by median+absdiff)
# ----------------------------------------------------------------------------
class PathParams:
    def __init__(self, crop_region, median_n=5, thresh=15, se=3,
                 do_open=True, do_close=False, label_mode="u8",
                 min_seg_size=(24, 24)):
        self.crop_region = crop_region      # [(x0, y0), (x1, y1)]
        self.median_n = median_n
        self.thresh = thresh
        self.se = se
        self.do_open = do_open
        self.do_close = do_close
        self.label_mode = label_mode        # "u8" (reference) | "i32"
        self.min_seg_size = min_seg_size


def filter_frame(gray_window, params):
    """gray window (oldest..newest) -> opened/closed grey image."""
    cur = gray_window[-1]
    bg = temporal_median(gray_window)
    fg = absdiff(cur, bg)
    th = thresh_to_zero(fg, params.thresh)
    out = th
    if params.do_open:
        out = grayscale_opening(out, (params.se, params.se))
    if params.do_close:
        out = grayscale_closing(out, (params.se, params.se))
    return out


def label_frame(filtered, params):
    if params.label_mode == "u8":
        return cc_labeling(filtered, 4)     # call site data_structures.py:206
    return cc_labeling_i32(filtered)


def run_path(frames_bgr, params, history=None, want_images=False):
    """Run the whole path over a list/array of full BGR frames.

    ``history``: optional list of up to N-1 BGR frames preceding
    ``frames_bgr[0]`` (the temporal halo).  Missing history replicates the
    first available frame (``window_indices``).

    Returns a list with one dict per output frame:
    ``mask`` (uint8 0/255), ``labels`` (uint8 or int32), ``props`` (list of
    RegionProperties) and, when ``want_images``, ``filtered`` and ``crops``."""
    n = params.median_n
    hist = list(history) if history is not None else []
    seq = hist + list(frames_bgr)
    n_hist = len(hist)
    grays = [convert_grayscale(crop_frame(f, params.crop_region)) for f in seq]
    out = []
    for t in range(n_hist, len(seq)):
        win = [grays[i] for i in window_indices(t, n)]
        filt = filter_frame(win, params)
        labels = label_frame(filt, params)
        props = get_segment_properties(labels)
        rec = {"mask": ((filt > 0).astype(np.uint8) * 255),
               "labels": labels, "props": props}
        if want_images:
            rec["filtered"] = filt
            rec["crops"] = extract_segment_images(
                props, seq[t], params.min_seg_size, params.crop_region)
        out.append(rec)
    return out


def props_table(props):
    """list[RegionProperties] -> (k, 9) float64 table
    [label, area, r0, c0, r1, c1, centroid_r, centroid_c, 0] for comparisons."""
    tab = np.zeros((len(props), 8), dtype=np.float64)
    for i, p in enumerate(props):
        tab[i] = (p.label, p.area, *p.bbox, *p.centroid)
    return tab


# ----------------------------------------------------------------------------
# The reference's own background model: rpca + bilateral_blur
# ----------------------------------------------------------------------------
def soft_threshold(values, t):
    """Elementwise shrinkage towards zero by t, written as the reference writes it (:283):
    max(v - t, 0) + min(v + t, 0)."""
    return np.maximum(values - t, 0) + np.minimum(values + t, 0)


def inexact_augmented_lagrange_multiplier(X, lmbda=0.01, tol=0.001, maxiter=100, skip_null=False):
    """Robust PCA by the inexact ALM iteration of image_filtering.py:256-301 — the same float64
    operations in the same order (numpy + LAPACK svd), so that the result is bit-identical to the
    reference's; ``oracle/make_golden.py rpca`` asserts that.  Returns (low_rank, sparse, iterations).

    Per iteration: sparse = shrink(X - low_rank + Y/mu, lmbda/mu); thin SVD of X - sparse + Y/mu;
    low_rank = U diag(S - 1/mu) V (all singular values are kept: ``(S > 1/mu).shape[0]`` at :285 is
    the length of S, not a count); residual Z = X - low_rank - sparse; Y += mu Z; mu *= 1.5.

    ``skip_null=True`` is NOT the reference: it leaves out the components whose singular value is
    (numerically) zero, which only exist when columns are exactly dependent — the all-zero frames the
    reader pads the last batch of a video with (io_video.py:40-44).  The reference adds
    -(1/mu) u_k v_k^T for them with whatever null-space basis LAPACK returns; the CUDA path skips them
    (DESIGN.md §8), and tests/test_rpca.py pins that behaviour against this variant."""
    from numpy.linalg import norm, svd
    flat = X.ravel()
    two_norm = norm(flat, 2)
    scale = np.max([two_norm, norm(flat, np.inf) / lmbda])     # dual norm (:271-273)
    multiplier = X / scale
    low_rank = np.zeros(multiplier.shape)
    sparse = np.zeros(multiplier.shape)
    x_norm = norm(X, 'fro')
    mu, growth, done = 1.25 / two_norm, 1.5, 0
    converged = False
    while not converged:
        sparse = soft_threshold(X - low_rank + (1 / mu) * multiplier, lmbda / mu)
        U, S, Vt = svd(X - sparse + (1 / mu) * multiplier, full_matrices=False)
        kept = int((S > 1e-12 * S[0]).sum()) if skip_null else S.shape[0]
        low_rank = np.dot(np.dot(U[:, :kept], np.diag(S[:kept] - 1 / mu)), Vt[:kept, :])
        residual = X - low_rank - sparse
        multiplier = multiplier + mu * residual
        mu = np.min([mu * growth, mu * 1e7])
        done += 1
        converged = (norm(residual, 'fro') / x_norm) < tol or done >= maxiter
    return low_rank, sparse, done


def rpca(frame_list, want_iters=False, skip_null=False):
    """image_filtering.py:220-253: list of gray frames -> list of uint8 "sparse" images
    (what is darker than the low-rank background)."""
    img_matrix = np.array(frame_list)
    col_matrix = np.transpose(img_matrix.reshape(img_matrix.shape[0], img_matrix.shape[1] * img_matrix.shape[2]))
    _, s_columns, itr = inexact_augmented_lagrange_multiplier(col_matrix, skip_null=skip_null)
    s_columns = np.negative(s_columns)
    s_columns = np.clip(s_columns, 0, 255).astype(np.uint8)
    out = [np.reshape(s_columns[:, i], (img_matrix.shape[1], img_matrix.shape[2]))
           for i in range(img_matrix.shape[0])]
    return (out, itr) if want_iters else out


def bilateral_blur(frame, d, sigmaColor, sigmaSpace):
    """image_filtering.py:304-307 (the reference's own cv2 call)."""
    return cv2.bilateralFilter(frame, d, sigmaColor, sigmaSpace).astype(np.uint8)


def bilateral_scalar(img, d=7, sigma_color=15.0, sigma_space=1.0):
    """OpenCV's scalar definition of bilateralFilter for 8-bit single-channel images, restated
    (library-independent): taps with sqrt(i^2 + j^2) <= d // 2 visited i-outer / j-inner (centre
    included), float32 weights (float)exp(.), BORDER_REFLECT_101, float32 accumulation in tap
    order, cvRound(sum / wsum).  Equals cv2 with setUseOptimized(False) bit for bit; cv2's SIMD
    body differs from it on about one pixel per million (exact .5 ties)."""
    radius = max(d // 2, 1)
    gc, gs = -0.5 / (sigma_color * sigma_color), -0.5 / (sigma_space * sigma_space)
    cw = np.array([np.float32(math.exp(i * i * gc)) for i in range(256)], dtype=np.float32)
    t = cv2.copyMakeBorder(img, radius, radius, radius, radius, cv2.BORDER_REFLECT_101)
    h, w = img.shape
    c = img.astype(np.int32)
    s = np.zeros((h, w), np.float32)
    ws = np.zeros((h, w), np.float32)
    for i in range(-radius, radius + 1):
        for j in range(-radius, radius + 1):
            r = math.sqrt(i * i + j * j)
            if r > radius:
                continue
            sw = np.float32(math.exp(r * r * gs))
            v = t[radius + i:radius + i + h, radius + j:radius + j + w]
            wk = (sw * cw[np.abs(v.astype(np.int32) - c)]).astype(np.float32)
            ws = (ws + wk).astype(np.float32)
            s = (s + (v.astype(np.float32) * wk).astype(np.float32)).astype(np.float32)
    return np.rint((s / ws).astype(np.float32)).astype(np.uint8)


def run_path_rpca(frames_bgr, params, want_images=False):
    """FrameQueue.preprocess_queue + segment_queue as the reference really runs them
    (data_structures.py:171-217) on ONE batch of frames given oldest first: crop, gray,
    rpca over the batch with the newest frame as column 0 (appendleft, :134,:160), bilateral
    (7, 15, 1), thresh_to_zero, opening, labelling, regionprops.  One dict per frame, oldest first."""
    grays = [convert_grayscale(crop_frame(f, params.crop_region)) for f in frames_bgr]
    sparse = rpca(grays[::-1])[::-1]
    out = []
    for t, sp in enumerate(sparse):
        bl = bilateral_blur(sp, 7, 15, 1)
        th = thresh_to_zero(bl, params.thresh)
        filt = th
        if params.do_open:
            filt = grayscale_opening(filt, (params.se, params.se))
        if params.do_close:
            filt = grayscale_closing(filt, (params.se, params.se))
        labels = label_frame(filt, params)
        props = get_segment_properties(labels)
        rec = {"mask": ((filt > 0).astype(np.uint8) * 255), "labels": labels, "props": props,
               "rpca": sp, "bilateral": bl}
        if want_images:
            rec["filtered"] = filt
            rec["crops"] = extract_segment_images(props, frames_bgr[t], params.min_seg_size, params.crop_region)
        out.append(rec)
    return out
