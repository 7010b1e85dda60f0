"""SegmentTracker with the cost matrix computed on the GPU.

Mirror of ``swiftwatcher/segment_tracking.py``: same class, method names, call order
(``set_current_frame`` -> ``formulate_cost_matrix`` -> ``store_assignments`` ->
``link_matching_segments`` -> ``check_for_events`` -> ``cache_current_frame``,
``__main__.py:85-91``), same cost model, same statuses ("A", "D", or the index of the
matched segment), same event rule.

The one thing that changes is how the (n_prev + n_curr)^2 matrix of
``formulate_cost_matrix`` (segment_tracking.py:46-102) is filled: the reference runs a Python
double loop with one ``scipy.spatial.distance.euclidean`` and two ``math.atan2`` calls per
pair — 250,000 iterations per frame at 500 segments, far slower than the filtering path that
feeds it — here one CUDA kernel (csrc/track.cu, ``swb_tracker_costs``) writes the whole matrix,
one thread per element, into page-locked memory:

    cost[i, n_prev + j] = 0.5 * 2^(dist(i, j) - 25) + 0.5 * angle_cost(i, j)
    angle_cost = 2^(min(|new - old|, 360 - |new - old|) - 90)  with a motion history, else 1

Same float64 formulas in the same operation order; CUDA's (and numpy's) ``atan2`` / ``exp2`` may
differ from libm's ``atan2`` / ``pow`` in the last bits, so parity is stated as 1e-12 relative on
the matrix and identical assignments (tests/test_tracking.py).  ``SegmentTracker(roi_mask,
device=None)`` keeps the whole tracker on the host (three numpy expressions for the match block):
the reference's tracker is host code, this is its vectorised form, not a fallback of the CUDA path.  The assignment problem itself
stays ``scipy.optimize.linear_sum_assignment`` on the host: the tracker is a strict
frame-to-frame recurrence over small matrices (SURVEY.md §8e "what stays serial").
"""

import ctypes as C
import sys

import numpy as np
from scipy.optimize import linear_sum_assignment

from . import _lib
from . import data_structures as ds
from ._lib import SwbError


class CostWorkspace:
    """Device + page-locked buffers of ``swb_tracker_costs`` (one per tracker / GPU)."""

    def __init__(self, device=0, max_segments=2048):
        self._lib = _lib.load()
        self._t = C.c_void_p()
        rc = self._lib.swb_tracker_create(int(device), int(max_segments), C.byref(self._t))
        if rc != 0:
            raise SwbError(rc, (self._lib.swb_tracker_last_error(None) or b"").decode())
        self.max_segments = int(max_segments)

    def costs(self, prev_yx, first_yx, has_history, curr_yx):
        """The (n_prev + n_curr)^2 matrix as a numpy view of the workspace's page-locked result buffer
        (valid until the next call)."""
        n_prev, n_curr = len(prev_yx), len(curr_yx)
        out = C.c_void_p()
        rc = self._lib.swb_tracker_costs(self._t, _lib.ptr(prev_yx), _lib.ptr(first_yx), _lib.ptr(has_history),
                                         n_prev, _lib.ptr(curr_yx), n_curr, C.byref(out))
        if rc != 0:
            raise SwbError(rc, (self._lib.swb_tracker_last_error(self._t) or b"").decode())
        n = n_prev + n_curr
        if n == 0:
            return np.zeros((0, 0))
        buf = (C.c_double * (n * n)).from_address(out.value)
        return np.frombuffer(buf, dtype=np.float64).reshape(n, n)

    def launch_count(self):
        return int(self._lib.swb_tracker_launch_count(self._t))

    def close(self):
        if getattr(self, "_t", None) and self._t.value:
            self._lib.swb_tracker_destroy(self._t)
            self._t = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _centroid_arrays(prev_segments, curr_segments):
    p = np.array([s.centroid for s in prev_segments], dtype=np.float64).reshape(-1, 2)
    c = np.array([s.centroid for s in curr_segments], dtype=np.float64).reshape(-1, 2)
    has = np.fromiter((len(s.segment_history) > 0 for s in prev_segments), dtype=np.uint8, count=len(prev_segments))
    first = p.copy()
    for i in np.flatnonzero(has):
        first[i] = prev_segments[i].segment_history[0].centroid
    return p, first, has, c


class SegmentTracker:
    """segment_tracking.py:17-152."""

    def __init__(self, roi_mask, device=0, max_segments=2048):
        """``device``: CUDA device whose kernel fills the cost matrix (default 0; raises without one —
        there is no silent fallback); ``None`` = the host (numpy) tracker.  ``max_segments``: the most
        segments two consecutive frames may hold together."""
        self.current_frame = None
        self.cached_frame = ds.Frame()
        self.roi_mask = roi_mask
        self.detected_events = []
        self._costs = CostWorkspace(device, max_segments) if device is not None else None

    def get_current_frame(self):
        return self.current_frame

    def get_cached_frame(self):
        return self.cached_frame

    def set_current_frame(self, frame):
        self.current_frame = frame

    def cache_current_frame(self):
        self.cached_frame = self.current_frame

    def formulate_cost_matrix(self):
        """segment_tracking.py:46-102, match block vectorised (layout: see the reference docstring)."""
        curr = self.current_frame.segments
        prev = self.cached_frame.segments
        n_curr, n_prev = len(curr), len(prev)
        if self._costs is not None:
            return self._costs.costs(*_centroid_arrays(prev, curr))
        cost = intialize_cost_matrix(n_curr, n_prev)
        if n_curr > 0 and n_prev > 0:
            cost[:n_prev, n_prev:] = match_costs(prev, curr)
        idx = np.arange(n_curr + n_prev)
        cost[idx, idx] = calculate_nonmatch_cost()
        return cost

    def store_assignments(self, assignments):
        """segment_tracking.py:104-131."""
        prev = self.cached_frame.segments
        curr = self.current_frame.segments
        n_prev = len(prev)
        for prev_label, v in enumerate(assignments[:n_prev]):
            a = int(v) - n_prev
            if a >= 0:
                prev[prev_label].status = a
                curr[a].status = prev_label
            else:
                prev[prev_label].status = "D"
        for curr_label, v in enumerate(assignments[n_prev:]):
            if int(v) - n_prev == curr_label:
                curr[curr_label].status = "A"

    def link_matching_segments(self):
        """segment_tracking.py:133-152: a matched segment takes over (and extends) the shared history list."""
        for segment in self.current_frame.segments:
            if segment.status != "A":
                matched = self.cached_frame.segments[segment.status]
                history = matched.segment_history
                history.append(matched)
                segment.segment_history = history

    def check_for_events(self):
        """segment_tracking.py:154-176."""
        for segment in self.cached_frame.segments:
            if segment.status != "D":
                continue
            pos = segment.centroid
            if self.roi_mask[int(pos[0]), int(pos[1])] != 255:
                continue
            if len(segment.segment_history) < 1:
                continue
            path = segment.segment_history
            path.append(segment)
            self.detected_events.append(path)


def intialize_cost_matrix(n_curr, n_prev):
    """segment_tracking.py:179-187 (the reference's spelling is kept: it is the public name)."""
    n_total = n_curr + n_prev
    return np.ones((n_total, n_total)) + sys.float_info.epsilon


def match_costs(prev_segments, curr_segments):
    """[n_prev, n_curr] float64: 0.5 * distance cost + 0.5 * angle cost
    (calculate_distance_cost :190-198, calculate_angle_cost :201-243) for every pair at once."""
    p = np.array([s.centroid for s in prev_segments], dtype=np.float64).reshape(-1, 2)
    c = np.array([s.centroid for s in curr_segments], dtype=np.float64).reshape(-1, 2)
    d = p[:, None, :] - c[None, :, :]                              # prev - curr: (del_y, del_x) of the new vector
    dist = np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1])
    with np.errstate(over="ignore"):                              # far-apart pairs: 2^x -> inf, as in the reference
        d_cost = np.exp2(dist - 25)
    has_hist = np.array([len(s.segment_history) > 0 for s in prev_segments], dtype=bool)
    a_cost = np.ones_like(d_cost)
    if has_hist.any():
        first = np.array([s.segment_history[0].centroid if h else s.centroid
                          for s, h in zip(prev_segments, has_hist)], dtype=np.float64).reshape(-1, 2)
        old = np.degrees(np.arctan2(first[:, 0] - p[:, 0], -1 * (first[:, 1] - p[:, 1])))   # per previous segment
        new = np.degrees(np.arctan2(d[..., 0], -1 * d[..., 1]))
        diff = np.abs(new - old[:, None])
        diff = np.minimum(diff, 360 - diff)
        a_cost = np.where(has_hist[:, None], np.exp2(diff - 90), 1.0)
    return 0.5 * d_cost + 0.5 * a_cost


def calculate_nonmatch_cost():
    """segment_tracking.py:246-250."""
    return 1


def apply_hungarian_algorithm(cost_matrix):
    """segment_tracking.py:253-260."""
    _, assignments = linear_sum_assignment(cost_matrix)
    return assignments
