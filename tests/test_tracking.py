"""Tracker with the vectorised cost matrix against the oracle's scalar
restatement and — in the build container — the reference's own SegmentTracker."""
import os

import numpy as np
import pytest

from oracle import reference_tracker as rt

REF = "/root/reference"
RTOL = 1e-12   # numpy arctan2 / exp2 vs libm atan2 / pow: last-bit differences only


class Seg:
    def __init__(self, centroid):
        self.centroid = (float(centroid[0]), float(centroid[1]))
        self.segment_history = []
        self.status = None


class Fr:
    def __init__(self, centroids, number):
        self.segments = [Seg(c) for c in centroids]
        self.frame_number = number

    def get_num_segments(self):
        return len(self.segments)


def swarm(rng, n_frames, n_birds, h=240, w=320):
    """Birds on straight paths with jitter; some vanish and new ones appear."""
    pos = rng.uniform([0, 0], [h, w], size=(n_birds, 2))
    vel = rng.normal(0, 6, size=(n_birds, 2))
    frames = []
    for t in range(n_frames):
        alive = rng.random(n_birds) > 0.1
        pts = pos[alive] + rng.normal(0, 0.7, size=(alive.sum(), 2))
        pts = pts[(pts[:, 0] >= 0) & (pts[:, 0] < h) & (pts[:, 1] >= 0) & (pts[:, 1] < w)]
        frames.append(pts)
        pos = pos + vel
        back = rng.random(n_birds) < 0.05
        pos[back] = rng.uniform([0, 0], [h, w], size=(back.sum(), 2))
    return frames


def run(tracker_cls, apply, frames, roi_mask, make_frame=Fr, on_matrix=None):
    tr = tracker_cls(roi_mask)
    statuses = []
    for t, pts in enumerate(frames):
        fr = make_frame(pts, t)
        tr.set_current_frame(fr)
        m = tr.formulate_cost_matrix()
        if on_matrix:
            on_matrix(t, tr, m)
        tr.store_assignments(apply(m))
        tr.link_matching_segments()
        tr.check_for_events()
        tr.cache_current_frame()
        statuses.append([s.status for s in fr.segments])
    events = [[(s.centroid[0], s.centroid[1]) for s in ev] for ev in tr.detected_events]
    return statuses, events


def test_cost_matrix_equals_the_scalar_restatement():
    from swiftwatcher_b200 import segment_tracking as st
    rng = np.random.default_rng(0)
    frames = swarm(rng, 12, 40)
    roi = np.full((240, 320), 255, np.uint8)
    roi[:, :80] = 0

    def check(t, tr, m):
        want = rt.cost_matrix(tr.cached_frame.segments, tr.current_frame.segments)
        assert m.shape == want.shape
        np.testing.assert_allclose(m, want, rtol=RTOL, atol=0)

    statuses, events = run(st.SegmentTracker, st.apply_hungarian_algorithm, frames, roi, on_matrix=check)
    assert any("A" in s for s in statuses) and any(len(e) >= 2 for e in events)


def test_empty_frames_and_single_segments():
    from swiftwatcher_b200 import segment_tracking as st
    roi = np.full((50, 50), 255, np.uint8)
    frames = [np.zeros((0, 2)), np.array([[10.0, 10.0]]), np.zeros((0, 2)), np.array([[5.0, 5.0], [30.0, 30.0]]),
              np.array([[6.0, 7.0]]), np.zeros((0, 2))]
    statuses, events = run(st.SegmentTracker, st.apply_hungarian_algorithm, frames, roi)
    assert statuses[1] == ["A"] and statuses[3] == ["A", "A"]
    assert len(events) == 1 and len(events[0]) == 2          # (5,5) -> (6,7) then disappears inside the ROI


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "swiftwatcher", "segment_tracking.py")),
                    reason="the reference only exists in the build container")
def test_same_statuses_and_events_as_the_reference_tracker():
    from swiftwatcher_b200 import segment_tracking as st
    ref = rt.reference_module(REF)
    rng = np.random.default_rng(1)
    frames = swarm(rng, 25, 60)
    roi = np.full((240, 320), 255, np.uint8)
    roi[100:, :] = 0
    mats = {}
    a = run(ref.SegmentTracker, ref.apply_hungarian_algorithm, frames, roi,
            on_matrix=lambda t, tr, m: mats.__setitem__(t, m.copy()))
    b = run(st.SegmentTracker, st.apply_hungarian_algorithm, frames, roi,
            on_matrix=lambda t, tr, m: np.testing.assert_allclose(m, mats[t], rtol=RTOL, atol=0))
    assert a[0] == b[0]
    assert a[1] == b[1] and len(a[1]) > 0
