"""Tracker cost matrix — the host (numpy) form and the CUDA kernel — against the
oracle's scalar restatement and, in the build container, the reference's own SegmentTracker."""
import functools
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

    host = functools.partial(st.SegmentTracker, device=None)
    statuses, events = run(host, st.apply_hungarian_algorithm, frames, roi, on_matrix=check)
    assert any("A" in s for s in statuses) and any(len(e) >= 2 for e in events)


@pytest.mark.parametrize("device", [None, pytest.param(0, marks=pytest.mark.gpu)])
def test_empty_frames_and_single_segments(device):
    from swiftwatcher_b200 import segment_tracking as st
    st_tracker = functools.partial(st.SegmentTracker, device=device)
    roi = np.full((50, 50), 255, np.uint8)
    frames = [np.zeros((0, 2)), np.array([[10.0, 10.0]]), np.zeros((0, 2)), np.array([[5.0, 5.0], [30.0, 30.0]]),
              np.array([[6.0, 7.0]]), np.zeros((0, 2))]
    statuses, events = run(st_tracker, st.apply_hungarian_algorithm, frames, roi)
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
    b = run(functools.partial(st.SegmentTracker, device=None), st.apply_hungarian_algorithm, frames, roi,
            on_matrix=lambda t, tr, m: np.testing.assert_allclose(m, mats[t], rtol=RTOL, atol=0))
    assert a[0] == b[0]
    assert a[1] == b[1] and len(a[1]) > 0


def test_tracker_needs_a_gpu_unless_told_otherwise():
    """The default tracker computes its matrix with the CUDA kernel and says so when there is no device."""
    import swiftwatcher_b200 as swb
    from swiftwatcher_b200 import segment_tracking as st
    if swb.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(swb.SwbError):
        st.SegmentTracker(np.zeros((4, 4), np.uint8))


@pytest.mark.gpu
def test_gpu_cost_matrix_equals_the_scalar_restatement_and_gives_the_same_assignments():
    """swb_tracker_costs against oracle/reference_tracker.py (the reference's scalar math / scipy calls) over a
    tracked swarm: matrix to 1e-12 relative, identical Hungarian assignments, statuses and events."""
    from swiftwatcher_b200 import segment_tracking as st
    rng = np.random.default_rng(2)
    frames = swarm(rng, 14, 120)
    roi = np.full((240, 320), 255, np.uint8)
    roi[:, :60] = 0
    seen = {"pairs": 0}

    def check(t, tr, m):
        want = rt.cost_matrix(tr.cached_frame.segments, tr.current_frame.segments)
        assert m.shape == want.shape
        np.testing.assert_allclose(m, want, rtol=RTOL, atol=0)
        assert np.array_equal(st.apply_hungarian_algorithm(m), st.apply_hungarian_algorithm(want))
        seen["pairs"] += (len(tr.cached_frame.segments) > 0) * (len(tr.current_frame.segments) > 0)

    gpu = run(functools.partial(st.SegmentTracker, device=0), st.apply_hungarian_algorithm, frames, roi, on_matrix=check)
    host = run(functools.partial(st.SegmentTracker, device=None), st.apply_hungarian_algorithm, frames, roi)
    assert gpu == host and seen["pairs"] >= 12 and len(gpu[1]) > 0


@pytest.mark.gpu
def test_gpu_cost_matrix_500_by_500_under_a_millisecond():
    """configs[4] scale: ~500 segments in each of two frames (the numpy form took 33 ms, the reference's
    Python loop 3 s); also exercises the capacity error."""
    import time
    from swiftwatcher_b200 import segment_tracking as st
    rng = np.random.default_rng(3)
    n = 500
    ws = st.CostWorkspace(0, max_segments=1024)
    p = rng.uniform(0, 1080, (n, 2))
    c = p + rng.normal(0, 5, (n, 2))
    first = p - rng.normal(0, 20, (n, 2))
    has = (rng.random(n) < 0.7).astype(np.uint8)
    m = ws.costs(p, first, has, c).copy()

    class S:
        def __init__(self, cen, hist):
            self.centroid, self.segment_history = tuple(cen), hist
    prev = [S(p[i], [S(first[i], [])] if has[i] else []) for i in range(n)]
    curr = [S(c[j], []) for j in range(n)]
    want = rt.cost_matrix(prev[:60], curr[:50])
    np.testing.assert_allclose(m[:60, n:n + 50], want[:60, 60:], rtol=RTOL, atol=0)
    assert np.all(m[np.arange(2 * n), np.arange(2 * n)] == 1.0)
    off = m[n:, :n]
    assert np.all(off == 1.0 + np.finfo(float).eps) and m[3, 7] == 1.0 + np.finfo(float).eps
    np.testing.assert_allclose(m[:n, n:], st.match_costs(prev, curr), rtol=RTOL, atol=0)
    for _ in range(3):
        ws.costs(p, first, has, c)
    t0 = time.perf_counter()
    for _ in range(20):
        ws.costs(p, first, has, c)
    dt = (time.perf_counter() - t0) / 20
    assert dt < 1e-3, "swb_tracker_costs took %.3f ms for 500 x 500" % (dt * 1e3)
    with pytest.raises(Exception):
        ws.costs(np.zeros((600, 2)), np.zeros((600, 2)), np.zeros(600, np.uint8), np.zeros((600, 2)))
    ws.close()

