"""Oracle for the tracker's cost matrix — TEST INFRASTRUCTURE ONLY.

``pair_cost`` restates ``calculate_distance_cost`` / ``calculate_angle_cost``
(swiftwatcher/segment_tracking.py:190-243) with the same scalar ``math`` / scipy calls, and
``cost_matrix`` the double loop of ``formulate_cost_matrix`` (:86-102).  ``reference_module()``
imports the reference's own segment_tracking.py unmodified (skimage import shim) so the tests
can run the real ``SegmentTracker`` next to the product wherever /root/reference exists.
"""
import importlib
import math
import os
import sys
import warnings

import numpy as np
from scipy.spatial import distance

HERE = os.path.dirname(os.path.abspath(__file__))


def pair_cost(curr, prev):
    dist = distance.euclidean(prev.centroid, curr.centroid)
    d_cost = 2 ** (dist - 25)
    if len(prev.segment_history) > 0:
        cp, pp, ip = curr.centroid, prev.centroid, prev.segment_history[0].centroid
        old = math.degrees(math.atan2(ip[0] - pp[0], -1 * (ip[1] - pp[1])))
        new = math.degrees(math.atan2(pp[0] - cp[0], -1 * (pp[1] - cp[1])))
        diff = abs(new - old)
        diff = min(diff, 360 - diff)
        a_cost = 2 ** (diff - 90)
    else:
        a_cost = 1
    return 0.5 * d_cost + 0.5 * a_cost


def cost_matrix(prev_segments, curr_segments):
    n_prev, n_curr = len(prev_segments), len(curr_segments)
    n = n_prev + n_curr
    cost = np.ones((n, n)) + sys.float_info.epsilon
    if n_curr > 0 and n_prev > 0:
        for i, sp in enumerate(prev_segments):
            for j, sc in enumerate(curr_segments):
                cost[i, j + n_prev] = pair_cost(sc, sp)
    for i in range(n):
        cost[i, i] = 1
    return cost


def reference_module(root="/root/reference"):
    sys.path.insert(0, os.path.join(HERE, "_shim"))
    sys.path.insert(0, root)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for m in ("swiftwatcher.segment_tracking", "swiftwatcher.data_structures"):
                sys.modules.pop(m, None)
            return importlib.import_module("swiftwatcher.segment_tracking")
    finally:
        sys.path.remove(root)
        sys.path.remove(os.path.join(HERE, "_shim"))
