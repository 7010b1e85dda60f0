"""Oracle for the frame reader semantics — TEST INFRASTRUCTURE ONLY.

``RefFrameReader`` restates ``FrameReader.get_frame`` / ``get_n_frames`` /
``frame_number_to_timestamp`` of swiftwatcher/io_video.py:11-82 line for line;
``reference_module()`` imports the reference's own io_video.py unmodified (with the h5py import
shim of oracle/_shim) so that the tests can pin the restatement and the product against the
real base class wherever /root/reference exists.
"""
import importlib
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))


class RefFrameReader:
    """io_video.py:11-82 with the frame source passed in as a callable."""

    def __init__(self, read, fps, start, end, frame_shape=(0, 0, 0)):
        self.fps = fps
        self.start_frame = start
        self.end_frame = end
        self.total_frames = end - start
        self.next_frame_number = start
        self.frame_shape = frame_shape
        self.last_read_frame = None
        self.frames_read = 0
        self.read_errors = 0
        self._read = read

    def read_frame(self, frame_number):
        frame = self._read(frame_number)
        self.next_frame_number += 1
        return frame

    def get_frame(self, frame_number=None):
        """io_video.py:32-58 as three cases."""
        k = self.next_frame_number if frame_number is None else frame_number
        inside = self.start_frame <= k <= self.end_frame
        if not inside:
            # :40-44 — a request past either end: black frame of the last known shape, -1, string stamp
            return np.zeros(self.frame_shape).astype(np.uint8), -1, "00:00:00.000"
        pixels = self.read_frame(k)
        stamp = self.frame_number_to_timestamp(k)
        if pixels is None:
            # :51-53 — the source failed: hand out the previous good frame again, count the error
            self.read_errors += 1
            return self.last_read_frame, k, stamp
        # :54-57 — a good frame updates shape, fallback frame and counter
        self.frame_shape, self.last_read_frame = pixels.shape, pixels
        self.frames_read += 1
        return pixels, k, stamp

    def get_n_frames(self, n):
        """io_video.py:60-72: n requests, transposed into three lists."""
        got = [self.get_frame() for _ in range(n)]
        return [g[0] for g in got], [g[1] for g in got], [g[2] for g in got]

    def frame_number_to_timestamp(self, frame_number):   # :74-82
        total_s = frame_number / self.fps
        timestamp = pd.Timestamp("00:00:00.000") + pd.Timedelta(total_s, 's')
        return timestamp.round(freq='us')


def reference_module(root="/root/reference"):
    sys.path.insert(0, os.path.join(HERE, "_shim"))
    sys.path.insert(0, root)
    try:
        sys.modules.pop("swiftwatcher.io_video", None)
        return importlib.import_module("swiftwatcher.io_video")
    finally:
        sys.path.remove(root)
        sys.path.remove(os.path.join(HERE, "_shim"))
