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

    def get_frame(self, frame_number=None):          # :32-58
        if frame_number is None:
            frame_number = self.next_frame_number
        if not self.start_frame <= frame_number <= self.end_frame:
            frame = np.zeros(self.frame_shape).astype(np.uint8)
            frame_number = -1
            timestamp = "00:00:00.000"
        else:
            frame = self.read_frame(frame_number)
            timestamp = self.frame_number_to_timestamp(frame_number)
            if frame is None:
                frame = self.last_read_frame
                self.read_errors += 1
            else:
                self.frame_shape = frame.shape
                self.last_read_frame = frame
                self.frames_read += 1
        return frame, frame_number, timestamp

    def get_n_frames(self, n):                       # :60-72
        frames, numbers, stamps = [], [], []
        for _ in range(n):
            f, k, t = self.get_frame()
            frames.append(f)
            numbers.append(k)
            stamps.append(t)
        return frames, numbers, stamps

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
