"""Frame ingest — mirror of ``swiftwatcher/io_video.py`` that decodes into pinned host memory.

Same classes and semantics as the reference:
* ``FrameReader.get_frame`` (io_video.py:32-58): a frame number outside
  ``[start_frame, end_frame]`` yields a zero frame of the last known shape, frame number -1
  and the string timestamp "00:00:00.000" (:40-44); a failed read returns the last good
  frame and counts a read error (:51-53); timestamps are ``pd.Timestamp`` values rounded to
  microseconds (:74-82);
* ``get_n_frames`` (:60-72) returns three lists;
* ``VideoReader`` (:133-165) wraps ``cv2.VideoCapture`` with the reference's
  grab-ahead / retrieve order.  (``HDF5Reader`` needs h5py, which is optional here exactly
  as it is a hard import in the reference.)

What is new: ``get_n_frames(n, out=batch)`` writes the frames straight into the rows of a
caller-supplied array — ``FrameQueue.pinned_batch`` hands out page-locked batches
(``swb_host_alloc``) — so the frames the queue later submits are already where the DMA
engine can take them: no ``np.stack`` copy, no pageable staging copy inside the driver.
"""

import numpy as np
import pandas as pd


class FrameReader:
    """Base class for reading frames from a video source (io_video.py:11-82)."""

    def __init__(self):
        self.fps = 0
        self.start_frame = 0
        self.end_frame = 0
        self.total_frames = 0
        self.next_frame_number = 0

        self.frame_shape = (0, 0, 0)
        self.last_read_frame = None
        self.frames_read = 0
        self.read_errors = 0

    def __init_subclass__(cls, **kwargs):
        super().__init_subclass__(**kwargs)
        if not hasattr(cls, "read_frame"):
            raise NotImplementedError("Derived FrameReader must implement read_frame() method.")

    def get_frame(self, frame_number=None, out=None):
        """Returns frame, frame_number, and timestamp while also handling read errors
        (io_video.py:32-58).  ``out``: optional array the frame is written into."""
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

        if out is not None and frame is not None and tuple(out.shape) == tuple(frame.shape):
            if frame is not out:
                np.copyto(out, frame)
            if frame is self.last_read_frame:
                self.last_read_frame = out
            frame = out
        return frame, frame_number, timestamp

    def get_n_frames(self, n, out=None):
        """Calls get_frame in batches of N, returning as lists (io_video.py:60-72).
        ``out``: optional [>= n, H, W, C] array (e.g. a pinned batch); frame i lands in out[i]."""
        frames, frame_numbers, timestamps = [], [], []
        for i in range(n):
            frame, frame_number, timestamp = self.get_frame(out=None if out is None else out[i])
            frames.append(frame)
            frame_numbers.append(frame_number)
            timestamps.append(timestamp)
        return frames, frame_numbers, timestamps

    def frame_number_to_timestamp(self, frame_number):
        """io_video.py:74-82 (constant-FPS assumption)."""
        total_s = frame_number / self.fps
        timestamp = pd.Timestamp("00:00:00.000") + pd.Timedelta(total_s, 's')
        timestamp = timestamp.round(freq='us')
        return timestamp


class VideoReader(FrameReader):
    """Subclass using OpenCV's VideoCapture as frame source (io_video.py:133-165)."""

    def __init__(self, filepath, end):
        super().__init__()
        import cv2
        self._cv2 = cv2
        self.filepath = filepath
        self.vid_cap = cv2.VideoCapture(str(filepath))
        self.vid_cap.grab()  # Load first frame so retrieve() won't fail

        self.fps = self.vid_cap.get(cv2.CAP_PROP_FPS)
        self.start_frame = 0
        if end > 0:
            self.end_frame = end
        else:
            self.end_frame = int(self.vid_cap.get(cv2.CAP_PROP_FRAME_COUNT))

        self.next_frame_number = self.start_frame
        self.total_frames = self.end_frame - self.start_frame

    def read_frame(self, frame_number, increment=True):
        _, frame = self.vid_cap.retrieve()
        if increment:
            self.vid_cap.grab()
            self.next_frame_number += 1
        return frame


class ArrayReader(FrameReader):
    """Frames from memory or from a callable ``frame_number -> ndarray | None`` (synthetic
    video, tests).  Same bookkeeping as the file readers."""

    def __init__(self, source, fps=30.0, start=0, end=0, total=None):
        super().__init__()
        self.filepath = None
        self._source = source
        n = len(source) if total is None and hasattr(source, "__len__") else int(total or 0)
        self.fps = fps
        self.start_frame = start
        self.end_frame = end if end > 0 else n
        self.next_frame_number = self.start_frame
        self.total_frames = self.end_frame - self.start_frame

    def read_frame(self, frame_number, increment=True):
        try:
            frame = self._source(frame_number) if callable(self._source) else self._source[frame_number]
        except (IndexError, ValueError):
            frame = None
        if increment:
            self.next_frame_number += 1
        return frame
