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


DUMMY_TIMESTAMP = "00:00:00.000"


class FrameReader:
    """Base of the readers: bookkeeping + the error policy of io_video.py:11-82; a subclass
    supplies ``read_frame(frame_number)`` -> ndarray or None."""

    def __init__(self):
        self.fps = 0
        self.start_frame = self.end_frame = self.total_frames = 0
        self.next_frame_number = 0
        self.frame_shape = (0, 0, 0)
        self.last_read_frame = None
        self.frames_read = self.read_errors = 0

    def __init_subclass__(cls, **kwargs):
        super().__init_subclass__(**kwargs)
        if getattr(cls, "read_frame", None) is None:
            raise NotImplementedError("Derived FrameReader must implement read_frame() method.")

    # -- the three outcomes of a request (io_video.py:40-56) --------------------------------
    def _outside(self):
        """frame number outside [start_frame, end_frame]: zero frame, number -1, string stamp"""
        return np.zeros(self.frame_shape).astype(np.uint8), -1, DUMMY_TIMESTAMP

    def _decoded(self, frame):
        self.frame_shape = frame.shape
        self.last_read_frame = frame
        self.frames_read += 1
        return frame

    def _failed(self):
        self.read_errors += 1
        return self.last_read_frame

    def get_frame(self, frame_number=None, out=None):
        """(frame, frame_number, timestamp) with the reference's handling of bad requests and
        read errors.  ``out``: optional array of the frame's shape that receives the pixels."""
        wanted = self.next_frame_number if frame_number is None else frame_number
        if wanted < self.start_frame or wanted > self.end_frame:
            frame, wanted, stamp = self._outside()
        else:
            raw = self.read_frame(wanted)
            stamp = self.frame_number_to_timestamp(wanted)
            frame = self._failed() if raw is None else self._decoded(raw)
        if out is not None and frame is not None and tuple(out.shape) == tuple(frame.shape):
            if frame is not out:
                np.copyto(out, frame)
            if frame is self.last_read_frame:          # keep the fallback frame alive inside the batch
                self.last_read_frame = out
            frame = out
        return frame, wanted, stamp

    def get_n_frames(self, n, out=None):
        """n consecutive ``get_frame`` calls as three lists (io_video.py:60-72).  ``out``: optional
        [>= n, H, W, C] array (e.g. a pinned batch); frame i lands in out[i]."""
        triples = [self.get_frame(out=None if out is None else out[i]) for i in range(n)]
        return [t[0] for t in triples], [t[1] for t in triples], [t[2] for t in triples]

    def frame_number_to_timestamp(self, frame_number):
        """Constant-FPS stamp, microsecond resolution (io_video.py:74-82)."""
        elapsed = pd.Timedelta(frame_number / self.fps, 's')
        return (pd.Timestamp(DUMMY_TIMESTAMP) + elapsed).round(freq='us')


class VideoReader(FrameReader):
    """cv2.VideoCapture source with the reference's grab-ahead order (io_video.py:133-165):
    one frame is always grabbed in advance, ``read_frame`` retrieves it and grabs the next."""

    def __init__(self, filepath, end):
        super().__init__()
        import cv2
        self.filepath = filepath
        cap = cv2.VideoCapture(str(filepath))
        cap.grab()
        self.vid_cap = cap
        self.fps = cap.get(cv2.CAP_PROP_FPS)
        self.start_frame = 0
        self.end_frame = end if end > 0 else int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
        self.next_frame_number = self.start_frame
        self.total_frames = self.end_frame - self.start_frame

    def read_frame(self, frame_number, increment=True):
        ok_and_frame = self.vid_cap.retrieve()
        if increment:
            self.vid_cap.grab()
            self.next_frame_number += 1
        return ok_and_frame[1]


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
