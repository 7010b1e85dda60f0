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
``IngestRing`` puts a decode-ahead thread and a ring of such batches in front of any reader:
batch k+1 is decoded (and, through ``FrameQueue.attach_ring``, already submitted to the GPU)
while the caller is still tracking the frames of batch k.
"""

import queue as _queue
import threading

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


class IngestRing:
    """Decode-ahead ring in front of a ``FrameReader`` (the reference decodes, filters
    and tracks strictly one after the other: io_video.py:60-72, __main__.py:71-92).

    A background thread keeps calling ``reader.get_n_frames(batch_frames, out=<page-locked batch>)`` into a
    ring of ``depth`` batches; ``get_n_frames(n)`` hands the caller the oldest finished batch — the same
    three lists the reader returns, the frames being the rows of one page-locked array, so
    ``FrameQueue.segment_queue`` submits them in place.  The caller's last ``depth - 2`` batches stay
    intact (default depth 4: the current batch and the one before it, which the tracker's cached frame and
    the on-demand stage images may still look at); older ones are recycled for decoding.
    Every other attribute (``total_frames``, ``fps``, ``filepath`` ...) is the reader's."""

    def __init__(self, reader, batch_frames=21, depth=4, frame_shape=None, allocate=None):
        """``allocate(shape) -> uint8 array``: where the batches live; default page-locked memory
        (``_lib.pinned_empty``, needs a CUDA device)."""
        self._alloc = allocate
        if depth < 3:
            raise ValueError("depth must be at least 3 (one batch decoding, one ready, one in use)")
        self._reader = reader
        self._n = int(batch_frames)
        self._depth = int(depth)
        self._shape = tuple(frame_shape) if frame_shape is not None else None
        self._free = _queue.Queue()
        self._ready = _queue.Queue()
        self._lent = []
        self._error = None
        self._stop = False
        self._peeked = None
        self._thread = None

    def __getattr__(self, name):                       # total_frames, fps, filepath, read_frame ...
        return getattr(self.__dict__["_reader"], name)

    def _allocate(self, shape):
        if self._alloc is None:
            from ._lib import pinned_empty
            self._alloc = pinned_empty
        self._shape = tuple(shape)
        for _ in range(self._depth):
            self._free.put(self._alloc((self._n,) + self._shape))

    def _decode_loop(self):
        try:
            reader = self._reader
            while not self._stop:
                head = None
                if self._shape is None:                # the frame shape is unknown until one frame has been decoded
                    head = reader.get_frame()
                    if head[0] is None:
                        raise RuntimeError("IngestRing: the first frame could not be read")
                    self._allocate(head[0].shape)
                batch = self._free.get()
                if batch is None:
                    return
                if head is not None:
                    np.copyto(batch[0], head[0])
                    if reader.last_read_frame is head[0]:
                        reader.last_read_frame = batch[0]
                    frames, numbers, stamps = reader.get_n_frames(self._n - 1, out=batch[1:])
                    frames, numbers, stamps = [batch[0]] + frames, [head[1]] + numbers, [head[2]] + stamps
                else:
                    frames, numbers, stamps = reader.get_n_frames(self._n, out=batch)
                self._ready.put((batch, frames, numbers, stamps))
        except BaseException as e:                      # surfaced by the next get_n_frames
            self._error = e
            self._ready.put(None)

    def _start(self):
        if self._thread is None:
            if self._shape is not None:
                self._allocate(self._shape)
            self._thread = threading.Thread(target=self._decode_loop, name="swb-ingest", daemon=True)
            self._thread.start()

    def peek_next(self):
        """(batch array, frame_numbers) of the next batch if it has already been decoded — without
        consuming it — else None.  ``FrameQueue`` uses it to submit the batch to the GPU ahead of time."""
        self._start()
        if self._peeked is None:
            try:
                item = self._ready.get_nowait()
            except _queue.Empty:
                return None
            if item is None:
                self._ready.put(None)
                return None
            self._peeked = item
        return self._peeked[0], self._peeked[2]

    def get_n_frames(self, n=None):
        """The next batch: (frames, frame_numbers, timestamps) exactly as ``reader.get_n_frames(n)`` would
        have returned them (dummy frames past the end, last good frame on read errors, io_video.py:40-56)."""
        if n is not None and n != self._n:
            raise ValueError("IngestRing decodes batches of %d frames (asked for %d)" % (self._n, n))
        self._start()
        item, self._peeked = self._peeked, None
        if item is None:
            item = self._ready.get()
            if item is None:
                self._ready.put(None)
                raise RuntimeError("IngestRing: the decode thread failed: %r" % (self._error,)) from self._error
        batch, frames, numbers, stamps = item
        self._lent.append(batch)
        while len(self._lent) > max(self._depth - 2, 1):   # recycle what the caller can no longer be using
            self._free.put(self._lent.pop(0))
        return frames, numbers, stamps

    def close(self):
        self._stop = True
        self._free.put(None)
        if self._thread is not None:
            self._thread.join(timeout=5)
