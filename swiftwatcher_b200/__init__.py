"""swiftwatcher_b200 — B200-native drop-in for swiftwatcher's per-frame filtering
and segmentation hot path (reference: swiftwatcher/image_filtering.py and its
only caller, FrameQueue.preprocess_queue/segment_queue in data_structures.py).

Layout
* ``csrc/``            hand-written sm_100a CUDA kernels + the C ABI (include/swb200.h)
* ``_lib``             ctypes binding of that ABI
* ``pipeline``         FilterContext: the fused, batched path (swb_submit/collect)
* ``image_filtering``  the reference's function signatures, one CUDA stage each
* ``data_structures``  Frame / Segment / FrameQueue mirror with the fused path inside
* ``chunking``         temporal-chunk partitioning across GPUs (halo N-1, no collective)
* ``io_video``         FrameReader / VideoReader mirror, pinned batches, decode-ahead IngestRing
* ``segment_classification``  SegmentClassifier fed batched crops (device crops: swb_gather_crops)
* ``segment_tracking`` SegmentTracker with the cost matrix as a CUDA kernel (swb_tracker_costs)
"""

from ._lib import (HALO_CARRY, LABELS_I32, LABELS_U8, MEM_DEVICE, MEM_HOST,  # noqa: F401
                   OUT_LABELS, OUT_MASK, SEGMENT_DTYPE, SwbError, device_count)
from .pipeline import FilterContext, RegionProperties, props_from_rows  # noqa: F401

__version__ = "0.2.0"
