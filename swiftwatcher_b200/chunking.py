"""Temporal-chunk partitioning of one video across GPUs.

Every output frame depends only on input frames [t-N+1, t]; morphology,
labelling and regionprops are per frame.  So rank r of R gets a contiguous
range of output frames plus the N-1 preceding input frames as a read-only
halo, and there is no collective on the hot path: the only cross-rank step is
a host-side gather of the per-chunk segment tables, concatenated in chunk
order (frame numbers are global, labels are per frame).

In the reference the analogue is ``--start/--end`` (io_video.py:107-114) and
the independent 21-frame batches of ``__main__.py:71-78``.
"""

import numpy as np

from ._lib import SEGMENT_DTYPE


def rank_range(total_frames, rank, world_size):
    """Contiguous output-frame range [t0, t1) of ``rank``; sizes differ by <= 1."""
    base, rem = divmod(int(total_frames), int(world_size))
    t0 = rank * base + min(rank, rem)
    return t0, t0 + base + (1 if rank < rem else 0)


def plan_chunks(t0, t1, chunk_frames, median_n):
    """Split [t0, t1) into submits of at most ``chunk_frames`` output frames.
    Yields (first_input_frame, n_halo, first_output_frame, n_frames): the
    submit must be given input frames [first_input_frame, first_output_frame +
    n_frames).  History before frame 0 of the video does not exist (the
    library replicates the earliest frame it is given)."""
    t = t0
    while t < t1:
        n = min(chunk_frames, t1 - t)
        halo = min(median_n - 1, t)
        yield t - halo, halo, t, n
        t += n


def run_rank(ctx, read_frames, total_frames, rank=0, world_size=1, chunk_frames=None):
    """Process this rank's share of a video.

    ``read_frames(a, b)`` returns input frames [a, b) as a numpy array or CUDA
    tensor.  Every chunk carries an explicit halo, so the ranks and chunks are
    independent.  Returns (rows, counts, t0, t1) with rows' ``frame`` field
    holding GLOBAL frame numbers."""
    chunk_frames = chunk_frames or ctx.max_frames
    t0, t1 = rank_range(total_frames, rank, world_size)
    all_rows, all_counts = [], []
    for first_in, halo, first_out, n in plan_chunks(t0, t1, chunk_frames, ctx.median_n):
        ctx.submit(read_frames(first_in, first_out + n), n_halo=halo)
        rows, counts = ctx.collect()
        rows = rows.copy()
        rows["frame"] += first_out
        all_rows.append(rows)
        all_counts.append(counts)
    rows = np.concatenate(all_rows) if all_rows else np.empty(0, SEGMENT_DTYPE)
    counts = np.concatenate(all_counts) if all_counts else np.empty(0, np.int32)
    return rows, counts, t0, t1


def gather_tables(rows, counts, group=None):
    """Host-side gather of every rank's segment table in rank (= frame) order
    via ``torch.distributed`` (NCCL or gloo); returns the global table on every
    rank.  The tables are tens of bytes per segment: latency-only traffic."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return rows, counts
    gathered = [None] * dist.get_world_size(group)
    dist.all_gather_object(gathered, (rows.tobytes(), counts.tobytes()), group=group)
    rows = np.concatenate([np.frombuffer(r, dtype=SEGMENT_DTYPE) for r, _ in gathered])
    counts = np.concatenate([np.frombuffer(c, dtype=np.int32) for _, c in gathered])
    return rows, counts
