"""GPU suite, part 2 (B200): parity at the GEOMETRY THE BENCH RUNS.

bench.py processes BASELINE.json's configs in large submits (1080p N=5: T = 1024; 4K N=9 5x5
open+close: T = 256), where the filtering kernel cuts the submit into temporal sub-chunks and runs
2-3 CTAs per SM through its mbarrier pipeline.  The small-size tests never reach that geometry, so
here the same submit the bench makes is compared with the oracle:

* masks, int32 labels and table rows of >= 12 frames (first, last, and the frames on either side of
  temporal sub-chunk boundaries) against ``oracle.reference_path.run_path`` on the same frames;
* every frame of the one-submit result against the same video fed in 8 submits with the carried
  history (``SWB_HALO_CARRY``) — on the device, plus SHA-256 of all masks on the host;
* the submit repeated 20 times must reproduce itself bit for bit (rare-race detector);
* size-independent properties over all T frames (area checksum, label <-> mask consistency).

Reference functions restated by the oracle: swiftwatcher/image_filtering.py:188-203, :310-335 and
the batch flow of swiftwatcher/data_structures.py:171-217.
"""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import reference_path as rp

import swiftwatcher_b200 as swb
from swiftwatcher_b200.pipeline import centroids, synth_frames

SEED = 2          # bench.py's seed and first frame: the very chunk the bench times
T0 = 1000


def boundary_frames(T, ts, want=12):
    """First / last frames and both sides of temporal sub-chunk boundaries (k * ts)."""
    picks = {0, 1, T - 2, T - 1}
    edges = list(range(ts, T, ts))
    # spread over the submit: first, middle and last boundaries first
    order = sorted(edges, key=lambda e: min(abs(e - edges[0]), abs(e - edges[-1]), abs(e - edges[len(edges) // 2]))) \
        if edges else []
    for e in order:
        for t in (e - 1, e, e + 1):
            if 0 <= t < T:
                picks.add(t)
        if len(picks) >= want:
            break
    t = T // 3
    while len(picks) < min(want, T):
        picks.add(t % T)
        t += 7
    return sorted(picks)


def oracle_frame(dev, t, halo, par):
    """The oracle's record of output frame t of the submit (input frames t .. t + halo of `dev`)."""
    host = dev[t:t + halo + 1].cpu().numpy()
    return rp.run_path(host[halo:], par, history=list(host[:halo]))[0]


def compare_frame(rec, mask, labels, rows_t, t):
    assert np.array_equal(mask, rec["mask"]), "mask differs at frame %d" % t
    assert np.array_equal(labels, rec["labels"]), "labels differ at frame %d" % t
    exp = rp.props_table(rec["props"])
    assert len(rows_t) == len(exp), "segment count differs at frame %d" % t
    got = np.zeros((len(rows_t), 8))
    got[:, 0], got[:, 1], got[:, 2:6] = rows_t["label"], rows_t["area"], rows_t["bbox"]
    if len(rows_t):
        got[:, 6:8] = centroids(rows_t)
    assert np.array_equal(got[:, :6], exp[:, :6]), "label/area/bbox differ at frame %d" % t
    np.testing.assert_allclose(got[:, 6:], exp[:, 6:], rtol=1e-5, atol=0)    # north_star tolerance
    assert np.array_equal(got[:, 6:], exp[:, 6:])                            # integer sums: in fact bit-exact


def bench_geometry(H, W, n, se, do_close, T, birds, repeats=20, pieces=8, subchunk=0):
    import torch
    halo = n - 1
    dev = torch.empty((halo + T, H, W, 3), dtype=torch.uint8, device="cuda")
    synth_frames(SEED, 0, T0 - halo, halo + T, H, W, birds, out=dev)
    par = rp.PathParams([(0, 0), (W, H)], n, 15, se, True, do_close, "i32")
    with swb.FilterContext((H, W, 3), None, median_n=n, morph_size=se, do_close=do_close, label_mode="i32",
                           max_frames=T, max_segments=T * 4096) as ctx:
        if subchunk:
            ctx.set_option("temporal_subchunk", subchunk)
        ctx.submit(dev, n_halo=halo)
        rows, counts = ctx.collect()
        ts = ctx.last_subchunk()
        assert 0 < ts <= T
        offs = np.concatenate([[0], np.cumsum(counts)])
        m_dev, l_dev = ctx.device_views()
        m_ref, l_ref = m_dev[:, :, :W].clone(), l_dev[:, :, :W].clone()

        # ---- the oracle on frames at the sub-chunk boundaries
        picks = boundary_frames(T, ts)
        assert len(picks) >= min(12, T)
        for t in picks:
            rec = oracle_frame(dev, t, halo, par)
            compare_frame(rec, m_ref[t].cpu().numpy(), l_ref[t].cpu().numpy(), rows[offs[t]:offs[t + 1]], t)

        # ---- size-independent properties over the whole submit
        assert torch.equal(l_ref > 0, m_ref == 255)
        assert int(rows["area"].sum()) == int((m_ref == 255).sum().item())            # checksum of checksums
        assert counts.min() > 0 and int(l_ref.amax().item()) == int(counts.max())

        # ---- the same submit again and again: bit-identical (pipeline races would show here)
        for i in range(repeats):
            ctx.submit(dev, n_halo=halo)
            rows_i, counts_i = ctx.collect()
            assert np.array_equal(counts_i, counts) and np.array_equal(rows_i, rows), "table changed in repeat %d" % i
            assert torch.equal(m_dev[:, :, :W], m_ref), "masks changed in repeat %d" % i
            assert torch.equal(l_dev[:, :, :W], l_ref), "labels changed in repeat %d" % i

        # ---- one submit == the same frames in `pieces` submits with the carried history
        sha_one = hashlib.sha256(m_ref.cpu().numpy().tobytes()).hexdigest()
        sha_parts = hashlib.sha256()
        step = T // pieces
        ctx.reset()
        got_rows = []
        for k in range(pieces):
            a, b = k * step, (T if k == pieces - 1 else (k + 1) * step)
            if k == 0:
                ctx.submit(dev[:halo + b], n_halo=halo)
            else:
                ctx.submit(dev[halo + a:halo + b])                       # SWB_HALO_CARRY
            r, c = ctx.collect()
            assert np.array_equal(c, counts[a:b]), "counts differ in piece %d" % k
            r = r.copy()
            r["frame"] += a
            got_rows.append(r)
            mk, lk = ctx.device_views()
            assert torch.equal(mk[:b - a, :, :W], m_ref[a:b]), "masks differ in piece %d" % k
            assert torch.equal(lk[:b - a, :, :W], l_ref[a:b]), "labels differ in piece %d" % k
            sha_parts.update(mk[:b - a, :, :W].cpu().numpy().tobytes())
        assert np.array_equal(np.concatenate(got_rows), rows)
        assert sha_parts.hexdigest() == sha_one
    return len(picks), ts


def test_bench_geometry_1080p_n5_open3_t1024():
    n, ts = bench_geometry(1080, 1920, 5, 3, False, 1024, 300)
    assert n >= 12 and ts < 1024          # the submit really was cut into temporal sub-chunks


def test_bench_geometry_4k_n9_openclose5_t256():
    n, ts = bench_geometry(2160, 3840, 9, 5, True, 256, 600)
    assert n >= 12 and ts == 256          # 4050 column blocks fill the GPU: no temporal split at this size


def test_bench_geometry_4k_n9_with_temporal_subchunks():
    """The same 4K submit with the filtering kernel forced to cut it into sub-chunks of 66 frames (what a
    smaller grid would get): boundaries at 66, 132, 198 against the oracle, and nothing else may change."""
    n, ts = bench_geometry(2160, 3840, 9, 5, True, 256, 600, repeats=3, pieces=2, subchunk=66)
    assert n >= 12 and ts == 66


def test_bench_geometry_dense_swarm_t512():
    """~2,600 segments per frame: the labelling tiles' overflow paths at full size."""
    bench_geometry(1080, 1920, 5, 3, False, 512, 2500, repeats=5, pieces=4)


def test_bench_geometry_u8_labels_match_i32_mod_256():
    """Reference-compat labels at bench geometry: the uint8 image is the int32 image mod 256 and the
    merged table equals regionprops of it (oracle on 4 frames; the swarm has > 255 components)."""
    import torch
    H, W, n, T, birds = 1080, 1920, 5, 256, 500
    halo = n - 1
    dev = torch.empty((halo + T, H, W, 3), dtype=torch.uint8, device="cuda")
    synth_frames(SEED, 0, T0 - halo, halo + T, H, W, birds, out=dev)
    with swb.FilterContext((H, W, 3), None, median_n=n, label_mode="i32", max_frames=T, max_segments=T * 4096) as c32:
        c32.submit(dev, n_halo=halo)
        _, counts32 = c32.collect()
        _, l32 = c32.device_views()
        l32 = l32[:, :, :W].clone()
    assert counts32.max() > 255
    par = rp.PathParams([(0, 0), (W, H)], n, 15, 3, True, False, "u8")
    with swb.FilterContext((H, W, 3), None, median_n=n, label_mode="u8", max_frames=T, max_segments=T * 4096) as c8:
        c8.submit(dev, n_halo=halo)
        rows, counts = c8.collect()
        m8, l8 = c8.device_views()
        assert torch.equal(l8[:, :, :W], (l32 & 0xFF).to(torch.uint8))
        offs = np.concatenate([[0], np.cumsum(counts)])
        for t in (0, 97, 98, T - 1):
            rec = oracle_frame(dev, t, halo, par)
            compare_frame(rec, m8[t, :, :W].cpu().numpy(), l8[t, :, :W].cpu().numpy(), rows[offs[t]:offs[t + 1]], t)
