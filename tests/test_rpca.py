"""The reference's own background model: rpca (IALM) + bilateral_blur.

Golden vectors tests/golden/rpca_*.npz were produced by the reference's own rpca(),
bilateral_blur(), thresh_to_zero(), grayscale_opening(), cc_labeling() and
get_segment_properties() (oracle/make_golden.py rpca).  Parity bar: the float64 IALM goes
through a Gram-matrix eigen-decomposition on the GPU instead of LAPACK's SVD, so the uint8
"RPCA" image may differ where -E is within rounding of an integer; on the golden batches and on
the seeded cases below it is required to be identical, the tolerance written in the asserts is
what the test would still accept (<= 1 grey level on <= 1e-5 of the pixels)."""
import os

import numpy as np
import pytest

from oracle import reference_path as rp
from oracle import synth

CASES = ["rpca_roi_batch21", "rpca_full_batch8"]


def load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    seed, video, H, W, birds, T, x0, y0, x1, y1, iters = [int(v) for v in z["cfg"]]
    frames = synth.synth_video(seed, video, 0, T, H, W, birds)
    return z, frames, [(x0, y0), (x1, y1)], iters


def close_u8(a, b, frac=1e-5):
    d = np.abs(a.astype(np.int16) - b.astype(np.int16))
    return d.max(initial=0) <= 1 and (d > 0).mean() <= frac


# ---------------------------------------------------------------------------------- CPU: the oracle
@pytest.mark.parametrize("name", CASES)
def test_oracle_restatement_reproduces_the_reference_outputs(golden_dir, name):
    z, frames, roi, iters = load(golden_dir, name)
    grays = [rp.convert_grayscale(rp.crop_frame(f, roi)) for f in frames]
    assert np.array_equal(np.stack(grays), z["gray"])
    sparse, it = rp.rpca(grays[::-1], want_iters=True)
    assert it == iters
    assert np.array_equal(np.stack(sparse[::-1]), z["rpca"])
    got = rp.run_path_rpca(frames, rp.PathParams(roi, 1, 15, 3, True, False, "u8"))
    o = 0
    for t, rec in enumerate(got):
        assert np.array_equal(rec["bilateral"], z["bilateral"][t])
        assert np.array_equal(rec["labels"], z["labels_u8"][t])
        assert np.array_equal(np.packbits(rec["mask"] > 0, axis=1, bitorder="little"), z["masks_packed"][t])
        k = int(z["counts_u8"][t])
        assert np.array_equal(rp.props_table(rec["props"]), z["props_u8"][o:o + k])
        o += k


def test_scalar_bilateral_definition_matches_cv2(golden_dir):
    z = np.load(os.path.join(golden_dir, "rpca_roi_batch21.npz"))
    rng = np.random.default_rng(3)
    imgs = [z["rpca"][0], z["rpca"][7], rng.integers(0, 256, (90, 130), dtype=np.uint8),
            (rng.integers(0, 20, (64, 64)) + 60).astype(np.uint8)]
    for im in imgs:
        a, b = rp.bilateral_scalar(im), rp.bilateral_blur(im, 7, 15, 1)
        d = np.abs(a.astype(int) - b.astype(int))
        assert d.max() <= 1 and (d > 0).sum() <= 2        # cv2's SIMD body vs its scalar definition: .5 ties only


# ---------------------------------------------------------------------------------- GPU: the CUDA path
@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_stage_rpca_and_bilateral_against_golden(golden_dir, name):
    from swiftwatcher_b200 import image_filtering as img
    z, frames, roi, iters = load(golden_dir, name)
    grays = list(z["gray"])
    sparse = img.rpca(grays[::-1])[::-1]                  # newest first, as the reference's queue
    assert close_u8(np.stack(sparse), z["rpca"])
    assert np.array_equal(np.stack(sparse), z["rpca"])    # and in fact identical
    for t in (0, len(grays) // 2, len(grays) - 1):
        b = img.bilateral_blur(z["rpca"][t], 7, 15, 1)
        assert np.array_equal(b, rp.bilateral_scalar(z["rpca"][t]))
        assert close_u8(b, z["bilateral"][t], frac=2e-6) or np.array_equal(b, z["bilateral"][t])


@pytest.mark.gpu
def test_zero_padded_last_batch_is_pinned(golden_dir):
    """io_video.py:40-44 pads the last batch of a video with all-zero frames; the reference's IALM then adds
    -(1/mu) u_k v_k^T for the exactly-zero singular values with LAPACK's arbitrary null-space vectors
    (image_filtering.py:284-290), which nothing else reproduces.  The CUDA path leaves those components out
    (DESIGN.md §8).  What it produces instead is pinned here: the oracle's IALM with the same rule
    (``skip_null=True``) — identical images for the real frames, all-zero images for the padding — and the
    distance to the reference-as-written result on the real frames is bounded (the two differ: 3,880 pixels of the 16 real frames, by up to 13)."""
    from swiftwatcher_b200 import image_filtering as img
    z = np.load(os.path.join(golden_dir, "rpca_roi_batch21.npz"))
    grays = list(z["gray"])
    for n_real in (16, 1):
        padded = grays[:n_real] + [np.zeros_like(grays[0])] * (21 - n_real)
        got = np.stack(img.rpca(padded[::-1])[::-1])
        want, iters = rp.rpca(padded[::-1], want_iters=True, skip_null=True)
        want = np.stack(want[::-1])
        assert close_u8(got, want)
        assert got[n_real:].max(initial=0) == 0
        as_written = np.stack(rp.rpca(padded[::-1])[::-1])
        d = np.abs(got[:n_real].astype(int) - as_written[:n_real].astype(int))
        assert d.max() <= 16        # measured on the oracle: 13 grey levels at most (n_real = 16), 0 for n_real = 1


@pytest.mark.gpu
def test_stage_bilateral_random_images():
    from swiftwatcher_b200 import image_filtering as img
    rng = np.random.default_rng(9)
    # widths that are multiples of four take the four-pixels-per-thread kernel (k_bilateral_r3x4), the others the
    # one-pixel kernels; heights below seven rows have no interior at all
    for shape in [(97, 131), (5, 9), (1, 40), (33, 1), (40, 64), (7, 16), (33, 132), (8, 20), (64, 256)]:
        a = rng.integers(0, 256, shape, dtype=np.uint8)
        assert np.array_equal(img.bilateral_blur(a, 7, 15, 1), rp.bilateral_scalar(a))
    a = (rng.integers(0, 24, (200, 300)) + 90).astype(np.uint8)
    assert np.array_equal(img.bilateral_blur(a, 5, 10.0, 2.0), rp.bilateral_scalar(a, 5, 10.0, 2.0))


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("mode", ["u8", "i32"])
def test_fused_rpca_path_against_golden(golden_dir, name, mode):
    import torch
    import swiftwatcher_b200 as swb
    from swiftwatcher_b200.pipeline import centroids
    z, frames, roi, iters = load(golden_dir, name)
    T = len(frames)
    for src in ("host", "device"):
        with swb.FilterContext(frames.shape[1:], roi, label_mode=mode, max_frames=T, bg_model="rpca") as ctx:
            if src == "host":
                ctx.set_option("rpca_device_loop", 0)          # the host loop (21-frame batches default to the device loop)
            ctx.submit(frames if src == "host" else torch.from_numpy(frames).cuda(), n_halo=0)
            rows, counts = ctx.collect()
            masks, labels, sparse = ctx.masks(), ctx.labels(), ctx.rpca_images()
            stats = ctx.rpca_stats()
            # the reference's own batch size runs its whole iteration loop on the device (CUDA-graph WHILE node)
            assert stats["iterations"] == iters and stats["device_loop"] == (T == 21 and src == "device")
            if stats["device_loop"]:                               # the graph is reused by the next submits
                assert 2 * iters <= stats["jacobi_sweeps"] <= 12 * iters
                for _ in range(2):
                    ctx.submit(frames, n_halo=0)
                    assert np.array_equal(ctx.rpca_images(), z["rpca"])
                    assert ctx.rpca_stats()["iterations"] == iters
                ctx.submit(np.zeros_like(frames), n_halo=0)       # all black: no iteration at all
                assert ctx.rpca_images().max() == 0 and ctx.rpca_stats()["iterations"] == 0 and len(ctx.collect()[0]) == 0
                ctx.submit(frames, n_halo=0)
                assert np.array_equal(ctx.rpca_images(), z["rpca"])
        assert np.array_equal(sparse, z["rpca"])
        # expectations come from the golden vectors only (no LAPACK on this machine): the golden
        # mask, labelled and measured by the oracle's integer stages
        w = roi[1][0] - roi[0][0]
        o = 0
        for t in range(T):
            assert np.array_equal(np.packbits(masks[t] > 0, axis=1, bitorder="little"), z["masks_packed"][t]), t
            gmask = (np.unpackbits(z["masks_packed"][t], axis=1, bitorder="little")[:, :w] * 255).astype(np.uint8)
            want_labels = rp.cc_labeling(gmask, 4) if mode == "u8" else rp.cc_labeling_i32(gmask)
            assert np.array_equal(labels[t], want_labels), t
            if mode == "u8":
                assert np.array_equal(labels[t], z["labels_u8"][t]), t
            k = counts[t]
            exp = rp.props_table(rp.get_segment_properties(want_labels))
            assert k == len(exp)
            r = rows[o:o + k]
            assert np.array_equal(r["label"], exp[:, 0]) and np.array_equal(r["area"], exp[:, 1])
            assert np.array_equal(r["bbox"], exp[:, 2:6])
            np.testing.assert_allclose(centroids(r), exp[:, 6:8], rtol=1e-5, atol=0)
            o += k


@pytest.mark.gpu
def test_rpca_queue_drop_in_and_limits(golden_dir):
    import swiftwatcher_b200 as swb
    import swiftwatcher_b200.data_structures as ds
    z, frames, region, _ = load(golden_dir, "rpca_roi_batch21")
    queue = ds.FrameQueue(queue_size=21, bg_model="rpca")
    queue.push_list_of_frames(list(frames), list(range(21)), ["00:00:00.000"] * 21)
    queue.preprocess_queue(region, None)
    queue.segment_queue((24, 24), region)
    while not queue.is_empty():
        f = queue.pop_frame()
        assert np.array_equal(f.get_processed_frame("RPCA"), z["rpca"][f.frame_number])
        assert np.array_equal(f.get_processed_frame("cc_labeling"), z["labels_u8"][f.frame_number])
        assert len(f.segments) == int(z["counts_u8"][f.frame_number])
    with pytest.raises(swb.SwbError):
        swb.FilterContext(frames.shape[1:], region, max_frames=64, bg_model="rpca")
    # an all-black batch decomposes to nothing
    with swb.FilterContext((40, 64, 3), None, max_frames=5, bg_model="rpca", label_mode="i32") as ctx:
        ctx.submit(np.zeros((5, 40, 64, 3), np.uint8), n_halo=0)
        rows, counts = ctx.collect()
        assert len(rows) == 0 and ctx.masks().max() == 0
