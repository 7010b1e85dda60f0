"""CPU suite: the oracle against the committed golden vectors (which
oracle/make_golden.py produced by running the reference's own functions)."""
import os

import numpy as np
import pytest

from oracle import reference_path as rp
from oracle import synth


@pytest.fixture(scope="module")
def kats(golden_dir):
    return np.load(os.path.join(golden_dir, "stage_kats.npz"))


def test_gray_matches_reference(kats):
    assert np.array_equal(rp.convert_grayscale(kats["gray_in"]), kats["gray_out"])
    assert np.array_equal(rp.gray_fixed_point(kats["gray_in"]), kats["gray_out"])


def test_gray_identity_on_2d(kats):
    g = kats["gray_out"]
    assert rp.convert_grayscale(g) is g


def test_crop(kats):
    region = [tuple(kats["crop_region"][0]), tuple(kats["crop_region"][1])]
    assert np.array_equal(rp.crop_frame(kats["gray_in"], region), kats["crop_out"])


def test_threshold_kat(kats):
    assert np.array_equal(rp.thresh_to_zero(kats["thresh_in"], 15), kats["thresh_out"])
    probe = np.array([[14, 15, 16, 255]], np.uint8)
    assert rp.thresh_to_zero(probe, 15).tolist() == [[0, 0, 16, 255]]


def test_morphology(kats):
    m = kats["open_in"]
    assert np.array_equal(rp.grayscale_opening(m, (3, 3)), kats["open3_out"])
    assert np.array_equal(rp.grayscale_opening(m, (5, 5)), kats["open5_out"])
    assert np.array_equal(rp.grayscale_closing(m, (3, 3)), kats["close3_out"])
    assert np.array_equal(rp.grayscale_closing(m, (5, 5)), kats["close5_out"])


def test_binary_morphology_equivalence(kats):
    """The fused kernels run binary morphology on [v > 0]; it must equal the
    grey result's support (min/max commute with v -> [v > 0])."""
    from scipy import ndimage
    m = kats["open_in"]
    for k in (3, 5):
        grey = rp.grayscale_opening(m, (k, k)) > 0
        se = np.ones((k, k), bool)
        b = m > 0
        er = ndimage.binary_erosion(b, se, border_value=1)
        op = ndimage.binary_dilation(er, se, border_value=0)
        assert np.array_equal(grey, op)
        grey_c = rp.grayscale_closing(m, (k, k)) > 0
        di = ndimage.binary_dilation(b, se, border_value=0)
        cl = ndimage.binary_erosion(di, se, border_value=1)
        assert np.array_equal(grey_c, cl)


def test_labelling_is_8_connected(kats):
    assert np.array_equal(rp.cc_labeling(kats["cc_diag_in"], 4), kats["cc_diag_out"])
    assert kats["cc_diag_out"].max() == 1


def test_labelling_golden_and_order_rule(kats):
    assert np.array_equal(rp.cc_labeling(kats["cc_blobs_in"], 4), kats["cc_blobs_out"])
    assert np.array_equal(rp.cc_labeling(kats["cc_many_in"], 4), kats["cc_many_out"])
    assert np.array_equal(rp.cc_labeling_i32(kats["cc_many_in"]), kats["cc_many_i32"])
    # OpenCV numbering == rank of the minimum 2x2-block raster index
    assert np.array_equal(rp.label_order_spec(kats["cc_many_in"]), kats["cc_many_i32"])
    assert np.array_equal(rp.label_order_spec(kats["cc_blobs_in"]).astype(np.uint8), kats["cc_blobs_out"])


def test_label_order_rule_random():
    rng = np.random.default_rng(7)
    for _ in range(30):
        h, w = rng.integers(1, 40), rng.integers(1, 70)
        img = (rng.random((h, w)) < rng.uniform(0.05, 0.6)).astype(np.uint8) * 255
        assert np.array_equal(rp.label_order_spec(img), rp.cc_labeling_i32(img))


def test_uint8_wrap_merges_regions(kats):
    l32, l8 = kats["cc_many_i32"], kats["cc_many_out"]
    assert l32.max() > 255
    assert np.array_equal(l32.astype(np.uint8), l8)
    props = rp.get_segment_properties(l8)
    assert np.array_equal(rp.props_table(props), kats["props_many"])
    area1 = int(((l32 > 0) & (l32 % 256 == 1)).sum())      # components 1, 257, 513, ...
    assert props[0].label == 1 and props[0].area == area1


def test_regionprops_golden(kats):
    props = rp.get_segment_properties(kats["cc_blobs_out"])
    assert np.array_equal(rp.props_table(props), kats["props_blobs"])


def test_extract_segment_images(kats):
    props = rp.get_segment_properties(kats["cc_blobs_out"])
    region = [tuple(kats["crops_region"][0]), tuple(kats["crops_region"][1])]
    crops = rp.extract_segment_images(props, kats["crops_frame"], (24, 24), region)
    assert np.array_equal(np.array([c.shape for c in crops]), kats["crops_shapes"])
    assert np.array_equal(np.array([int(c.sum()) for c in crops]), kats["crops_sums"])


def test_window_policy():
    assert rp.window_indices(0, 5) == [0, 0, 0, 0, 0]
    assert rp.window_indices(2, 5) == [0, 0, 0, 1, 2]
    assert rp.window_indices(7, 5) == [3, 4, 5, 6, 7]


def test_median_is_order_statistic():
    rng = np.random.default_rng(3)
    stack = rng.integers(0, 256, (9, 13, 17), dtype=np.uint8)
    for n in (1, 3, 5, 7, 9):
        want = np.sort(stack[:n], axis=0)[n // 2]
        assert np.array_equal(rp.temporal_median(list(stack[:n])), want)


PATH_CASES = ["path_roi_n5_open3", "path_full_n9_oc5", "path_dense_n5_open3"]


def load_case(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    c = [int(v) for v in z["cfg"]]
    cfg = dict(seed=c[0], video=c[1], H=c[2], W=c[3], birds=c[4], T=c[5],
               roi=[(c[6], c[7]), (c[8], c[9])], N=c[10], thresh=c[11], se=c[12],
               do_open=c[13], do_close=c[14])
    return z, cfg


@pytest.mark.parametrize("name", PATH_CASES)
def test_path_against_golden(golden_dir, name):
    z, c = load_case(golden_dir, name)
    frames = synth.synth_video(c["seed"], c["video"], 0, c["T"], c["H"], c["W"], c["birds"])
    assert np.array_equal(frames.reshape(c["T"], -1).sum(axis=1), z["frame_sums"])
    for mode, lab_key, cnt_key, tab_key in (("u8", "labels_u8", "counts_u8", "props_u8"),
                                            ("i32", "labels_i32", "counts_i32", "props_i32")):
        par = rp.PathParams(c["roi"], c["N"], c["thresh"], c["se"], bool(c["do_open"]),
                            bool(c["do_close"]), mode)
        got = rp.run_path(frames, par)
        w = c["roi"][1][0] - c["roi"][0][0]
        tabs = []
        for t, rec in enumerate(got):
            bits = np.unpackbits(z["masks_packed"][t], axis=1, bitorder="little")[:, :w]
            assert np.array_equal(rec["mask"] > 0, bits.astype(bool)), (name, mode, t)
            assert np.array_equal(rec["labels"], z[lab_key][t]), (name, mode, t)
            assert len(rec["props"]) == z[cnt_key][t]
            tabs.append(rp.props_table(rec["props"]))
        assert np.array_equal(np.concatenate(tabs), z[tab_key])


def test_path_halo_equals_whole(golden_dir):
    """Temporal chunks with an N-1 halo reproduce the unchunked result."""
    z, c = load_case(golden_dir, "path_roi_n5_open3")
    frames = synth.synth_video(c["seed"], c["video"], 0, c["T"], c["H"], c["W"], c["birds"])
    par = rp.PathParams(c["roi"], c["N"], c["thresh"], c["se"], True, False, "i32")
    whole = rp.run_path(frames, par)
    cut = 6
    a = rp.run_path(frames[:cut], par)
    b = rp.run_path(frames[cut:], par, history=list(frames[cut - (c["N"] - 1):cut]))
    for t, rec in enumerate(a + b):
        assert np.array_equal(rec["labels"], whole[t]["labels"])


def test_synth_is_deterministic_and_hash_based():
    a = synth.synth_frame(9, 1, 5, 40, 64, 6)
    b = synth.synth_video(9, 1, 4, 3, 40, 64, 6)[1]
    assert np.array_equal(a, b)
    assert synth.mix32(0) == 0
