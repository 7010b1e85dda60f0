"""Segment classifier fed batched crops against the oracle's
per-segment restatement of swiftwatcher/segment_classification.py, and — where
/root/reference exists — against the reference's own class with the real model.pt."""
import os

import numpy as np
import pytest
import torch

from oracle import reference_classifier as rc
from oracle import reference_path as rp
from oracle import synth

REF = "/root/reference"
TOL = 1e-4   # float32 scores: batched vs one-by-one forward passes (different conv algorithms)


class Seg:
    def __init__(self, image, label):
        self.segment_image = image
        self.label = label


def make_segments(rng, n, odd=()):
    segs = []
    for i in range(n):
        shape = odd[i] if i < len(odd) else (24, 24)
        segs.append(Seg(rng.integers(0, 256, size=shape + (3,), dtype=np.uint8), 100 + i))
    return segs


def test_preprocess_equals_the_transform_chain_bit_for_bit():
    from swiftwatcher_b200.segment_classification import SegmentClassifier
    sd = rc.random_state_dict(1)
    clf = SegmentClassifier(sd, device="cpu")
    ref = rc.RefSegmentClassifier(sd, "cpu")
    rng = np.random.default_rng(0)
    crops = rng.integers(0, 256, size=(9, 24, 24, 3), dtype=np.uint8)
    crops[0] = 0
    crops[1] = 255
    got = clf.preprocess(crops)
    for i in range(len(crops)):
        assert torch.equal(got[i], ref.transform(crops[i])), i


def test_batched_call_equals_per_segment_loop_cpu():
    from swiftwatcher_b200.segment_classification import SegmentClassifier
    sd = rc.random_state_dict(2)
    rng = np.random.default_rng(1)
    odd = [(30, 24), (24, 41), (11, 24)]          # go through the PIL resize, as in the reference
    a = make_segments(rng, 40, odd)
    b = [Seg(s.segment_image.copy(), s.label) for s in a]
    clf = SegmentClassifier(sd, device="cpu", batch_size=16)
    ref = rc.RefSegmentClassifier(sd, "cpu")
    kept = clf(a)
    want = ref(b)
    assert [id(s) for s in kept] == [id(a[b.index(w)]) for w in want]
    assert [s.label for s in kept] == list(range(1, len(kept) + 1))
    assert 0 < len(kept) < len(a)                 # the seeded weights split the crops
    exact = np.stack([s.segment_image for s in a[3:]])
    got = clf.scores(exact)
    exp = torch.cat([ref.score(s.segment_image) for s in a[3:]])
    assert torch.allclose(got, exp, rtol=TOL, atol=TOL)
    assert clf(a[:0]) == []


def test_windowed_evaluation_equals_the_full_forward_pass():
    """WindowedSqueezeNet computes each layer only where the 24x24 crop can matter and takes
    the rest from the blank canvas: same scores as model(x) on the padded 224x224 canvas."""
    from swiftwatcher_b200.segment_classification import SegmentClassifier
    rng = np.random.default_rng(5)
    crops = rng.integers(0, 256, size=(12, 24, 24, 3), dtype=np.uint8)
    crops[0] = 0
    crops[1] = 255
    states = [rc.random_state_dict(6)]
    if os.path.exists(os.path.join(REF, "swiftwatcher", "model.pt")):
        states.append(torch.load(os.path.join(REF, "swiftwatcher", "model.pt"), map_location="cpu"))
    for sd in states:
        full = SegmentClassifier(sd, device="cpu", windowed=False)
        win = SegmentClassifier(sd, device="cpu", windowed=True)
        a, b = full.scores(crops), win.scores(crops)
        assert torch.allclose(a, b, rtol=TOL, atol=TOL), (a - b).abs().max()
        assert a.abs().max() > 0.1                     # not a dead network


def test_empty_crop_is_an_error():
    from swiftwatcher_b200.segment_classification import SegmentClassifier
    clf = SegmentClassifier(rc.random_state_dict(3), device="cpu")
    with pytest.raises(ValueError):
        clf([Seg(np.zeros((0, 24, 3), dtype=np.uint8), 1)])


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "swiftwatcher", "model.pt")),
                    reason="the reference (and its model.pt) only exists in the build container")
def test_against_the_reference_class_and_model_pt():
    """The reference's own SegmentClassifier (unmodified module, real weights, eval mode
    for determinism) keeps exactly the segments the batched classifier keeps."""
    from swiftwatcher_b200.segment_classification import SegmentClassifier
    mod = rc.reference_module(REF)
    try:
        path = os.path.join(REF, "swiftwatcher", "model.pt")
        their = rc.build_reference_classifier(mod, path)
        their.model.eval()
        dev = str(mod.device)
        ours = SegmentClassifier(path, device=dev, batch_size=32)
        oracle = rc.RefSegmentClassifier(torch.load(path, map_location=dev), dev)
        # bird-like crops from the synthetic generator + noise crops
        frames = synth.synth_video(5, 0, 0, 6, 120, 200, 40)
        par = rp.PathParams([(0, 0), (200, 120)], 5, 15, 3, True, False, "u8")
        out = rp.run_path(frames, par, want_images=True)
        images = [im for o in out for im in o.get("crops", []) if im.shape == (24, 24, 3)][:14]
        rng = np.random.default_rng(2)
        images += [rng.integers(0, 256, size=(24, 24, 3), dtype=np.uint8) for _ in range(6)]
        a = [Seg(im, 7) for im in images]
        b = [Seg(im, 7) for im in images]
        c = [Seg(im, 7) for im in images]
        kept_theirs = their(a)
        kept_ours = ours(b)
        kept_oracle = oracle(c)
        assert [a.index(s) for s in kept_theirs] == [b.index(s) for s in kept_ours] == [c.index(s) for s in kept_oracle]
        assert [s.label for s in kept_ours] == list(range(1, len(kept_ours) + 1))
        got = ours.scores(np.stack(images))
        exp = torch.cat([oracle.score(im) for im in images])
        assert torch.allclose(got, exp, rtol=TOL, atol=TOL)
    finally:
        mod._restore()


@pytest.mark.gpu
def test_gpu_batched_scores_and_device_crops():
    import swiftwatcher_b200 as swb
    from swiftwatcher_b200 import image_filtering as img
    from swiftwatcher_b200.segment_classification import SegmentClassifier
    # full float32 convolutions on both sides (torch's default lets cuDNN use TF32, whose
    # rounding noise is larger than the margins of the random-weight model)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sd = rc.random_state_dict(4)
    clf = SegmentClassifier(sd, device="cuda:0", batch_size=64)
    ref = rc.RefSegmentClassifier({k: v.cuda() for k, v in sd.items()}, "cuda:0")
    frames = synth.synth_video(6, 0, 0, 8, 160, 256, 60)
    region = [(0, 0), (256, 160)]
    dev = torch.from_numpy(frames).cuda()
    with swb.FilterContext(frames.shape[1:], region, label_mode="i32", max_frames=8) as ctx:
        ctx.submit(dev, n_halo=0)
        rows, counts = ctx.collect()
        n = len(rows)
        assert n > 20
        crops = torch.empty((n, 24, 24, 3), dtype=torch.uint8, device="cuda")
        ctx.gather_crops(n, 24, out=crops)
        with pytest.raises(ValueError):                          # birds at the top / left edge: empty segment images,
            clf.classify_submit(ctx, n)                          # the reference's ToPILImage raises there too
        keep_dev = clf.classify_submit(ctx, n, empty="drop").cpu().numpy()
    # host crops with the reference's slicing; interior segments up to 24 px give the same 24x24 tile
    props = swb.props_from_rows(rows)
    same = []
    for i, (p, r) in enumerate(zip(props, rows)):
        im = img.extract_segment_images([p], frames[r["frame"]], (24, 24), region)[0]
        if im.shape == (24, 24, 3):
            assert np.array_equal(im, crops[i].cpu().numpy()), i
            same.append(i)
    assert len(same) > 10
    got = clf.scores(crops[same])
    exp = torch.cat([ref.score(crops[i].cpu().numpy()) for i in same])
    assert torch.allclose(got, exp, rtol=TOL, atol=TOL)          # batch 64 vs batch 1
    margin = (exp[:, 1] - exp[:, 0]).abs().cpu().numpy()
    pred_ref = (exp[:, 1] > exp[:, 0]).cpu().numpy()
    decided = margin > 10 * TOL
    assert decided.sum() > 5 and 0 < pred_ref[decided].sum() < decided.sum()
    assert np.array_equal(keep_dev[same][decided], pred_ref[decided])
    full = SegmentClassifier(sd, device="cuda:0", batch_size=64, windowed=False)
    assert clf.windowed is not None and full.windowed is None
    assert torch.allclose(full.scores(crops[same]), exp, rtol=TOL, atol=TOL)   # plain batched forward pass
    assert torch.allclose(full.scores(crops[same]), got, rtol=TOL, atol=TOL)   # == windowed evaluation
    segs = [Seg(crops[i].cpu().numpy(), 9) for i in same]
    kept = clf(segs)
    kept_idx = np.zeros(len(segs), dtype=bool)
    kept_idx[[segs.index(s) for s in kept]] = True
    assert np.array_equal(kept_idx[decided], keep_dev[same][decided])


@pytest.mark.gpu
def test_gpu_classify_submit_keeps_what_the_per_segment_reference_keeps():
    """Large birds (> 24 px: resized by transforms.Resize in the reference) and birds cut by the frame edge go
    through the device path with the same tiles, hence the same decisions, as the per-segment oracle."""
    import swiftwatcher_b200 as swb
    from swiftwatcher_b200.segment_classification import SegmentClassifier
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sd = rc.random_state_dict(5)
    clf = SegmentClassifier(sd, device="cuda:0", batch_size=64)
    ref = rc.RefSegmentClassifier({k: v.cuda() for k, v in sd.items()}, "cuda:0")
    rng = np.random.default_rng(7)
    H, W = 240, 400
    base = rng.integers(100, 256, (H, W, 3), dtype=np.uint8)
    frames = np.repeat(base[None], 7, axis=0)
    boxes = []
    for _ in range(60):
        h, w = int(rng.integers(4, 70)), int(rng.integers(4, 70))
        y, x = int(rng.integers(0, H - h + 1)), int(rng.integers(0, W - w + 1))
        if any(y < b[0] + b[2] + 2 and b[0] < y + h + 2 and x < b[1] + b[3] + 2 and b[1] < x + w + 2 for b in boxes):
            continue
        boxes.append((y, x, h, w))
        frames[5:, y:y + h, x:x + w] = rng.integers(0, 60, (h, w, 3), dtype=np.uint8)
    region = [(0, 0), (W, H)]
    dev = torch.from_numpy(frames).cuda()
    with swb.FilterContext(frames.shape[1:], region, label_mode="i32", max_frames=7) as ctx:
        ctx.submit(dev, n_halo=0)
        rows, counts = ctx.collect()
        keep_dev = clf.classify_submit(ctx, len(rows), empty="drop").cpu().numpy()
    props = swb.props_from_rows(rows)
    big = decided = 0
    for i, (p, r) in enumerate(zip(props, rows)):
        im = rp.extract_segment_images([p], frames[r["frame"]], (24, 24), region)[0]
        if im.size == 0:
            assert not keep_dev[i]
            continue
        score = ref.score(np.ascontiguousarray(im))[0]
        if abs(float(score[1] - score[0])) > 10 * TOL:
            assert bool(keep_dev[i]) == bool(score[1] > score[0]), (i, im.shape)
            decided += 1
            big += im.shape[:2] != (24, 24)
    assert decided > 20 and big > 10


def test_pil_bilinear_restatement_equals_pillow():
    """oracle.reference_classifier.pil_bilinear_resize (Pillow's Resample.c restated; the arithmetic the device
    crop kernel implements) against transforms.Resize((24, 24)) of the installed Pillow, bit for bit."""
    from torchvision import transforms
    rng = np.random.default_rng(3)
    shapes = [(24, 24), (24, 30), (15, 16), (40, 30), (1, 1), (1, 50), (50, 1), (24, 1), (100, 237), (333, 24),
              (25, 25), (23, 24), (47, 49), (7, 640), (500, 3), (301, 3), (300, 3), (1000, 9), (1000, 10), (101, 1)]
    for h, w in shapes:
        for ch in (3, 1):
            a = rng.integers(0, 256, (h, w, ch) if ch == 3 else (h, w), dtype=np.uint8)
            want = np.asarray(transforms.Resize((24, 24))(transforms.ToPILImage()(a)))
            assert np.array_equal(rc.pil_bilinear_resize(a, (24, 24)), want), (h, w, ch)


@pytest.mark.gpu
def test_gpu_glue_kernels_and_buffered_evaluation():
    """swb_nhwc_paste / swb_nhwc_maxpool against the torch operations they replace (channel counts that take the
    float4 and the scalar path, with and without ReLU, partial batches), and the buffered windowed evaluation against
    the plain one and against the full forward pass, in both memory formats."""
    import torch.nn.functional as F
    from swiftwatcher_b200 import _lib
    from swiftwatcher_b200._lib import check
    from swiftwatcher_b200.segment_classification import SegmentClassifier, setup_model
    lib = _lib.load()
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(5)
    stream = torch.cuda.current_stream(dev).cuda_stream
    for C, h, H, off, relu in [(3, 24, 36, 6, 0), (16, 8, 10, 1, 1), (96, 15, 17, 1, 1), (64, 5, 9, 3, 0)]:
        for B in (1, 7):
            x = torch.randn((B, C, h, h), generator=g).to(dev).contiguous(memory_format=torch.channels_last)
            buf = torch.randn((B + 2, C, H, H), generator=g).to(dev).contiguous(memory_format=torch.channels_last)
            want = buf.clone()
            want[:B, :, off:off + h, off:off + h] = torch.relu(x) if relu else x
            check(lib.swb_nhwc_paste(x.data_ptr(), buf.data_ptr(), B, C, h, h, H, H, off, off, None, relu, stream))
            assert torch.equal(buf, want), (C, h, H, off, relu, B)
            bias = torch.randn((C,), generator=g).to(dev)
            want[:B, :, off:off + h, off:off + h] = torch.relu(x + bias.view(1, C, 1, 1)) if relu else x + bias.view(1, C, 1, 1)
            check(lib.swb_nhwc_paste(x.data_ptr(), buf.data_ptr(), B, C, h, h, H, H, off, off, bias.data_ptr(), relu, stream))
            assert torch.equal(buf, want), (C, h, H, off, relu, B, "bias")
            y = x.clone()                                          # in place: bias + ReLU of a convolution output
            check(lib.swb_nhwc_paste(y.data_ptr(), y.data_ptr(), B, C, h, h, h, h, 0, 0, bias.data_ptr(), 1, stream))
            assert torch.equal(y, torch.relu(x + bias.view(1, C, 1, 1)))
    for C, H, k, s in [(96, 17, 3, 2), (256, 9, 3, 2), (8, 7, 2, 1)]:
        x = torch.randn((5, C, H, H), generator=g).to(dev).contiguous(memory_format=torch.channels_last)
        x[0, :, 0, :] = float("-inf")
        n_out = (H - k) // s + 1
        out = torch.empty((5, C, n_out, n_out), device=dev).contiguous(memory_format=torch.channels_last)
        check(lib.swb_nhwc_maxpool(x.data_ptr(), out.data_ptr(), 5, C, H, H, k, s, stream))
        assert torch.equal(out, F.max_pool2d(x, k, s)), (C, H, k, s)
    assert lib.swb_nhwc_maxpool(x.data_ptr(), out.data_ptr(), 5, 6, 7, 7, 2, 1, stream) != 0   # channels not a multiple of 4
    torch.manual_seed(3)
    sd = setup_model(2, "cpu").state_dict()
    crops = torch.randint(0, 256, (150, 24, 24, 3), dtype=torch.uint8, generator=g)
    full = SegmentClassifier(sd, device="cuda:0", batch_size=64, windowed=False).scores(crops)
    for cl in (True, False):
        plain = SegmentClassifier(sd, device="cuda:0", batch_size=64, channels_last=cl, buffered=False)
        buffered = SegmentClassifier(sd, device="cuda:0", batch_size=64, channels_last=cl, buffered=True)
        a, b = plain.scores(crops), buffered.scores(crops)          # 150 = two full batches + a partial one
        assert torch.allclose(a, b, rtol=TOL, atol=TOL) and torch.allclose(b, full, rtol=TOL, atol=TOL), cl
        assert torch.allclose(buffered.scores(crops[:9]), full[:9], rtol=TOL, atol=TOL)
