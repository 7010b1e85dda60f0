"""Oracle for the segment classifier — TEST INFRASTRUCTURE ONLY.

Restates ``swiftwatcher/segment_classification.py`` one segment at a time, exactly as the
reference does it: the five torchvision transforms through PIL (:16-23), one forward pass
on a [1, 3, 224, 224] tensor per segment (:30-36), ``torch.max(score, 1)`` (:37), keep
class 1 (:39-40), relabel the kept segments 1..k (:42-43).

Two things cannot be taken over literally (both recorded in DESIGN.md):
* ``models.squeezenet1_0(pretrained=True)`` (:51) needs a download; every parameter is
  overwritten by ``load_state_dict`` (:15), so ``weights=None`` builds the same module;
* the reference never calls ``model.eval()``, so Dropout(0.5) is live and its output is
  random; ``eval_mode=True`` (default) is the deterministic function the product is
  compared against, ``eval_mode=False`` reproduces the reference's stochastic behaviour.

Pinning: ``reference_module()`` imports the reference's own, unmodified
``segment_classification.py`` from /root/reference (with the constructor's download
disabled) so that the tests can check this restatement and the product against the real
class and the real ``model.pt`` wherever /root/reference exists (the build container).
"""

import importlib
import sys

import torch
from torch import nn
from torchvision import models, transforms


def setup_model(num_classes, device):
    """segment_classification.py:48-67."""
    model = models.squeezenet1_0(weights=None)
    for param in model.parameters():
        param.requires_grad = False
    model.classifier[1] = nn.Conv2d(512, num_classes, kernel_size=1)
    model.num_classes = num_classes
    return model.to(device)


class RefSegmentClassifier:
    """segment_classification.py:13-45, per-segment loop."""

    def __init__(self, state_dict, device, eval_mode=True):
        self.device = torch.device(device)
        self.model = setup_model(2, self.device)
        self.model.load_state_dict(state_dict)
        if eval_mode:
            self.model.eval()
        self.transforms = [
            transforms.ToPILImage(),
            transforms.Resize((24, 24)),
            transforms.Pad((224 - 24) // 2),
            transforms.ToTensor(),
            transforms.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225]),
        ]

    def transform(self, image):
        x = image
        for t in self.transforms:
            x = t(x)
        return x

    @torch.no_grad()
    def score(self, segment_image):
        return self.model(self.transform(segment_image).unsqueeze(0).to(self.device))

    @torch.no_grad()
    def __call__(self, segments):
        keep = []
        for segment in segments:
            score = self.score(segment.segment_image)
            _, y_pred = torch.max(score, 1)
            if y_pred == 1:
                keep.append(segment)
        for i, segment in enumerate(keep):
            segment.label = i + 1
        return keep


_STATE_CACHE = {}


def random_state_dict(seed):
    if seed in _STATE_CACHE:
        return {k: v.clone() for k, v in _STATE_CACHE[seed].items()}
    sd = _random_state_dict(seed)
    _STATE_CACHE[seed] = {k: v.clone() for k, v in sd.items()}
    return sd


def _random_state_dict(seed):
    """Seeded random-init weights of the reference architecture (there is no network for
    checkpoints on the GPU box; model.pt itself is not redistributed).  The bias of the
    class-0 output is centred on a fixed set of noise crops so that the two classes both
    occur (with plain random weights the constant padding decides every crop the same way)."""
    g = torch.Generator().manual_seed(seed)
    model = setup_model(2, "cpu")
    sd = model.state_dict()
    for k, v in sd.items():
        if v.dtype.is_floating_point:
            fan = v[0].numel() if v.dim() > 1 else 16
            sd[k] = (torch.randn(v.shape, generator=g) * (2.0 / max(fan, 1)) ** 0.5).to(v.dtype)
    model.load_state_dict(sd)
    model.eval()
    crops = torch.randint(0, 256, (48, 3, 24, 24), generator=g).float().div(255)
    x = torch.zeros((48, 3, 224, 224))
    x[:, :, 100:124, 100:124] = crops
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    with torch.no_grad():
        feat = model.features((x - mean) / std)
        # both outputs on the linear side of the final ReLU, then bisection on the class-0
        # bias until the median margin over the noise crops is zero
        conv = model.classifier[1]
        conv.bias[1] = 16.0 - float((conv(feat) - conv.bias.view(1, 2, 1, 1))[:, 1].min())
        lo, hi = float(conv.bias[1]) - 256.0, float(conv.bias[1]) + 256.0
        for _ in range(40):
            mid = 0.5 * (lo + hi)
            model.classifier[1].bias[0] = mid
            s = torch.flatten(model.classifier(feat), 1)
            if float((s[:, 1] - s[:, 0]).median()) > 0:
                lo = mid
            else:
                hi = mid
    sd["classifier.1.bias"] = model.classifier[1].bias.detach().clone()
    return sd


def reference_module(root="/root/reference"):
    """The reference's own segment_classification module, imported unmodified
    (torchvision's constructor is told not to download)."""
    real = models.squeezenet1_0

    def no_download(pretrained=False, **kw):
        kw.pop("weights", None)
        return real(weights=None, **kw)

    models.squeezenet1_0 = no_download
    sys.path.insert(0, root)
    try:
        sys.modules.pop("swiftwatcher.segment_classification", None)
        mod = importlib.import_module("swiftwatcher.segment_classification")
    finally:
        sys.path.remove(root)
    mod._restore = lambda: setattr(models, "squeezenet1_0", real)
    return mod


def build_reference_classifier(mod, model_path):
    """``mod.SegmentClassifier(model_path)`` on this machine: model.pt holds CUDA tensors
    and the reference's bare ``torch.load`` (:15) cannot map them on a CPU-only host, so
    ``torch.load`` is given ``map_location=mod.device`` for the duration of the call."""
    real_load = torch.load

    def load(f, *a, **kw):
        kw.setdefault("map_location", mod.device)
        return real_load(f, *a, **kw)

    torch.load = load
    try:
        return mod.SegmentClassifier(model_path)
    finally:
        torch.load = real_load


# ----------------------------------------------------------------------------------------------------
# transforms.Resize((24, 24)) on a PIL image (segment_classification.py:20): Pillow's Image.resize with the
# BILINEAR filter.  Pillow is a third-party dependency that is not under /root/reference (reference pins
# Pillow==6.1.0, requirements.txt; installed here: 12.2.0); its published algorithm (src/libImaging/
# Resample.c: precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal/Vertical_8bpc) restated in
# numpy.  tests/test_classifier.py pins it bit for bit against the installed Pillow.
# ----------------------------------------------------------------------------------------------------
PIL_PRECISION_BITS = 32 - 8 - 2


def pil_bilinear_coeffs(in_size, out_size):
    """precompute_coeffs + normalize_coeffs_8bpc for the whole-image box: per output index the first input
    index, the tap count and the 22-bit fixed-point weights."""
    import numpy as np
    scale = float(np.float32(in_size) - np.float32(0)) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale                       # bilinear: support 1
    ss = 1.0 / filterscale
    out = []
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)    # C cast: truncation
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = []
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            w.append(1.0 - a if a < 1.0 else 0.0)
        ww = 0.0
        for v in w:
            ww += v
        k = [(v / ww if ww != 0.0 else v) for v in w]
        out.append((xmin, [int(0.5 + v * (1 << PIL_PRECISION_BITS)) for v in k]))
    return out


def pil_bilinear_resize(image, size):
    """``np.asarray(Image.fromarray(image).resize((size[1], size[0]), BILINEAR))`` for a uint8 image
    [h, w] or [h, w, c]: horizontal pass, rounding to uint8, then the vertical pass; a pass whose size
    already matches is skipped (ImagingResample's need_horizontal / need_vertical)."""
    import numpy as np

    def one_pass(a, out_size):                        # resample axis 0
        if a.shape[0] == out_size:
            return a
        res = np.empty((out_size,) + a.shape[1:], np.uint8)
        for i, (lo, k) in enumerate(pil_bilinear_coeffs(a.shape[0], out_size)):
            acc = np.full(a.shape[1:], 1 << (PIL_PRECISION_BITS - 1), np.int64)
            for j, kv in enumerate(k):
                acc += a[lo + j].astype(np.int64) * kv
            res[i] = np.clip(acc >> PIL_PRECISION_BITS, 0, 255).astype(np.uint8)
        return res
    a = np.asarray(image)
    if a.shape[0] > a.shape[1] * 100 and size[0] < a.shape[0]:
        # Image.resize (Pillow >= 9.1): an image more than 100 times taller than wide that shrinks vertically is
        # resampled vertically FIRST (two im.resize calls); every other shape goes horizontal first
        a = one_pass(a, size[0])
        return np.ascontiguousarray(np.swapaxes(one_pass(np.swapaxes(a, 0, 1), size[1]), 0, 1))
    a = np.swapaxes(one_pass(np.swapaxes(a, 0, 1), size[1]), 0, 1)       # horizontal first
    return np.ascontiguousarray(one_pass(a, size[0]))
