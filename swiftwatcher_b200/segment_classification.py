"""SegmentClassifier — the reference's segment classifier fed batched crops.

Mirror of ``swiftwatcher/segment_classification.py``: same
class name, constructor argument and call semantics (``classifier(segments)``
returns the segments predicted as class 1, relabelled 1..k, :27-45).  The model
is the reference's own (torchvision SqueezeNet1.0 with a 2-class 1x1 classifier
conv, :48-67, weights from the user's ``model.pt``) and it stays a PyTorch
module — BASELINE.json keeps the classifier as it is.  What changes is how it
is fed:

* the reference runs one forward pass per segment on a [1, 3, 224, 224] tensor
  built by five torchvision transforms through PIL (:16-23, :30-36);
* here all crops of a frame (or of a whole submit) are preprocessed as one
  tensor with the same arithmetic (uint8 -> float32 / 255, zero pad 100 px,
  ``(x - mean) / std``: bit-identical to ToTensor + Pad + Normalize) and go
  through the model in large batches; crops can come straight from the device
  (``FilterContext.gather_crops``: 24x24 tiles cut from the BGR frames by a CUDA
  kernel) without touching the host.

Deliberate differences:
* ``model.eval()`` — the reference never calls it, so its Dropout(0.5) is live
  and its predictions are random from run to run;
* ``squeezenet1_0(weights=None)`` — the reference's ``pretrained=True`` needs a
  download and every tensor is overwritten by ``model.pt`` anyway (52/52 keys);
* crops that are not 24x24 (segments larger than 24 px, or cut by the frame
  border) go through the reference's PIL resize, one by one, as before.
"""

import numpy as np
import torch
from torch import nn
from torchvision import models, transforms

CROP = 24
PAD = (224 - CROP) // 2
_MEAN = (0.485, 0.456, 0.406)
_STD = (0.229, 0.224, 0.225)


def setup_model(num_classes, device):
    """segment_classification.py:48-67 (architecture only; no download)."""
    model = models.squeezenet1_0(weights=None)
    for param in model.parameters():
        param.requires_grad = False
    model.classifier[1] = nn.Conv2d(512, num_classes, kernel_size=1)
    model.num_classes = num_classes
    return model.to(device)


class WindowedSqueezeNet:
    """The same SqueezeNet1.0, evaluated only where a 24x24 crop can influence it.

    Every input of the classifier is a 224x224 canvas that is constant (the normalised zero
    padding) except for the 24x24 crop in its centre.  A convolution / pooling output whose
    receptive field does not touch the crop has the value it has for the blank canvas, whatever
    the crop is — so the blank canvas is pushed through the network once, every intermediate
    map is kept, and for a batch of crops each layer is computed only on the window of positions
    that can differ (15x15 of 109x109 after the first convolution, 8x8 of 54x54 after the first
    pooling, ... 11x11 of 13x13 at the end), reading the blank activations as the halo.  The
    final average pool adds the blank values of the positions outside the window.  This is the
    same function as ``model(x)`` up to float summation order (about 6x fewer multiply-adds and
    20-50x less activation traffic in the early, memory-heavy layers); tests compare it with
    the full forward pass.
    """

    def __init__(self, model, blank):
        """model: torchvision SqueezeNet (eval mode); blank: [1, 3, 224, 224] normalised blank canvas."""
        self.model = model
        self.H = int(blank.shape[-1])
        f = model.features
        self.plan = []
        x = blank
        for layer in f:
            name = type(layer).__name__
            if name == "Conv2d":
                self.plan.append(("conv", layer, x))
                x = layer(x)
            elif name == "ReLU":
                self.plan.append(("relu", layer, None))
                x = layer(x)
            elif name == "MaxPool2d":
                self.plan.append(("pool", layer, x))
                x = layer(x)
            elif name == "Fire":
                sq = layer.squeeze_activation(layer.squeeze(x))
                self.plan.append(("fire", layer, sq))
                x = torch.cat([layer.expand1x1_activation(layer.expand1x1(sq)),
                               layer.expand3x3_activation(layer.expand3x3(sq))], 1)
            else:
                raise TypeError("unexpected layer %s in SqueezeNet.features" % name)
        self.head = model.classifier[1]
        self.blank_head = torch.relu(self.head(x))              # [1, 2, h, w]
        self.blank_feat_hw = int(x.shape[-1])

    @staticmethod
    def _out_window(win, k, s, p, n_out):
        a, b = win                                              # input window [a, b)
        lo = max(0, -((-(a - (k - 1) + p)) // s))               # ceil((a - k + 1 + p) / s)
        hi = min(n_out - 1, (b - 1 + p) // s)
        return lo, hi + 1

    @staticmethod
    def _patch(blank, xw, win, need, fill):
        """[B, C, need, need] input patch: blank activations (``fill`` outside the map) with the window pasted in."""
        (a, b), (na, nb) = win, need
        n = blank.shape[-1]
        B, C = xw.shape[0], xw.shape[1]
        ia, ib = max(na, 0), min(nb, n)
        if na < 0 or nb > n:
            patch = xw.new_full((B, C, nb - na, nb - na), fill)
            patch[:, :, ia - na:ib - na, ia - na:ib - na] = blank[:, :, ia:ib, ia:ib]
        else:
            patch = blank[:, :, na:nb, na:nb].expand(B, C, nb - na, nb - na).clone(
                memory_format=torch.channels_last if xw.is_contiguous(memory_format=torch.channels_last)
                else torch.contiguous_format)
        patch[:, :, a - na:b - na, a - na:b - na] = xw
        return patch

    # -- the same evaluation with persistent patch buffers ---------------------------------------------
    # Profiling the plain version below (profiles/classifier_timing.py) showed that the convolutions are ~15 % of its
    # GPU time: the rest was glue — cloning the blank halo into a fresh patch for every layer and call, pasting the
    # window in, concatenating the two expand branches of every Fire module, NHWC max-pooling.  Here every layer owns
    # a patch buffer that is filled with the blank activations ONCE (the halo never changes); per call the producer
    # writes its (ReLU'd) output straight into the buffer's interior (``clamp_min(out=view)``), and the 1x1 expand
    # branch of a Fire module is the centre tap of a 3x3 kernel, so both branches are one convolution whose output is
    # already the concatenation.  Same function, same float32 / TF32 library kernels.
    def _build(self, offset, c, bmax, like):
        F = torch.nn.functional
        cl = like.is_contiguous(memory_format=torch.channels_last)
        fmt = torch.channels_last if cl else torch.contiguous_format
        glue = self._glue(like)

        def buffer(blank, win, need, fill):
            (a, b), (na, nb) = win, need
            n = blank.shape[-1]
            C = blank.shape[1]
            buf = torch.empty((bmax, C, nb - na, nb - na), dtype=like.dtype, device=like.device).contiguous(memory_format=fmt)
            buf.fill_(fill)
            ia, ib = max(na, 0), min(nb, n)
            buf[:, :, ia - na:ib - na, ia - na:ib - na] = blank[:, :, ia:ib, ia:ib]
            return buf, (a - na, b - na)

        steps = []
        win = (offset, offset + c)
        pending_relu = False                                   # the producer's ReLU is applied when its output is consumed
        for kind, layer, blank in self.plan:
            if kind == "relu":
                pending_relu = True
            elif kind == "conv":
                k, st, p = layer.kernel_size[0], layer.stride[0], layer.padding[0]
                n_out = (blank.shape[-1] + 2 * p - k) // st + 1
                ow = self._out_window(win, k, st, p, n_out)
                need = (ow[0] * st - p, (ow[1] - 1) * st - p + k)
                buf, inner = buffer(blank, win, need, 0.0)
                steps.append(("conv", buf, inner, pending_relu, layer.weight, layer.bias, st))
                pending_relu, win = False, ow
            elif kind == "pool":
                k, st = layer.kernel_size, layer.stride
                n_in = blank.shape[-1]
                n_out = -((-(n_in - k)) // st) + 1 if layer.ceil_mode else (n_in - k) // st + 1
                if layer.ceil_mode and (n_out - 1) * st >= n_in:
                    n_out -= 1
                ow = self._out_window(win, k, st, 0, n_out)
                need = (ow[0] * st, (ow[1] - 1) * st + k)
                if glue:                                       # swb_nhwc_maxpool on a channels-last patch
                    buf, inner = buffer(blank, win, need, float("-inf"))
                else:                                          # torch's NHWC max-pool kernel is 2-3x slower than its NCHW one here
                    (a, b), (na, nb) = win, need
                    n = blank.shape[-1]
                    buf = torch.full((bmax, blank.shape[1], nb - na, nb - na), float("-inf"), dtype=like.dtype, device=like.device)
                    ia, ib = max(na, 0), min(nb, n)
                    buf[:, :, ia - na:ib - na, ia - na:ib - na] = blank[:, :, ia:ib, ia:ib]
                    inner = (a - na, b - na)
                steps.append(("pool", buf, inner, pending_relu, k, st, fmt))
                pending_relu, win = False, ow
            else:   # fire
                n = blank.shape[-1]
                ow = (max(win[0] - 1, 0), min(win[1] + 1, n))
                need = (ow[0] - 1, ow[1] + 1)
                buf, inner = buffer(blank, win, need, 0.0)     # blank = the squeeze activations of the blank canvas
                w3, w1 = layer.expand3x3.weight, layer.expand1x1.weight
                w1p = torch.zeros((w1.shape[0], w1.shape[1], 3, 3), dtype=w1.dtype, device=w1.device)
                w1p[:, :, 1:2, 1:2] = w1
                w = torch.cat([w1p, w3], 0).contiguous(memory_format=fmt)
                bias = torch.cat([layer.expand1x1.bias, layer.expand3x3.bias], 0)
                steps.append(("fire", buf, inner, pending_relu, layer.squeeze.weight, layer.squeeze.bias, w, bias))
                pending_relu, win = True, ow                   # the module ends with ReLU on the concatenation
        a, b = win
        outside = self.blank_head.sum((2, 3)) - self.blank_head[:, :, a:b, a:b].sum((2, 3))
        return steps, outside, pending_relu

    @staticmethod
    def _glue(like):
        """The library's NHWC glue kernels (swb_nhwc_paste / swb_nhwc_maxpool) for channels-last CUDA float32 tensors."""
        if like.device.type != "cuda" or like.dtype != torch.float32 or not like.is_contiguous(memory_format=torch.channels_last):
            return None
        from . import _lib
        return _lib.load()

    @staticmethod
    def _paste(lib, x, buf, B, inner, relu, bias=None):
        """buf[:B, :, inner, inner] = relu?(x + bias?) with one streaming kernel (x: [B, C, h, w] channels-last;
        buf may be x itself with inner = (0, h): the in-place bias + ReLU of a convolution output)."""
        from ._lib import check
        Bx, C, h, w = x.shape
        assert Bx == B and C == buf.shape[1] and x.is_contiguous(memory_format=torch.channels_last)
        check(lib.swb_nhwc_paste(x.data_ptr(), buf.data_ptr(), B, C, h, w, buf.shape[2], buf.shape[3], inner[0], inner[0],
                                 None if bias is None else bias.data_ptr(), 1 if relu else 0,
                                 torch.cuda.current_stream(x.device).cuda_stream))

    @torch.no_grad()
    def forward_buffered(self, crops_norm, offset, bmax=2048):
        """The evaluation of ``__call__`` on persistent buffers (at most ``bmax`` crops per call)."""
        F = torch.nn.functional
        B, c = int(crops_norm.shape[0]), int(crops_norm.shape[-1])
        lib = self._glue(crops_norm)
        key = (offset, c, crops_norm.device, crops_norm.is_contiguous(memory_format=torch.channels_last))
        if getattr(self, "_built_key", None) != key or self._built_bmax < B:
            self._steps, self._outside, self._last_relu = self._build(offset, c, max(bmax, B), crops_norm)
            self._built_key, self._built_bmax = key, max(bmax, B)
        cl = torch.channels_last
        x = crops_norm
        xb = None          # glue path: the bias of the convolution that produced x, not added yet (the library adds a
                           # bias with a separate strided elementwise kernel: a third of the time before this)
        for step in self._steps:
            kind, buf, (ia, ib) = step[0], step[1], step[2]
            view = buf[:B, :, ia:ib, ia:ib]
            if kind == "fire":
                _, _, _, pre_relu, sw, sb, w, bias = step
                if lib is not None:
                    x = x.contiguous(memory_format=cl)
                    if pre_relu or xb is not None:                             # in place: + bias, ReLU
                        self._paste(lib, x, x, B, (0, int(x.shape[2])), pre_relu, xb)
                    sq = F.conv2d(x, sw, None).contiguous(memory_format=cl)
                    self._paste(lib, sq, buf, B, (ia, ib), True, sb)           # squeeze + bias -> ReLU -> into the halo'd patch
                    x, xb = F.conv2d(buf[:B], w, None), bias                   # [expand1x1 | expand3x3], bias and ReLU pending
                    continue
                if pre_relu:
                    x = torch.relu_(x)
                torch.clamp_min(F.conv2d(x, sw, sb), 0.0, out=view)
                x = F.conv2d(buf[:B], w, bias)
            elif kind == "conv":
                _, _, _, pre_relu, w, bias, st = step
                if lib is not None:
                    self._paste(lib, x.contiguous(memory_format=cl), buf, B, (ia, ib), pre_relu, xb)
                    x, xb = F.conv2d(buf[:B], w, None, stride=st), bias
                    continue
                if pre_relu:
                    torch.clamp_min(x, 0.0, out=view)
                else:
                    view.copy_(x)
                x = F.conv2d(buf[:B], w, bias, stride=st)
            else:   # pool
                _, _, _, pre_relu, k, st, fmt = step
                if lib is not None:
                    from ._lib import check
                    self._paste(lib, x.contiguous(memory_format=cl), buf, B, (ia, ib), pre_relu, xb)
                    n_in = int(buf.shape[2])
                    n_out = (n_in - k) // st + 1
                    x = torch.empty((B, buf.shape[1], n_out, n_out), dtype=buf.dtype, device=buf.device).contiguous(memory_format=cl)
                    check(lib.swb_nhwc_maxpool(buf.data_ptr(), x.data_ptr(), B, int(buf.shape[1]), n_in, n_in, k, st,
                                               torch.cuda.current_stream(buf.device).cuda_stream))
                    xb = None
                    continue
                if pre_relu:
                    torch.clamp_min(x, 0.0, out=view)
                else:
                    view.copy_(x)
                x = F.max_pool2d(buf[:B], k, st).contiguous(memory_format=fmt)
        if lib is not None and (self._last_relu or xb is not None):
            x = x.contiguous(memory_format=cl)
            self._paste(lib, x, x, B, (0, int(x.shape[2])), self._last_relu, xb)
        elif self._last_relu:
            x = torch.relu_(x)
        head = torch.relu(F.conv2d(x, self.head.weight, self.head.bias))       # [B, 2, w, w]
        return (head.sum((2, 3)) + self._outside) / float(self.blank_feat_hw * self.blank_feat_hw)

    @torch.no_grad()
    def __call__(self, crops_norm, offset):
        """crops_norm: [B, 3, c, c] normalised crops sitting at rows/cols [offset, offset + c) of the canvas."""
        F = torch.nn.functional
        xw = crops_norm
        win = (offset, offset + int(crops_norm.shape[-1]))
        for kind, layer, blank in self.plan:
            if kind == "relu":
                xw = torch.relu(xw)
            elif kind == "conv":
                k, s, p = layer.kernel_size[0], layer.stride[0], layer.padding[0]
                n_out = (blank.shape[-1] + 2 * p - k) // s + 1
                ow = self._out_window(win, k, s, p, n_out)
                need = (ow[0] * s - p, (ow[1] - 1) * s - p + k)
                xw = F.conv2d(self._patch(blank, xw, win, need, 0.0), layer.weight, layer.bias, stride=s)
                win = ow
            elif kind == "pool":
                k, s = layer.kernel_size, layer.stride
                n_in = blank.shape[-1]
                n_out = -((-(n_in - k)) // s) + 1 if layer.ceil_mode else (n_in - k) // s + 1
                if layer.ceil_mode and (n_out - 1) * s >= n_in:
                    n_out -= 1
                ow = self._out_window(win, k, s, 0, n_out)
                need = (ow[0] * s, (ow[1] - 1) * s + k)
                xw = F.max_pool2d(self._patch(blank, xw, win, need, float("-inf")), k, s)
                win = ow
            else:   # fire: squeeze 1x1 -> relu -> {expand 1x1, expand 3x3 (pad 1)} -> relu -> cat
                sq = torch.relu(F.conv2d(xw, layer.squeeze.weight, layer.squeeze.bias))
                n = blank.shape[-1]
                ow = (max(win[0] - 1, 0), min(win[1] + 1, n))
                need = (ow[0] - 1, ow[1] + 1)
                patch = self._patch(blank, sq, win, need, 0.0)
                e3 = F.conv2d(patch, layer.expand3x3.weight, layer.expand3x3.bias)
                e1 = F.conv2d(patch[:, :, 1:-1, 1:-1], layer.expand1x1.weight, layer.expand1x1.bias)
                xw = torch.relu(torch.cat([e1, e3], 1))
                win = ow
        head = torch.relu(F.conv2d(xw, self.head.weight, self.head.bias))      # [B, 2, w, w]
        a, b = win
        outside = self.blank_head.sum((2, 3)) - self.blank_head[:, :, a:b, a:b].sum((2, 3))
        return (head.sum((2, 3)) + outside) / float(self.blank_feat_hw * self.blank_feat_hw)


class SegmentClassifier:
    def __init__(self, model_path, device=None, batch_size=2048, channels_last=True, windowed=True, buffered=True):
        if device is None:   # the reference's module-level choice (:10)
            device = torch.device("cuda:0" if torch.cuda.is_available() else "cpu")
        self.device = torch.device(device)
        self.model = setup_model(2, self.device)
        state = model_path if isinstance(model_path, dict) else torch.load(model_path, map_location=self.device)
        self.model.load_state_dict(state)
        self.model.eval()
        for p in self.model.parameters():
            p.requires_grad = False
        self.channels_last = channels_last and self.device.type == "cuda"
        if self.channels_last:
            self.model = self.model.to(memory_format=torch.channels_last)
        self.batch_size = int(batch_size)
        self.buffered = bool(buffered)       # persistent patch buffers (WindowedSqueezeNet.forward_buffered)
        self._mean = torch.tensor(_MEAN, dtype=torch.float32, device=self.device).view(1, 3, 1, 1)
        self._std = torch.tensor(_STD, dtype=torch.float32, device=self.device).view(1, 3, 1, 1)
        self.windowed = None
        if windowed:
            with torch.no_grad():
                blank = torch.zeros((1, 3, 224, 224), dtype=torch.float32, device=self.device)
                blank = blank.sub_(self._mean).div_(self._std)
                self.windowed = WindowedSqueezeNet(self.model, blank)
        # the reference's per-image transform chain, for crops that need the PIL resize
        self.transforms = [
            transforms.ToPILImage(),
            transforms.Resize((CROP, CROP)),
            transforms.Pad(PAD),
            transforms.ToTensor(),
            transforms.Normalize(list(_MEAN), list(_STD)),
        ]

    # -- preprocessing -----------------------------------------------------------
    def preprocess(self, crops):
        """[B, 24, 24, 3] uint8 (numpy or torch, any device) -> [B, 3, 224, 224]
        float32 on the model's device; same values as ToPILImage -> Resize(24)
        (identity) -> Pad(100) -> ToTensor -> Normalize applied to each crop."""
        x = torch.as_tensor(crops)
        if x.dim() != 4 or tuple(x.shape[1:]) != (CROP, CROP, 3) or x.dtype != torch.uint8:
            raise ValueError("crops must be [B, %d, %d, 3] uint8" % (CROP, CROP))
        x = x.to(self.device, non_blocking=True).permute(0, 3, 1, 2).to(torch.float32).div(255)
        out = torch.zeros((x.shape[0], 3, 224, 224), dtype=torch.float32, device=self.device)
        out[:, :, PAD:PAD + CROP, PAD:PAD + CROP] = x
        out.sub_(self._mean).div_(self._std)
        if self.channels_last:
            out = out.contiguous(memory_format=torch.channels_last)
        return out

    def preprocess_crops(self, crops):
        """[B, 24, 24, 3] uint8 -> [B, 3, 24, 24] float32 normalised (the centre of ``preprocess``)."""
        x = torch.as_tensor(crops)
        if x.dim() != 4 or tuple(x.shape[1:]) != (CROP, CROP, 3) or x.dtype != torch.uint8:
            raise ValueError("crops must be [B, %d, %d, 3] uint8" % (CROP, CROP))
        x = x.to(self.device, non_blocking=True).permute(0, 3, 1, 2).to(torch.float32).div(255)
        x = x.sub_(self._mean).div_(self._std)
        return x.contiguous(memory_format=torch.channels_last) if self.channels_last else x.contiguous()

    def _transform_one(self, image):
        x = np.ascontiguousarray(image)
        for t in self.transforms:
            x = t(x)
        return x

    # -- scoring -----------------------------------------------------------------
    @torch.no_grad()
    def scores(self, crops):
        """Class scores [B, 2] (float32, on the model's device) for 24x24 crops."""
        n = int(crops.shape[0])
        out = torch.empty((n, 2), dtype=torch.float32, device=self.device)
        for a in range(0, n, self.batch_size):
            b = min(n, a + self.batch_size)
            if self.windowed is not None and self.buffered:
                out[a:b] = self.windowed.forward_buffered(self.preprocess_crops(crops[a:b]), PAD, self.batch_size)
            elif self.windowed is not None:
                out[a:b] = self.windowed(self.preprocess_crops(crops[a:b]), PAD)
            else:
                out[a:b] = self.model(self.preprocess(crops[a:b]))
        return out

    @torch.no_grad()
    def predict(self, crops):
        """argmax class per crop (torch.max(score, 1), :37-38) as a bool 'keep' mask."""
        if int(crops.shape[0]) == 0:
            return torch.zeros((0,), dtype=torch.bool, device=self.device)
        _, y = torch.max(self.scores(crops), 1)
        return y == 1

    @torch.no_grad()
    def __call__(self, segments):
        """segment_classification.py:27-45 with one batched forward pass."""
        segments = list(segments)
        keep = np.zeros(len(segments), dtype=bool)
        exact, odd = [], []
        for i, s in enumerate(segments):
            im = s.segment_image
            (exact if tuple(im.shape) == (CROP, CROP, 3) else odd).append(i)
        if exact:
            crops = np.stack([segments[i].segment_image for i in exact])
            keep[exact] = self.predict(crops).cpu().numpy()
        for i in odd:   # resized through PIL exactly as the reference does
            im = segments[i].segment_image
            if im.size == 0:
                raise ValueError("segment %d has an empty segment_image (bbox outside the frame)" % i)
            x = self._transform_one(im).unsqueeze(0).to(self.device)
            _, y = torch.max(self.model(x), 1)
            keep[i] = bool(y.item() == 1)
        kept = [s for s, k in zip(segments, keep) if k]
        for i, s in enumerate(kept):
            s.label = i + 1
        return kept

    # -- device-resident path -------------------------------------------------------
    @torch.no_grad()
    def classify_submit(self, ctx, n_rows, empty="raise"):
        """Keep mask [n_rows] (bool, device) for the segment table of ``ctx``'s last submit: the
        reference's segment images are cut AND (where they are not 24 x 24: a bird larger than 24 px, a
        bbox truncated by the frame edge) resized on the device exactly as ``transforms.Resize`` would
        (``swb_gather_crops``), and never visit the host.  A segment whose image is EMPTY (bbox within
        12 px of the frame's top / left edge: the reference's unclamped slice wraps around,
        image_filtering.py:363-365) makes the reference's ``ToPILImage`` raise: ``empty="raise"`` does the
        same (ValueError), ``empty="drop"`` classifies such segments as not kept.
        Needs device-resident full frames (device submit, or host submit of a full-frame ROI)."""
        if empty not in ("raise", "drop"):
            raise ValueError("empty must be 'raise' or 'drop'")
        if n_rows == 0:
            return torch.zeros((0,), dtype=torch.bool, device=self.device)
        crops = torch.empty((n_rows, CROP, CROP, 3), dtype=torch.uint8, device=self.device)
        _, rects = ctx.gather_crops(n_rows, CROP, out=crops)
        nonempty = (rects[:, 2] > rects[:, 0]) & (rects[:, 3] > rects[:, 1])
        if empty == "raise" and not bool(nonempty.all().item()):
            bad = int((~nonempty).nonzero()[0].item())
            raise ValueError("segment %d has an empty segment_image (bbox outside the frame)" % bad)
        return self.predict(crops) & nonempty
