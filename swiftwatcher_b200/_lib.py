"""ctypes binding of include/swb200.h (the C-ABI drop-in boundary).

The library is the product: if ``libswb200.so`` is missing this module raises
immediately — there is no Python / CPU fallback for any compute entry point.
"""

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libswb200.so")

SWB_OK = 0
ERR_INVALID, ERR_CUDA, ERR_CAPACITY, ERR_STATE = -1, -2, -3, -4
MEM_HOST, MEM_DEVICE = 0, 1
LABELS_I32, LABELS_U8 = 0, 1
HALO_CARRY = -1
OUT_MASK, OUT_LABELS = 1, 2
BG_MEDIAN, BG_RPCA = 0, 1


class SwbConfig(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("frame_h", C.c_int32), ("frame_w", C.c_int32),
        ("channels", C.c_int32), ("frame_pitch", C.c_int64), ("frame_stride", C.c_int64),
        ("roi_x0", C.c_int32), ("roi_y0", C.c_int32), ("roi_x1", C.c_int32), ("roi_y1", C.c_int32),
        ("median_n", C.c_int32), ("threshold", C.c_int32), ("morph_size", C.c_int32),
        ("do_open", C.c_int32), ("do_close", C.c_int32), ("label_mode", C.c_int32),
        ("out_flags", C.c_int32), ("max_frames", C.c_int32), ("max_segments", C.c_int32),
        ("bg_model", C.c_int32), ("gpu_share", C.c_int32), ("reserved", C.c_int32),
    ]


SEGMENT_DTYPE = np.dtype([
    ("frame", "<i4"), ("label", "<i4"), ("area", "<i4"), ("bbox", "<i4", (4,)),
    ("reserved", "<i4"), ("sum_row", "<i8"), ("sum_col", "<i8"),
])
assert SEGMENT_DTYPE.itemsize == 48

# every symbol include/swb200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_I32, _I64, _U32 = C.c_int32, C.c_int64, C.c_uint32
SYMBOLS = {
    "swb_version": (C.c_char_p, []),
    "swb_last_error": (C.c_char_p, [_P]),
    "swb_device_count": (C.c_int, [C.POINTER(_I32)]),
    "swb_create": (C.c_int, [C.POINTER(SwbConfig), C.POINTER(_P)]),
    "swb_destroy": (C.c_int, [_P]),
    "swb_reset": (C.c_int, [_P]),
    "swb_set_stream": (C.c_int, [_P, _P]),
    "swb_submit": (C.c_int, [_P, _P, _I32, _I32, _I32]),
    "swb_collect": (C.c_int, [_P, _P, _I64, C.POINTER(_I64), _P]),
    "swb_sync": (C.c_int, [_P]),
    "swb_collect_all": (C.c_int, [_P, _P, _I64, C.POINTER(_I64), _P, _P, _P]),
    "swb_collect_begin": (C.c_int, [_P, _P, _I64, _P, _P]),
    "swb_collect_end": (C.c_int, [_P, C.POINTER(_I64), _P]),
    "swb_set_option": (C.c_int, [_P, C.c_char_p, _I64]),
    "swb_last_subchunk": (C.c_int, [_P, C.POINTER(_I32)]),
    "swb_get_masks": (C.c_int, [_P, _I32, _I32, _P, _I32]),
    "swb_get_labels": (C.c_int, [_P, _I32, _I32, _P, _I32]),
    "swb_get_mask_bits": (C.c_int, [_P, _I32, _I32, _P, _I32]),
    "swb_device_views": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_I64), C.POINTER(_P),
                                   C.POINTER(_I64), C.POINTER(_P), C.POINTER(_P)]),
    "swb_enable_timing": (C.c_int, [_P, _I32]),
    "swb_get_timing": (C.c_int, [_P, C.POINTER(C.c_char_p), C.POINTER(C.c_float), _I32,
                                 C.POINTER(_I32)]),
    "swb_launch_count": (_I64, [_P]),
    "swb_gather_crops": (C.c_int, [_P, _I32, _P, _P, _I32]),
    "swb_stage_gray": (C.c_int, [_I32, _P, _I32, _I32, _P]),
    "swb_stage_median": (C.c_int, [_I32, _P, _I32, _I32, _I32, _P]),
    "swb_stage_absdiff": (C.c_int, [_I32, _P, _P, _I32, _I32, _P]),
    "swb_stage_thresh_to_zero": (C.c_int, [_I32, _P, _I32, _I32, _I32, _P]),
    "swb_stage_grey_morph": (C.c_int, [_I32, _P, _I32, _I32, _I32, _I32, _I32, _P]),
    "swb_stage_cc_label": (C.c_int, [_I32, _P, _I32, _I32, _P, _P, C.POINTER(_I32)]),
    "swb_stage_regionprops": (C.c_int, [_I32, _P, _I32, _I32, _I32, _P, _I32, C.POINTER(_I32)]),
    "swb_stage_rpca": (C.c_int, [_I32, _P, _I32, _I32, _I32, _P, C.POINTER(_I32)]),
    "swb_stage_bilateral": (C.c_int, [_I32, _P, _I32, _I32, _I32, C.c_double, C.c_double, _P]),
    "swb_rpca_stats": (C.c_int, [_P, C.POINTER(_I32), C.POINTER(_I32), C.POINTER(_I32)]),
    "swb_get_rpca": (C.c_int, [_P, _I32, _I32, _P, _I32]),
    "swb_tracker_create": (C.c_int, [_I32, _I32, C.POINTER(_P)]),
    "swb_tracker_destroy": (C.c_int, [_P]),
    "swb_tracker_costs": (C.c_int, [_P, _P, _P, _P, _I32, _P, _I32, C.POINTER(_P)]),
    "swb_tracker_last_error": (C.c_char_p, [_P]),
    "swb_tracker_launch_count": (_I64, [_P]),
    "swb_host_gather_tiles": (C.c_int, [_P, _I64, _I32, _I32, _I64, _P]),
    "swb_nhwc_paste": (C.c_int, [_P, _P, _I64, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P, _I32, _P]),
    "swb_nhwc_maxpool": (C.c_int, [_P, _P, _I64, _I32, _I32, _I32, _I32, _I32, _P]),
    "swb_host_alloc": (C.c_int, [C.POINTER(_P), C.c_uint64]),
    "swb_host_free": (C.c_int, [_P]),
    "swb_synth_frames": (C.c_int, [_I32, _P, _I32, _U32, _U32, _I32, _I32, _I32, _I32, _I32]),
}

_lib = None


def load():
    """Load libswb200.so (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "swiftwatcher_b200: %s is missing — build it with "
            "`python -m swiftwatcher_b200.build` (there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class SwbError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("swb200 error %d: %s" % (code, message))
        self.code = code


def check(rc, ctx=None):
    if rc != SWB_OK:
        msg = load().swb_last_error(ctx)
        raise SwbError(rc, (msg or b"").decode("utf-8", "replace"))


def device_count():
    n = _I32(0)
    rc = load().swb_device_count(C.byref(n))
    return n.value if rc == SWB_OK else 0


def gather_tiles(addresses, pitch, rows, row_bytes, out):
    """``out[i] = rows x row_bytes bytes starting at host address addresses[i]`` (row pitch ``pitch``)."""
    check(load().swb_host_gather_tiles(C.c_void_p(addresses.ctypes.data), int(pitch), int(rows), int(row_bytes),
                                       len(addresses), C.c_void_p(out.ctypes.data)))


def ptr(a):
    """Raw address of a numpy array / torch tensor / int."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError("cannot take the address of %r" % type(a))


_PINNED = []          # (first address, end address) of the live swb_host_alloc allocations


def is_pinned(address, nbytes):
    """True when [address, address + nbytes) lies inside one live swb_host_alloc allocation."""
    return any(a <= address and address + nbytes <= b for a, b in _PINNED)


class _PinnedOwner:
    """Owns one swb_host_alloc allocation; freed when the last numpy view goes away."""

    def __init__(self, nbytes):
        self.ptr = _P()
        check(load().swb_host_alloc(C.byref(self.ptr), nbytes))
        self.nbytes = nbytes
        self.range = (self.ptr.value, self.ptr.value + nbytes)
        _PINNED.append(self.range)

    def __del__(self):
        try:
            if self.ptr:
                if self.range in _PINNED:
                    _PINNED.remove(self.range)
                load().swb_host_free(self.ptr)
                self.ptr = _P()
        except Exception:
            pass


def pinned_empty(shape, dtype=np.uint8):
    """numpy array in page-locked host memory (swb_host_alloc): the host side of the
    frame ingest — decode into it, submit it, and the copy to the device is a DMA."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    owner = _PinnedOwner(max(n, 1))
    buf = (C.c_uint8 * max(n, 1)).from_address(owner.ptr.value)
    buf._owner = owner                      # keeps the allocation alive as long as any view exists
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
