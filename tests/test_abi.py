"""CPU suite: the C-ABI library builds, loads and exports every symbol that
include/swb200.h declares; no compute call succeeds without a GPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "swb200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(swb_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    names = declared_functions()
    for must in ("swb_create", "swb_submit", "swb_collect", "swb_destroy", "swb_last_error",
                 "swb_get_masks", "swb_get_labels", "swb_stage_gray", "swb_stage_cc_label"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    from swiftwatcher_b200 import _lib
    names = declared_functions()
    for n in names:
        assert hasattr(lib, n), "libswb200.so does not export %s" % n
    assert set(_lib.SYMBOLS) == set(names), "ctypes binding and header disagree"


def test_header_is_plain_c_and_layouts_match(tmp_path):
    """The header compiles as C (no C++/torch types) and struct layouts equal
    the ctypes / numpy mirrors."""
    from swiftwatcher_b200 import _lib
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "swb200.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu\\n", sizeof(swb_segment), sizeof(swb_config),'
                   ' offsetof(swb_segment, sum_row), offsetof(swb_segment, bbox), offsetof(swb_config, roi_x0));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    seg, cfg, off_sum, off_bbox, off_roi = map(int, out)
    assert seg == _lib.SEGMENT_DTYPE.itemsize == 48
    assert off_sum == _lib.SEGMENT_DTYPE.fields["sum_row"][1]
    assert off_bbox == _lib.SEGMENT_DTYPE.fields["bbox"][1]
    assert cfg == C.sizeof(_lib.SwbConfig)
    assert off_roi == _lib.SwbConfig.roi_x0.offset


def test_version_and_error_plumbing(lib):
    assert lib.swb_version().startswith(b"swb200")
    assert lib.swb_create(None, None) != 0
    assert b"null" in lib.swb_last_error(None)


def test_invalid_config_is_rejected_before_touching_cuda(lib):
    from swiftwatcher_b200 import _lib
    cfg = _lib.SwbConfig()
    cfg.frame_h, cfg.frame_w, cfg.channels = 100, 100, 3
    cfg.median_n, cfg.max_frames = 4, 8            # even window
    ctx = C.c_void_p()
    assert lib.swb_create(C.byref(cfg), C.byref(ctx)) == _lib.ERR_INVALID
    assert b"median_n" in lib.swb_last_error(None)
    cfg.median_n = 5
    cfg.roi_x0, cfg.roi_x1, cfg.roi_y0, cfg.roi_y1 = 50, 150, 0, 10   # outside the frame
    assert lib.swb_create(C.byref(cfg), C.byref(ctx)) == _lib.ERR_INVALID


def test_no_cpu_fallback_without_gpu(lib):
    """Without a CUDA device every compute entry point must fail loudly."""
    import swiftwatcher_b200 as swb
    from swiftwatcher_b200 import image_filtering as img
    if swb.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(swb.SwbError) as e:
        swb.FilterContext((64, 64, 3))
    assert e.value.code == -2
    with pytest.raises(swb.SwbError):
        img.convert_grayscale(np.zeros((8, 8, 3), np.uint8))
    with pytest.raises(swb.SwbError):
        img.cc_labeling(np.zeros((8, 8), np.uint8), 4)


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: the package must never import it."""
    pkg = os.path.join(ROOT, "swiftwatcher_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
