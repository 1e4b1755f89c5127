"""The C-ABI library loads and exports every symbol include/gm_capi.h declares (no compute calls)."""
import ctypes as C
import os
import re

import pytest

from geometric_mapping_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "gm_capi.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.findall(r"GM_API\s+[\w\s\*]+?\b(gm_\w+)\s*\(", text)


def test_header_symbols_are_exported_and_bound():
    names = _declared()
    assert len(names) >= 40 and len(set(names)) == len(names)
    assert set(names) == set(capi.SYMBOLS), set(names) ^ set(capi.SYMBOLS)
    lib = C.CDLL(capi.library_path())
    for n in names:
        assert getattr(lib, n) is not None


def test_library_is_in_tree_and_has_sm100a_code():
    path = capi.library_path()
    assert path.startswith(ROOT) and os.path.exists(path)
    blob = open(path, "rb").read()
    assert b"sm_100a" in blob
    # the hot kernels are present in the fat binary
    for k in (b"k_count_plane", b"k_count_cyl", b"k_normals", b"k_rs2_down", b"k_crop", b"k_voxel_keys"):
        assert k in blob


def test_struct_layouts_match_header():
    assert C.sizeof(capi.gm_params) == 4 * 8 + 6 * 4 + 3 * 8 + 2 * 4 + 8 + 2 * 4
    assert capi.gm_params.weight_mode.offset == 96 and capi.gm_params.arrow_mode.offset == 100
    assert C.sizeof(capi.gm_counts) == 32
    assert C.sizeof(capi.gm_frame) == 84
    assert C.sizeof(capi.gm_arrow) == 56 and capi.ARROW_DTYPE.itemsize == 56
    assert C.sizeof(capi.gm_model) == 88
    assert C.sizeof(capi.gm_slice) == 40 and capi.SLICE_DTYPE.itemsize == 40
    assert C.sizeof(capi.gm_scan_summary) == 32 + 84 + 88 + 88 + 8
    assert C.sizeof(capi.gm_compression) == 160 and capi.gm_compression.bytes_in.offset == 144


def test_defaults_mirror_paramhandler():
    p = capi.default_params()
    # include/geometric_mapping/paramHandler.hpp:26-36
    assert (p.boxFilterBound, p.voxelGridLeafSize, p.neighborRadius, p.weightingFactor) == (5.0, 0.1, 0.03, 0.2)
    assert (p.displayCloud, p.displayNormals, p.displayCenterAxis, p.usePCLViz) == (1, 1, 1, 0)
    assert p.is_dense == 1 and p.nn_index_mode == 0 and p.ransacThreshold == 0.05
    assert p.weight_mode == 0 and p.arrow_mode == 0     # reference-faithful quirks by default


def test_status_strings_and_version():
    lib = capi._lib()
    assert lib.gm_version() >= 1
    assert b"no CPU path" in lib.gm_status_string(capi.GM_ERR_NO_DEVICE)
    assert lib.gm_profile_num_segments() == 21
    assert lib.gm_profile_segment_name(4) == b"normals"


def test_no_cpu_fallback_without_device():
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.GmError) as e:
        capi.Context()
    assert e.value.status == capi.GM_ERR_NO_DEVICE


def test_invalid_arguments_are_rejected_without_a_device():
    lib = capi._lib()
    h = C.c_void_p()
    bad = capi.default_params(neighborRadius=-1.0)
    assert lib.gm_create(C.byref(bad), 1000, 16, C.byref(h)) == capi.GM_ERR_INVALID_ARG
    assert lib.gm_create(C.byref(capi.default_params()), 0, 16, C.byref(h)) == capi.GM_ERR_INVALID_ARG
    assert lib.gm_crop(None) == capi.GM_ERR_INVALID_ARG


def test_product_does_not_import_the_oracle():
    """The product path may not import, call, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "geometric_mapping_b200")
    pat = re.compile(r"(^\s*(from|import)\s+oracle\b)|libgm_oracle|\bgmo_\w+\s*\(|#include\s+\".*oracle", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not pat.search(text), (dirpath, f)
    # and the shared library does not link it
    blob = open(capi.library_path(), "rb").read()
    assert b"libgm_oracle" not in blob and b"gmo_" not in blob
