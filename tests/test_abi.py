"""CPU-only checks of the C-ABI boundary: the library builds, loads and exports every symbol
include/cairo_zstd_b200.h declares; header-only entry points work without a GPU."""
import os
import re

import pytest

import cairo_zstd_b200 as czb
from cairo_zstd_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from cairo_zstd_b200 import build
    build.build()
    return czb.load_library()


def test_header_declares_exactly_the_exported_symbols(lib):
    text = open(os.path.join(ROOT, "include", "cairo_zstd_b200.h")).read()
    text += open(os.path.join(ROOT, "include", "czstd_status.h")).read()
    declared = set(re.findall(r"\b(cz[bs]_[a-z0-9_]+)\s*\(", text)) - {"czs_status"}  # "(czs_status)" appears in a comment
    assert declared == set(api.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.czb_abi_version() == 1


def test_status_names_match_header(lib):
    for code, name in api.STATUS_NAMES.items():
        assert lib.czs_status_name(code).decode() == name


def test_frame_header_info_matches_oracle(lib, corpus):
    import oracle_lib as O
    for i, e in enumerate(corpus.index):
        f = corpus.frame(i)
        info = czb.frame_header_info(f)
        st, out, res = O.decode_frame(f, dst_cap=e["orig_len"] + 8)
        assert info.status == 0
        assert info.window_size == res.window_size and info.content_size == res.content_size
        assert info.has_checksum_flag == 1
        st2, end = czb.find_frame_end(f + b"trailing")
        assert st2 == 0 and end == len(f) == res.bytes_read


def test_header_errors_match_oracle(lib):
    import struct
    import oracle_lib as O
    cases = [b"", b"\x00\x01\x02\x03\x04", struct.pack("<II", 0x184D2A50, 4), struct.pack("<I", 0x184D2A5F),
             struct.pack("<I", 0xFD2FB528), struct.pack("<IB", 0xFD2FB528, 0x00), struct.pack("<IBB", 0xFD2FB528, 0x03, 0),
             struct.pack("<IB", 0xFD2FB528, 0xE0), struct.pack("<IBB", 0xFD2FB528, 0x00, 0xFF)]
    for c in cases:
        assert czb.frame_header_info(c).status == O.decode_frame(c)[0], c


def test_no_gpu_means_loud_failure(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(czb.CzbError):
        czb.Context(0)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under cairo_zstd_b200/ or include/ may reference it."""
    for base in ("cairo_zstd_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in dirpath.split(os.sep):
                continue
            for fn in files:
                if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                    text = open(os.path.join(dirpath, fn), errors="ignore").read()
                    assert "zstd_oracle" not in text and "oracle_lib" not in text, (dirpath, fn)
