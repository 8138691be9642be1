"""include/frame_decoder.hpp (header-only C++ mirror of the reference's FrameDecoder traits) compiles against the C ABI
(CPU check) and, on a GPU box, decodes the reference's own 29-file test set through it (-m gpu)."""
import os
import struct
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_frame_decoder.cpp")
LIBDIR = os.path.join(ROOT, "cairo_zstd_b200")


def _build(tmp_path):
    from cairo_zstd_b200 import build
    build.build()
    exe = str(tmp_path / "test_frame_decoder")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
                           "-L", LIBDIR, "-lcairo_zstd_b200", f"-Wl,-rpath,{LIBDIR}"])
    return exe


def test_frame_decoder_hpp_compiles_and_links(tmp_path):
    exe = _build(tmp_path)
    assert os.path.exists(exe)
    # plain C must be able to include the ABI header too
    c = tmp_path / "abi.c"
    c.write_text('#include "cairo_zstd_b200.h"\nint main(void) { return czb_abi_version() == CZB_ABI_VERSION ? 0 : 1; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(tmp_path / "abi"),
                           "-L", LIBDIR, "-lcairo_zstd_b200", f"-Wl,-rpath,{LIBDIR}"])
    assert subprocess.call([str(tmp_path / "abi")]) == 0


@pytest.mark.gpu
def test_frame_decoder_hpp_decodes_the_reference_test_set(tmp_path, corpus):
    exe = _build(tmp_path)
    pack = tmp_path / "pack.bin"
    items = [(corpus.frame(i), corpus.small_original(i)) for i, e in enumerate(corpus.index) if e["in_reference_test_set"]]
    with open(pack, "wb") as f:
        f.write(struct.pack("<I", len(items)))
        for fr, orig in items:
            f.write(struct.pack("<I", len(fr))); f.write(fr)
            f.write(struct.pack("<I", len(orig))); f.write(orig)
    out = subprocess.run([exe, str(pack)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert f"{len(items)} frames, 0 failures" in out.stdout
