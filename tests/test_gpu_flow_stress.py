"""-m gpu: the data-flow executor (k_exec_flow, both shapes) and the in-order one (k_exec_big) on frames of very different kinds --
text, long-window copies, periodic patterns with tiny offsets (self-overlapping matches, decode_buffer.cairo:101-120), literal-heavy
blocks with zero runs and raw blocks, low-entropy binary -- every frame forced through the CTA-per-frame executor, outputs compared
with the originals, XXH64 with the trailer, and the spin-limit safety net must not have fired (scripts/flow_stress.py has more seeds)."""
import ctypes as C
import os
import sys

import pytest

import cairo_zstd_b200 as czb
from cairo_zstd_b200 import api

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["flow-narrow", "flow-wide", "in-order"])
def test_cta_per_frame_executors_on_many_kinds_of_frames(monkeypatch, mode):
    import flow_stress
    monkeypatch.setenv("CZB_BIG_CLS", "0"); monkeypatch.setenv("CZB_BIG_SEQ_BYTES", "0")
    if mode == "flow-wide":
        monkeypatch.setenv("CZB_FLOW_WIDE", "1")
    if mode == "in-order":
        monkeypatch.setenv("CZB_BIG_FLOW", "0")
    ctx = czb.Context(0)
    frames, origs = flow_stress.make(7)
    outs, res = ctx.decode_batch(frames, [len(o) for o in origs], api.FLAG_VERIFY_CHECKSUM)
    for i, (o, g, r) in enumerate(zip(origs, outs, res)):
        assert r.status == 0, (i, czb.status_name(r.status))
        assert g == o, (i, len(o))
        assert r.checksum_calculated == r.checksum_from_data, i
    wd = (C.c_uint32 * 16)()
    czb.load_library().czb_debug_flow_watchdog(wd)
    assert wd[0] == 0, list(wd)
    ctx.close()
