"""Helpers for the -m gpu parity tests: run frames through the C ABI and through the oracle."""
import hashlib

import oracle_lib as O
import cairo_zstd_b200 as czb
from cairo_zstd_b200 import api

_ctx = None


def ctx():
    global _ctx
    if _ctx is None:
        _ctx = czb.Context(0)
    return _ctx


def gpu_decode(frames, caps, flags=api.FLAG_VERIFY_CHECKSUM):
    return ctx().decode_batch(frames, caps, flags)


def compare_with_oracle(frames, caps, check_status_code=True, label=""):
    """Decode on the GPU and with the oracle; assert identical observable results per frame."""
    outs, res = gpu_decode(frames, caps)
    for i, f in enumerate(frames):
        st, want, ores = O.decode_frame(f, dst_cap=caps[i])
        r = res[i]
        tag = f"{label}[{i}] gpu={czb.status_name(r.status)} oracle={czb.status_name(st)}"
        if st == 0:
            assert r.status == 0, tag
            assert outs[i] == want, f"{tag}: output differs (len {len(outs[i])} vs {len(want)}, first diff at " \
                                    f"{next((k for k in range(min(len(outs[i]), len(want))) if outs[i][k] != want[k]), -1)})"
            assert r.bytes_written == ores.bytes_written, tag
            assert r.bytes_read == ores.bytes_read, tag
            assert r.blocks_decoded == ores.blocks_decoded, tag
            assert r.content_size == ores.content_size and r.window_size == ores.window_size, tag
            assert bool(r.has_checksum) == bool(ores.has_checksum), tag
            assert r.checksum_from_data == ores.checksum_from_data, tag
            assert r.checksum_calculated == ores.checksum_calculated, tag
            assert r.finished == ores.finished == 1, tag
        else:
            assert r.status != 0, tag
            if check_status_code:
                assert r.status == st, tag
    return outs, res


def sha(b):
    return hashlib.sha256(b).hexdigest()
