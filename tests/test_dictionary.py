"""Dictionaries.  The reference parses a dictionary (Dictionary::decode_dict,
src/decoding/dictionary.cairo:35-90) and never applies it (frame_decoder.cairo:73); the same parse exists here on the device
(czb_dictionary_parse_host) and in the oracle (oracle_dict_decode).  Dictionaries come from libzstd's trainer (ZDICT_trainFromBuffer:
input generation only).  CPU part: the oracle accepts them and reads the id libzstd reports.  GPU part: every field and the hash of
the four decoding tables equal the oracle's, and so do the statuses of corrupted dictionaries."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
import cairo_zstd_b200 as czb
from cairo_zstd_b200 import workloads as W


def train_dictionaries():
    z = W.libzstd()
    z.ZDICT_trainFromBuffer.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_size_t), C.c_uint]
    z.ZDICT_trainFromBuffer.restype = C.c_size_t
    z.ZDICT_getDictID.argtypes = [C.c_void_p, C.c_size_t]
    z.ZDICT_getDictID.restype = C.c_uint
    dicts = []
    for seed, (n_samples, sample_len, cap) in enumerate([(400, 900, 8192), (800, 300, 4096), (300, 2000, 32768), (600, 500, 16384)]):
        text = W.synth_text(n_samples * sample_len, 1000 + seed)
        sizes = (C.c_size_t * n_samples)(*([sample_len] * n_samples))
        buf = C.create_string_buffer(cap)
        n = z.ZDICT_trainFromBuffer(buf, cap, text, sizes, n_samples)
        if z.ZSTD_isError(n):
            continue
        d = buf.raw[:n]
        dicts.append((d, z.ZDICT_getDictID(d, len(d))))
    assert len(dicts) >= 2
    return dicts


def oracle_dict(d):
    info = O.DictInfo()
    return O.lib().oracle_dict_decode(d, len(d), C.byref(info)), info


def test_oracle_parses_libzstd_dictionaries():
    for d, dict_id in train_dictionaries():
        st, info = oracle_dict(d)
        assert st == 0
        assert info.id == dict_id
        assert 5 <= info.of_log <= 8 and 5 <= info.ml_log <= 9 and 5 <= info.ll_log <= 9 and 1 <= info.huf_max_bits <= 11
        assert info.content_off == 8 + info.huf_bytes + info.of_bytes + info.ml_bytes + info.ll_bytes + 12
        assert info.content_off + info.content_len == len(d) and info.content_len > 0
        assert all(1 <= o <= info.content_len for o in info.offset_hist)   # a trained dictionary's repeat offsets point into its content
    assert oracle_dict(b"\x37\xa4\x30")[0] == 100                           # .expect() on a short read traps
    assert oracle_dict(b"\x00" * 16)[0] == 55                               # BadMagicNum


FIELDS = ["id", "huf_bytes", "of_bytes", "ml_bytes", "ll_bytes", "huf_max_bits", "n_weights", "of_log", "ml_log", "ll_log", "table_hash",
          "content_off", "content_len"]


@pytest.mark.gpu
def test_dictionary_parse_matches_oracle():
    from gpu_common import ctx
    rng = np.random.default_rng(4)
    cases = []
    for d, _ in train_dictionaries():
        cases.append(d)
        hdr_end = oracle_dict(d)[1].content_off
        for _ in range(60):      # corrupt the header / table region, truncate, or stomp the magic
            b = bytearray(d)
            kind = rng.integers(0, 4)
            if kind == 0:
                b[int(rng.integers(8, hdr_end))] ^= 1 << int(rng.integers(0, 8))
            elif kind == 1:
                b[int(rng.integers(8, hdr_end))] = int(rng.integers(0, 256))
            elif kind == 2:
                b = b[:int(rng.integers(0, hdr_end + 4))]
            else:
                b[int(rng.integers(0, 8))] ^= 0xFF
            cases.append(bytes(b))
    cases += [b"", b"\x37\xa4\x30\xec", b"\x37\xa4\x30\xec\x01\x00\x00\x00"]
    n_ok = 0
    for k, d in enumerate(cases):
        st, want = oracle_dict(d)
        got = ctx().dictionary_parse(d)
        assert got.status == st, (k, czb.status_name(st), czb.status_name(got.status))
        if st == 0:
            n_ok += 1
            for f in FIELDS:
                assert getattr(got, f) == getattr(want, f), (k, f, getattr(got, f), getattr(want, f))
            assert list(got.offset_hist) == list(want.offset_hist)
    assert n_ok >= 4
