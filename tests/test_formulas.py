"""CPU checks of the closed forms the kernels rely on, against the oracle's restatement of the reference."""
import ctypes as C

import numpy as np

import oracle_lib as O


def test_fse_entry_next_state_form_equals_calc_baseline_and_numbits():
    """czb_fse_build.cuh stores (symbol, next_state) with num_bits = log - floor(log2(next_state)) and
    base_line = (next_state << num_bits) - (1 << log).  Check it reproduces the reference's table
    (fse_decoder.cairo:156-256, :377-400) for random distributions at every accuracy log 5..9."""
    rng = np.random.default_rng(7)
    L = O.lib()
    for log in range(5, 10):
        size = 1 << log
        for trial in range(40):
            n = int(rng.integers(2, min(60, size - 8)))
            # random normalized counts summing to size, with some "-1" (less than one) symbols
            neg = int(rng.integers(0, min(n - 1, 6)))
            cuts = np.sort(rng.choice(np.arange(1, size - neg), size=n - neg - 1, replace=False)) if n - neg > 1 else np.array([], dtype=int)
            pos = np.diff(np.concatenate([[0], cuts, [size - neg]])).astype(np.int32)
            probs = np.concatenate([pos, -np.ones(neg, dtype=np.int32)])
            rng.shuffle(probs)
            bl = (C.c_uint32 * size)(); nb = (C.c_uint8 * size)(); sy = (C.c_uint8 * size)()
            arr = (C.c_int32 * len(probs))(*[int(p) for p in probs])
            assert L.oracle_fse_build_from_probs(log, arr, len(probs), bl, nb, sy) == 0
            counter = {}
            for i in range(size):
                s = sy[i]
                if probs[s] == -1:
                    ns = 1
                else:
                    k = counter.get(s, 0); counter[s] = k + 1
                    ns = int(probs[s]) + k
                num_bits = log - (ns.bit_length() - 1)
                base = (ns << num_bits) - size
                assert (num_bits, base) == (nb[i], bl[i]), (log, trial, i)
                assert ns < 1024


def test_symbolic_offset_encoding_roundtrip():
    """SYM_* arithmetic in czb_internal.cuh: subtracting 1 from enc(k, c) is enc(k, c + 1); resolution gives h[k] - c."""
    SYM_BASE, SYM_MID = 0xF0000000, 0x00800000
    enc = lambda k: SYM_BASE + (k << 24) + SYM_MID
    def resolve(v, h):
        t = v - SYM_BASE
        return (h[t >> 24] - (SYM_MID - (t & 0xFFFFFF))) & 0xFFFFFFFF
    def pack29(v):
        if v >= SYM_BASE:
            t = v - SYM_BASE
            return (1 << 28) | ((t >> 24) << 24) | (SYM_MID - (t & 0xFFFFFF))
        return min(v, (1 << 28) - 1)
    def resolve29(f, h):
        if f >> 28:
            return (h[(f >> 24) & 3] - (f & 0xFFFFFF)) & 0xFFFFFFFF
        return f
    h = (1000, 2000, 3000)
    for k in range(3):
        v = enc(k)
        for c in range(0, 100000, 7919):
            assert resolve(v - c, h) == (h[k] - c) & 0xFFFFFFFF
            assert resolve29(pack29(v - c), h) == (h[k] - c) & 0xFFFFFFFF
    assert pack29(12345) == 12345 and pack29(1 << 30) == (1 << 28) - 1
