#!/usr/bin/env python3
"""Generate tests/golden/ from the reference's own fixtures.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box).  Source of truth: /root/reference/data/decode_corpus/zNNNNNN{,.zst}
-- the 100 (original, zstd frame) pairs the reference's generated tests embed
(script/generate_decode_corpus_tests.js:6 keeps those with original <= 1 KiB).

Outputs (committed):
  tests/golden/corpus_frames.bin    concatenated .zst frames (compressed side only)
  tests/golden/corpus_small.bin     concatenated originals of the <= 1 KiB pairs (the reference's own test set)
  tests/golden/corpus_index.json    per pair: name, offsets, lengths, sha256 of the original,
                                    checksum trailer (XXH64 low 32), in_reference_test_set
"""
import glob
import hashlib
import json
import os
import struct

REF = "/root/reference/data/decode_corpus"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SIZE_LIMIT = 1024  # script/generate_decode_corpus_tests.js:6


def main():
    os.makedirs(OUT, exist_ok=True)
    frames = bytearray()
    small = bytearray()
    index = []
    for zst in sorted(glob.glob(os.path.join(REF, "*.zst"))):
        name = os.path.basename(zst)[:-4]
        comp = open(zst, "rb").read()
        orig = open(zst[:-4], "rb").read()
        in_set = len(orig) <= SIZE_LIMIT
        ent = dict(
            name=name,
            frame_off=len(frames), frame_len=len(comp),
            orig_len=len(orig), orig_sha256=hashlib.sha256(orig).hexdigest(),
            trailer_xxh64_low32=struct.unpack("<I", comp[-4:])[0],
            in_reference_test_set=in_set,
        )
        if in_set:
            ent["small_off"] = len(small)
            small += orig
        frames += comp
        index.append(ent)
    open(os.path.join(OUT, "corpus_frames.bin"), "wb").write(frames)
    open(os.path.join(OUT, "corpus_small.bin"), "wb").write(small)
    json.dump(index, open(os.path.join(OUT, "corpus_index.json"), "w"), indent=0)
    print(f"{len(index)} pairs, {sum(e['in_reference_test_set'] for e in index)} in the reference's own test set, "
          f"{len(frames)} B of frames, {len(small)} B of small originals")


if __name__ == "__main__":
    main()
