"""The committed Cairo-side harness (harness/cairo, generated on a B200 by scripts/gen_cairo_harness.py) is consistent
with the reference's fixtures: same 29 files as the reference's generator picks (script/generate_decode_corpus_tests.js:6),
the embedded compressed bytes are the corpus frames, and the embedded GPU outputs are the corpus originals."""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H = os.path.join(ROOT, "harness", "cairo")


def _arrays(text):
    return [bytes(int(x, 16) for x in re.findall(r"0x([0-9a-f]{2})\b", m)) for m in re.findall(r"array!\[(.*?)\]", text, re.S)]


def test_committed_harness_matches_the_fixtures(corpus):
    names = [e["name"] for e in corpus.index if e["in_reference_test_set"]]
    mods = re.findall(r"mod (\w+);", open(os.path.join(H, "src", "tests", "gpu_parity.cairo")).read())
    assert sorted(mods) == sorted(names) and len(names) == 29
    rec = {r["name"]: r for r in json.load(open(os.path.join(H, "gpu_outputs.json")))["files"]}
    for i, e in enumerate(corpus.index):
        if not e["in_reference_test_set"]:
            continue
        text = open(os.path.join(H, "src", "tests", "gpu_parity", e["name"] + ".cairo")).read()
        comp, out = _arrays(text)
        assert comp == corpus.frame(i) and out == corpus.small_original(i), e["name"]
        assert bytes.fromhex(rec[e["name"]]["gpu_output_hex"]) == out and rec[e["name"]]["status"] == 0
        assert f"0x{rec[e['name']]['gpu_checksum']:08x}" in text
        assert f"fn test_gpu_parity_{e['name']}()" in text and "FrameDecoderStateTrait::new(ref source)" in text
