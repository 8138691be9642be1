#!/bin/bash
# GPU box: compute-sanitizer (memcheck / racecheck / synccheck) over the GPU tests that stress the executors
# (k_exec_big forced on everything, giant sequences, tightly packed buffers, the corpus).  Summaries go to gpurun_out/.
# usage: scripts/r02_sanitize.sh <tag>
tag="${1:-r02}"
T="tests/test_gpu_parity.py"
SEL="test_cta_per_frame_executor_forced_on_everything or test_giant_sequences_through_the_cta_per_frame_executor or test_tightly_packed_buffers_any_alignment or test_corpus_all_100_one_batch or test_giant_sequences_and_mixed_chunk_lengths"
for tool in memcheck racecheck synccheck; do
    out="gpurun_out/${tag}_sanitizer_${tool}.log"
    echo "== compute-sanitizer --tool $tool" > "$out"
    timeout 1500 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 0 \
        python -m pytest $T -x -q -m gpu -k "$SEL" -p no:cacheprovider >> "$out" 2>&1
    echo "exit $?" >> "$out"
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit " "$out" | tail -5
done
