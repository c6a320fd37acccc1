#!/bin/bash
# Decode-path timing (CUDA-graph decode, 7B, 336 px + 256-token prompt), same box:
#   A3_DEC_GATED=0    gated SiLU as the prologue of the w2 projection instead of the epilogue of w1|w3 (round 1)
# To compare against an earlier commit on the SAME box (boxes differ by 5-30 % on this path):
#   git worktree add _old <commit> && (cd _old && python -c "import __graft_entry__ as g; g.build()")
#   cp scripts/gpu_decode_sweep.py _old/scripts/ && gpurun -- '(cd _old && python scripts/gpu_decode_sweep.py 16 48); python scripts/gpu_decode_sweep.py 16 48'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/decode_sweep.jsonl
: > $out
run() { echo "$1" | tee -a $out; shift; env "$@" python scripts/gpu_decode_sweep.py ${BATCH:-1} ${TOK:-128} | tail -1 | tee -a $out; }
run "batch 1: gate in the w2 prologue (round 1)" A3_DEC_GATED=0
run "batch 1: default (gate as the w1|w3 epilogue)"
for b in 4 16 64; do BATCH=$b TOK=48 run "batch $b"; done
