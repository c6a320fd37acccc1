#!/bin/bash
# Decode-path timing (CUDA-graph decode, 7B, 336 px + 256-token prompt): same-box A/B of the round-2 changes.
#   A3_DEC_PREWAIT=0  the attention kernel waits for the QKV projection before touching the KV cache (round 1)
#   A3_DEC_GATED=0    gated SiLU as the prologue of the w2 projection instead of the epilogue of w1|w3 (round 1)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/decode_sweep.jsonl
: > $out
run() { echo "$1" | tee -a $out; shift; env "$@" python scripts/gpu_decode_sweep.py ${BATCH:-1} ${TOK:-128} | tail -1 | tee -a $out; }
run "batch 1: round-1 schedule" A3_DEC_PREWAIT=0 A3_DEC_GATED=0
run "batch 1: early KV reads" A3_DEC_PREWAIT=1 A3_DEC_GATED=0
run "batch 1: default (early KV reads + gate as the w1|w3 epilogue)"
for b in 4 16 64; do
  BATCH=$b TOK=48 run "batch $b: no early KV reads" A3_DEC_PREWAIT=0
  BATCH=$b TOK=48 run "batch $b: default"
done
