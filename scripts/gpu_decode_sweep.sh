#!/bin/bash
# Decode-path timing experiments (CUDA-graph decode, 7B): same-box A/B of the round-2 changes.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/decode_sweep.jsonl
: > $out
V=$PWD/a3vlm_b200/lib/variants
run() { echo "$1" | tee -a $out; shift; env "$@" python scripts/gpu_decode_sweep.py ${BATCH:-1} ${TOK:-128} | tail -1 | tee -a $out; }
run "batch 1: round-1 schedule (no early KV reads, gate in the w2 prologue)" A3_DEC_PREWAIT=0 A3_DEC_GATED=0
run "batch 1: default (early KV reads + gate as the w1|w3 epilogue)" A3_DEC_PREWAIT=1 A3_DEC_GATED=1
for b in 4 16 64; do
  BATCH=$b TOK=48 run "batch $b: round-1 split heuristic, no early KV reads" A3_DEC_PREWAIT=0 A3_LIB_PATH=$V/oldsplit.so
  BATCH=$b TOK=48 run "batch $b: round-1 split heuristic, early KV reads" A3_DEC_PREWAIT=1 A3_LIB_PATH=$V/oldsplit.so
  BATCH=$b TOK=48 run "batch $b: new split heuristic, no early KV reads" A3_DEC_PREWAIT=0
  BATCH=$b TOK=48 run "batch $b: default" A3_DEC_PREWAIT=1
done
