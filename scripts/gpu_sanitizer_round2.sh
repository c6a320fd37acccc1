#!/bin/bash
# compute-sanitizer over the kernel tests at the end of round 2 (after the cross_entropy fix and the decode / patch-embed work).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_kernels_gpu.py tests/test_round2_gpu.py -q -x \
  -k "gemm or attn or rmsnorm or layernorm or embed or decode or cross_entropy or patch or sumsq or swiglu or rope" > gpurun_out/r02c_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/r02c_sanitizer_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_kernels_gpu.py -q -x \
  -k "rmsnorm or layernorm or colsum or swiglu or cross_entropy or argmax or decode_linear or attn_decode" > gpurun_out/r02c_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -c "Race reported" gpurun_out/r02c_sanitizer_racecheck.log; grep "Race reported" gpurun_out/r02c_sanitizer_racecheck.log | sed 's/.*access at //' | cut -c1-80 | sort | uniq -c | head; tail -3 gpurun_out/r02c_sanitizer_racecheck.log
