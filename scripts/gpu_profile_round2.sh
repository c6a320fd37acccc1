#!/bin/bash
# Round-2 evidence: ncu launch list of one step, ncu --set full of the attention and GEMM kernels, compute-sanitizer.
cd "$(dirname "$0")/.."
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
    --csv --log-file gpurun_out/r02_launches_step.csv python bench.py --ncu > gpurun_out/r02_ncu_step.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 5 -c 1 -o gpurun_out/r02_attn_fwd python scripts/gpu_attn_fwd_only.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -s 3 -c 1 -o gpurun_out/r02_attn_bwd python scripts/gpu_attn_bwd_only.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16_kernel -s 150 -c 1 -o gpurun_out/r02_gemm2 python scripts/gpu_llm_gemm_bench.py > /dev/null 2>&1
echo "ncu full done"
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm or attn_fwd or attn_bwd or rmsnorm or layernorm or embed" > gpurun_out/r02_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/r02_sanitizer_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_kernels_gpu.py -q -x -k "rmsnorm or layernorm or colsum or swiglu or cross_entropy or argmax" > gpurun_out/r02_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -3 gpurun_out/r02_sanitizer_racecheck.log
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_kernels_gpu.py -q -x -k "test_attn_fwd or test_gemm_fp32_out" > gpurun_out/r02_sanitizer_racecheck_tc.log 2>&1
echo "racecheck(tcgen05 kernels) rc=$?"; tail -3 gpurun_out/r02_sanitizer_racecheck_tc.log
