#!/bin/bash
# One 8-GPU box: BASELINE config 4 (Llama-2-13B, batch 8/GPU, T 1024, sharded DP) at N = 8 / 4 / 2, the 2-GPU DP parity
# test, the 7B headline step at N = 8, and the config-5 replica-parallel inference sweep.  Outputs under gpurun_out/.
cd "$(dirname "$0")/.."
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for N in 8 4 2; do
  timeout 170 $TR --nproc-per-node $N --master-port $((29500 + N)) bench.py --gpus $N --model 13b --text-len 1024 --steps 5 --warmup 3 --no-cpu-baseline \
    > gpurun_out/r02_cfg4_13b_${N}gpu.json 2> gpurun_out/r02_cfg4_13b_${N}gpu.err
  echo "cfg4 N=$N rc=$? $(head -c 300 gpurun_out/r02_cfg4_13b_${N}gpu.json)"
done
timeout 120 python -m pytest tests/test_dp_gpu.py -q 2>&1 | tail -2
timeout 120 $TR --nproc-per-node 8 --master-port 29520 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline \
  > gpurun_out/r02_bench_7b_8gpu.json 2> gpurun_out/r02_bench_7b_8gpu.err
echo "7b N=8 rc=$? $(head -c 200 gpurun_out/r02_bench_7b_8gpu.json)"
timeout 240 python scripts/gpu_sweep_config5.py --gpus 8 --no-cpu > gpurun_out/r02_sweep5_8gpu.log 2>&1
echo "sweep rc=$?"; tail -3 gpurun_out/r02_sweep5_8gpu.log | cut -c1-300
