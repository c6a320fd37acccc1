#!/bin/bash
# N-GPU step time vs the number of CTAs NCCL may use (NCCL_MAX_CTAS; our GEMM grids unchanged).  The exchange needs
# ~50 GB/s of NVLink per GPU -- a handful of CTAs -- while every SM NCCL occupies delays a persistent GEMM CTA.
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in default 8 4 2; do
  if [ "$c" = default ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$c; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 6 --warmup 3 --no-dp-check 2>/dev/null \
    | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read().splitlines()[0]); print('NCCL_MAX_CTAS=$c', d['ms_per_step'], d['value'], d['roofline']['frac'])"
done
