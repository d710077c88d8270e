#!/bin/bash
# Multi-GPU bench lines of BASELINE.json's configs at N GPUs of one box (run through `gpurun --gpus N`): the same command line the
# driver uses for the default workload, one line per workload appended to gpurun_out/r02_scale_${N}gpu.jsonl.
N=${1:-8}
mkdir -p gpurun_out
run() {
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --no-cpu-baseline "$@" >> gpurun_out/r02_scale_${N}gpu.jsonl 2>> gpurun_out/r02_scale_${N}gpu.err
}
run
run --workload cyclegan_2p6B
run --workload gaugan_5p6B
run --height 256 --width 512 --batch 16 --no-parity-probe
cat gpurun_out/r02_scale_${N}gpu.jsonl | cut -c1-200
