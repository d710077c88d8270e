#!/bin/bash
# Round-2 measurement session on one B200 (run through gpurun): ncu captures of the v3 kernels, launch list of one eager
# step of the benchmark, driver-comparable bench lines of the other workloads, the stock-PyTorch context number.
mkdir -p gpurun_out
# 1. full ncu sets: v3 (TMA) forward on a wide PatchGAN layer and on the teacher's fused 1x1, TMA weight gradient
timeout 300 ncu --set full --clock-control none --import-source on -k regex:igemm_halo_persist -s 2 -c 1 -f -o gpurun_out/r02_ncu_persist_d2 \
    python tools/profile_gemm.py d2 --tiling 32,1,0,2 --iters 2 > gpurun_out/r02_ncu_persist_d2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:igemm_halo_persist -s 2 -c 1 -f -o gpurun_out/r02_ncu_persist_t1c \
    python tools/profile_gemm.py t1c --tiling 32,1,0,2 --iters 2 > gpurun_out/r02_ncu_persist_t1c.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:igemm_halo_wgrad -s 6 -c 1 -f -o gpurun_out/r02_ncu_wgrad_d2w \
    python tools/profile_gemm.py d2w --iters 2 > gpurun_out/r02_ncu_wgrad_d2w.log 2>&1
# 2. launch list of ONE eager step of the benchmark (kernels inside the NVTX range only)
timeout 600 ncu --nvtx --nvtx-include "catb_step/" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/r02_launches_step.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --nvtx --no-cpu-baseline --no-parity-probe > gpurun_out/r02_launches_bench.log 2>&1
# 3. the other workloads of BASELINE.json at one GPU (configs[2], [3], [4])
timeout 400 python bench.py --workload gaugan_5p6B --no-cpu-baseline > gpurun_out/r02_bench_gaugan_1gpu.json 2> gpurun_out/r02_bench_gaugan_1gpu.err
timeout 300 python bench.py --workload cyclegan_2p6B --no-cpu-baseline > gpurun_out/r02_bench_cyclegan_1gpu.json 2> gpurun_out/r02_bench_cyclegan_1gpu.err
for b in 1 4 16 32; do
  timeout 300 python bench.py --height 256 --width 512 --batch $b --no-cpu-baseline --no-parity-probe >> gpurun_out/r02_sweep_pix2pix_256x512.jsonl 2>> gpurun_out/r02_sweep.err
done
# 4. stock PyTorch + cuDNN on the same GPU (context)
timeout 400 python tools/torch_gpu_baseline.py > gpurun_out/r02_torch_gpu_baseline.json 2> gpurun_out/r02_torch_gpu_baseline.err
ls -la gpurun_out | tail -30
