#!/bin/bash
# Round-2 GPU call 1: microbenchmarks, loss/metric kernel timings + ncu captures, fresh baseline bench line.
set -x
mkdir -p gpurun_out/r2
cd profiles/microbench
for f in hbm_rw mma_sync_tf32 attn_fwd_mma_sync; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o /tmp/$f $f.cu && timeout 120 /tmp/$f > ../../gpurun_out/r2/mb_$f.log 2>&1
done
cd ../..
timeout 300 python profiles/run_slate_kernels.py --json gpurun_out/r2/slate_kernels.json > gpurun_out/r2/slate_kernels.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:'metrics_kernel|listnet|listmle_kernel|approx_ndcg_kernel|lambda_loss_kernel|ranknet_kernel|neural_ndcg' \
  -o gpurun_out/r2/slate_kernels python profiles/run_slate_kernels.py --batch 4096 > gpurun_out/r2/ncu_slate.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2/bench_base.json 2> gpurun_out/r2/bench_base.err
ls -la gpurun_out/r2
