#!/bin/bash
# Round-2 GPU call 3: re-run the test files that failed, ncu captures of the short-K GEMM (tile vs persistent kernel),
# of both attention kernels with the padding skip, and the per-launch DRAM traffic of one default step.
mkdir -p gpurun_out/r3
: > gpurun_out/r3/pytest_gpu.log
for f in tests/test_gpu_gemm_bf16.py tests/test_gpu_fc_block.py tests/test_shipped_configs.py tests/test_gpu_bf16.py \
         tests/test_gpu_losses.py tests/test_gpu_scorer.py; do
  echo "=== $f" >> gpurun_out/r3/pytest_gpu.log
  timeout 900 python -m pytest $f -m gpu -q -s >> gpurun_out/r3/pytest_gpu.log 2>&1
  echo "=== $f rc=$?" >> gpurun_out/r3/pytest_gpu.log
done
grep -E "^=== |passed|failed|^E  |FAILED" gpurun_out/r3/pytest_gpu.log | cut -c1-220 | tail -60
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch 1024"
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:gemm_tf32_kernel -s 3 -c 1 -o gpurun_out/r3/gemm_tile_w1 $B > gpurun_out/r3/ncu_tile.log 2>&1
ARB_GEMM_PERSISTENT=1 timeout 300 $NCU -k regex:gemm_tf32_persistent -s 3 -c 1 -o gpurun_out/r3/gemm_persist_w1 $B > gpurun_out/r3/ncu_persist.log 2>&1
timeout 300 $NCU -k regex:attn_bwd_kernel -s 2 -c 1 -o gpurun_out/r3/attn_bwd $B > gpurun_out/r3/ncu_attn_bwd.log 2>&1
timeout 300 $NCU -k regex:attn_fwd2_kernel -s 2 -c 1 -o gpurun_out/r3/attn_fwd2 $B > gpurun_out/r3/ncu_attn_fwd2.log 2>&1
# per-launch duration + DRAM bytes of one step of the default workload (B = 4096): skip the 3 warm-up steps (52 launches each)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 156 -c 60 \
  --csv --log-file gpurun_out/r3/launches_dram_b4096.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r3/ncu_launches.log 2>&1
ls -la gpurun_out/r3
