#!/bin/bash
# Round-2 GPU call 2: full GPU test suite, then A/B bench lines (padding skip, persistent GEMM everywhere, PDL), other
# workloads, bf16.
mkdir -p gpurun_out/r2
# one pytest process per file: a faulting kernel poisons only its own CUDA context
: > gpurun_out/r2/pytest_gpu.log
for f in tests/test_gpu_gemm.py tests/test_gpu_gemm_bf16.py tests/test_gpu_scorer.py tests/test_gpu_dropout.py \
         tests/test_gpu_fc_block.py tests/test_shipped_configs.py tests/test_gpu_bf16.py tests/test_gpu_losses.py \
         tests/test_gpu_metrics.py tests/test_gpu_optim.py tests/test_gpu_slates.py tests/test_gpu_l3_training.py \
         tests/test_gpu_ddp.py; do
  echo "=== $f" >> gpurun_out/r2/pytest_gpu.log
  timeout 900 python -m pytest $f -m gpu -q -x -s >> gpurun_out/r2/pytest_gpu.log 2>&1
  echo "=== $f rc=$?" >> gpurun_out/r2/pytest_gpu.log
done
grep -E "^=== |passed|failed|error|Error|assert" gpurun_out/r2/pytest_gpu.log | cut -c1-200 | tail -80
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
timeout 300 $B > gpurun_out/r2/bench_cfg2_default.json 2> gpurun_out/r2/bench_cfg2_default.err
ARB_PDL=0 timeout 300 $B > gpurun_out/r2/bench_cfg2_nopdl.json 2>&1
ARB_ATTN_SKIP_PADDING=0 timeout 300 $B > gpurun_out/r2/bench_cfg2_dense.json 2>&1
ARB_GEMM_PERSISTENT=1 timeout 300 $B > gpurun_out/r2/bench_cfg2_persist1.json 2>&1
ARB_GEMM_PERSISTENT=0 timeout 300 $B > gpurun_out/r2/bench_cfg2_persist0.json 2>&1
timeout 300 $B --batch 64 > gpurun_out/r2/bench_cfg2_b64.json 2>&1
ARB_PDL=0 timeout 300 $B --batch 64 > gpurun_out/r2/bench_cfg2_b64_nopdl.json 2>&1
timeout 300 $B --batch 64 --optimizer torch > gpurun_out/r2/bench_cfg2_b64_torchadam.json 2>&1
timeout 300 $B --workload cfg3 --batch 1024 > gpurun_out/r2/bench_cfg3_tf32.json 2>&1
timeout 300 $B --workload cfg3 --batch 1024 --dtype bf16 > gpurun_out/r2/bench_cfg3_bf16.json 2>&1
timeout 300 $B --dtype bf16 > gpurun_out/r2/bench_cfg2_bf16.json 2>&1
timeout 300 $B --workload cfg4 --batch 4096 > gpurun_out/r2/bench_cfg4.json 2>&1
timeout 300 $B --workload cfg5 --batch 1024 > gpurun_out/r2/bench_cfg5.json 2>&1
timeout 600 python bench.py --impl reference-gpu --steps 5 --warmup 3 --batch 64 > gpurun_out/r2/refgpu_cfg2_b64.json 2>&1
timeout 600 python bench.py --impl reference-gpu --steps 3 --warmup 3 --batch 512 > gpurun_out/r2/refgpu_cfg2_b512.json 2>&1
for f in gpurun_out/r2/bench_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads([l for l in open('$f') if l.startswith('{')][-1]); print(round(d['value']), round(d['ms_per_step'],3), d['roofline']['kernel'], d['roofline']['frac'])" 2>&1 | tail -1)"; done
