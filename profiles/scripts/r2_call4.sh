#!/bin/bash
# Round-2 GPU call 4: localise the case-2 gradient discrepancy, re-run the touched test files, A/B bench lines with the
# compile-time specialised GEMM epilogue and the multi-row SIMT kernels.
mkdir -p gpurun_out/r4
timeout 600 python profiles/scripts/debug_case2.py > gpurun_out/r4/debug_case2.log 2>&1
cat gpurun_out/r4/debug_case2.log | cut -c1-260
: > gpurun_out/r4/pytest_gpu.log
for f in tests/test_gpu_gemm.py tests/test_gpu_gemm_bf16.py tests/test_gpu_scorer.py tests/test_gpu_dropout.py tests/test_gpu_fc_block.py \
         tests/test_shipped_configs.py tests/test_gpu_bf16.py tests/test_gpu_losses.py; do
  echo "=== $f" >> gpurun_out/r4/pytest_gpu.log
  timeout 900 python -m pytest $f -m gpu -q -s >> gpurun_out/r4/pytest_gpu.log 2>&1
  echo "=== $f rc=$?" >> gpurun_out/r4/pytest_gpu.log
done
grep -E "^=== |passed|failed|^E  |FAILED" gpurun_out/r4/pytest_gpu.log | cut -c1-220 | tail -40
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
timeout 300 $B > gpurun_out/r4/bench_cfg2_default.json 2> gpurun_out/r4/bench_cfg2_default.err
ARB_GEMM_PERSISTENT=1 timeout 300 $B > gpurun_out/r4/bench_cfg2_persist1.json 2>&1
ARB_GEMM_PERSISTENT=0 timeout 300 $B > gpurun_out/r4/bench_cfg2_persist0.json 2>&1
timeout 300 $B --batch 64 > gpurun_out/r4/bench_cfg2_b64.json 2>&1
timeout 300 $B --workload cfg3 --batch 1024 --dtype bf16 > gpurun_out/r4/bench_cfg3_bf16.json 2>&1
ARB_GEMM_PERSISTENT=1 timeout 300 $B --workload cfg3 --batch 1024 > gpurun_out/r4/bench_cfg3_tf32_persist1.json 2>&1
timeout 300 $B --workload cfg3 --batch 1024 > gpurun_out/r4/bench_cfg3_tf32.json 2>&1
for f in gpurun_out/r4/bench_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads([l for l in open('$f') if l.startswith('{')][-1]); print(round(d['value']), round(d['ms_per_step'],3), d['roofline']['kernel'], d['roofline']['frac'])" 2>&1 | tail -1)"; done
