#!/bin/bash
# Round-2 GPU call 7: packed rows (padding removal) -- parity tests, A/B bench lines, the whole suite with packing on;
# persistent-GEMM aux-buffer modes.
mkdir -p gpurun_out/r7
timeout 600 python -m pytest tests/test_gpu_pack_rows.py -m gpu -q -x > gpurun_out/r7/pytest_pack.log 2>&1
tail -25 gpurun_out/r7/pytest_pack.log | cut -c1-220
timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -q > gpurun_out/r7/pytest_gemm.log 2>&1
tail -5 gpurun_out/r7/pytest_gemm.log | cut -c1-220
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
ARB_PACK_ROWS=1 timeout 300 $B > gpurun_out/r7/bench_cfg2_pack.json 2> gpurun_out/r7/bench_cfg2_pack.err
ARB_PACK_ROWS=0 timeout 300 $B > gpurun_out/r7/bench_cfg2_dense.json 2>&1
ARB_PACK_ROWS=1 ARB_GEMM_PERSISTENT=3 timeout 300 $B > gpurun_out/r7/bench_cfg2_pack_p3.json 2>&1
ARB_PACK_ROWS=1 ARB_GEMM_PERSISTENT=4 timeout 300 $B > gpurun_out/r7/bench_cfg2_pack_p4.json 2>&1
ARB_PACK_ROWS=0 ARB_GEMM_PERSISTENT=3 timeout 300 $B > gpurun_out/r7/bench_cfg2_dense_p3.json 2>&1
ARB_PACK_ROWS=1 timeout 300 $B --batch 64 > gpurun_out/r7/bench_cfg2_b64_pack.json 2>&1
ARB_PACK_ROWS=0 timeout 300 $B --batch 64 > gpurun_out/r7/bench_cfg2_b64_dense.json 2>&1
ARB_PACK_ROWS=1 timeout 300 $B --batch 512 > gpurun_out/r7/bench_cfg2_b512_pack.json 2>&1
ARB_PACK_ROWS=1 timeout 300 $B --dtype bf16 > gpurun_out/r7/bench_cfg2_bf16_pack.json 2>&1
ARB_PACK_ROWS=1 timeout 300 $B --workload cfg3 --batch 1024 --dtype bf16 > gpurun_out/r7/bench_cfg3_bf16_pack.json 2>&1
ARB_PACK_ROWS=1 timeout 300 $B --workload cfg4 --batch 4096 > gpurun_out/r7/bench_cfg4_pack.json 2>&1
ARB_PACK_ROWS=1 timeout 300 $B --workload cfg5 --batch 1024 > gpurun_out/r7/bench_cfg5_pack.json 2>&1
for f in gpurun_out/r7/bench_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads([l for l in open('$f') if l.startswith('{')][-1]); print(round(d['value']), round(d['ms_per_step'],3), d.get('roofline',{}).get('kernel'), d.get('roofline',{}).get('frac'), d.get('e2e',{}).get('value'))" 2>&1 | tail -1)"; done
ARB_PACK_ROWS=1 timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r7/pytest_gpu_packed.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r7/pytest_gpu_packed.log | cut -c1-220 | tail -40
