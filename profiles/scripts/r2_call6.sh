#!/bin/bash
# Round-2 GPU call 6: the whole GPU test suite in one pytest process (what the driver runs), smoke(), the default bench
# line with its CPU leg, the reference arm, and the A/B lines of the current code.
mkdir -p gpurun_out/r6
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r6/pytest_gpu.log 2>&1
tail -15 gpurun_out/r6/pytest_gpu.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r6/smoke.log 2>&1; tail -2 gpurun_out/r6/smoke.log
timeout 600 python bench.py > gpurun_out/r6/bench_default.json 2> gpurun_out/r6/bench_default.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r6/bench_reference.json 2> gpurun_out/r6/bench_reference.err
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
ARB_GEMM_PERSISTENT=1 timeout 300 $B > gpurun_out/r6/bench_cfg2_persist1.json 2>&1
ARB_ATTN_SKIP_PADDING=0 timeout 300 $B > gpurun_out/r6/bench_cfg2_dense.json 2>&1
timeout 300 $B --optimizer torch > gpurun_out/r6/bench_cfg2_torchadam.json 2>&1
timeout 300 $B --batch 64 > gpurun_out/r6/bench_cfg2_b64.json 2>&1
timeout 300 $B --batch 1024 > gpurun_out/r6/bench_cfg2_b1024.json 2>&1
timeout 300 $B --dtype bf16 > gpurun_out/r6/bench_cfg2_bf16.json 2>&1
timeout 300 $B --workload cfg3 --batch 1024 > gpurun_out/r6/bench_cfg3_tf32.json 2>&1
timeout 300 $B --workload cfg3 --batch 1024 --dtype bf16 > gpurun_out/r6/bench_cfg3_bf16.json 2>&1
timeout 300 $B --workload cfg4 --batch 4096 > gpurun_out/r6/bench_cfg4.json 2>&1
timeout 300 $B --workload cfg5 --batch 1024 > gpurun_out/r6/bench_cfg5.json 2>&1
for f in gpurun_out/r6/bench_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads([l for l in open('$f') if l.startswith('{')][-1]); print(round(d['value']), round(d['ms_per_step'],3), d.get('roofline',{}).get('kernel'), d.get('roofline',{}).get('frac'), d.get('cpu_baseline',{}).get('value'))" 2>&1 | tail -1)"; done
