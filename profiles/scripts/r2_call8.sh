#!/bin/bash
# Round-2 GPU call 8: packed rows as the default -- the whole GPU suite, smoke, the default bench with its CPU leg;
# A/B: persistent attention backward (item stream), rows per warp of the backward row kernels; launch list (ncu).
mkdir -p gpurun_out/r8
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r8/pytest_gpu.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r8/pytest_gpu.log | cut -c1-220 | tail -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r8/smoke.log 2>&1; tail -2 gpurun_out/r8/smoke.log
ARB_ATTN_BWD_PERSISTENT=1 timeout 600 python -m pytest tests/test_gpu_scorer.py tests/test_gpu_pack_rows.py tests/test_gpu_bf16.py tests/test_gpu_dropout.py -m gpu -q > gpurun_out/r8/pytest_bwd_persist.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r8/pytest_bwd_persist.log | cut -c1-220 | tail -20
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
timeout 600 python bench.py > gpurun_out/r8/bench_default.json 2> gpurun_out/r8/bench_default.err
ARB_ATTN_BWD_PERSISTENT=1 timeout 300 $B > gpurun_out/r8/bench_cfg2_bwdpersist.json 2>&1
ARB_ATTN_BWD_PERSISTENT=1 ARB_PACK_ROWS=0 timeout 300 $B > gpurun_out/r8/bench_cfg2_dense_bwdpersist.json 2>&1
ARB_ROWS_PER_WARP=16 timeout 300 $B > gpurun_out/r8/bench_cfg2_rpw16.json 2>&1
ARB_ROWS_PER_WARP=32 timeout 300 $B > gpurun_out/r8/bench_cfg2_rpw32.json 2>&1
ARB_ROWS_PER_WARP=4 timeout 300 $B > gpurun_out/r8/bench_cfg2_rpw4.json 2>&1
ARB_ATTN_BWD_PERSISTENT=1 timeout 300 $B --batch 64 > gpurun_out/r8/bench_cfg2_b64_bwdpersist.json 2>&1
timeout 300 $B --batch 64 > gpurun_out/r8/bench_cfg2_b64.json 2>&1
timeout 300 $B --batch 1024 > gpurun_out/r8/bench_cfg2_b1024.json 2>&1
for f in gpurun_out/r8/bench_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads([l for l in open('$f') if l.startswith('{')][-1]); print(round(d['value']), round(d['ms_per_step'],3), d.get('roofline',{}).get('kernel'), d.get('roofline',{}).get('frac'), d.get('e2e',{}).get('value'), d.get('cpu_baseline',{}).get('value'))" 2>&1 | tail -1)"; done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r8/launches_packed_b4096.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r8/bench_under_ncu.log 2>&1
tail -2 gpurun_out/r8/bench_under_ncu.log | cut -c1-200
