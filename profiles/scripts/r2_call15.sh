#!/bin/bash
# Round-2 GPU call 15: attention forward as an item stream (two CTAs per SM walk the items): parity + A/B.
mkdir -p gpurun_out/r15
for p in 1 0; do
ARB_ATTN_FWD_PERSISTENT=$p timeout 600 python -m pytest tests/test_gpu_scorer.py tests/test_gpu_pack_rows.py tests/test_gpu_bf16.py tests/test_gpu_dropout.py tests/test_shipped_configs.py -m gpu -q -x > gpurun_out/r15/pytest_fwdpersist$p.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r15/pytest_fwdpersist$p.log | cut -c1-300 | tail -8
done
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
ARB_ATTN_FWD_PERSISTENT=1 timeout 300 $B > gpurun_out/r15/bench_cfg2_fp1.json 2>&1
ARB_ATTN_FWD_PERSISTENT=0 timeout 300 $B > gpurun_out/r15/bench_cfg2_fp0.json 2>&1
ARB_ATTN_FWD_PERSISTENT=1 timeout 300 $B --batch 64 > gpurun_out/r15/bench_cfg2_b64_fp1.json 2>&1
ARB_ATTN_FWD_PERSISTENT=1 ARB_PACK_ROWS=0 timeout 300 $B > gpurun_out/r15/bench_cfg2_dense_fp1.json 2>&1
for f in gpurun_out/r15/bench_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads([l for l in open('$f') if l.startswith('{')][-1]); k=[x for x in d['roofline']['kernels'] if 'fwd2' in x['kernel']][0]; print(round(d['value']), round(d['ms_per_step'],3), 'fwd2', k['us_per_step'], d.get('e2e',{}).get('value'))" 2>&1 | tail -1)"; done
