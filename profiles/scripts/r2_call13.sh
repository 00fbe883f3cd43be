#!/bin/bash
# Round-2 GPU call 13: row layout on by default, attention backward with deferred output drain; GEMM mode 3 A/B.
mkdir -p gpurun_out/r13
timeout 900 python -m pytest tests/test_gpu_scorer.py tests/test_gpu_pack_rows.py tests/test_gpu_fc_block.py tests/test_gpu_bf16.py tests/test_gpu_dropout.py tests/test_shipped_configs.py -m gpu -q > gpurun_out/r13/pytest_sel.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r13/pytest_sel.log | cut -c1-300 | tail -20
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
timeout 300 $B > gpurun_out/r13/bench_cfg2.json 2>&1
ARB_GEMM_PERSISTENT=3 timeout 300 $B > gpurun_out/r13/bench_cfg2_p3.json 2>&1
ARB_PACK_ROWS=0 timeout 300 $B > gpurun_out/r13/bench_cfg2_dense.json 2>&1
timeout 300 $B --batch 64 > gpurun_out/r13/bench_cfg2_b64.json 2>&1
for f in gpurun_out/r13/bench_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads([l for l in open('$f') if l.startswith('{')][-1]); print(round(d['value']), round(d['ms_per_step'],3), d.get('roofline',{}).get('kernel'), d.get('roofline',{}).get('frac'), d.get('e2e',{}).get('value'))" 2>&1 | tail -1)"; done
python - <<'PY'
import json
for f in ('cfg2','cfg2_p3'):
    d=json.loads([l for l in open(f'gpurun_out/r13/bench_{f}.json') if l.startswith('{')][-1])
    print(f)
    for k in d['roofline']['kernels'][:12]:
        print(f"  {k['kernel']:58s} n={k['launches_per_step']:<3} {k['us_per_step']:8.1f} frac={k['frac']} bound={k['bound']}")
PY
