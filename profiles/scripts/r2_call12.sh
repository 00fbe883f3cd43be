#!/bin/bash
# Round-2 GPU call 12: row kernels with several rows per warp step (W = 128 / 256): parity tests under the new layout, A/B.
mkdir -p gpurun_out/r12
ARB_ROW_LAYOUT=15 timeout 900 python -m pytest tests/test_gpu_scorer.py tests/test_gpu_pack_rows.py tests/test_gpu_fc_block.py tests/test_gpu_bf16.py tests/test_gpu_dropout.py tests/test_shipped_configs.py tests/test_gpu_l3_training.py -m gpu -q > gpurun_out/r12/pytest_sel.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r12/pytest_sel.log | cut -c1-300 | tail -20
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
ARB_ROW_LAYOUT=0 timeout 300 $B > gpurun_out/r12/bench_cfg2_l0.json 2>&1
ARB_ROW_LAYOUT=15 timeout 300 $B > gpurun_out/r12/bench_cfg2_l15.json 2>&1
ARB_ROW_LAYOUT=15 timeout 300 $B --batch 64 > gpurun_out/r12/bench_cfg2_b64_l15.json 2>&1
ARB_ROW_LAYOUT=15 timeout 300 $B --workload cfg3 --batch 1024 --dtype bf16 > gpurun_out/r12/bench_cfg3_bf16_l15.json 2>&1
for f in gpurun_out/r12/bench_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads([l for l in open('$f') if l.startswith('{')][-1]); print(round(d['value']), round(d['ms_per_step'],3), d.get('roofline',{}).get('kernel'), d.get('roofline',{}).get('frac'), d.get('e2e',{}).get('value'))" 2>&1 | tail -1)"; done
python - <<'PY'
import json
for f in ('l0','l15'):
    d=json.loads([l for l in open(f'gpurun_out/r12/bench_cfg2_{f}.json') if l.startswith('{')][-1])
    print(f)
    for k in d['roofline']['kernels']:
        if any(t in k['kernel'] for t in ('ln_','head_','attn')): print(f"  {k['kernel']:40s} n={k['launches_per_step']:<3} {k['us_per_step']:8.1f} frac={k['frac']} bound={k['bound']}")
PY
