#!/bin/bash
# Round-2 GPU call 19: small-batch work -- pack_rows with four rows in flight, approxNDCG with four lanes per item.
mkdir -p gpurun_out/r19
timeout 900 python -m pytest tests/test_gpu_losses.py tests/test_gpu_pack_rows.py tests/test_gpu_scorer.py tests/test_gpu_l3_training.py -m gpu -q > gpurun_out/r19/pytest_sel.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r19/pytest_sel.log | cut -c1-300 | tail -12
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
timeout 300 $B --batch 64 --steps 50 > gpurun_out/r19/bench_cfg2_b64.json 2>&1
timeout 300 $B --batch 256 > gpurun_out/r19/bench_cfg2_b256.json 2>&1
timeout 300 $B > gpurun_out/r19/bench_cfg2.json 2>&1
for f in gpurun_out/r19/bench_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads([l for l in open('$f') if l.startswith('{')][-1]); print(round(d['value']), round(d['ms_per_step'],3), d.get('e2e',{}).get('value'))" 2>&1 | tail -1)"; done
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r19/bench_cfg2_b64.json') if l.startswith('{')][-1])
for k in d['roofline']['kernels']:
    if any(t in k['kernel'] for t in ('pack','approx','zero','attn','extent')): print(f"  {k['kernel']:40s} n={k['launches_per_step']:<3} {k['us_per_step']:8.1f}")
PY
