#!/bin/bash
# Round-2 GPU call 21: device-side row counts / extents / offsets read through the read-only cache (__ldg).
mkdir -p gpurun_out/r21
timeout 900 python -m pytest tests/test_gpu_pack_rows.py tests/test_gpu_scorer.py tests/test_gpu_bf16.py tests/test_gpu_gemm.py tests/test_gpu_graph.py -m gpu -q > gpurun_out/r21/pytest_sel.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r21/pytest_sel.log | cut -c1-300 | tail -12
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
timeout 300 $B > gpurun_out/r21/bench_cfg2.json 2>&1
timeout 300 $B --steps 50 --batch 64 > gpurun_out/r21/bench_cfg2_b64.json 2>&1
timeout 300 $B --workload cfg3 --batch 1024 --dtype bf16 > gpurun_out/r21/bench_cfg3_bf16.json 2>&1
for f in gpurun_out/r21/bench_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads([l for l in open('$f') if l.startswith('{')][-1]); print(round(d['value']), round(d['ms_per_step'],3), d.get('e2e',{}).get('value'))" 2>&1 | tail -1)"; done
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r21/bench_cfg2.json') if l.startswith('{')][-1])
for k in d['roofline']['kernels'][:24]:
    print(f"  {k['kernel']:58s} n={k['launches_per_step']:<3} {k['us_per_step']:8.1f} frac={k['frac']}")
PY
