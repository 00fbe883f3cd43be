#!/bin/bash
# Round-2 GPU call 11: attention backward skips a dead second query half; the whole GPU suite; default bench + CPU leg.
mkdir -p gpurun_out/r11
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r11/pytest_gpu.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r11/pytest_gpu.log | cut -c1-300 | tail -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r11/smoke.log 2>&1; tail -2 gpurun_out/r11/smoke.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
timeout 600 python bench.py > gpurun_out/r11/bench_default.json 2> gpurun_out/r11/bench_default.err
timeout 300 $B --batch 64 > gpurun_out/r11/bench_cfg2_b64.json 2>&1
ARB_PACK_ROWS=0 timeout 300 $B > gpurun_out/r11/bench_cfg2_dense.json 2>&1
timeout 300 $B --dtype bf16 > gpurun_out/r11/bench_cfg2_bf16.json 2>&1
timeout 300 $B --workload cfg3 --batch 1024 --dtype bf16 > gpurun_out/r11/bench_cfg3_bf16.json 2>&1
timeout 300 $B --workload cfg3 --batch 1024 > gpurun_out/r11/bench_cfg3_tf32.json 2>&1
timeout 300 $B --workload cfg4 --batch 4096 > gpurun_out/r11/bench_cfg4.json 2>&1
timeout 300 $B --workload cfg5 --batch 1024 > gpurun_out/r11/bench_cfg5.json 2>&1
for f in gpurun_out/r11/bench_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads([l for l in open('$f') if l.startswith('{')][-1]); print(round(d['value']), round(d['ms_per_step'],3), d.get('roofline',{}).get('kernel'), d.get('roofline',{}).get('frac'), d.get('e2e',{}).get('value'), d.get('cpu_baseline',{}).get('value'))" 2>&1 | tail -1)"; done
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r11/bench_default.json') if l.startswith('{')][-1])
for k in d['roofline']['kernels']:
    print(f"{k['kernel']:58s} n={k['launches_per_step']:<3} {k['us_per_step']:8.1f} frac={k['frac']}")
PY
