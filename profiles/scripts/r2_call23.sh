#!/bin/bash
# Round-2 GPU call 23: ReLU bit mask with the persistent kernel fetching its mask words ahead of the accumulator wait.
mkdir -p gpurun_out/r23
ARB_RELU_BITS=1 timeout 900 python -m pytest tests/test_gpu_scorer.py tests/test_gpu_pack_rows.py tests/test_gpu_gemm.py -m gpu -q > gpurun_out/r23/pytest_bits.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r23/pytest_bits.log | cut -c1-300 | tail -12
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
ARB_RELU_BITS=1 timeout 300 $B > gpurun_out/r23/bench_cfg2_bits1.json 2>&1
ARB_RELU_BITS=0 timeout 300 $B > gpurun_out/r23/bench_cfg2_bits0.json 2>&1
ARB_RELU_BITS=1 timeout 300 $B --workload cfg3 --batch 1024 > gpurun_out/r23/bench_cfg3_tf32_bits1.json 2>&1
ARB_RELU_BITS=0 timeout 300 $B --workload cfg3 --batch 1024 > gpurun_out/r23/bench_cfg3_tf32_bits0.json 2>&1
for f in gpurun_out/r23/bench_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads([l for l in open('$f') if l.startswith('{')][-1]); print(round(d['value']), round(d['ms_per_step'],3), d.get('e2e',{}).get('value'))" 2>&1 | tail -1)"; done
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r23/bench_cfg2_bits1.json') if l.startswith('{')][-1])
for k in d['roofline']['kernels'][:8]:
    print(f"  {k['kernel']:58s} n={k['launches_per_step']:<3} {k['us_per_step']:8.1f} frac={k['frac']}")
PY
