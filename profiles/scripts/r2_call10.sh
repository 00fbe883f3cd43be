#!/bin/bash
# Round-2 GPU call 10: row kernels without per-element divisions, delta kernel with 4 rows per warp, attention forward
# with its loads issued before the setup barrier, key-mask prefetch in the attention backward; ncu captures.
mkdir -p gpurun_out/r10
timeout 900 python -m pytest tests/test_gpu_scorer.py tests/test_gpu_pack_rows.py tests/test_gpu_fc_block.py tests/test_gpu_bf16.py tests/test_gpu_dropout.py tests/test_shipped_configs.py -m gpu -q > gpurun_out/r10/pytest_sel.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r10/pytest_sel.log | cut -c1-300 | tail -20
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
timeout 300 $B > gpurun_out/r10/bench_cfg2.json 2>&1
timeout 300 $B --batch 64 > gpurun_out/r10/bench_cfg2_b64.json 2>&1
timeout 300 $B --batch 1024 > gpurun_out/r10/bench_cfg2_b1024.json 2>&1
for f in gpurun_out/r10/bench_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads([l for l in open('$f') if l.startswith('{')][-1]); print(round(d['value']), round(d['ms_per_step'],3), d.get('roofline',{}).get('kernel'), d.get('roofline',{}).get('frac'), d.get('e2e',{}).get('value'))" 2>&1 | tail -1)"; done
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r10/bench_cfg2.json') if l.startswith('{')][-1])
for k in d['roofline']['kernels']:
    print(f"{k['kernel']:58s} n={k['launches_per_step']:<3} {k['us_per_step']:8.1f} frac={k['frac']}")
PY
NCU="ncu --set full --clock-control none --import-source on"
for k in head_fwd_kernel head_bwd_kernel ln_fwd_kernel attn_fwd2_kernel attn_bwd_kernel attn_delta_kernel; do
  timeout 300 $NCU -k regex:$k -s 4 -c 1 -o gpurun_out/r10/$k python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r10/ncu_$k.log 2>&1
done
ls -la gpurun_out/r10 | head -30
