#!/bin/bash
# Round-2 GPU call 20: CUDA-graph replay of the training step (small batches), one zero_rows launch per call.
mkdir -p gpurun_out/r20
timeout 900 python -m pytest tests/test_gpu_graph.py tests/test_gpu_optim.py tests/test_gpu_pack_rows.py tests/test_gpu_scorer.py tests/test_gpu_bf16.py -m gpu -q > gpurun_out/r20/pytest_sel.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r20/pytest_sel.log | cut -c1-300 | tail -12
grep -n "Error" gpurun_out/r20/pytest_sel.log | head -5
B="python bench.py --steps 50 --warmup 5 --no-cpu-baseline"
timeout 300 $B --batch 64 > gpurun_out/r20/bench_cfg2_b64.json 2>&1
timeout 300 $B --batch 64 --cuda-graph > gpurun_out/r20/bench_cfg2_b64_graph.json 2>&1
timeout 300 $B --batch 256 --cuda-graph > gpurun_out/r20/bench_cfg2_b256_graph.json 2>&1
timeout 300 $B --batch 256 > gpurun_out/r20/bench_cfg2_b256.json 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r20/bench_cfg2.json 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --cuda-graph > gpurun_out/r20/bench_cfg2_graph.json 2>&1
for f in gpurun_out/r20/bench_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads([l for l in open('$f') if l.startswith('{')][-1]); print(round(d['value']), round(d['ms_per_step'],3), d.get('e2e',{}).get('value'), d.get('gpu_launches'))" 2>&1 | tail -1)"; done
tail -3 gpurun_out/r20/bench_cfg2_b64_graph.json | cut -c1-300
