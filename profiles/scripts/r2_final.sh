#!/bin/bash
# Round-2 final GPU call: what the driver runs (whole GPU suite, smoke, default bench, reference arm) on the final code,
# plus the evidence: launch list of one step with DRAM bytes, ncu --set full of the two attention kernels, other workloads.
O=gpurun_out/rfinal
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed" $O/pytest_gpu.log | cut -c1-300 | tail -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --batch 64 --cuda-graph > $O/bench_cfg2_b64_graph.json 2>&1
timeout 300 $B --batch 64 > $O/bench_cfg2_b64.json 2>&1
timeout 300 $B --batch 1024 > $O/bench_cfg2_b1024.json 2>&1
timeout 300 $B --dtype bf16 > $O/bench_cfg2_bf16.json 2>&1
ARB_PACK_ROWS=0 timeout 300 $B > $O/bench_cfg2_dense.json 2>&1
timeout 300 $B --workload cfg3 --batch 1024 --dtype bf16 > $O/bench_cfg3_bf16.json 2>&1
timeout 300 $B --workload cfg3 --batch 1024 > $O/bench_cfg3_tf32.json 2>&1
timeout 300 $B --workload cfg4 --batch 4096 > $O/bench_cfg4.json 2>&1
timeout 300 $B --workload cfg5 --batch 1024 > $O/bench_cfg5.json 2>&1
for f in $O/bench_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads([l for l in open('$f') if l.startswith('{')][-1]); print(round(d['value']), round(d['ms_per_step'],3), d.get('roofline',{}).get('kernel'), d.get('roofline',{}).get('frac'), d.get('e2e',{}).get('value'), d.get('cpu_baseline',{}).get('value'))" 2>&1 | tail -1)"; done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $O/launches_packed_b4096.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
for k in attn_bwd_kernel attn_fwd2_kernel ln_bwd_r_kernel; do
  timeout 300 $NCU -k regex:$k -s 4 -c 1 -o $O/$k python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_$k.log 2>&1
done
# the W1 forward linear (N512 K128 + ReLU, persistent kernel): 13 persistent launches per step, 3 warm-up steps, third forward one
timeout 300 $NCU -k regex:gemm_tf32_persistent -s 41 -c 1 -o $O/gemm_persist_w1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_gemm_persist_w1.log 2>&1
ls $O | head -40
