#!/bin/bash
# Round-2 multi-GPU call (gpurun --gpus 8): strong scaling at global batch 512 (cfg2, cfg5), weak scaling of cfg5
# (12.8 MB gradient bucket), neuralNDCG weak scaling (count-weighted reduction), and the 2-GPU gradient-parity test.
mkdir -p gpurun_out/r2/scale
run() {  # run <n> <name> <bench args...>
  n=$1; name=$2; shift 2
  if [ "$n" = 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r2/scale/${name}_n1.json 2> gpurun_out/r2/scale/${name}_n1.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r2/scale/${name}_n$n.json 2> gpurun_out/r2/scale/${name}_n$n.err
  fi
}
timeout 600 python -m pytest tests/test_gpu_ddp.py -m gpu -q -s > gpurun_out/r2/scale/pytest_ddp.log 2>&1
for n in 1 2 4 8; do
  run $n strong_cfg2 --scaling strong --global-batch 512
  run $n strong_cfg5 --workload cfg5 --scaling strong --global-batch 512
  run $n weak_cfg5 --workload cfg5 --batch 1024
done
for n in 1 8; do
  run $n weak_cfg4 --workload cfg4 --batch 4096
  run $n weak_cfg2 --batch 4096
done
tail -3 gpurun_out/r2/scale/pytest_ddp.log
for f in gpurun_out/r2/scale/*.json; do echo "$f: $(python -c "import json; d=json.loads([l for l in open('$f') if l.startswith('{')][-1]); print(round(d['value']), round(d['ms_per_step'],3), d.get('allreduce'))" 2>&1 | tail -1)"; done
