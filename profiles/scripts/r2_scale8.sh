#!/bin/bash
# Round-2 eight-GPU call (gpurun --gpus 8): weak scaling of the default workload at 1/2/4/8 GPUs, strong scaling at
# global batch 512 (8 GPUs), cfg5 weak at 8 GPUs (12.8 MB gradient bucket), and the 2-GPU gradient-parity test.
mkdir -p gpurun_out/r2/scale8
run() {  # run <n> <name> <bench args...>
  n=$1; name=$2; shift 2
  if [ "$n" = 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r2/scale8/${name}_n1.json 2> gpurun_out/r2/scale8/${name}_n1.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r2/scale8/${name}_n$n.json 2> gpurun_out/r2/scale8/${name}_n$n.err
  fi
}
for n in 1 2 4 8; do run $n weak_cfg2; done
run 8 strong_cfg2 --scaling strong --global-batch 512
run 8 weak_cfg5 --workload cfg5 --batch 1024
timeout 300 python -m pytest tests/test_gpu_ddp.py -m gpu -q > gpurun_out/r2/scale8/pytest_ddp.log 2>&1; tail -1 gpurun_out/r2/scale8/pytest_ddp.log
for f in gpurun_out/r2/scale8/*.json; do echo "$f: $(python -c "import json; d=json.loads([l for l in open('$f') if l.startswith('{')][-1]); print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d.get('allreduce'))" 2>&1 | tail -1)"; done
