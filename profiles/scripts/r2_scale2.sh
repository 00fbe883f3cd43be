#!/bin/bash
# Round-2 two-GPU call (gpurun --gpus 2): the 2-GPU gradient-parity test, strong scaling at global batch 512 and weak
# scaling at 4096 slates per GPU (cfg2), neuralNDCG count-weighted reduction (cfg4).
mkdir -p gpurun_out/r2/scale2
run() {  # run <n> <name> <bench args...>
  n=$1; name=$2; shift 2
  if [ "$n" = 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r2/scale2/${name}_n1.json 2> gpurun_out/r2/scale2/${name}_n1.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r2/scale2/${name}_n$n.json 2> gpurun_out/r2/scale2/${name}_n$n.err
  fi
}
timeout 600 python -m pytest tests/test_gpu_ddp.py -m gpu -q -s > gpurun_out/r2/scale2/pytest_ddp.log 2>&1
for n in 1 2; do
  run $n strong_cfg2 --scaling strong --global-batch 512
done
run 2 weak_cfg2 --batch 4096
run 2 weak_cfg4 --workload cfg4 --batch 4096
tail -3 gpurun_out/r2/scale2/pytest_ddp.log
for f in gpurun_out/r2/scale2/*.json; do echo "$f: $(python -c "import json; d=json.loads([l for l in open('$f') if l.startswith('{')][-1]); print(round(d['value']), round(d['ms_per_step'],3), d.get('allreduce'))" 2>&1 | tail -1)"; done
