#!/bin/bash
# shipped build on 8 GPUs: the driver's own launch line (config 2: 25 views per rank; weak line and e2e inside)
mkdir -p gpurun_out
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 3 --warmup 3 2> gpurun_out/r2n_bench_8gpu.err | grep '^{' > gpurun_out/r2n_bench_gs_8gpu.json; echo "rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2n_bench_gs_8gpu.json').read()); print(d['value'], d['ms_per_step'], d['config'].get('weak_8_views_per_rank'), d['e2e'] and d['e2e']['value'])"; tail -2 gpurun_out/r2n_bench_8gpu.err | cut -c1-200
