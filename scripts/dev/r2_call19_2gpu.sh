#!/bin/bash
# shipped build on 2 GPUs: the NCCL test, then the driver's own launch of the bench (config 2 dealing, e2e included)
mkdir -p gpurun_out
echo "=== dp test"; timeout 600 python -m pytest tests/test_gpu_dp.py -x -q > gpurun_out/r2n_dp_test.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2n_dp_test.log | cut -c1-300; cat gpurun_out/dp_test_result.json 2>/dev/null | tr '\n' ' ' | cut -c1-900; echo
echo "=== bench --gpus 2 (driver launch line)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 2>/dev/null | grep '^{' > gpurun_out/r2n_bench_gs_2gpu.json; python -c "
import json; d=json.load(open('gpurun_out/r2n_bench_gs_2gpu.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config'].get('weak_8_views_per_rank'))"
