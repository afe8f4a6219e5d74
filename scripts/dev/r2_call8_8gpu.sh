#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
echo "=== bench --gpus 8 (config 2: 25 views/rank, chunks 4)"; timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 2> gpurun_out/r2h_bench_8gpu.err | grep '^{' > gpurun_out/r2h_bench_8gpu.json; echo "rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2h_bench_8gpu.json').read()); print(d['value'], d['ms_per_step'], d['config'].get('weak_8_views_per_rank'), d['e2e'] and d['e2e']['value'])"; tail -3 gpurun_out/r2h_bench_8gpu.err | cut -c1-300
echo "=== weak 8 views/rank, plain all-reduce (chunks 1)"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 5 --warmup 3 --views 8 --ar-chunks 1 --no-e2e 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
echo "=== weak 8 views/rank, chunks 2"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --steps 5 --warmup 3 --views 8 --ar-chunks 2 --no-e2e 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
