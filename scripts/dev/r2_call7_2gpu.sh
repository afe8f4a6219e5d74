#!/bin/bash
mkdir -p gpurun_out
echo "=== dp test"; timeout 600 python -m pytest tests/test_gpu_dp.py -x -q > gpurun_out/r2g_dp_test.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r2g_dp_test.log | cut -c1-400; cat gpurun_out/dp_test_result.json 2>/dev/null
for c in 2 4; do
echo "=== bench --gpus 2 --ar-chunks $c (8 views/rank)"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$c bench.py --gpus 2 --steps 5 --warmup 3 --views 8 --ar-chunks $c --no-e2e 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
done
echo "=== chunks 1"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29529 bench.py --gpus 2 --steps 5 --warmup 3 --views 8 --ar-chunks 1 --no-e2e 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
echo "=== chunks 8"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus 2 --steps 5 --warmup 3 --views 8 --ar-chunks 8 --no-e2e 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
