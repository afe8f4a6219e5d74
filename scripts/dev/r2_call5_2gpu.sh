#!/bin/bash
# 2-GPU call: NCCL tests + bench.py --gpus 2 (config 2 dealing, overlapped all-reduce)
mkdir -p gpurun_out
nvidia-smi -L
echo "=== dp test"; timeout 900 python -m pytest tests/test_gpu_dp.py -x -q > gpurun_out/r2e_dp_test.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r2e_dp_test.log; cat gpurun_out/dp_test_result.json 2>/dev/null
echo "=== bench --gpus 2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2e_bench_2gpu.json 2> gpurun_out/r2e_bench_2gpu.err; echo "rc=$?"; cut -c1-2500 gpurun_out/r2e_bench_2gpu.json; tail -5 gpurun_out/r2e_bench_2gpu.err
echo "=== bench --gpus 2, plain all-reduce for comparison (ar-chunks 1)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --ar-chunks 1 --no-e2e > gpurun_out/r2e_bench_2gpu_chunks1.json 2> gpurun_out/r2e_bench_2gpu_chunks1.err; echo "rc=$?"; cut -c1-600 gpurun_out/r2e_bench_2gpu_chunks1.json
echo "=== reference arm under torchrun"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>/dev/null | cut -c1-300
