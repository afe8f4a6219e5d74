#!/bin/bash
# A/B of the forward stop-vote variant (+ bfind bit iteration in the backward), then full GPU suite + smoke + bench lines
mkdir -p gpurun_out
echo "=== default (stop vote on)"; timeout 300 python scripts/dev/r2_ab.py --skip-r1 2>&1 | tail -1
echo "=== GS_B200_FWD_STOPVOTE=0"; GS_B200_FWD_STOPVOTE=0 timeout 300 python scripts/dev/r2_ab.py --skip-r1 2>&1 | tail -1
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "=== bench gs"; timeout 600 python bench.py > gpurun_out/r2k_bench_gs.json 2>/dev/null; cut -c1-200 gpurun_out/r2k_bench_gs.json
echo "=== bench train"; timeout 600 python scripts/bench_train.py > gpurun_out/bench_train_r2.json 2>gpurun_out/bench_train_r2.err; cut -c1-500 gpurun_out/bench_train_r2.json
echo "=== bench mesh"; timeout 600 python bench.py --workload mesh > gpurun_out/r2k_bench_mesh.json 2>/dev/null; cut -c1-200 gpurun_out/r2k_bench_mesh.json
echo "=== bench ngp"; timeout 900 python bench.py --workload ngp > gpurun_out/r2k_bench_ngp.json 2>/dev/null; cut -c1-200 gpurun_out/r2k_bench_ngp.json
echo "=== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
