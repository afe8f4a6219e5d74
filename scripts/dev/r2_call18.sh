#!/bin/bash
# verification of the shipped build: trainer tests first (the order that flaked), full GPU suite, smoke, bench lines
mkdir -p gpurun_out
echo "=== trainer first"; timeout 900 python -m pytest tests/test_gpu_trainer.py -q 2>&1 | tail -4; cat gpurun_out/densify_timing.json; echo
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -4; cat gpurun_out/densify_timing.json; echo
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
echo "=== bench gs"; timeout 600 python bench.py > gpurun_out/r2n_bench_gs.json 2>/dev/null; cut -c1-200 gpurun_out/r2n_bench_gs.json
echo "=== bench train"; timeout 600 python scripts/bench_train.py > gpurun_out/bench_train_r2.json 2>gpurun_out/bench_train_r2.err; cut -c1-300 gpurun_out/bench_train_r2.json
echo "=== launch list (ncu, shares only)"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2n_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2n_ncu_bench.log 2>&1; echo "ncu rc=$?"
echo "=== ncu full on the composites"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:composite_ -s 6 -c 2 -o gpurun_out/prof_r2n python scripts/dev/r2_ab.py --skip-r1 --steps 1 > gpurun_out/r2n_ncu.log 2>&1; echo "ncu rc=$?"
