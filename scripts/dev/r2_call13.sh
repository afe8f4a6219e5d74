#!/bin/bash
# final-build verification: A/B line, full GPU suite, smoke, bench line, sanitizer passes with the full racecheck log
mkdir -p gpurun_out
echo "=== default"; timeout 300 python scripts/dev/r2_ab.py --skip-r1 2>&1 | tail -1
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
echo "=== bench gs"; timeout 600 python bench.py > gpurun_out/r2m_bench_gs.json 2>/dev/null; cut -c1-200 gpurun_out/r2m_bench_gs.json
echo "=== bench train"; timeout 600 python scripts/bench_train.py > gpurun_out/bench_train_r2.json 2>gpurun_out/bench_train_r2.err; cut -c1-300 gpurun_out/bench_train_r2.json
echo "=== sanitizer"; timeout 2400 bash scripts/sanitize_r2.sh 2>&1 | tail -60
