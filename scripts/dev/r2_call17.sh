#!/bin/bash
# PAIR build: parity + trainer tests in one process (the order that flaked), then the sanitizer passes
mkdir -p gpurun_out
echo "=== parity+trainer"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_trainer.py -x -q 2>&1 | tail -25
cat gpurun_out/densify_timing.json; echo
echo "=== again, trainer first"; timeout 900 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_parity.py -x -q 2>&1 | tail -4
echo "=== sanitizer"; timeout 2400 bash scripts/sanitize_r2.sh 2>&1 | tail -40
