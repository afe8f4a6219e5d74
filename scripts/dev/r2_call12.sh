#!/bin/bash
# backward without the staged id array (list position in the slot, id fetched in the second phase): A/B, parity, sanitizer
mkdir -p gpurun_out
echo "=== default"; timeout 300 python scripts/dev/r2_ab.py --skip-r1 2>&1 | tail -1
echo "=== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_trainer.py -x -q 2>&1 | tail -3
echo "=== sanitizer"; timeout 2400 bash scripts/sanitize_r2.sh 2>&1 | tail -45
