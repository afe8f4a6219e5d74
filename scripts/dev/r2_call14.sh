#!/bin/bash
# backward queue slots = list position only (mean2D / opacity re-read from the record in the second phase): A/B + parity
mkdir -p gpurun_out
echo "=== default"; timeout 300 python scripts/dev/r2_ab.py --skip-r1 2>&1 | tail -1
echo "=== again"; timeout 300 python scripts/dev/r2_ab.py --skip-r1 2>&1 | tail -1
echo "=== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_trainer.py -x -q 2>&1 | tail -3
