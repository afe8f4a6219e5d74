#!/bin/bash
# backward PAIR loop (both visits unconditional, one queue check per iteration): A/B + parity
mkdir -p gpurun_out
echo "=== default (PAIR=1)"; timeout 300 python scripts/dev/r2_ab.py --skip-r1 2>&1 | tail -1
echo "=== GS_B200_BWD_PAIR=0"; GS_B200_BWD_PAIR=0 timeout 300 python scripts/dev/r2_ab.py --skip-r1 2>&1 | tail -1
echo "=== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_trainer.py -x -q 2>&1 | tail -3
