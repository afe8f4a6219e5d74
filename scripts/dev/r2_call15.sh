#!/bin/bash
# + queue stores as v2.f32 (no pair copies): A/B, parity, trainer tests
mkdir -p gpurun_out
echo "=== default"; timeout 300 python scripts/dev/r2_ab.py --skip-r1 2>&1 | tail -1
echo "=== again"; timeout 300 python scripts/dev/r2_ab.py --skip-r1 2>&1 | tail -1
echo "=== trainer tests"; timeout 900 python -m pytest tests/test_gpu_trainer.py -q 2>&1 | tail -15
echo "=== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
cat gpurun_out/densify_timing.json
