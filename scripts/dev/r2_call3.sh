#!/bin/bash
mkdir -p gpurun_out
echo "=== A/B (merged producer, slots 8 occ 6)"; timeout 300 python scripts/dev/r2_ab.py > gpurun_out/r2c_ab.log 2>&1; echo "ab rc=$?"; tail -3 gpurun_out/r2c_ab.log
for v in "GS_B200_BWD_OCC=5" "GS_B200_BWD_OCC=4" "GS_B200_BWD_SLOTS=16 GS_B200_BWD_OCC=5" "GS_B200_BWD_SLOTS=16 GS_B200_BWD_OCC=4" "GS_B200_BWD_STAGES=3 GS_B200_FWD_STAGES=4" "GS_B200_FWD_STAGES=2"; do
  echo "=== variant $v"; env $v timeout 200 python scripts/dev/r2_ab.py --skip-r1 2>&1 | tail -1
done
echo "=== parity"; timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r2c_parity.log 2>&1; echo "parity rc=$?"; tail -5 gpurun_out/r2c_parity.log
echo "=== ncu"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:composite_ -s 6 -c 2 -o gpurun_out/prof_r2c python scripts/dev/r2_ab.py --skip-r1 --steps 1 > gpurun_out/r2c_ncu.log 2>&1; echo "ncu rc=$?"
