#!/bin/bash
mkdir -p gpurun_out
echo "=== A/B (2-visit pipeline)"; timeout 300 python scripts/dev/r2_ab.py > gpurun_out/r2b_ab.log 2>&1; echo "ab rc=$?"; tail -4 gpurun_out/r2b_ab.log
for v in "GS_B200_BWD_OCC=3" "GS_B200_BWD_OCC=5" "GS_B200_FWD_STAGES=2 GS_B200_BWD_STAGES=3"; do
  echo "=== variant $v"; env $v timeout 200 python scripts/dev/r2_ab.py --skip-r1 2>&1 | tail -1
done
echo "=== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_suite.log 2>&1; echo "suite rc=$?"; tail -15 gpurun_out/r2b_suite.log
echo "=== ncu"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:composite_ -s 6 -c 2 -o gpurun_out/prof_r2b python scripts/dev/r2_ab.py --skip-r1 --steps 1 > gpurun_out/r2b_ncu.log 2>&1; echo "ncu rc=$?"
