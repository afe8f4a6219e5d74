#!/bin/bash
# first GPU call of round 2: parity of the new composite kernels, A/B timing, stage variants, one ncu capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
echo "=== A/B" ; timeout 420 python scripts/dev/r2_ab.py > gpurun_out/r2a_ab.log 2>&1; echo "ab rc=$?"; tail -5 gpurun_out/r2a_ab.log
echo "=== parity"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r2a_parity.log 2>&1; echo "parity rc=$?"; tail -5 gpurun_out/r2a_parity.log
for v in "GS_B200_FWD_STAGES=2 GS_B200_BWD_STAGES=3" "GS_B200_FWD_STAGES=4 GS_B200_BWD_OCC=3" "GS_B200_BWD_OCC=5"; do
  echo "=== variant $v"; env $v timeout 200 python scripts/dev/r2_ab.py --skip-r1 2>&1 | tail -2
done
echo "=== ncu"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:composite_ -s 6 -c 4 -o gpurun_out/prof_r2a python scripts/dev/r2_ab.py --skip-r1 --steps 1 > gpurun_out/r2a_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out | tail -8
