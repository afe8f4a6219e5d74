#!/bin/bash
mkdir -p gpurun_out
echo "=== A/B default"; timeout 300 python scripts/dev/r2_ab.py --skip-r1 2>&1 | tail -1
for v in "GS_B200_BWD_VOTE=0" "GS_B200_FWD_STAGES=2" "GS_B200_FWD_STAGES=4" "GS_B200_BWD_STAGES=3" "GS_B200_STEP_OVERLAP=1" "GS_B200_STEP_OVERLAP=1 GS_B200_BWD_VOTE=0"; do
  echo "=== variant $v"; env $v timeout 200 python scripts/dev/r2_ab.py --skip-r1 2>&1 | tail -1
done
echo "=== trainer + mesh + parity gpu tests"; timeout 900 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_mesh.py tests/test_gpu_parity.py -q -x > gpurun_out/r2f_tests.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2f_tests.log; cat gpurun_out/densify_timing.json
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2f_launch_bench.log 2>&1; echo "rc=$?"; tail -c 300 gpurun_out/r2f_launch_bench.log
echo "=== ncu full"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:composite_ -s 6 -c 2 -o gpurun_out/prof_r2f python scripts/dev/r2_ab.py --skip-r1 --steps 1 > gpurun_out/r2f_ncu.log 2>&1; echo "ncu rc=$?"
echo "=== bench"; timeout 600 python bench.py > gpurun_out/r2f_bench_gs.json 2> gpurun_out/r2f_bench_gs.err; echo "rc=$?"; cut -c1-400 gpurun_out/r2f_bench_gs.json
