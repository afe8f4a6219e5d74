#!/bin/bash
mkdir -p gpurun_out
echo "=== A/B (v4 reds, fwd scalar accumulators; slots 16 occ 4)"; timeout 300 python scripts/dev/r2_ab.py > gpurun_out/r2d_ab.log 2>&1; echo "ab rc=$?"; tail -3 gpurun_out/r2d_ab.log
for v in "GS_B200_BWD_SLOTS=8 GS_B200_BWD_OCC=5" "GS_B200_BWD_SLOTS=8 GS_B200_BWD_OCC=6" "GS_B200_BWD_OCC=5"; do
  echo "=== variant $v"; env $v timeout 200 python scripts/dev/r2_ab.py --skip-r1 2>&1 | tail -1
done
echo "=== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2d_suite.log 2>&1; echo "suite rc=$?"; tail -25 gpurun_out/r2d_suite.log
echo "=== bench gs"; timeout 600 python bench.py > gpurun_out/r2d_bench_gs.json 2> gpurun_out/r2d_bench_gs.err; echo "rc=$?"; cut -c1-1500 gpurun_out/r2d_bench_gs.json; tail -3 gpurun_out/r2d_bench_gs.err
echo "=== bench ngp"; timeout 900 python bench.py --workload ngp > gpurun_out/r2d_bench_ngp.json 2> gpurun_out/r2d_bench_ngp.err; echo "rc=$?"; cut -c1-1500 gpurun_out/r2d_bench_ngp.json; tail -3 gpurun_out/r2d_bench_ngp.err
echo "=== bench mesh"; timeout 900 python bench.py --workload mesh > gpurun_out/r2d_bench_mesh.json 2> gpurun_out/r2d_bench_mesh.err; echo "rc=$?"; cut -c1-1500 gpurun_out/r2d_bench_mesh.json; tail -3 gpurun_out/r2d_bench_mesh.err
