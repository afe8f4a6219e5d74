#!/bin/bash
mkdir -p gpurun_out
echo "=== default (TMA gather)"; timeout 300 python scripts/dev/r2_ab.py --skip-r1 2>&1 | tail -1
echo "=== GS_B200_GATHER=ldgsts"; GS_B200_GATHER=ldgsts timeout 300 python scripts/dev/r2_ab.py 2>&1 | tail -2
echo "=== parity with ldgsts gather"; GS_B200_GATHER=ldgsts timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "forward_backward_parity or golden or config1" 2>&1 | tail -3
echo "=== ncu ldgsts"; GS_B200_GATHER=ldgsts timeout 300 ncu --set full --clock-control none -k regex:composite_ -s 6 -c 2 -o gpurun_out/prof_r2i_ldgsts python scripts/dev/r2_ab.py --skip-r1 --steps 1 > gpurun_out/r2i_ncu.log 2>&1; echo "ncu rc=$?"
echo "=== bench default"; timeout 600 python bench.py > gpurun_out/r2i_bench_gs.json 2> gpurun_out/r2i_bench_gs.err; echo "rc=$?"; cut -c1-300 gpurun_out/r2i_bench_gs.json
