#!/bin/bash
# cp.async gather is the default now: ring-depth sweeps; host-buffer step with the late SH upload; bench line
mkdir -p gpurun_out
echo "=== host-step tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "host_buffer_step or multiview_step_entries or chunking" 2>&1 | tail -5
echo "=== default"; timeout 300 python scripts/dev/r2_ab.py --skip-r1 2>&1 | tail -1
echo "=== FWD_STAGES=2"; GS_B200_FWD_STAGES=2 timeout 300 python scripts/dev/r2_ab.py --skip-r1 2>&1 | tail -1
echo "=== FWD_STAGES=4"; GS_B200_FWD_STAGES=4 timeout 300 python scripts/dev/r2_ab.py --skip-r1 2>&1 | tail -1
echo "=== BWD_STAGES=3"; GS_B200_BWD_STAGES=3 timeout 300 python scripts/dev/r2_ab.py --skip-r1 2>&1 | tail -1
echo "=== bench (late SH)"; timeout 600 python bench.py > gpurun_out/r2j_bench_gs.json 2> gpurun_out/r2j_bench_gs.err; echo "rc=$?"; python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r2j_bench_gs.json') if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'])
P
echo "=== bench e2e with GS_B200_HOST_LATE_SH=0"; GS_B200_HOST_LATE_SH=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2j_bench_gs_nolate.json 2>/dev/null; python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r2j_bench_gs_nolate.json') if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'])
P
