#!/bin/bash
# single GPU, more views per step than one chunk holds: isolates the chunked path from multi-process effects
mkdir -p gpurun_out
for v in 8 16 25; do
  echo "=== N=1 --views $v"; timeout 200 python bench.py --views $v --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['ms_per_step']/$v)"
done
