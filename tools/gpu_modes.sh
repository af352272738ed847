#!/bin/bash
# two-phase vs single-pass YIN on the default workload, and the streaming configuration
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for m in 2 1; do
VP_YIN_PHASES=$m timeout 600 python bench.py --no-e2e --no-cpu --no-stream --steps 3 --warmup 3 2>/dev/null | python -c "
import sys,json; j=json.loads(sys.stdin.read()); print('phases $m', round(j['value']), round(j['ms_per_step'],1), {k: round(v/j['steps'],1) for k,v in j['roofline']['stage_ms'].items() if k.startswith('yin')})"
done
timeout 600 python bench.py --stream-only 2>/dev/null | python -c "
import sys,json; j=json.loads(sys.stdin.read()); s=j.get('streaming', j); print({k: s[k] for k in ('p50_ms','p99_ms','mean_ms','max_ms') if k in s})"
