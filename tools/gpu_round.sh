#!/bin/bash
# One gpurun call: box facts, GPU tests, smoke, bench (reference arm first), launch list, ncu capture.
# Usage (from the repo root on the GPU box): bash tools/gpu_round.sh [tag] [bench args...]
TAG=${1:-r01}; shift
mkdir -p gpurun_out
O=gpurun_out
{ nvidia-smi; nproc; lscpu | head -20; free -g; ls /root/reference baseline/_ref 2>&1 | head; } > $O/box_$TAG.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke_$TAG.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu_$TAG.log
tail -5 $O/pytest_gpu_$TAG.log
[ -n "$SKIP_REF" ] || timeout 900 python bench.py --impl reference "$@" > $O/bench_ref_$TAG.json 2> $O/bench_ref_$TAG.err; echo "bench ref rc=$?"
timeout 1200 python bench.py "$@" > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"
cat $O/bench_$TAG.json | cut -c1-3000; tail -5 $O/bench_$TAG.err
