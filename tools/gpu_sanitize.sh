#!/bin/bash
# compute-sanitizer over the GPU tests of the build that travelled. memcheck: every GPU test (about 3 minutes for the 58 of
# round 2, the 64 streams x 60 s case included). racecheck + synccheck: the golden and multi-call cases.
# Usage: bash tools/gpu_sanitize.sh TAG [memcheck timeout s]
TAG=${1:-r02}; TM=${2:-1500}; O=gpurun_out; mkdir -p $O
timeout $TM compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q --durations=8 > $O/memcheck_$TAG.log 2>&1
echo "memcheck rc=$?" | tee -a $O/memcheck_$TAG.log; tail -4 $O/memcheck_$TAG.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q \
    -k "golden or consecutive or automation" > $O/racecheck_$TAG.log 2>&1
echo "racecheck rc=$?" | tee -a $O/racecheck_$TAG.log; tail -4 $O/racecheck_$TAG.log
timeout 400 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q \
    -k "golden" > $O/synccheck_$TAG.log 2>&1
echo "synccheck rc=$?" | tee -a $O/synccheck_$TAG.log; tail -4 $O/synccheck_$TAG.log
