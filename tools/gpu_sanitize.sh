#!/bin/bash
# compute-sanitizer over a subset of the parity tests. Usage: bash tools/gpu_sanitize.sh TAG
TAG=${1:-r01}; O=gpurun_out; mkdir -p $O
timeout 500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
    -k "golden or known_answer or consecutive or two_phase or voiced_silent" > $O/memcheck_$TAG.log 2>&1; echo "memcheck rc=$?"; tail -4 $O/memcheck_$TAG.log
timeout 400 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
    -k "golden and (cmaj or chrom or automation)" > $O/racecheck_$TAG.log 2>&1; echo "racecheck rc=$?"; tail -4 $O/racecheck_$TAG.log
