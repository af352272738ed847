#!/bin/bash
# Round-end evidence in one gpurun call: box facts, smoke, GPU tests, reference arm, full bench, launch list, ncu --set full
# of every stage kernel, and the other BASELINE configurations. Usage: bash tools/gpu_final.sh TAG
TAG=${1:-r01fin}
O=gpurun_out; mkdir -p $O
bash tools/gpu_round.sh $TAG
bash tools/gpu_ncu.sh $TAG "k_gate_partial|k_yin_corr|k_yin_decide|k_marks|k_voc_autocorr2|k_voc_levinson|k_voc_synth|k_pitch_autocorr|k_pitch_psola|k_pitch_iir|k_mix"
for wl in voc44 pitch44 chain44; do
  timeout 600 python bench.py --workload $wl --no-cpu --no-stream > $O/bench_${wl}_$TAG.json 2> $O/bench_${wl}_$TAG.err; echo "$wl rc=$?"
done
