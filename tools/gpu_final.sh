#!/bin/bash
# Round-end evidence in one gpurun call: box facts, smoke, GPU tests, reference arm, full bench (tools/gpu_round.sh), launch
# list + ncu --set full of every stage kernel (tools/gpu_ncu.sh), executed instruction mix of the four workloads
# (tools/gpu_mix.sh). Usage: bash tools/gpu_final.sh TAG
TAG=${1:-r02fin}
O=gpurun_out; mkdir -p $O
bash tools/gpu_round.sh $TAG
bash tools/gpu_ncu.sh $TAG "k_gate_partial|k_yin_corr|k_yin_decide|k_marks|k_voc_autocorr2|k_voc_levinson|k_voc_synth_rows|k_voc_orphans|k_pitch_autocorr|k_pitch_psola|k_pitch_iir|k_mix"
bash tools/gpu_mix.sh $TAG
