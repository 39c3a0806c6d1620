#!/bin/bash
# Round-2 ncu evidence (one gpurun call; every ncu run follows a plain run of the same command line that exited 0):
#   1. launch list of one timed bench step  2. --set full of one denoiser layer's kernels  3. --set full of the predictor's kernels
TAG=${1:-r2}
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --ncu"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo "launch list: exit $? lines $(wc -l < gpurun_out/launches_$TAG.csv)"
# one denoiser layer in the middle of the second evaluation: skip the first 100 matching launches, take 12
ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'gemm2_kernel|gemmln3_kernel|attention_tc' -s 100 -c 12 -f -o gpurun_out/prof_${TAG}_layer $CMD > gpurun_out/ncu_layer_$TAG.log 2>&1
echo "layer: exit $?"; ls -la gpurun_out/prof_${TAG}_layer.ncu-rep
ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'lstm_tc_kernel|adaln_pred_kernel|dur_head2_kernel|style_pool_attn2_kernel|split3_rows_kernel|lens_perm_kernel|ln_mod_kernel|cast_pool_kernel|init_state_kernel|cvec_kernel' \
    -c 30 -f -o gpurun_out/prof_${TAG}_pred $CMD > gpurun_out/ncu_pred_$TAG.log 2>&1
echo "predictor: exit $?"; ls -la gpurun_out/prof_${TAG}_pred.ncu-rep
