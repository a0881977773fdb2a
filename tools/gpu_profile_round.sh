# Round profile: plain bench run, ncu launch list of the same command, and --set full captures of the top kernels.
# usage: bash tools/gpu_profile_round.sh <tag>      (outputs under gpurun_out/<tag>_*)
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
TAG=${1:-r1b}
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
# warm-up launches to skip: weight packing (~64) + 3 warm-up predicts (131 each)
ncu --metrics gpu__time_duration.sum --clock-control none -s 460 -c 262 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
for spec in "gemm_bf16_2sm_kernel<0>:gemm_qkv" "gemm_bf16_2sm_kernel<1>:gemm_fc" "gemm_bf16_2sm_kernel<3>:gemm_resid" "mha_fwd_tc2:mha" "dec_attn_stream:dec_attn" "layernorm_kernel<6, 0>:ln" "linear_partial:dec_linear"; do
  k="${spec%%:*}"; n="${spec##*:}"
  ncu --set full --clock-control none --import-source on -k "regex:${k//</.}" -s 6 -c 2 -o gpurun_out/${TAG}_prof_${n} $CMD > gpurun_out/${TAG}_ncu_${n}.log 2>&1
  tail -1 gpurun_out/${TAG}_ncu_${n}.log
done
ls -la gpurun_out | grep ${TAG}
