# Round-1c profile: plain bench run, ncu launch list of the same command, --set full captures of the top kernels,
# plus a launch list of the shipped-config variant (last-6 taps + nln adapter).
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
TAG=r1c
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
# launches to skip: weight packing (~64) + 3 warm-up predicts (131 each); then 2 timed predicts = 262 launches
ncu --metrics gpu__time_duration.sum --clock-control none -s 460 -c 262 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
for spec in "gemm_bf16_2sm_kernel<0>:gemm_qkv:6" "gemm_bf16_2sm_kernel<1>:gemm_fc:6" "gemm_bf16_2sm_kernel<3>:gemm_resid:12" "mha_fwd_tc2:mha:6" "dec_attn_stream:dec_attn:6" "layernorm_kernel<6, 0>:ln:6" "linear_partial:dec_linear:6"; do
  k="${spec%%:*}"; rest="${spec#*:}"; n="${rest%%:*}"; skip="${rest##*:}"
  ncu --set full --clock-control none --import-source on -k "regex:${k//</.}" -s $skip -c 2 -o gpurun_out/${TAG}_prof_${n} $CMD > gpurun_out/${TAG}_ncu_${n}.log 2>&1
  tail -1 gpurun_out/${TAG}_ncu_${n}.log
done
CMD2="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --adapter 768-x-768-nln --taps 6,7,8,9,10,11"
$CMD2 > gpurun_out/${TAG}_plain_adapter.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 400 --csv --log-file gpurun_out/${TAG}_launches_adapter.csv $CMD2 > gpurun_out/${TAG}_ncu_launches_adapter.log 2>&1
CMD3="python bench.py --arch ViT-L/14 --clips 32 --frames 16 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:mha_fwd_tc3 -s 30 -c 1 -o gpurun_out/${TAG}_prof_mha_l $CMD3 > gpurun_out/${TAG}_ncu_mha_l.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_mha_l.log
ls -la gpurun_out | grep ${TAG}
