cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
TAG=r1c
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
# GEMM launches per predict: patch-embed <2>, then per layer QKV <0>, out-proj <3>, c_fc <1>, c_proj <3> (42 per predict)
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_2sm_kernel -s 127 -c 4 -o gpurun_out/${TAG}_prof_gemm $CMD > gpurun_out/${TAG}_ncu_gemm.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_gemm.log
# LayerNorm launches per predict: ln_pre, 21 encoder, 13 decoder
ncu --set full --clock-control none --import-source on -k regex:layernorm_kernel -s 106 -c 2 -o gpurun_out/${TAG}_prof_ln $CMD > gpurun_out/${TAG}_ncu_ln.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_ln.log
