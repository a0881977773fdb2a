cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
TAG=${1:-r1b}
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_2sm_kernel -s 9 -c 4 -o gpurun_out/${TAG}_prof_gemm $CMD > gpurun_out/${TAG}_ncu_gemm.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_gemm.log
ncu --set full --clock-control none --import-source on -k regex:layernorm_kernel -s 5 -c 2 -o gpurun_out/${TAG}_prof_ln $CMD > gpurun_out/${TAG}_ncu_ln.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_ln.log
