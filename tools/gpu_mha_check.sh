set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "mha" 2>&1 | tail -15 > gpurun_out/t_mha.log
cat gpurun_out/t_mha.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v2.log 2> gpurun_out/bench_v2.err
cat gpurun_out/bench_v2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['by_kernel_ms_per_step'], d['clocks'])"
tail -3 gpurun_out/bench_v2.err
