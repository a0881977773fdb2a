cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "resid_ln or lnfold or gemm" 2>&1 | tail -8
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu 2>&1 | tail -8
for mode in 0 1; do
  DFD_LN_FUSE=$mode timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/lnfuse_$mode.json 2> gpurun_out/lnfuse_$mode.err
  tail -2 gpurun_out/lnfuse_$mode.err
  python -c "
import json; d=json.loads(open('gpurun_out/lnfuse_$mode.json').read().strip().splitlines()[-1]); print('fuse=$mode', 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), d['clocks']); print(d['roofline']['by_kernel_ms_per_step'])"
done
