cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for rep in 1 2 3; do for mode in 0 1; do
  DFD_LN_FUSE=$mode timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "
import json; d=json.loads(open('gpurun_out/ab.json').read().strip().splitlines()[-1]); bk=d['roofline']['by_kernel_ms_per_step']; print('rep $rep fuse=$mode', 'ms', round(d['ms_per_step'],3), 'clk', d['clocks']['sm_mhz'], 'kernel-sum', round(sum(bk.values()),2), 'gemm', round(sum(v for k,v in bk.items() if k.startswith('gemm')),2))"
done; done
