#!/bin/bash
# round 2, pass e: uniform-operand MMA issue in every tcgen05 kernel: kernel tests + golden parity + bench c2 / c4
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_fullsize_gpu.py tests/test_adapter_gpu.py -m gpu -q -x > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -5 gpurun_out/r2e_pytest.log
for k in mha mha_l gemm_qkv gemm_fc gemm_proj gemm_out; do timeout 120 python tools/prof_kernel.py $k 30 2>&1 | tail -1; done | tee gpurun_out/r2e_kernels.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench_c2.json 2> gpurun_out/r2e_bench_c2.err; echo "c2 rc=$?"
timeout 600 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench_c4.json 2> gpurun_out/r2e_bench_c4.err; echo "c4 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2e_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get('roofline') or {}
        print(f, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', (d.get('e2e') or {}).get('value'), 'frac', r.get('frac'), 'whole', r.get('whole_step_frac'), r.get('by_kernel_ms_per_step'), d['clocks'])
    except Exception as e:
        print(f, 'ERR', e)
PY
