#!/bin/bash
# round 2, pass b: all GPU tests (new goldens, native decoder backward, resize), C5 / C3 workloads
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -15 gpurun_out/r2b_pytest.log
timeout 600 python bench.py --workload c5 --steps 30 --warmup 5 > gpurun_out/r2b_bench_c5.json 2> gpurun_out/r2b_bench_c5.err; echo "c5 rc=$?"; tail -3 gpurun_out/r2b_bench_c5.err
DFD_NATIVE_DECODER_BWD=0 timeout 600 python bench.py --workload c5 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_bench_c5_torchchain.json 2> gpurun_out/r2b_bench_c5_torchchain.err; echo "c5 torch rc=$?"
timeout 600 python bench.py --workload c3 --steps 2 --warmup 1 > gpurun_out/r2b_bench_c3.json 2> gpurun_out/r2b_bench_c3.err; echo "c3 rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2b_bench_ref.json 2> gpurun_out/r2b_bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2b_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get('roofline') or {}
        print(f, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', (d.get('e2e') or {}).get('value'), 'frac', r.get('frac'), 'whole', r.get('whole_step_frac'), (d.get('cpu_baseline') or {}).get('kind'))
    except Exception as e:
        print(f, 'ERR', e)
PY
