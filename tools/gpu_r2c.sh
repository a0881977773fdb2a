#!/bin/bash
# round 2, pass c: full GPU suite again (epilogue refactor, tmap cache, capture fixes) + C5 / C2 benches
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -12 gpurun_out/r2c_pytest.log
timeout 600 python bench.py --workload c5 --steps 30 --warmup 5 > gpurun_out/r2c_bench_c5.json 2> gpurun_out/r2c_bench_c5.err; echo "c5 rc=$?"; tail -3 gpurun_out/r2c_bench_c5.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_c2.json 2> gpurun_out/r2c_bench_c2.err; echo "c2 rc=$?"
timeout 300 python bench.py --clips 8 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench_c2_8clips.json 2> gpurun_out/r2c_bench_c2_8clips.err; echo "c2-8 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get('roofline') or {}
        print(f, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', (d.get('e2e') or {}).get('value'), 'frac', r.get('frac'), 'whole', r.get('whole_step_frac'), r.get('by_kernel_ms_per_step'))
    except Exception as e:
        print(f, 'ERR', e)
PY
