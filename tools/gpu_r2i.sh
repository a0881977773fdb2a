#!/bin/bash
# round 2, pass i: full GPU suite on the final binary + every bench workload at N=1 + reference arm
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -m gpu -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
tail -6 gpurun_out/r2i_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2i_smoke.log 2>&1; tail -1 gpurun_out/r2i_smoke.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2i_bench_reference_arm.json 2> gpurun_out/r2i_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2i_bench_c2_n1.json 2> gpurun_out/r2i_c2.err; echo "c2 rc=$?"
timeout 600 python bench.py --workload c4 --steps 10 --warmup 3 > gpurun_out/r2i_bench_c4_n1.json 2> gpurun_out/r2i_c4.err; echo "c4 rc=$?"
timeout 600 python bench.py --workload c3 --steps 2 --warmup 1 > gpurun_out/r2i_bench_c3_n1.json 2> gpurun_out/r2i_c3.err; echo "c3 rc=$?"
timeout 600 python bench.py --workload c5 --steps 30 --warmup 5 > gpurun_out/r2i_bench_c5_n1.json 2> gpurun_out/r2i_c5.err; echo "c5 rc=$?"
timeout 300 python bench.py --clips 8 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2i_bench_c2_8clips.json 2> /dev/null; echo "c2-8 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2i_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get('roofline') or {}
        print(f, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', (d.get('e2e') or {}).get('value'), 'frac', r.get('frac'), 'whole', r.get('whole_step_frac'), (d.get('cpu_baseline') or {}).get('value'), (d.get('clocks') or {}).get('sm_mhz'))
    except Exception as e:
        print(f, 'ERR', e)
PY
