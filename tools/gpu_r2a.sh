#!/bin/bash
# round 2, pass a: new full-size parity tests + bench workloads + LN-fold A/B on one box
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench_c2.json 2> gpurun_out/r2a_bench_c2.err; echo "c2 rc=$?"
for mode in 0 2 1 0 2; do
  DFD_LN_FUSE=$mode timeout 300 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2a_lnfuse_${mode}_$RANDOM.json 2>> gpurun_out/r2a_lnfuse.err; echo "lnfuse $mode rc=$?"
done
timeout 600 python bench.py --workload c4 --steps 10 --warmup 3 > gpurun_out/r2a_bench_c4.json 2> gpurun_out/r2a_bench_c4.err; echo "c4 rc=$?"
timeout 600 python bench.py --workload c3 --steps 3 --warmup 1 > gpurun_out/r2a_bench_c3.json 2> gpurun_out/r2a_bench_c3.err; echo "c3 rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2a_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get('roofline') or {}
        print(f, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', (d.get('e2e') or {}).get('value'), 'frac', r.get('frac'), 'whole', r.get('whole_step_frac'), d.get('clocks',{}).get('sm_mhz'))
    except Exception as e:
        print(f, 'ERR', e)
PY
