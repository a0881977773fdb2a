#!/bin/bash
# round 2, pass m: TrainStep(pipeline=True) — tests, then C5 A/B (pipeline off / on, alternating) on one box
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_train_gpu.py -m gpu -q -x > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log
tail -15 gpurun_out/r2m_pytest.log
for rep in 1 2; do for pl in 0 1; do
  timeout 120 python bench.py --workload c5 --c5-pipeline $pl --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2m_c5_p${pl}_$rep.json 2> gpurun_out/r2m_c5_p${pl}_$rep.err
  echo "c5 pipeline=$pl rep=$rep rc=$? $(grep 'c5 e2e' gpurun_out/r2m_c5_p${pl}_$rep.err) $(tail -1 gpurun_out/r2m_c5_p${pl}_$rep.err | cut -c1-300)"
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2m_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round((d.get('e2e') or {}).get('value') or 0,1), (d.get('clocks') or {}).get('sm_mhz'), d['config'].get('final_loss'))
    except Exception as e:
        print(f, 'ERR', e)
PY
