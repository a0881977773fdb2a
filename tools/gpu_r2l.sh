#!/bin/bash
# round 2, pass l: single-call training forward (dfd_train_forward) + weight gradients on the side stream:
# GPU suite, then C5 A/B on one box (alternating), C2 sanity
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
tail -5 gpurun_out/r2l_pytest.log
for rep in 1 2; do
  for cfg in "0 0" "1 0" "0 1" "1 1"; do
    set -- $cfg
    DFD_OVERLAP=$1 DFD_BWD_STREAMS=$2 timeout 200 python bench.py --workload c5 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2l_c5_o$1_s$2_$rep.json 2> gpurun_out/r2l_c5_o$1_s$2_$rep.err
    echo "c5 overlap=$1 bwd_streams=$2 rep=$rep rc=$? $(grep 'c5 e2e' gpurun_out/r2l_c5_o$1_s$2_$rep.err)"
  done
done
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_c2.json 2> gpurun_out/r2l_c2.err; echo "c2 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2l_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round((d.get('e2e') or {}).get('value') or 0,1), (d.get('clocks') or {}).get('sm_mhz'))
    except Exception as e:
        print(f, 'ERR', e)
PY
