#!/bin/bash
# round 2, pass f: LayerNorm-fold A/B on the current binary (alternating, one box)
mkdir -p gpurun_out
for rep in 1 2; do for mode in 0 1 2; do
  DFD_LN_FUSE=$mode timeout 300 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2f_lnfuse_${mode}_${rep}.json 2>> gpurun_out/r2f_lnfuse.err; echo "lnfuse $mode rc=$?"
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2f_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get('roofline') or {}
        print(f, round(d['value'],1), round(d['ms_per_step'],3), 'frac', round(r.get('frac'),3), 'whole', round(r.get('whole_step_frac'),3), r.get('by_kernel_ms_per_step'), d['clocks']['sm_mhz'])
    except Exception as e:
        print(f, 'ERR', e)
PY
