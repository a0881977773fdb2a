#!/bin/bash
# bench.py workloads at N GPUs of one box (launched the way the driver launches them): bash tools/gpu_scale.sh N [workloads]
mkdir -p gpurun_out
N=${1:-2}; shift
WL=${@:-c2 c3 c5 c4}
for w in $WL; do
  case $w in c2) A="--steps 20 --warmup 3";; c3) A="--workload c3 --steps 2 --warmup 1";; c5) A="--workload c5 --steps 30 --warmup 5";; c4) A="--workload c4 --steps 10 --warmup 3";; esac
  timeout ${SCALE_TIMEOUT:-240} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N $A > gpurun_out/r2_bench_${w}_n$N.json 2> gpurun_out/r2_bench_${w}_n$N.err
  echo "$w n$N rc=$?"; tail -2 gpurun_out/r2_bench_${w}_n$N.err | cut -c1-200
done
python - "$N" <<'PY'
import json,glob,sys
for f in sorted(glob.glob('gpurun_out/r2_bench_*_n%s.json' % sys.argv[1])):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get('roofline') or {}
        print(f, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round((d.get('e2e') or {}).get('value') or 0,1), 'frac', r.get('frac'), 'whole', r.get('whole_step_frac'), {k:v for k,v in d['config'].items() if k in ('replicas_identical','scores_identical_on_all_ranks')}, d['clocks'])
    except Exception as e:
        print(f, 'ERR', e)
PY
