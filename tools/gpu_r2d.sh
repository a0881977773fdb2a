#!/bin/bash
# round 2, pass d (2 GPUs): multi-rank workloads — C5 with the all-reduce inside the graph, C3 job, C2 with one gather
mkdir -p gpurun_out
N=${1:-2}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@"; }
timeout 600 python -m pytest tests/test_train_gpu.py -m gpu -q -x > gpurun_out/r2d_pytest_train.log 2>&1; tail -3 gpurun_out/r2d_pytest_train.log
timeout 300 python bench.py --workload c5 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2d_bench_c5_n1.json 2> gpurun_out/r2d_bench_c5_n1.err; echo "c5 n1 rc=$?"
timeout 600 bash -c "$(declare -f run); N=$N; run --workload c5 --steps 30 --warmup 5" > gpurun_out/r2d_bench_c5_n$N.json 2> gpurun_out/r2d_bench_c5_n$N.err; echo "c5 n$N rc=$?"; tail -5 gpurun_out/r2d_bench_c5_n$N.err
timeout 600 bash -c "$(declare -f run); N=$N; run --workload c3 --steps 2 --warmup 1" > gpurun_out/r2d_bench_c3_n$N.json 2> gpurun_out/r2d_bench_c3_n$N.err; echo "c3 n$N rc=$?"; tail -3 gpurun_out/r2d_bench_c3_n$N.err
timeout 600 bash -c "$(declare -f run); N=$N; run --steps 20 --warmup 3" > gpurun_out/r2d_bench_c2_n$N.json 2> gpurun_out/r2d_bench_c2_n$N.err; echo "c2 n$N rc=$?"; tail -3 gpurun_out/r2d_bench_c2_n$N.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2d_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get('roofline') or {}
        print(f, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', (d.get('e2e') or {}).get('value'), 'frac', r.get('frac'), 'whole', r.get('whole_step_frac'), {k:v for k,v in d['config'].items() if k in ('replicas_identical','scores_identical_on_all_ranks','final_loss')})
    except Exception as e:
        print(f, 'ERR', e)
PY
