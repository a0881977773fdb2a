#!/bin/bash
# round 2, pass n (final binary): GPU suite, smoke, default bench line, launch lists (C2 predict, C5 training step),
# --set full of the encoder layer GEMMs and the encoder LayerNorm (DFD_OVERLAP=0: launch order = program order)
mkdir -p gpurun_out
T0=$(date +%s); lap() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log
tail -4 gpurun_out/r2n_pytest.log; lap pytest
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2n_smoke.log 2>&1; tail -1 gpurun_out/r2n_smoke.log; lap smoke
timeout 300 python bench.py > gpurun_out/r2n_bench_default.json 2> gpurun_out/r2n_bench_default.err; lap "default bench rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2n_launches.csv $CMD > /dev/null 2>&1; lap "launch list c2 rc=$?"
cap() { # name regex skip count env
  env $5 timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o gpurun_out/r2n_prof_$1 $CMD > gpurun_out/r2n_ncu_$1.log 2>&1
  ncu -i gpurun_out/r2n_prof_$1.ncu-rep --page raw --csv > gpurun_out/r2n_raw_$1.csv 2>/dev/null; lap "set full $1 rc=$?"
}
cap gemm gemm_bf16_2sm_kernel 5 4 DFD_X=0
cap ln layernorm_kernel 3 1 DFD_OVERLAP=0
rm -f gpurun_out/r2n_prof_gemm.ncu-rep
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node --csv --log-file gpurun_out/r2n_launches_c5.csv python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; lap "launch list c5 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2n_bench_default.json').read().strip().splitlines()[-1])
r=d['roofline']
print('default bench', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'frac', round(r['frac'],3), 'whole', round(r['whole_step_frac'],3), 'traffic', r['traffic'], d['clocks'], 'cpu', d['cpu_baseline']['value'])
PY
