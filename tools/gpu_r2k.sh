#!/bin/bash
# round 2, pass k (evidence pass on the final binary, 1 GPU): every bench workload at N=1 + the reference arm, the ncu
# launch list of the bench command, `--set full` captures of the kernels DESIGN.md §4 names, then the GPU test suite.
# Ordered by value: each stage leaves its files in gpurun_out/ even if a later one runs out of time.
mkdir -p gpurun_out
T0=$(date +%s); lap() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2k_smoke.log 2>&1; tail -1 gpurun_out/r2k_smoke.log; lap smoke
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_c2_n1.json 2> gpurun_out/r2k_c2.err; lap "c2 rc=$?"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2k_ref.err; lap "ref rc=$?"
timeout 400 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c4_n1.json 2> gpurun_out/r2k_c4.err; lap "c4 rc=$?"
timeout 400 python bench.py --workload c5 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_c5_n1.json 2> gpurun_out/r2k_c5.err; lap "c5 rc=$?"
timeout 500 python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench_c3_n1.json 2> gpurun_out/r2k_c3.err; lap "c3 rc=$?"
timeout 200 python bench.py --clips 8 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_c2_8clips.json 2> /dev/null; lap "c2-8 rc=$?"
# launch list of the bench command (cold-cache, serialised: shares only)
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2k_ncu_launches.log 2>&1; lap "launch list rc=$?"
# --set full captures inside the same bench command: one encoder layer's four GEMMs (layer 1: launches 5..8 after the
# patch GEMM + layer 0), attention, LayerNorm, decoder attention, decoder linear
cap() { # name regex skip count
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o gpurun_out/r2_prof_$1 $CMD > gpurun_out/r2k_ncu_$1.log 2>&1
  ncu -i gpurun_out/r2_prof_$1.ncu-rep --page raw --csv > gpurun_out/r2_raw_$1.csv 2>/dev/null; lap "set full $1 rc=$?"
}
cap gemm gemm_bf16_2sm_kernel 5 4
cap mha mha_fwd_tc2_kernel 2 1
cap ln layernorm_kernel 3 1
cap dec dec_attn_stream_kernel 1 1
cap declin linear_partial_kernel 4 2
ncu -i gpurun_out/r2_prof_mha.ncu-rep --page source --csv > gpurun_out/r2_src_mha.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
# keep the merge under 64 MiB: drop the largest reports if needed (the raw CSVs carry the numbers)
SZ=$(du -sm gpurun_out | cut -f1); if [ "$SZ" -gt 55 ]; then rm -f gpurun_out/r2_prof_gemm.ncu-rep gpurun_out/r2_prof_ln.ncu-rep; fi
timeout ${PYTEST_TIMEOUT:-900} python -m pytest tests -m gpu -q --durations=12 > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
tail -18 gpurun_out/r2k_pytest.log; lap pytest
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get('roofline') or {}
        print(f, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', (d.get('e2e') or {}).get('value'), 'frac', r.get('frac'), 'whole', r.get('whole_step_frac'), r.get('by_kernel_ms_per_step'), (d.get('cpu_baseline') or {}).get('value'), (d.get('clocks') or {}).get('sm_mhz'))
    except Exception as e:
        print(f, 'ERR', e)
PY
