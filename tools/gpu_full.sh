# full GPU check: all gpu tests, smoke, bench (both arms)
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; tail -2 gpurun_out/bench_full.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_full.log').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'], 'clocks', d['clocks'])
print('roofline', {k:v for k,v in d['roofline'].items() if k!='by_kernel_ms_per_step'})
print(d['roofline']['by_kernel_ms_per_step'])
print('cpu', d['cpu_baseline'])
PY
