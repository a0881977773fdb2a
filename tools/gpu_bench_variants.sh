# extra bench lines: shipped-config shape (last-6 taps + nln adapter) at N=1
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --adapter 768-x-768-nln --taps 6,7,8,9,10,11 > gpurun_out/bench_adapter_nln.json 2> gpurun_out/bench_adapter_nln.err; tail -3 gpurun_out/bench_adapter_nln.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --taps 6,7,8,9,10,11 > gpurun_out/bench_last6.json 2> gpurun_out/bench_last6.err; tail -3 gpurun_out/bench_last6.err
python - <<'PY'
import json
for f in ("bench_adapter_nln","bench_last6"):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(f,'value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1))
    print(d['roofline']['by_kernel_ms_per_step'])
PY
