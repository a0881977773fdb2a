# ViT-L/14 (config C4) check: attention kernel tests, golden parity, bench line
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k mha 2>&1 | tail -8
timeout 300 python -m pytest tests/test_parity_gpu.py -q -m gpu -k vitl14 2>&1 | tail -3
timeout 600 python bench.py --arch ViT-L/14 --clips 32 --frames 16 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_vitl.json 2> gpurun_out/bench_vitl.err; tail -3 gpurun_out/bench_vitl.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_vitl.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["whole_step_frac"], d["roofline"]["frac"])
print(d["roofline"]["by_kernel_ms_per_step"])
PY
