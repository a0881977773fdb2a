#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_modes_gpu.py tests/test_host_logic.py tests/test_train_gpu.py -m gpu -q > gpurun_out/r2j_pytest.log 2>&1; tail -4 gpurun_out/r2j_pytest.log
for i in 1 2 3; do timeout 300 python bench.py --workload c5 --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c5', round(d['ms_per_step'],3), round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"; done
