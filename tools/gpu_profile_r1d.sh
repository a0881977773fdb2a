# Round-1d measurement pass (after the host-pipeline / single-call work): headline bench line with both baselines,
# reference arm, ncu launch list of the same bench command, the other configurations.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
TAG=r1d
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err || { echo "bench failed"; tail -5 gpurun_out/${TAG}_bench_n1.err; exit 1; }
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_reference_arm.err
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_launches.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --taps 6,7,8,9,10,11 > gpurun_out/${TAG}_bench_n1_last6taps.json 2>/dev/null
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --taps 6,7,8,9,10,11 --adapter 768-x-768-nln > gpurun_out/${TAG}_bench_n1_last6taps_adapter_nln.json 2>/dev/null
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --taps 6,7,8,9,10,11 --adapter 768-x-768-nln --frames 20 --clips 24 > gpurun_out/${TAG}_bench_n1_shipped_t20_last6_nln.json 2>/dev/null
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --arch ViT-L/14 --clips 32 --frames 16 > gpurun_out/${TAG}_bench_n1_vitl14_c4.json 2>/dev/null
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --clips 8 > gpurun_out/${TAG}_bench_n1_8clips.json 2>/dev/null
DFD_OVERLAP=0 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --clips 8 --no-e2e > gpurun_out/${TAG}_bench_n1_8clips_two_call.json 2>/dev/null
python tools/bench_configs.py both > gpurun_out/${TAG}_configs_c3_c5.json 2> gpurun_out/${TAG}_configs.err
for f in gpurun_out/${TAG}_bench_*.json; do python - "$f" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
e = d.get("e2e") or {}
r = d.get("roofline") or {}
print(sys.argv[1].split("/")[-1], "value %.1f" % d["value"], "ms %.2f" % d["ms_per_step"], "e2e %.1f" % e.get("value", 0),
      "gemm_frac %s" % r.get("frac"), "step_frac %s" % r.get("whole_step_frac"), d.get("clocks"))
PY
done
cat gpurun_out/${TAG}_configs_c3_c5.json
