# usage: bash tools/gpu_prof.sh <kernel tag for prof_kernel.py> <ncu kernel regex> <out name>
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
python tools/prof_kernel.py $1 5 > gpurun_out/plain_$3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -o gpurun_out/prof_$3 python tools/prof_kernel.py $1 2 > gpurun_out/ncu_$3.log 2>&1
cat gpurun_out/plain_$3.log; tail -3 gpurun_out/ncu_$3.log
