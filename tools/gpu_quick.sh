# usage: bash tools/gpu_quick.sh "<pytest -k expr or empty>" "<prof_kernel tags...>"
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
if [ -n "$1" ]; then timeout 900 python -m pytest tests -x -q -m gpu -k "$1" 2>&1 | tail -8; fi
for k in $2; do timeout 120 python tools/prof_kernel.py $k 10; done
if [ -n "$3" ]; then eval "$3"; fi
