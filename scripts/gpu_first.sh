set -x
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv
nproc; free -g | head -2
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -5 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40
