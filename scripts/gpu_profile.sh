#!/bin/bash
# Runs on the GPU box (via gpurun): tests, bench line, ncu launch list + full capture of the top
# kernel.  Usage: bash scripts/gpu_profile.sh <round-tag> [skip-tests]
TAG=${1:-r01}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
if [ "$2" != "skip-tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15
fi
# the bench line (N=1)
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench exit $?"; tail -3 gpurun_out/bench_${TAG}.err; cat gpurun_out/bench_${TAG}.json
# launch list of one whole single-stream transcribe (eager launches so every kernel is a launch)
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
  python scripts/profile_decode.py --eager > gpurun_out/profile_${TAG}.log 2>&1
echo "ncu launch list exit $?"; tail -2 gpurun_out/profile_${TAG}.log
# full capture of the dominant kernel (decode matvec, M=1): 3 launches from the decode loop
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:q4_matvec -s 400 -c 3 -o gpurun_out/matvec_${TAG} \
  python scripts/profile_decode.py --eager >> gpurun_out/profile_${TAG}.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out
