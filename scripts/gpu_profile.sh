#!/bin/bash
# Runs on the GPU box (via gpurun): full GPU test-suite, the bench line (ours + reference arm), an ncu
# launch list of one single-stream transcribe and full captures of the two dominant kernels.
# Usage: bash scripts/gpu_profile.sh <tag> [skip-tests]
TAG=${1:-r01}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
if [ "$2" != "skip-tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15
fi
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench exit $?"; tail -3 gpurun_out/bench_${TAG}.err; cat gpurun_out/bench_${TAG}.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2>> gpurun_out/bench_${TAG}.err
echo "reference arm exit $?"; cat gpurun_out/bench_${TAG}_reference.json
# launch list: mel + encoder + prefill + the first decode steps of one single-stream transcribe
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_issued.avg.pct_of_peak_sustained_active \
  --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv -c 1400 \
  python scripts/profile_decode.py --eager > gpurun_out/profile_${TAG}.log 2>&1
echo "ncu launch list exit $?"; tail -1 gpurun_out/profile_${TAG}.log
# full captures: decode matvec (M=1) from the decode loop, and the tcgen05 GEMM from the encoder
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:q4_matvec_tc -s 300 -c 4 -o gpurun_out/matvec_tc_${TAG} \
  python scripts/profile_decode.py --eager >> gpurun_out/profile_${TAG}.log 2>&1
echo "ncu matvec exit $?"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:gemm_tc5 -s 4 -c 3 -o gpurun_out/gemm_tc5_${TAG} \
  python scripts/profile_decode.py --region encode --streams 8 >> gpurun_out/profile_${TAG}.log 2>&1
echo "ncu gemm exit $?"
# the persistent decode-step kernel (B = 8: the roofline launch of bench.py)
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:decode_mega -s 20 -c 1 -o gpurun_out/mega_${TAG}_b8 \
  python scripts/profile_decode.py --streams 8 >> gpurun_out/profile_${TAG}.log 2>&1
echo "ncu mega exit $?"
timeout 300 python scripts/mega_trace.py --streams 8 > gpurun_out/mega_trace_${TAG}_b8.txt 2>&1
timeout 300 python scripts/mega_trace.py --streams 1 > gpurun_out/mega_trace_${TAG}_b1.txt 2>&1
tail -12 gpurun_out/mega_trace_${TAG}_b8.txt
ls -la gpurun_out | tail -12
