#!/bin/bash
# round-2 ncu evidence (never a bench value): launch list of one single-stream transcribe, full captures of the persistent
# decode kernel (B = 8: the roofline launch; DRAM bytes -> profiles/traffic.json), the tcgen05 GEMM (tensor-pipe
# utilisation), the mel front-end kernels (achieved HBM GB/s) and the decode matvec (M = 1).
TAG=${1:-r02}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
MET=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_issued.avg.pct_of_peak_sustained_active
timeout 900 ncu --profile-from-start off --metrics $MET --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv -c 1500 \
  python scripts/profile_decode.py --eager > gpurun_out/profile_${TAG}.log 2>&1
echo "launch list exit $?"; tail -1 gpurun_out/profile_${TAG}.log
timeout 900 ncu --profile-from-start off --metrics $MET --clock-control none --csv --log-file gpurun_out/launches_${TAG}_b8.csv -c 700 \
  python scripts/profile_decode.py --streams 8 >> gpurun_out/profile_${TAG}.log 2>&1
echo "launch list b8 exit $?"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:decode_mega -s 20 -c 1 -o gpurun_out/mega_${TAG}_b8 \
  python scripts/profile_decode.py --streams 8 >> gpurun_out/profile_${TAG}.log 2>&1
echo "ncu mega exit $?"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc5 -s 4 -c 3 -o gpurun_out/gemm_tc5_${TAG} \
  python scripts/profile_decode.py --region encode --streams 8 >> gpurun_out/profile_${TAG}.log 2>&1
echo "ncu gemm exit $?"
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:'mel_kernel|peak_max|scale_pad' -c 3 -o gpurun_out/mel_${TAG} \
  python scripts/profile_decode.py --streams 8 >> gpurun_out/profile_${TAG}.log 2>&1
echo "ncu mel exit $?"
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:q4_matvec_tc -s 300 -c 4 -o gpurun_out/matvec_tc_${TAG} \
  python scripts/profile_decode.py --eager >> gpurun_out/profile_${TAG}.log 2>&1
echo "ncu matvec exit $?"
ls -la gpurun_out | tail -8
