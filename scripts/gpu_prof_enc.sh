#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,sm__inst_issued.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum \
  --clock-control none --csv --log-file gpurun_out/enc_b8.csv -c 120 \
  python scripts/profile_decode.py --region encode --streams 8 > gpurun_out/prof_enc.log 2>&1
echo "exit $?"; tail -2 gpurun_out/prof_enc.log
