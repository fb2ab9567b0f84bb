#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,sm__inst_issued.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none --csv --log-file gpurun_out/step3_b8.csv -s 700 -c 150 \
  python scripts/profile_decode.py --eager --streams 8 > gpurun_out/prof3_b8.log 2>&1
echo "list exit $?"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:q4_matvec_tc -s 400 -c 4 -o gpurun_out/tc3_b8 \
  python scripts/profile_decode.py --eager --streams 8 >> gpurun_out/prof3_b8.log 2>&1
echo "full exit $?"
