#!/bin/bash
# per-kernel timing of the fused decode step (B=1 and B=8), ~2 steps each, + full capture of the M=8 w13 matvec
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for B in 1 8; do
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,sm__inst_issued.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none --csv --log-file gpurun_out/step_b${B}.csv -s 520 -c 290 \
  python scripts/profile_decode.py --eager --streams $B > gpurun_out/prof2_b${B}.log 2>&1
echo "B=$B exit $?"; tail -1 gpurun_out/prof2_b${B}.log
done
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:q4_matvec_tc -s 300 -c 6 -o gpurun_out/tc_b8 \
  python scripts/profile_decode.py --eager --streams 8 >> gpurun_out/prof2_b8.log 2>&1
echo "full exit $?"
ls -la gpurun_out | tail -8
