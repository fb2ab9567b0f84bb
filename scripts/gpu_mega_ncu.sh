#!/bin/bash
# ncu full captures (with source counters) of one persistent decode-step kernel at B=1 and B=8.
TAG=${1:-mega}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for B in 1 8; do
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:decode_mega -s 20 -c 1 -o gpurun_out/mega_${TAG}_b${B} \
  python scripts/profile_decode.py --streams $B > gpurun_out/ncu_${TAG}_b${B}.log 2>&1
echo "ncu B=$B exit $?"; tail -2 gpurun_out/ncu_${TAG}_b${B}.log
done
ls -la gpurun_out | tail -5
