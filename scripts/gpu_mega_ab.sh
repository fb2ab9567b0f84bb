#!/bin/bash
# A/B of the persistent kernel's experiment switches through the phase trace (B=1 and B=8).
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 300 python -m pytest tests/test_model_gpu.py -x -q -k "persistent" 2>&1 | tail -3
for F in ${@:-0 8}; do
  VOX_MEGA_FLAGS=$F timeout 300 python scripts/mega_trace.py --streams 1 2>&1 | tail -11
  VOX_MEGA_FLAGS=$F timeout 300 python scripts/mega_trace.py --streams 8 2>&1 | tail -11
done
