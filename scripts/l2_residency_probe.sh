#!/bin/bash
# L2 residency experiments for the streaming matvec on one GPU: eviction hints (MB kept) x traversal direction
set -u
timeout 400 python -m pytest tests/test_gpu_rhs.py tests/test_gpu_greens.py tests/test_gpu_solve.py -m gpu -q -x -k "not twins" 2>&1 | tail -3
for cfg in "0 0" "0 1" "96 1" "96 0" "64 0" "112 0" "48 0"; do
  set -- $cfg
  echo "keep_mb=$1 pingpong=$2:"; OQ_MATVEC_KEEP_MB=$1 OQ_MATVEC_PINGPONG=$2 timeout 200 python scripts/shard_probe.py 2>&1 | tail -4
done
